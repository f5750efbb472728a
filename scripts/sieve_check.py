"""First-contact check of the sieve kernel on a GPU: small seeded cases against the oracle with readable diagnostics
(the parity tests proper are tests/test_gpu_parity.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ahocorasick_rs_b200 import AhoCorasick, BytesAhoCorasick, MatchKind, workloads as W, _capi
from oracle import Oracle

KINDS = [MatchKind.Standard, MatchKind.LeftmostFirst, MatchKind.LeftmostLongest]
bad = 0


def check(name, pats, kind, data, offs, overlapping=False, codepoints=False, task=0):
    global bad
    _capi.set_tuning(5, 0, task, 0)
    orc = Oracle(pats, kind.name)
    total, counts, rec = orc.scan_batch(data, offs, overlapping=overlapping, codepoints=codepoints)
    ac = AhoCorasick([p.decode() for p in pats], kind) if codepoints else BytesAhoCorasick(pats, kind)
    m, mo, t = ac.scan_device(torch.from_numpy(np.ascontiguousarray(data)).cuda(), torch.from_numpy(offs).cuda(), overlapping)
    got = m.cpu().numpy().view(np.uint32)
    ok = t == total and np.array_equal(got, rec) and np.array_equal(np.diff(mo.cpu().numpy()), counts.astype(np.int64))
    print(f"{'ok ' if ok else 'BAD'} {name}: kind={kind.name} ovl={overlapping} cp={codepoints} task={task} total={t} expect={total} {ac._ac.last_stats}", flush=True)
    if not ok:
        bad += 1
        n = min(len(got), len(rec))
        diff = np.nonzero((got[:n] != rec[:n]).any(axis=1))[0]
        if len(diff):
            i = int(diff[0])
            print("   first difference at row", i, "got", got[max(0, i - 2):i + 3].tolist(), "expect", rec[max(0, i - 2):i + 3].tolist(), flush=True)
        else:
            print("   common prefix equal; lengths", len(got), len(rec), flush=True)


rng = np.random.default_rng(11)
pats = sorted({bytes(rng.integers(97, 100, size=rng.integers(1, 6)).astype(np.uint8)) for _ in range(40)})
pats += pats[:3]
data, offs = W.ragged(3000, 300, b"abc", seed=12)
for task in (0, 512):
    for kind in KINDS:
        check("ragged abc", pats, kind, data, offs, task=task)
    check("ragged abc", pats, MatchKind.Standard, data, offs, overlapping=True, task=task)
p2, d2, o2 = W.config2(300)
pb = [p.encode() for p in p2]
for task in (0, 512):
    for kind in KINDS:
        check("config2 x300", pb, kind, d2, o2, codepoints=True, task=task)
    check("config2 x300", pb, MatchKind.Standard, d2, o2, overlapping=True, codepoints=True, task=task)
p3, d3, o3 = W.config3(n_patterns=2000, n_lines=4000)
check("config3", p3, MatchKind.LeftmostLongest, d3, o3)
p5, d5, o5 = W.config5(n_patterns=20000, n_haystacks=256, hay_bytes=4096)
check("config5", p5, MatchKind.Standard, d5, o5)
p4, d4 = W.config4(n_patterns=20000, hay_bytes=2_000_017)
check("config4", p4, MatchKind.Standard, d4, np.array([0, len(d4)], dtype=np.int64), overlapping=True)
# unicode-heavy text with tiny haystacks (code point bookkeeping across windows / many haystack starts per window)
text = "".join(rng.choice(list("ab—é☃cd"), size=40_000))
raw = text.encode()
cuts = sorted(set(int(x) for x in rng.integers(0, len(text), size=3000)))
pieces = [text[a:b].encode() for a, b in zip([0] + cuts, cuts + [len(text)])]
dd = np.frombuffer(b"".join(pieces), dtype=np.uint8)
oo = np.zeros(len(pieces) + 1, dtype=np.int64); np.cumsum([len(x) for x in pieces], out=oo[1:])
up = [u.encode() for u in ["a—", "—é", "☃c", "b", "é☃c", "dd"]]
for task in (0, 512):
    for kind in KINDS:
        check("unicode tiny haystacks", up, kind, dd, oo, codepoints=True, task=task)
    check("unicode tiny haystacks", up, MatchKind.Standard, dd, oo, overlapping=True, codepoints=True, task=task)
one = np.array([0, len(dd)], dtype=np.int64)
check("unicode one haystack", up, MatchKind.Standard, dd, one, overlapping=True, codepoints=True)
check("unicode one haystack", up, MatchKind.LeftmostLongest, dd, one, codepoints=True, task=1024)
print("FAILURES:", bad, flush=True)
sys.exit(1 if bad else 0)
