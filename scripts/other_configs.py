"""Device-resident throughput on the shapes of BASELINE configs 3, 4 and 5 (scaled to fit a quick run).
Not part of bench.py's contract: a sanity check that the path does not fall off a cliff off config 2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ahocorasick_rs_b200 import AhoCorasick, BytesAhoCorasick, MatchKind, Implementation, workloads as W, _capi
import ctypes as C

def run(name, ac, data, offs, overlapping=False, steps=10):
    d = torch.from_numpy(data).cuda(); o = torch.from_numpy(offs).cuda()
    m, mo, total = ac.scan_device(d, o, overlapping)
    cap = int(total * 1.2) + 1024
    for _ in range(3): ac.scan_device(d, o, overlapping, capacity=cap, sync=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): ac.scan_device(d, o, overlapping, capacity=cap, sync=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = ac._ac.last_stats if hasattr(ac, "_ac") else {}
    print(f"{name}: {len(data)/1e6:.0f} MB, {total} matches, {ms:.3f} ms/step, {len(data)/ms/1e6:.1f} GB/s, stats {st}", flush=True)

only = [a for a in sys.argv[1:] if a in ("c3", "c4", "c5")]
steps = 2 if "--short" in sys.argv else 10
if not only or "c3" in only:
    pats, data, offs = W.config3(n_patterns=10_000, n_lines=400_000)
    run("config3 (10k tokens, LeftmostLongest, 400k x 256 B)", BytesAhoCorasick(pats, MatchKind.LeftmostLongest), data, offs, steps=steps)
if not only or "c5" in only:
    pats, data, offs = W.config5(n_patterns=50_000, n_haystacks=25_000, hay_bytes=4096)
    run("config5 (50k patterns a-z, 25k x 4 KiB)", BytesAhoCorasick(pats), data, offs, steps=steps)
if not only or "c4" in only:
    pats, data = W.config4(n_patterns=100_000, hay_bytes=100_000_000)
    run("config4 (100k patterns, one 100 MB haystack, overlapping)", BytesAhoCorasick(pats, implementation=Implementation.ContiguousNFA), data,
        np.array([0, len(data)], dtype=np.int64), overlapping=True, steps=min(steps, 5))

if "--large" in sys.argv:
    # one haystack above the 2 GiB window (BASELINE config 4's shape at 2.3 GB): two windows, checked against the oracle
    from oracle import Oracle
    pats, data = W.config4(n_patterns=100_000, hay_bytes=2_300_000_000)
    t0 = time.time()
    ac = BytesAhoCorasick(pats, implementation=Implementation.ContiguousNFA)
    d = torch.from_numpy(data).cuda()
    o = torch.tensor([0, len(data)], dtype=torch.int64).cuda()
    m, mo, total = ac.scan_device(d, o, overlapping=True)
    torch.cuda.synchronize()
    print(f"2.3 GB single haystack, overlapping: {total} matches, dtype {m.dtype}, {time.time() - t0:.1f} s incl. build/profile", flush=True)
    tot, counts, rec = Oracle(pats, "Standard").scan_batch(data, np.array([0, len(data)], dtype=np.int64), overlapping=True)
    got = m.cpu().numpy()
    assert tot == total, (tot, total)
    assert np.array_equal(got, rec.astype(np.int64)), "windowed result differs from the oracle"
    print("matches the oracle exactly", flush=True)
