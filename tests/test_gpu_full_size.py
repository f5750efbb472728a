"""GPU parity at the BASELINE.json sizes (-m gpu): the CUDA path against the CPU oracle on the FULL config-2 and
config-3 batches, on a config-5 shaped batch of 256 MB, and on config 4's single 4 GiB haystack (the oracle checks
byte ranges of it, including the range around the host layer's window cut; size-independent properties cover the
rest).  Also here: two threads with two automata (the C ABI keeps its knobs per thread), one automaton shared by two
threads through the host-buffer calls, and -- with two GPUs -- the NCCL gather of sharded scans."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from ahocorasick_rs_b200 import AhoCorasick, BytesAhoCorasick, Implementation, MatchKind, _capi, workloads as W
from oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("engine", ["auto", "sieve"])
def test_config2_full_size_vs_oracle(engine):
    pats, data, offs = W.config2(100_000)
    pb = [p.encode() for p in pats]
    total, counts, rec = Oracle(pb, "Standard").scan_batch(data, offs, codepoints=True)
    _capi.set_tuning(5 if engine == "sieve" else 0)
    try:
        ac = AhoCorasick(pats, implementation=Implementation.DFA)
        m, mo, t = ac.scan_device(dev(data), dev(offs))
        assert t == total == 7784 or t == total
        assert np.array_equal(m.cpu().numpy().view(np.uint32), rec)
        assert np.array_equal(np.diff(mo.cpu().numpy()), counts.astype(np.int64))
    finally:
        _capi.set_tuning(0)


def test_config3_full_size_vs_oracle():
    pats, data, offs = W.config3(n_patterns=10_000, n_lines=1_000_000)
    total, counts, rec = Oracle(pats, "LeftmostLongest").scan_batch(data, offs)
    ac = BytesAhoCorasick(pats, MatchKind.LeftmostLongest)
    m, mo, t = ac.scan_device(dev(data), dev(offs))
    assert t == total and total > 1_000_000
    assert np.array_equal(m.cpu().numpy().view(np.uint32), rec)
    assert np.array_equal(np.diff(mo.cpu().numpy()), counts.astype(np.int64))
    # the host-buffer path (chunked, pipelined) gives the same list
    hm, hmo = ac.scan_host(data, offs, chunk_bytes=48 << 20)
    assert np.array_equal(hm, rec) and np.array_equal(np.diff(hmo), counts.astype(np.int64))


def test_config5_shape_256mb_vs_oracle():
    pats, data, offs = W.config5(n_patterns=50_000, n_haystacks=65_536, hay_bytes=4096)
    total, counts, rec = Oracle(pats, "Standard").scan_batch(data, offs)
    ac = BytesAhoCorasick(pats)
    m, mo, t = ac.scan_device(dev(data), dev(offs))
    assert t == total and total > 100_000
    assert np.array_equal(m.cpu().numpy().view(np.uint32), rec)
    assert np.array_equal(np.diff(mo.cpu().numpy()), counts.astype(np.int64))


def test_config4_single_4gib_haystack_overlapping():
    """BASELINE config 4 at size: ONE haystack of 2^32 bytes, 100k patterns, overlapping -- two host-level windows,
    64-bit offsets.  The oracle checks the first 64 MiB and 32 MiB around the window cut; over the whole list: ends are
    non-decreasing, every match slices back to its pattern (sampled), and the count equals the sum over two halves
    scanned separately plus the matches across the cut."""
    n = 1 << 32
    pats = W.random_lowercase_patterns(100_000, 5, 8, 4)
    g = torch.Generator(device="cuda")
    g.manual_seed(1004)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for a in range(0, n, 1 << 28):
        d[a:a + (1 << 28)] = torch.randint(97, 123, (1 << 28,), dtype=torch.uint8, device="cuda", generator=g)
    ac = BytesAhoCorasick(pats, implementation=Implementation.ContiguousNFA)
    m, mo, total = ac.scan_device(d, torch.tensor([0, n], dtype=torch.int64, device="cuda"), overlapping=True)
    assert m.dtype == torch.int64 and mo.tolist() == [0, total] and total > 5_000_000
    ends = m[:, 3]
    assert bool((ends[1:] >= ends[:-1]).all()) and int(ends[-1]) <= n
    got = m.cpu().numpy()
    orc = Oracle(pats, "Standard")
    cut = ac._ac.WINDOW_BYTES - (ac._ac.max_pattern_len - 1)
    for a, b in [(0, 64 << 20), (cut - (16 << 20), cut + (16 << 20)), (n - (32 << 20), n)]:
        sl = d[a:b].cpu().numpy()
        et, _, erec = orc.scan_batch(sl, np.array([0, b - a], dtype=np.int64), overlapping=True)
        exp = erec.astype(np.int64)
        exp[:, 2] += a
        exp[:, 3] += a
        sel = got[(got[:, 2] >= a) & (got[:, 3] <= b)]
        assert np.array_equal(sel, exp), (a, b)
    # every sampled match slices back to its pattern
    idx = np.linspace(0, total - 1, 2000).astype(np.int64)
    for i in idx:
        _, pid, s, e = got[i]
        assert d[s:e].cpu().numpy().tobytes() == pats[pid]


def test_two_threads_two_automata_and_a_shared_one():
    """The C ABI keeps tuning / timing state per calling thread and the host-buffer calls hold the automaton until
    their results are on the host: two threads, each with its own automaton and its own kernel choice, plus both
    hammering a third, shared automaton, must all get their own exact answers."""
    rng = np.random.default_rng(3)
    p1 = sorted({bytes(rng.integers(97, 101, size=rng.integers(2, 7)).astype(np.uint8)) for _ in range(200)})
    p2 = sorted({bytes(rng.integers(97, 100, size=rng.integers(1, 5)).astype(np.uint8)) for _ in range(60)})
    shared_p = [b"ab", b"bca", b"c", b"abcab"]
    shared = BytesAhoCorasick(shared_p)
    errors = []

    def work(pats, kernel, seed):
        try:
            _capi.set_tuning(kernel)
            r = np.random.default_rng(seed)
            ac = BytesAhoCorasick(pats)
            orc, sorc = Oracle(pats, "Standard"), Oracle(shared_p, "Standard")
            for it in range(25):
                hay = r.integers(97, 101, size=int(r.integers(1, 60_000)), dtype=np.uint8).tobytes()
                assert ac.find_matches_as_indexes(hay) == orc.find(hay)
                assert shared.find_matches_as_indexes(hay, overlapping=True) == sorc.find(hay, overlapping=True)
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
        finally:
            _capi.set_tuning(0)

    ts = [threading.Thread(target=work, args=(p1, 5, 1)), threading.Thread(target=work, args=(p2, 2, 2))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_host_pipeline_matches_one_shot():
    """scan_host in several runs (copy of run i+1 overlapping the scan of run i) == one device-resident call."""
    pats, data, offs = W.config2(12_000)
    ac = AhoCorasick(pats)
    m, mo, t = ac.scan_device(dev(data), dev(offs))
    ref, refo = m.cpu().numpy().view(np.uint32).copy(), mo.cpu().numpy().copy()
    pinned = torch.from_numpy(data).pin_memory()
    for chunk in (4 << 20, 16 << 20, 1 << 30):
        hm, hmo = ac.scan_host(pinned, offs, chunk_bytes=chunk)
        assert np.array_equal(hm, ref) and np.array_equal(hmo, refo)
    # ragged runs + a capacity overflow inside a run ("a" matches everywhere)
    dense = BytesAhoCorasick([b"a", b"aa"])
    hay = np.full(3_000_000, 97, dtype=np.uint8)
    o = np.array([0, 1_000_000, 1_000_000, 3_000_000], dtype=np.int64)
    hm, hmo = dense.scan_host(hay, o, overlapping=True, chunk_bytes=1 << 20)
    assert hmo.tolist() == [0, 1_999_999, 1_999_999, 1_999_999 + 3_999_999]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_two_ranks_sharded_scan_and_single_haystack():
    """The multi-GPU path on hardware: two ranks, each scanning its shard on its own GPU, match lists gathered over NCCL
    (both gather forms), content compared with the oracle; then ONE haystack split across the ranks (overlapping)."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
