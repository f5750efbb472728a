"""Host-buffer path: throughput of scan_host by run size on config 2, and per-call latency of the drop-in methods
(one haystack per call, the reference's usage: benchmarks/test_comparison.py:119-122) next to the oracle's."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ahocorasick_rs_b200 import AhoCorasick, BytesAhoCorasick, workloads as W
from oracle import Oracle

pats, data, offs = W.config2(100_000)
ac = AhoCorasick(pats)
pinned = torch.from_numpy(data).pin_memory()
for chunk in (32 << 20, 64 << 20, 128 << 20, 256 << 20, 1 << 30):
    ac.scan_host(pinned, offs, chunk_bytes=chunk)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        m, mo = ac.scan_host(pinned, offs, chunk_bytes=chunk)
    dt = (time.perf_counter() - t0) / 5
    print(f"scan_host config 2, runs of {chunk >> 20} MiB: {dt * 1e3:.2f} ms, {len(data) / dt / 1e9:.1f} GB/s", flush=True)
# plain H2D for reference
d = torch.empty(len(data), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(pinned, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"plain H2D of the same bytes: {dt * 1e3:.2f} ms, {len(data) / dt / 1e9:.1f} GB/s", flush=True)

# per-call latency
_, hay1k = W.config1()
cases = [("config 1: 3 patterns, 1 KB ASCII haystack", AhoCorasick(["hello", "world", "fish"]), Oracle([b"hello", b"world", b"fish"], "Standard"), hay1k),
         ("names.txt patterns, one 4 KiB haystack", ac, Oracle([p.encode() for p in pats], "Standard"), data[:4096].tobytes().decode("utf-8", "ignore")),
         ("10 patterns, 75-byte haystack (benchmarks 'short')", AhoCorasick(["hello", "world", "fish", "kw1", "kw2", "abc", "def", "ghi", "jkl", "mno"]),
          Oracle([b"hello", b"world", b"fish", b"kw1", b"kw2", b"abc", b"def", b"ghi", b"jkl", b"mno"], "Standard"), "x" * 30 + "hello world fish" + "y" * 29)]
for name, a, o, hay in cases:
    for _ in range(20): a.find_matches_as_indexes(hay)
    t0 = time.perf_counter()
    for _ in range(300): r = a.find_matches_as_indexes(hay)
    gpu_us = (time.perf_counter() - t0) / 300 * 1e6
    t0 = time.perf_counter()
    for _ in range(300): e = o.find_str(hay)
    cpu_us = (time.perf_counter() - t0) / 300 * 1e6
    assert r == e
    print(f"latency, {name}: GPU path {gpu_us:.0f} us per call, oracle via ctypes {cpu_us:.0f} us per call", flush=True)
