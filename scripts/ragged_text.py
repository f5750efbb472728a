"""Config 2's text with ragged haystack lengths (lanes of a warp no longer walk the same text in lock step):
what the shared-memory table scan does when its loads stop being broadcasts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ahocorasick_rs_b200 import BytesAhoCorasick, workloads as W

pats, data, offs = W.config2(100_000)
rng = np.random.default_rng(1)
for name, lens in [("equal 4096", np.full(100_000, 4096)), ("uniform 2048..4096", rng.integers(2048, 4097, size=100_000)),
                   ("uniform 64..4096", rng.integers(64, 4097, size=100_000))]:
    rows = data.reshape(100_000, 4096)
    keep = np.arange(4096)[None, :] < lens[:, None]
    flat = rows[keep]
    o = np.zeros(100_001, dtype=np.int64); np.cumsum(lens, out=o[1:])
    ac = BytesAhoCorasick([p.encode() for p in pats])
    d = torch.from_numpy(np.ascontiguousarray(flat)).cuda(); od = torch.from_numpy(o).cuda()
    m, mo, total = ac.scan_device(d, od)
    cap = int(total * 1.2) + 1024
    for _ in range(3): ac.scan_device(d, od, capacity=cap, sync=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ac.scan_device(d, od, capacity=cap, sync=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: {len(flat)/1e6:.0f} MB, {total} matches, {ms:.3f} ms/step, {len(flat)/ms/1e6:.0f} GB/s, engine {ac._ac.last_stats.get('engine')}", flush=True)
