"""Worker of tests/test_gpu_full_size.py::test_nccl_two_ranks_...: run under torchrun with one rank per GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ahocorasick_rs_b200 import AhoCorasick, BytesAhoCorasick, workloads as W  # noqa: E402
from ahocorasick_rs_b200.sharding import (MatchListGather, decode_gathered, gather_match_lists, partition_by_bytes,  # noqa: E402
                                          scan_sharded_single)
from oracle import Oracle  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)

# (1) a batch sharded by haystack index
pats, data, offs = W.config2(6000)
pb = [p.encode() for p in pats]
etotal, ecounts, erec = Oracle(pb, "Standard").scan_batch(data, offs, codepoints=True)
lo, hi = partition_by_bytes(offs, world)[rank]
ac = AhoCorasick(pats)
d = torch.from_numpy(data[offs[lo]:offs[hi]]).to(dev)
o = torch.from_numpy(offs[lo:hi + 1] - offs[lo]).to(dev)
m, mo, total = ac.scan_device(d, o)
glob = gather_match_lists(m.clone(), lo)
assert np.array_equal(glob.cpu().numpy().view(np.uint32), erec), "two-collective gather differs from the oracle"
# the fixed-block gather, two steps in flight on alternating slots
g = MatchListGather(4096, dev)
res = None
for i in range(4):
    out, _, st = ac.scan_device(d, o, capacity=8192, sync=False, ws_slot=i & 1)
    res = g(out, st, lo, slot=i & 1)
g.finish()
torch.cuda.synchronize()
assert np.array_equal(decode_gathered(res).cpu().numpy().view(np.uint32), erec), "block gather differs from the oracle"

# (2) one haystack across the ranks, overlapping
rng = np.random.default_rng(5)
p2 = sorted({bytes(rng.integers(97, 101, size=rng.integers(2, 9)).astype(np.uint8)) for _ in range(300)})
hay = rng.integers(97, 101, size=3_000_001, dtype=np.uint8)
exp = Oracle(p2, "Standard").find(hay.tobytes(), overlapping=True)
bac = BytesAhoCorasick(p2)


def scan_fn(window):
    mm, _, _ = bac.scan_device(torch.from_numpy(np.ascontiguousarray(window)).to(dev), torch.tensor([0, len(window)], dtype=torch.int64, device=dev), True)
    return mm.cpu().numpy().view(np.uint32)[:, 1:].astype(np.int64)


got = scan_sharded_single(scan_fn, hay, bac._ac.max_pattern_len, device=dev)
assert [tuple(int(x) for x in r[1:]) for r in got.numpy()] == exp, "single haystack across ranks differs from the oracle"
print(f"rank {rank} ok", flush=True)
dist.destroy_process_group()
