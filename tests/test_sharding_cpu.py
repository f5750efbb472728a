"""CPU tests of the multi-GPU host logic with world_size-2 gloo: partition ->
per-shard scan -> gather == unsharded result.  The per-shard scan is stood in
for by the oracle here (no GPU in this container); on GPUs bench.py --gpus N
drives the same gather over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ahocorasick_rs_b200 import workloads as W
from ahocorasick_rs_b200.sharding import (decode_gathered, gather_match_lists, gather_match_lists_async, partition_by_bytes,
                                          scan_sharded, scan_sharded_single)


def test_partition_by_bytes_covers_everything():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        lens = rng.integers(0, 1000, size=257)
        offs = np.zeros(258, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        offs += 13
        parts = partition_by_bytes(offs, world)
        assert parts[0][0] == 0 and parts[-1][1] == 257
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        sizes = [offs[hi] - offs[lo] for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 2000
    assert partition_by_bytes(np.array([0, 0, 0], dtype=np.int64), 4)[-1][1] == 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import Oracle
    pats = [b"ab", b"abc", b"ca", b"b"]
    data, offs = W.ragged(500, 200, b"abc", seed=3)
    orc = Oracle(pats, "Standard")

    def scan_fn(d, o):
        _, _, rec = orc.scan_batch(d, o)
        return torch.from_numpy(rec.astype(np.int64).astype(np.int32).reshape(-1, 4))

    full = scan_sharded(scan_fn, data, offs)
    _, _, exp = orc.scan_batch(data, offs)
    ok = np.array_equal(full.numpy().astype(np.uint32), exp)
    # gather to one destination only
    lo, hi = partition_by_bytes(offs, world)[rank]
    local = scan_fn(data[offs[lo]:offs[hi]], offs[lo:hi + 1] - offs[lo])
    only0 = gather_match_lists(local, lo, dst=0)
    ok = ok and ((only0 is None) == (rank != 0))
    if rank == 0:
        ok = ok and np.array_equal(only0.numpy().astype(np.uint32), exp)
    # the fixed-block gather that never looks at the counts on the host (what bench.py uses per step)
    cap = 1 << 15
    buf = torch.zeros((cap + 100, 4), dtype=torch.int32)
    buf[: local.shape[0]] = local
    status = torch.tensor([local.shape[0], 1, 0, 0, 0, 0, 0, 0], dtype=torch.int64)
    glob = decode_gathered(gather_match_lists_async(buf, status, lo, cap))
    ok = ok and np.array_equal(glob.numpy().astype(np.uint32), exp)
    # a block that is too small is reported, not truncated
    try:
        decode_gathered(gather_match_lists_async(buf, status, lo, 8))
        ok = False
    except RuntimeError:
        pass
    # one large haystack, overlapping, split across the ranks with a halo (bytes, then code points)
    rng = np.random.default_rng(9)
    hay = rng.integers(97, 100, size=50_001, dtype=np.uint8).astype(np.uint8)
    exp1 = orc.find(hay.tobytes(), overlapping=True)
    got1 = scan_sharded_single(lambda w: np.array(orc.find(w.tobytes(), overlapping=True), dtype=np.int64).reshape(-1, 3),
                               hay, max_pattern_len=3)
    ok = ok and [tuple(int(x) for x in r[1:]) for r in got1.tolist()] == exp1
    text = "".join(rng.choice(list("ab—é☃c"), size=20_000))
    upats = ["a—", "—é", "☃c", "b", "é☃c"]
    uorc = Oracle([u.encode() for u in upats], "Standard")
    exp2 = uorc.find_str(text, overlapping=True)
    raw = np.frombuffer(text.encode(), dtype=np.uint8)

    def scan_cp(w):
        # window-relative code point indexes: the oracle's own byte -> code point map over the window's bytes
        wb = w.tobytes()
        cont = np.cumsum(np.concatenate([[0], (np.frombuffer(wb, dtype=np.uint8) & 0xC0) == 0x80]))
        return np.array([(p, s - cont[s], e - cont[e]) for (p, s, e) in uorc.find(wb, overlapping=True)], dtype=np.int64).reshape(-1, 3)

    got2 = scan_sharded_single(scan_cp, raw, max_pattern_len=max(len(u.encode()) for u in upats), codepoints=True)
    ok = ok and [tuple(int(x) for x in r[1:]) for r in got2.tolist()] == exp2 and len(exp2) > 1000
    q.put((rank, bool(ok), int(full.shape[0])))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_scan_and_gather_equals_unsharded():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2] > 100
