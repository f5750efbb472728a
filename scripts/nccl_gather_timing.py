"""Where the time of the exact-size match-list gather goes (run under torchrun, one rank per GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from ahocorasick_rs_b200.sharding import gather_match_lists

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
k = 4_750_000
local_rows = torch.randint(0, 1 << 20, (k, 4), dtype=torch.int32, device=dev)
buf = torch.empty(world * k * 4, dtype=torch.int32, device=dev)


def timed(fn, n=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


t_ag = timed(lambda: dist.all_gather_into_tensor(buf, local_rows.view(-1)))
t_full = timed(lambda: gather_match_lists(local_rows, rank * 1000))
t_small = timed(lambda: dist.all_gather_into_tensor(buf[: world * 1024], local_rows.view(-1)[:1024]), n=50)
if rank == 0:
    mb = k * 16 / 1e6
    print(f"world {world}: all_gather_into_tensor of {mb:.0f} MB per rank: {t_ag:.2f} ms ({mb * (world - 1) / t_ag / 1e3 * 1e3 / 1e3:.1f} GB/s received per rank); "
          f"gather_match_lists: {t_full:.2f} ms; 4 KB all-gather: {t_small * 1e3:.0f} us", flush=True)
dist.destroy_process_group()
