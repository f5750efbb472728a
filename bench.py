#!/usr/bin/env python
"""bench.py -- haystack GB/s scanned (+ matches/s) for find_matches_as_indexes
on the BASELINE.json config-2 workload (names.txt patterns, 100k x 4 KiB UTF-8
haystacks per GPU, MatchKind.Standard, code point indexes).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one device-resident batch (409.6 MB,
larger than L2; two batches alternate).  `value` is device-resident throughput
(CUDA events on the launching stream, max over ranks); `e2e` is the same work
through the public host-buffer API with H2D/D2H inside the timed region.
`--impl reference` times the reference's CPU path: the Rust crate cannot be
built in this image, so that arm runs the C oracle port (oracle/) on all host
cores and says so in cpu_baseline.kind = "port".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "haystack_GB_per_s_scanned_find_matches_as_indexes"
N_HAY = 100_000
HAY_BYTES = 4096
WORKLOAD = "config2: benchmarks/names.txt patterns (4244, Implementation.DFA, MatchKind.Standard), 100k x 4 KiB synthetic UTF-8 haystacks, AhoCorasick (code point indexes)"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = [float(r[1]) for r in rows]
            out["samples"] = len(rows)
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][2])
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for i, nm in enumerate(names):
                    if any("Active" in r[5 + i] and "Not" not in r[5 + i] for r in rows):
                        out["reasons"].append(nm)
            os.unlink(self.path)
        except Exception:
            pass
        return out


def usable_cores() -> int:
    """Host threads this process may really run: the affinity mask, cut down to the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def cpu_port_throughput(pats_bytes, data, offs, budget_s=12.0, threads=None):
    """Times the oracle port (dense DFA, one contiguous shard of haystacks per host
    thread) on a bounded sample of the same workload.  Returns (GB/s, matches/s, cores, sample)."""
    from oracle import Oracle

    threads = threads or usable_cores()
    orc = Oracle(pats_bytes, "Standard")
    n = len(offs) - 1
    orc.time_batch(data, offs, codepoints=True, nthreads=threads, reps=1)  # warm + calibrate
    t0 = time.perf_counter()
    orc.time_batch(data, offs, codepoints=True, nthreads=threads, reps=1)
    one = max(time.perf_counter() - t0, 1e-4)
    reps = int(max(1, min(2000, budget_s / one)))
    t0 = time.perf_counter()
    matches = orc.time_batch(data, offs, codepoints=True, nthreads=threads, reps=reps)
    dt = time.perf_counter() - t0
    gbs = reps * float(offs[-1] - offs[0]) / dt / 1e9
    return gbs, matches / dt, threads, (f"{n} haystacks x {HAY_BYTES} B ({(offs[-1] - offs[0]) / 1e6:.1f} MB) x {reps} passes inside one "
                                        f"thread launch, {threads} threads, one contiguous shard per thread; per haystack the "
                                        f"byte->code-point map is rebuilt like the reference does (src/lib.rs:235)")


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port), rank 0 only."""
    if rank != 0:
        return
    from ahocorasick_rs_b200 import workloads as W

    n_ref = 100_000
    pats, data, offs = W.config2(n_ref)
    pb = [p.encode() for p in pats]
    from oracle import Oracle

    threads = usable_cores()
    orc = Oracle(pb, "Standard")
    orc.time_batch(data, offs, codepoints=True, nthreads=threads, reps=max(args.warmup, 1))
    t0 = time.perf_counter()
    matches = orc.time_batch(data, offs, codepoints=True, nthreads=threads, reps=args.steps)
    dt = time.perf_counter() - t0
    gbs = args.steps * float(offs[-1]) / dt / 1e9
    sample = (f"each step = {n_ref} haystacks x {HAY_BYTES} B ({offs[-1] / 1e6:.1f} MB) of the config-2 workload, {threads} host "
              f"threads, one contiguous shard per thread, all steps inside one thread launch")
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_arm": "C oracle port of the reference's CPU path (Rust aho-corasick 1.1.4 cannot be built here: no rustc/cargo)"},
        "matches_per_s": matches / dt,
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--haystacks", type=int, default=N_HAY, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--segment-bytes", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--kernel", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--hot-rows", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--table", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from ahocorasick_rs_b200 import AhoCorasick, Implementation, _capi, workloads as W
    from ahocorasick_rs_b200.sharding import MatchListGather, decode_gathered

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries ONE JSON line: keep NCCL's "NCCL version ..." banner (printed to stdout at the VERSION debug
        # level) out of it; anything more verbose that the caller asked for is left alone
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    L = _capi.lib()
    if args.segment_bytes or args.hot_rows or args.table or args.kernel:
        _capi.set_tuning(args.kernel, args.hot_rows, args.segment_bytes, args.table)
    n_hay = args.haystacks
    # two different batches per rank (rank r owns haystack indices [r*2*n, (r+1)*2*n)): weak scaling
    batches = []
    pats = None
    for b in range(2):
        pats, data, offs = W.config2(n_hay, HAY_BYTES, first_index=(rank * 2 + b) * n_hay)
        batches.append((data, offs))
    ac = AhoCorasick(pats, implementation=Implementation.DFA)
    d_batches = [(torch.from_numpy(d).to(dev), torch.from_numpy(o).to(dev)) for d, o in batches]
    bytes_per_step = int(batches[0][1][-1])
    cap = 1 << 16

    def step(i):
        d, o = d_batches[i & 1]
        return ac.scan_device(d, o, capacity=cap, sync=False)

    # ---- device-resident throughput ------------------------------------------------
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    totals = [int(ac.scan_device(*d_batches[b], capacity=cap)[2]) for b in range(2)]
    scan_stats = dict(ac._ac.last_stats)
    gather_cap = max(4096, -(-2 * max(totals) // 4096) * 4096)  # rows per rank in the match-list gather (multi-GPU)
    gather = MatchListGather(gather_cap, dev, overlap=True) if world > 1 else None
    if world > 1:
        for i in range(max(args.warmup, 3)):  # warm the exchange too (communicator set-up, buffers)
            o_, _, t_ = step(i)
            gather(o_, t_, (rank * 2 + (i & 1)) * n_hay)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    L.acb_timing_enable(1)
    launches0 = L.acb_launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    host_t0 = time.perf_counter()
    for i in range(args.steps):
        out, moffs, tot = step(i)
        if world > 1:
            # the only exchange of the path: gather the per-shard match lists (sparse, a few KB)
            # (fixed-size blocks, no host round trip: the scans of the next steps are enqueued meanwhile)
            gathered = gather(out, tot, (rank * 2 + (i & 1)) * n_hay)
    if world > 1:
        gather.finish()  # the exchanges ran on a side stream: the timed region ends when the last one has
    ev1.record()
    host_enqueue_ms = (time.perf_counter() - host_t0) * 1e3 / max(args.steps, 1)
    if world > 1 and args.steps > 0:
        # the last step's gathered lists, decoded after the timed region: every rank's list must be whole
        glob = decode_gathered(gathered)
        assert glob.shape[0] >= totals[(args.steps - 1) & 1], "gathered match list is short"
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = int(L.acb_launch_count() - launches0)
    kms, kn = __import__("ctypes").c_double(0), __import__("ctypes").c_uint64(0)
    L.acb_timing_read(__import__("ctypes").byref(kms), __import__("ctypes").byref(kn))
    L.acb_timing_enable(0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    matches_per_step = sum(totals) / 2.0

    # ---- end to end through the public host API (rank-local) ---------------------------
    h_data = [torch.from_numpy(d).pin_memory() for d, _ in batches]
    h_offs = [torch.from_numpy(o).pin_memory() for _, o in batches]

    def e2e_step(i):
        d = h_data[i & 1].to(dev, non_blocking=True)
        o = h_offs[i & 1].to(dev, non_blocking=True)
        m, mo, total = ac.scan_device(d, o, capacity=cap)       # syncs to read the total
        return m.cpu(), mo.cpu(), total                        # result on the host

    e2e_steps = max(3, min(args.steps, 10))
    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    d2h = 0
    for i in range(e2e_steps):
        m, mo, total = e2e_step(i)
        d2h += m.numel() * 4 + mo.numel() * 8 + 16
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    total_bytes = bytes_per_step * args.steps * world
    value = total_bytes / (ms_max * 1e-3) / 1e9
    # algorithmic bytes of the scan kernel per launch: haystack bytes + int64 offsets + 16 B per match
    alg_bytes = bytes_per_step + 8 * (n_hay + 1) + 16 * matches_per_step
    k_ms = kms.value / max(kn.value, 1)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    # DRAM bytes per launch of that kernel from the committed ncu capture (profiles/summarize.py), at the default
    # workload size only: the capture is of this workload
    traffic, traffic_src = None, None
    tj = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_scan_kernel.json")
    if os.path.exists(tj) and n_hay == 100_000:
        with open(tj) as f:
            t = json.load(f)
        traffic, traffic_src = t["dram_traffic_bytes_per_launch"], "profiles/r01_scan_kernel.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "haystacks_per_gpu": n_hay, "haystack_bytes": HAY_BYTES,
                   "l2": "inputs (409.6 MB per batch, two batches alternating) are larger than L2; no flush needed",
                   "multi_gpu": "one process per GPU, batch sharded by haystack index, table replicated; per step one all-gather of the match lists (NCCL, fixed-size blocks, on a side stream so that it overlaps the next scan)"},
        "matches_per_s": matches_per_step * args.steps * world / (ms_max * 1e-3),
        "matches_per_step_per_gpu": matches_per_step,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "scan_staged_kernel",
                     "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes},
        "e2e": {"value": bytes_per_step * e2e_steps * world / e2e_s / 1e9, "unit": "GB/s",
                "h2d_bytes_per_step": bytes_per_step + 8 * (n_hay + 1), "d2h_bytes_per_step": d2h // e2e_steps,
                "steps": e2e_steps, "api": "AhoCorasick.scan_device on pinned host tensors copied H2D inside the timed region, matches copied back"},
        "gpu_launches": launches,
        "scan_stats": scan_stats,
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        gbs, mps, cores, sample = cpu_port_throughput([p.encode() for p in pats], batches[0][0], batches[0][1])
        line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample,
                                "matches_per_s": mps}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
