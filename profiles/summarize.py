#!/usr/bin/env python3
"""Turns an ncu report (`ncu --set full --import-source on`, brought back from
the GPU box in gpurun_out/) into the small text/JSON summaries committed here.

    python profiles/summarize.py gpurun_out/prof.ncu-rep scan_staged profiles/r01_scan_kernel

writes <out>.md (key metrics, stall reasons, instruction mix, hottest SASS
lines) and <out>.json (the numbers bench.py quotes: DRAM traffic per launch).
Reads the report with `ncu -i ... --page raw|source --csv`; needs no GPU."""
import collections
import csv
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("sm__warps_active.avg.per_cycle_active", "warps active per SM"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts (LSU data pipe)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "  ... % of peak"),
    ("smsp__sass_inst_executed_op_shared_ld.sum", "LDS instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "LDS bank conflicts"),
    ("smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum", "LDGSTS (cp.async) shared wavefronts"),
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(out.splitlines()))


def to_bytes(value, unit):
    v = float(value)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, kernel, out = sys.argv[1:4]
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    row = next(r for r in raw[2:] if kernel in r[hdr.index("Kernel Name")])
    val = dict(zip(hdr, row))
    unit = dict(zip(hdr, units))
    lines = [f"# {val['Kernel Name'][:100]}", "", f"source report: `{rep}` (ncu --set full --clock-control none)", "", "| metric | value |", "|---|---|"]
    for k, label in KEYS:
        if k in val:
            lines.append(f"| {label} | {val[k]} {unit.get(k, '')} |")
    stalls = [(float(v), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k, v in val.items()
              if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v]
    tot = sum(v for v, _ in stalls) or 1
    lines += ["", "## warp stall reasons (pc samples)", "", "| reason | samples | share |", "|---|---|---|"]
    for v, k in sorted(stalls, reverse=True)[:10]:
        lines.append(f"| {k} | {v:.0f} | {100 * v / tot:.1f} % |")

    src = ncu_csv(rep, "source")
    h = src[1]
    ix = {k: i for i, k in enumerate(h)}
    data = src[2:]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError, IndexError):
            return 0.0

    mix = collections.Counter()
    for r in data:
        toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
        if toks:
            op = toks[0]
            mix[op if op.startswith(("LDS", "LDG", "STG", "LDL", "STL", "ATOM", "RED")) else op.split(".")[0]] += f(r, "Instructions Executed")
    total = sum(mix.values()) or 1
    lines += ["", "## executed warp instructions by opcode", "", "| opcode | executed | share |", "|---|---|---|"]
    for op, n in mix.most_common(14):
        lines.append(f"| {op} | {n:.0f} | {100 * n / total:.1f} % |")
    lines += ["", "## SASS lines with the most stall samples", "", "| samples | executed | avg threads | instruction |", "|---|---|---|---|"]
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:16]:
        lines.append(f"| {f(r, '# Samples'):.0f} | {f(r, 'Instructions Executed'):.0f} | {f(r, 'Avg. Threads Executed'):.0f} | `{r[ix['Source']].strip()[:90]}` |")
    with open(out + ".md", "w") as fh:
        fh.write("\n".join(lines) + "\n")
    rd = to_bytes(val["dram__bytes_read.sum"], unit["dram__bytes_read.sum"])
    wr = to_bytes(val["dram__bytes_write.sum"], unit["dram__bytes_write.sum"])
    dur = float(val["gpu__time_duration.sum"]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(unit["gpu__time_duration.sum"], 1e-6)
    with open(out + ".json", "w") as fh:
        json.dump({"kernel": kernel, "report": rep, "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "dram_traffic_bytes_per_launch": rd + wr, "duration_s_under_ncu": dur,
                   "registers_per_thread": int(float(val["launch__registers_per_thread"])),
                   "block_size": int(float(val["launch__block_size"]))}, fh, indent=1)
        fh.write("\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
