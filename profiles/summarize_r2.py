#!/usr/bin/env python
"""Round-2 profile summaries: reads gpurun_out/*.ncu-rep (brought back from the B200 box) with the ncu CLI and writes
the text files under profiles/ that DESIGN.md and profiles/README.md quote.
    python profiles/summarize_r2.py gpurun_out/r2_gear_v1.ncu-rep profiles/r2_ncu_gear.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    lines = []
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[-1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines.append("kernel: " + vals[ix["Kernel Name"]])
    for m in RAW:
        if m in ix:
            lines.append(f"  {m:86s} {vals[ix[m]]} {units[ix[m]]}")
    src = page(rep, "source")
    h2, data = src[1], src[2:]
    jx = {h: i for i, h in enumerate(h2)}

    def f(r, k):
        try:
            return float(r[jx[k]])
        except Exception:
            return 0.0
    tot_i = sum(f(r, "Instructions Executed") for r in data)
    tot_s = sum(f(r, "# Samples") for r in data)
    lines.append(f"source page: {tot_i:.0f} warp instructions executed in the sampled pass, {tot_s:.0f} stall samples")
    st = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
    tots = sorted(((sum(f(r, h) for r in data), h) for h in st), reverse=True)
    lines.append("  stall reasons (all samples): " + ", ".join(f"{h[6:]} {v / tot_s * 100:.1f}%" for v, h in tots[:9]))
    ops = Counter()
    for r in data:
        s = r[jx["Source"]].split()
        if not s:
            continue
        op = s[1] if s[0].startswith("@") and len(s) > 1 else s[0]
        ops[op.split(".")[0]] += f(r, "Instructions Executed")
    lines.append("  executed by opcode: " + ", ".join(f"{op} {v / tot_i * 100:.1f}%" for op, v in ops.most_common(14)))
    bad = [(f(r, "L1 Wavefronts Shared Excessive"), r[jx["Source"]][:60]) for r in data if f(r, "L1 Wavefronts Shared Excessive") > 0]
    lines.append("  shared-memory instructions with excessive wavefronts (bank conflicts attributed to an instruction): "
                 + (", ".join(f"{s} (+{v:.0f})" for v, s in sorted(bad, reverse=True)[:5]) if bad else "none"))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
