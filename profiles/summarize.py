#!/usr/bin/env python
"""Turn ncu outputs into the small text/CSV summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
  python profiles/summarize.py full     gpurun_out/prof_r1_final.ncu-rep > profiles/r1_ncu_full.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("mk::", "")
    return name.strip()


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot = OrderedDict()
    n = 0
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        k = short(r[ik])
        t = float(r[iv].replace(",", ""))
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
        n += 1
    unit = rows[1][hdr.index("Metric Unit")] if len(rows) > 1 else "ns"
    s = sum(v[1] for v in tot.values())
    print(f"# {n} launches, total {s:.0f} {unit} (ncu per-launch times are cold-cache and serialised: compare SHARES)")
    print(f"{'kernel':44s} {'launches':>8s} {'time':>14s} {'share':>7s}")
    for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {c:8d} {t:14.0f} {t / s:7.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==", short(r[idx["Kernel Name"]]))
        for k in KEYS:
            if k in idx:
                print(f"   {k:86s} {r[idx[k]]:>18s} {units[idx[k]]}")
        st = []
        for h in hdr:
            if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[idx[h]].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("   top stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:6]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
