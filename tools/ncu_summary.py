#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                d[h] = {"value": vals[i], "unit": units[i]}
            elif h.startswith(STALLS) and h.endswith("_per_issue_active.ratio"):
                d.setdefault("stalls_per_issue", {})[h[len(STALLS):-len("_per_issue_active.ratio")]] = float(vals[i])
        res.append(d)
    json.dump({"source": rep, "launches": res}, open(out, "w"), indent=1)
    for d in res:
        print(d["kernel"])
        for k in KEYS:
            if k in d:
                print("  %-78s %s %s" % (k, d[k]["value"], d[k]["unit"]))
        st = sorted(d.get("stalls_per_issue", {}).items(), key=lambda kv: -kv[1])[:6]
        print("  top stalls/issue:", ", ".join("%s=%.2f" % kv for kv in st))


if __name__ == "__main__":
    main()
