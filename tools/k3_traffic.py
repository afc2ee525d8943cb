#!/usr/bin/env python
"""profiles/k3_traffic.json (what bench.py reports as roofline.traffic / binding_units) from an ncu summary written by
tools/ncu_summary.py.  usage: python tools/k3_traffic.py profiles/<summary>.json [pairs_per_launch]"""
import json
import os
import sys

src = sys.argv[1]
pairs = float(sys.argv[2]) if len(sys.argv) > 2 else 100000 * 99999 / 2
d = json.load(open(src))
kname = sys.argv[3] if len(sys.argv) > 3 else "k3s_tile_kernel"
k = [l for l in d["launches"] if kname in l["kernel"]][0]


def val(name):
    v, u = float(k[name]["value"].replace(",", "")), k[name]["unit"]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
inst = val("smsp__inst_executed.sum")
out = {
    "kernel": kname,
    "source": "%s (ncu --set full --clock-control none, bench.py --steps 1 --warmup 3 --no-sketch --no-semb --no-cpu, "
              "one launch = one dense pass over %.6g pairs, generation order)" % (src, pairs),
    "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
    "kernel_ms_under_ncu": val("gpu__time_duration.sum"),
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "alu_pipe_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "lsu_pipe_pct": val("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "shared_wavefronts_pct": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
    "warp_instructions": inst, "warp_instructions_per_pair": inst / pairs,
}
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "k3_traffic.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
