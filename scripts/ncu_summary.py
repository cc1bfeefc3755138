#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into a small JSON: python scripts/ncu_summary.py rep [out.json]"""
import csv, json, subprocess, sys
WANT = [
 "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed",
 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
 "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
 "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
 "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
 "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg",
 "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
 "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second", "gpc__cycles_elapsed.avg.per_second",
 "smsp__issue_active.avg.pct_of_peak_sustained_active",
 "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")][:90]}
    for i, n in enumerate(hdr):
        if n in WANT or "warp_issue_stalled" in n and n.endswith("per_warp_active.pct"):
            d[n] = f"{r[i]} {units[i]}"
    out.append(d)
txt = json.dumps(out, indent=1)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
print(txt)
