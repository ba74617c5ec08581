#!/usr/bin/env python
"""Concise summary of an ncu report: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-row]"""
import csv, subprocess, sys
rep = sys.argv[1]; row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, r = rows[0], rows[1], rows[2 + row]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for w in want:
    if w in h:
        i = h.index(w); print("%-75s %s %s" % (w, r[i], units[i]))
print("-- warp stall reasons (pct of warp-active cycles per issue) --")
st = [(float(r[i].replace(",", "")), n) for i, n in enumerate(h) if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio") and r[i]]
if not st:
    st = [(float(r[i].replace(",", "")), n) for i, n in enumerate(h) if "issue_stalled" in n and n.endswith("per_warp_active.pct") and r[i]]
for v, n in sorted(st, reverse=True)[:8]:
    print("%-75s %.3f" % (n, v))
