#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel headline metrics, pipe utilisation and warp stall reasons (text, for profiles/)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:90])
    for k in keys:
        if k in d:
            print(f"  {k:70s} {d[k]:>16s} {units[hdr.index(k)]}")
    print("  -- pipes (pct of peak, active):")
    for k in hdr:
        if k.startswith("sm__inst_executed_pipe_") and k.endswith("pct_of_peak_sustained_active") or \
           (k.startswith("sm__pipe_") and k.endswith("pct_of_peak_sustained_active")):
            try:
                if float(d[k]) >= 2: print(f"     {k:75s} {d[k]}")
            except ValueError: pass
    print("  -- warp stalls (per warp active, pct):")
    st = [(float(d[k]), k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and d[k] not in ("", "n/a")]
    if not st:
        st = [(float(d[k]), k) for k in hdr if "warp_issue_stalled" in k and k.endswith("per_warp_active.pct") and d[k] not in ("", "n/a")]
    for v, k in sorted(st, reverse=True)[:10]:
        print(f"     {k:85s} {v:.2f}")
