#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics + hottest SASS lines.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25] > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== kernel:", d.get("Kernel Name", "?"), " id", d.get("ID", "?"))
    for k in KEYS:
        if k in d:
            print("  %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
    stalls = [(float(d[h]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h]]
    print("  stall reasons (warps stalled per issue-active cycle):")
    for v, h in sorted(stalls, reverse=True)[:8]:
        print("    %-40s %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
for si, start in enumerate(starts):
    end = starts[si + 1] - 1 if si + 1 < len(starts) else len(srows)
    name = srows[start - 1][1] if start > 0 and len(srows[start - 1]) > 1 else "?"
    sh = srows[start]
    ix = {h: i for i, h in enumerate(sh)}
    body = [r for r in srows[start + 1:end] if len(r) == len(sh) and r[0] != "Address"]
    if not body:
        continue
    tot = sum(int(r[ix["# Samples"]]) for r in body) or 1
    warps = max(int(r[ix["Instructions Executed"]]) for r in body) or 1
    print("== source page of %s: %d SASS lines, %d stall samples, %.1f instructions per warp" % (
        name, len(body), tot, sum(int(r[ix["Instructions Executed"]]) for r in body) / warps))
    print("  hottest lines (line, samples %, executed, shared N-way conflicts, SASS):")
    for n, r in sorted(enumerate(body), key=lambda x: -int(x[1][ix["# Samples"]]))[:top]:
        print("    %4d %5.1f%% %9s %6s  %s" % (n, 100.0 * int(r[ix["# Samples"]]) / tot, r[ix["Instructions Executed"]],
                                             r[ix["L1 Conflicts Shared N-Way"]], r[ix["Source"]].strip()[:90]))
