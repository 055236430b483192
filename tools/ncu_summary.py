"""Turn `ncu -i X.ncu-rep --page raw --csv` output (stdin) into the short per-kernel summary kept under profiles/."""
import csv
import sys

WANT = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max")
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
units = rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print(d.get("Kernel Name", "?")[:110])
    for k in WANT:
        if k in d:
            print(f"    {k:75s} {d[k]} {u.get(k, '')}")
    stalls = sorted(((float(v), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v), reverse=True)[:6]
    for v, k in stalls:
        print(f"    stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:36s} {v:.2f} warps/issue")
    print()
