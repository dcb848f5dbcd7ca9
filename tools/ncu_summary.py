"""Summarise an .ncu-rep: per-kernel headline metrics + top stall sites (run where ncu is installed)."""
import csv, subprocess, sys, io

rep = sys.argv[1]
kfilter = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]
stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if kfilter and kfilter not in name:
        continue
    print("==", name[:100])
    for w in want:
        if w in idx:
            print(f"   {w:80s} {r[idx[w]]} {units[idx[w]]}")
    sv = sorted(((float(r[idx[s]] or 0), s.replace("smsp__pcsamp_warps_issue_stalled_", "")) for s in stalls), reverse=True)
    tot = sum(v for v, _ in sv) or 1
    print("   stalls: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sv[:7]))
