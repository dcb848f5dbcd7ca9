"""Summarise an ncu report (--page raw --csv) per kernel launch: the metrics DESIGN.md / VERDICT quote.  python tools/ncu_summary.py <rep>"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]
stall = [n for n in hdr if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")] or \
        [n for n in hdr if n.startswith("smsp__average_warp_latency_issue_stalled_") ]
for r in data:
    name = r[col["Kernel Name"]][:110]
    print("==", name)
    for w in want:
        if w in col:
            print(f"   {w:84s} {r[col[w]]} {units[col[w]]}")
    st = []
    for n in stall:
        try: st.append((float(r[col[n]].replace(",", "")), n.split("stalled_")[1].split("_per_")[0].replace(".ratio", "")))
        except ValueError: pass
    st.sort(reverse=True)
    tot = sum(v for v, _ in st) or 1.0
    print("   stalls: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in st[:7]))
