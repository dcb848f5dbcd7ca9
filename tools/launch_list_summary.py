"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python bench.py ...`).
python tools/launch_list_summary.py gpurun_out/r02c_launches.csv "<command>" > profiles/r02c_launch_list_summary.txt"""
import csv, re, sys
from collections import defaultdict
path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "python bench.py --steps 4 --warmup 3 --skip-cpu")
rows = [r for r in csv.reader(l for l in open(path, errors="replace") if not l.startswith("=="))]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= vi or "gpu__time_duration.sum" not in r: continue
    name = re.sub(r"^void ", "", r[ki]); name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|at::|kltdev::|ekfvio::", "", name); name = re.sub(r"\(.*", "", name)
    v = float(r[vi].replace(",", "")); u = r[ui]
    us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    tot[name] += us; cnt[name] += 1
total = sum(tot.values())
print(f"# ncu launch list of `{cmd}`: gpu__time_duration.sum per kernel, --clock-control none.")
print("# Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.  Raw list:", path.replace("gpurun_out", "profiles"))
print(f"# {sum(cnt.values())} launches, {total / 1e3:.2f} ms of kernel time in total\n")
print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'mean us':>10s} {'share':>7s}")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k[:60]:60s} {cnt[k]:8d} {tot[k]:12.1f} {tot[k] / cnt[k]:10.1f} {100 * tot[k] / total:6.1f}%")
