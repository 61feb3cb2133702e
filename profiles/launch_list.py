"""Condense an `ncu --metrics gpu__time_duration.sum --csv` log into launches / total time / share per kernel.
usage: python profiles/launch_list.py gpurun_out/p_bench_launches.csv "<command profiled>" > profiles/<name>.txt"""
import csv, re, sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    name = re.sub(r"^void ", "", r[4])
    name = re.sub(r"\(.*", "", name)[:70]
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + float(r[-1]) / 1e3)
total = sum(t for _, t in agg.values())
print(f"# ncu launch list of `{sys.argv[2]}` (B200, --clock-control none)")
print("# per-launch times are cold-cache and serialised under the profiler: compare SHARES, not absolutes")
print(f"# {'kernel':70s} launches   total_us   share")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {name:70s} {n:8d} {t:10.1f} {100 * t / total:6.1f}%")
