"""Summarise `ncu --page source --csv` output: per SASS instruction samples and stall reasons,
printing the instructions that hold the most warp-stall samples and totals per code region.
usage: ncu -i X.ncu-rep --page source --csv | python profiles/ncu_source_hotspots.py [top_n]"""
import csv, sys

rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
recs = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    recs.append((r[col["Address"]], " ".join(r[col["Source"]].split()), n, int(r[col["Instructions Executed"]] or 0),
                 {s: int(r[col[s]] or 0) for s in stalls}))
total = sum(r[2] for r in recs)
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
print(f"total samples {total}, instructions {len(recs)}")
print("---- instructions by samples ----")
for idx, (a, src, n, ex, st) in sorted(enumerate(recs), key=lambda x: -x[1][2])[:top]:
    why = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{idx:5d} {n:6d} ({100*n/total:4.1f}%) exec {ex:9d}  {src[:70]:70s} {why}")
print("---- running profile (every 50 instructions) ----")
for i in range(0, len(recs), 50):
    blk = recs[i:i + 50]
    n = sum(r[2] for r in blk)
    ex = sum(r[3] for r in blk)
    agg = {}
    for r in blk:
        for k, v in r[4].items():
            agg[k] = agg.get(k, 0) + v
    why = ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:4] if v)
    print(f"{i:5d}-{i+len(blk)-1:5d} samples {n:6d} ({100*n/total:4.1f}%) exec {ex:10d}  {why}")
