"""Top stall sites of a kernel from `ncu --page source --csv` (gzipped): python tools/ncu_source_top.py file.csv.gz [N]"""
import csv, gzip, sys
rd = list(csv.reader(gzip.open(sys.argv[1], "rt")))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
starts = [i for i, r in enumerate(rd) if r and r[0] == "Kernel Name"]
lo = starts[-1]                                   # the last profiled launch of the kernel (warm)
rd = rd[lo:]
head = rd[1]
ix = {h: i for i, h in enumerate(head)}
rows = [r for r in rd[2:] if len(r) == len(head) and r[0] != "Address"]
tot = sum(int(r[ix["# Samples"]]) for r in rows)
stalls = [h for h in head if h.startswith("stall_") and "Not Issued" not in h]
print(f"kernel: {rd[0][1][:90]}   total samples {tot}")
agg = {h: sum(int(r[ix[h]]) for r in rows) for h in stalls}
print("stall totals:", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
rows.sort(key=lambda r: -int(r[ix["# Samples"]]))
for r in rows[:N]:
  n = int(r[ix["# Samples"]])
  top = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
  print(f"{100*n/tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
