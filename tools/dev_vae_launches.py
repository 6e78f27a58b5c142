"""Summarise an ncu launch list (gpu__time_duration) of one VAE decode by kernel name and grid size (dev tool)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
kn, gs, mv, mu = H.index("Kernel Name"), H.index("Grid Size"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    v = float(r[mv].replace(",", ""))
    v = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[mu], 1e-3) * v
    name = r[kn].split("(")[0][-60:]
    agg[(name, r[gs])][0] += 1; agg[(name, r[gs])][1] += v; tot += v
print(f"total {tot/1e3:.1f} ms over {sum(a[0] for a in agg.values())} launches")
byname = collections.defaultdict(float)
for (n, g), (c, t) in agg.items(): byname[n] += t
for n, t in sorted(byname.items(), key=lambda kv: -kv[1])[:12]: print(f"  {t/1e3:8.1f} ms {100*t/tot:5.1f}%  {n}")
print("top (kernel, grid):")
for (n, g), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]: print(f"  {t/1e3:8.1f} ms {100*t/tot:5.1f}%  n={c:5d} avg {t/c:8.1f} us  {n} grid {g}")
