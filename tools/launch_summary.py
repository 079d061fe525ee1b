"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  Usage: launch_summary.py <csv> [out.md]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; i_name = hdr.index("Kernel Name"); i_val = hdr.index("Metric Value"); i_grid = hdr.index("Grid Size"); i_blk = hdr.index("Block Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    key = (r[i_name].split("(")[0].replace("void ", ""), r[i_grid], r[i_blk])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += float(r[i_val].replace(",", ""))
tot = sum(a[1] for a in agg.values())
lines = [f"launches captured: {len(rows)-1}, total {tot/1e3:.1f} us (cold-cache, serialised under ncu: compare shares, not absolutes)", "",
         "| kernel | grid | block | launches | total us | avg us | share |", "|---|---|---|---|---|---|---|"]
for (k, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| {k} | {g} | {b} | {n} | {t/1e3:.1f} | {t/n/1e3:.2f} | {100*t/tot:.1f} % |")
out = "\n".join(lines); print(out)
if len(sys.argv) > 2: open(sys.argv[2], "w").write(out + "\n")
