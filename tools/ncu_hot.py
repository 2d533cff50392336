"""Print the hottest SASS instructions of an ncu source-page CSV (ncu -i X --page source --csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ie, src, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
lines = []
for r in rows[2:]:
    try:
        lines.append((int(r[ie].replace(",", "")), int(r[smp].replace(",", "") or 0), r[src]))
    except ValueError:
        pass
tot = sum(n for n, _, _ in lines); mx = max(n for n, _, _ in lines); ts = sum(s for _, s, _ in lines)
print("total executed", tot, "samples", ts)
hot = [(n, s, t) for n, s, t in lines if n > frac * mx]
print("hot instructions:", len(hot), "executed in hot:", sum(n for n, _, _ in hot))
for n, s, t in hot:
    print(f"{n:10d} {s:6d}  {t[:120]}")
