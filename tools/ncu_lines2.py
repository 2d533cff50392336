"""Per-source-line totals of an ncu SASS source page (ncu -i X --page source --csv) joined with the line table of the
same kernel from `nvdisasm -g -c` (lines `//## File "...", line N`), with the top stall reasons of every line.
usage: ncu_lines2.py src.csv kernel.sass <mangled kernel name fragment> [top]"""
import csv, re, sys, collections
src, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src)))
h = rows[1]
iA, iN, iE = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[2:]:
    if r and r[iA].startswith("0x"):
        data.append((int(r[iA], 16), int(r[iN] or 0), int(r[iE] or 0), [int(r[i] or 0) for i, _ in stall_cols]))
base = data[0][0]
line_of = {}
cur = None
inside = False
for ln in open(sass):
    if ln.startswith(".text."):
        inside = kern in ln
        continue
    if inside:
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
tot_s = sum(d[1] for d in data); tot_e = sum(d[2] for d in data)
agg = collections.defaultdict(lambda: [0, 0, [0] * len(stall_cols)])
for a, n, e, st in data:
    k = line_of.get(a - base, ("?", 0))
    agg[k][0] += n; agg[k][1] += e
    for i, v in enumerate(st): agg[k][2][i] += v
print("samples", tot_s, "warp instructions", tot_e)
for k, (n, e, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = sorted(zip(st, [nm for _, nm in stall_cols]), reverse=True)[:3]
    print(f"{k[0]:>20s}:{k[1]:<5d} samples {100*n/tot_s:5.1f}%  instr {100*e/tot_e:5.1f}%   " + ", ".join(f"{nm[6:]} {100*v/max(n,1):.0f}%" for v, nm in tops))
