"""Stall samples and warp instructions of k_chain_step by kernel phase: ncu SASS source page joined with the
`nvdisasm -g` line table of the same build; phases are the `// ----` section comments of chain_kernels.cu.
usage: ncu_phases.py src.csv sass.txt <mangled kernel> chain_kernels.cu"""
import collections, csv, re, sys
src, sass, kern, cu = sys.argv[1:5]
rows = list(csv.reader(open(src))); h = rows[1]
iA, iN, iE = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
names = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
idx = {n: h.index(n) for n in names}
data = [(int(r[iA], 16), int(r[iN] or 0), int(r[iE] or 0), {n: int(r[idx[n]] or 0) for n in names})
        for r in rows[2:] if r and r[iA].startswith("0x")]
base = data[0][0]
line_of, cur, inside = {}, None, False
for ln in open(sass):
    if ln.startswith(".text." + kern + ":"):
        inside = True; continue
    if inside:
        if ln.startswith("//-----") or ln.startswith("\t.section"): break
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m: line_of[int(m.group(1), 16)] = cur
marks = {"tables: proposed": "tables", "change-local rasterisation: which": "rasterisation", "re-derivation: 16 lanes": "re-derivation",
         "change-local forward: fixed": "forward (columns)", "misfit: phi": "misfit", "likelihood ratio, decision": "decision + commit",
         "One block per chain.": "kernel entry (zeroing, dependency wait, block order)"}
bounds, kstart = [], None
for i, ln in enumerate(open(cu), 1):
    if "chain_step_body(const EngineParams &ep" in ln and kstart is None: kstart = i
    for k, v in marks.items():
        if k in ln: bounds.append((i, v))
bounds = [(kstart, "entry / gather")] + bounds
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for a, n, e, st in data:
    k = line_of.get(a - base)
    if not k: ph = "?"
    elif k[0] != "chain_kernels.cu": ph = "inlined headers (%s)" % k[0]
    elif k[1] < kstart: ph = "helpers (fx_add4 atomics, block sum, index vector)"
    else: ph = [nm for lo, nm in bounds if k[1] >= lo][-1]
    agg[ph][0] += n; agg[ph][1] += e; agg[ph][2].update(st)
ts = sum(v[0] for v in agg.values()); te = sum(v[1] for v in agg.values())
print("samples", ts, "warp instructions", te)
print("| phase | stall samples | warp instructions | top stall reasons |\n|---|---:|---:|---|")
for ph, (n, e, st) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    top = ", ".join("%s %.0f %%" % (k.replace("stall_", ""), 100 * v / max(1, sum(st.values()))) for k, v in st.most_common(3))
    print("| %s | %.1f %% | %.1f %% | %s |" % (ph, 100 * n / ts, 100 * e / te, top))
