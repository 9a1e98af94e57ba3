#!/usr/bin/env python
"""Join an ncu SASS-page CSV (ncu -i X.ncu-rep --page source --csv --print-source sass) with
nvdisasm -g line info of the same cubin and aggregate executed instructions / stall samples per
solver phase (function of gmr_solver.cuh).  Usage: by_phase.py sass.csv dis.txt <kernel-substring>"""
import csv, re, sys, collections

sass_csv, dis_txt, kern = sys.argv[1:4]
SRC = "/root/repo/general_motion_retargeting_b200/csrc/gmr_solver.cuh"
# function line ranges of gmr_solver.cuh
funcs = []
with open(SRC) as f:
    for n, l in enumerate(f, 1):
        m = re.match(r"\s*(?:template <[^>]*>\s*)?GMR_FN\s+[\w:<>\*& ]+?\s+(\w+)\(", l)
        if m:
            funcs.append((n, m.group(1)))
def func_of(fname, line):
    if not fname.endswith("gmr_solver.cuh"):
        return "kernel.cu" if fname.endswith("gmr_kernels.cu") else fname.split("/")[-1]
    name = "?"
    for n, f in funcs:
        if n <= line: name = f
        else: break
    return name

# offset -> (file, line) for the chosen kernel
off2loc, infn, cur = {}, False, ("?", 0)
for l in open(dis_txt):
    if l.startswith(".text."):
        infn = kern in l
        continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m: off2loc[int(m.group(1), 16)] = (cur, m.group(2).strip())

rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
base = None
agg = collections.defaultdict(lambda: collections.Counter())
ops = collections.defaultdict(lambda: collections.Counter())
lines = collections.defaultdict(lambda: collections.Counter())
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = int(r[col["Address"]], 16)
    if base is None: base = a
    loc = off2loc.get(a - base)
    if loc is None: continue
    (fn, ln), text = loc
    f = func_of(fn, ln)
    ie = int(r[col["Instructions Executed"]]); te = int(r[col["Thread Instructions Executed"]])
    agg[f]["static"] += 1
    agg[f]["inst"] += ie; agg[f]["thr"] += te
    agg[f]["samples"] += int(r[col["# Samples"]])
    for k in ("stall_barrier", "stall_no_inst", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_math", "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_dispatch"):
        if k in col: agg[f][k] += int(r[col[k]])
    op = re.sub(r"^@!?U?P\d\s+", "", text).split()[0].split(".")[0]
    ops[f][op] += ie
    lines[(fn.split("/")[-1], ln)]["inst"] += ie; lines[(fn.split("/")[-1], ln)]["samples"] += int(r[col["# Samples"]]); lines[(fn.split("/")[-1], ln)]["static"] += 1
tot = sum(v["inst"] for v in agg.values()); tots = sum(v["samples"] for v in agg.values())
print(f"{'phase':18s} {'static':>6s} {'inst%':>6s} {'lanes':>5s} {'smp%':>6s}  barrier no_inst long_sb short_sb wait  mio  math notsel sel  branch")
for f, v in sorted(agg.items(), key=lambda kv: -kv[1]["inst"]):
    s = max(v["samples"], 1)
    print(f"{f:18s} {v['static']:6d} {100*v['inst']/tot:6.2f} {v['thr']/max(v['inst'],1):5.1f} {100*v['samples']/tots:6.2f}  " +
          " ".join(f"{100*v[k]/s:5.1f}" for k in ("stall_barrier", "stall_no_inst", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_math", "stall_not_selected", "stall_selected", "stall_branch_resolving")))
print("total inst", tot, "samples", tots)
if len(sys.argv) > 4 and sys.argv[4] == "--lines":
    lo, hi = int(sys.argv[5]), int(sys.argv[6])
    src = open(SRC).read().split("\n")
    for (fn, ln), v in sorted(lines.items()):
        if fn.endswith("gmr_solver.cuh") and lo <= ln <= hi and v["inst"]:
            print(f"{ln:5d} {v['static']:4d} {100*v['inst']/tot:6.2f}% {100*v['samples']/tots:6.2f}%  {src[ln-1][:110]}")
elif len(sys.argv) > 4:
    for f in sys.argv[4:]:
        print(f, ops[f].most_common(25))
