#!/usr/bin/env python
"""Shared-memory wavefronts per CALL SITE: joins an ncu SASS-page CSV (ncu -i X.ncu-rep --page source --csv --print-source sass) with
nvdisasm -gi line info (inline chains) and charges every LDS / STS to the first source line outside the g_ld*/g_st* helpers.
Usage: smem_by_caller.py sass.csv dis_gi.txt <kernel-substring> <top-n> [helper-line-range lo-hi]"""
import csv, re, sys, collections
sass_csv, dis_txt, kern = sys.argv[1:4]
HELP = set(range(254,286))   # g_ld4/g_st4 helper lines (base source)
if len(sys.argv) > 5: HELP = set(range(*map(int, sys.argv[5].split('-'))))
off2loc, infn = {}, False
chain = []
for l in open(dis_txt):
    if l.startswith(".text."):
        infn = kern in l; continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        chain.append((m.group(1).split('/')[-1], int(m.group(2)))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        # the chain entries since the last instruction: innermost first
        loc = None
        for f, n in chain:
            if f == "gmr_solver.cuh" and n in HELP: continue
            loc = (f, n); break
        if loc is None and chain: loc = chain[0]
        if not chain: loc = last
        last = loc
        off2loc[int(m.group(1), 16)] = (loc, m.group(2).strip())
        chain = []
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
byline = collections.defaultdict(lambda: [0,0,0,0])
tot=[0,0,0]; base=None
for r in rows[2:]:
    a = int(r[col["Address"]],16) if not r[col["Address"]].isdigit() else int(r[col["Address"]])
    if base is None: base=a
    loc = off2loc.get(a-base, (("?",0),""))
    wf = int(r[col["L1 Wavefronts Shared"]] or 0); ideal=int(r[col["L1 Wavefronts Shared Ideal"]] or 0); ex=int(r[col["Instructions Executed"]] or 0)
    if wf:
        k = loc[0]
        byline[k][0]+=wf; byline[k][1]+=ideal; byline[k][2]+=ex
        tot[0]+=wf; tot[1]+=ideal; tot[2]+=ex
print("total wavefronts / ideal / smem instructions", tot)
for loc,v in sorted(byline.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[4])]:
    print(f"{loc[0]}:{loc[1]:5d}  wf {v[0]/tot[0]*100:5.2f}%  inst {v[2]/tot[2]*100:5.2f}%  wf/inst {v[0]/max(v[2],1):5.2f} ideal/inst {v[1]/max(v[2],1):5.2f}  excess {100*(v[0]-v[1])/tot[0]:5.2f}%")
