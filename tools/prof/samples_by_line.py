#!/usr/bin/env python
"""Executed instructions and stall samples (~time) per source line of one kernel: joins an ncu SASS-page CSV with nvdisasm -g line info.
Usage: samples_by_line.py sass.csv dis.txt <kernel-substring> <min % of samples>"""
import csv, re, sys, collections
sass_csv, dis_txt, kern, thr = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
off2loc, infn, cur = {}, False, ("?", 0)
for l in open(dis_txt):
    if l.startswith(".text."):
        infn = kern in l; continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m: off2loc[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
base=None; tot=0; tots=0
by=collections.defaultdict(lambda:[0,0,0])
for r in rows[2:]:
    a = int(r[col["Address"]],16) if not r[col["Address"]].isdigit() else int(r[col["Address"]])
    if base is None: base=a
    loc = off2loc.get(a-base, (("?",0),""))
    ex=int(r[col["Instructions Executed"]] or 0); sm=int(r[col["# Samples"]] or 0); tot+=ex; tots+=sm
    by[loc[0]][0]+=ex; by[loc[0]][1]+=1; by[loc[0]][2]+=sm
for k in sorted(by, key=lambda k:(k[0],k[1])):
    v=by[k]
    if v[2]/tots*100>=thr: print(f"{k[0]}:{k[1]:5d} inst {v[0]/tot*100:5.2f}%  time {v[2]/tots*100:5.2f}%  static {v[1]}")
