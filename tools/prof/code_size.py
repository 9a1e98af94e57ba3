#!/usr/bin/env python
"""Static SASS instruction count of one kernel per source function (nvdisasm -g line info): where the code size is.
   Usage: code_size.py <lib.so> <kernel-substring>"""
import collections, re, subprocess, sys, tempfile, os
lib, kern = sys.argv[1], sys.argv[2]
SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "general_motion_retargeting_b200/csrc/gmr_solver.cuh")
funcs = []
for n, l in enumerate(open(SRC), 1):
    m = re.match(r"\s*(?:template <[^>]*>\s*)?GMR_FN\s+[\w:<>\*& ]+?\s+(\w+)\(", l)
    if m: funcs.append((n, m.group(1)))
def func_of(fname, line):
    if not fname.endswith("gmr_solver.cuh"): return fname.split("/")[-1]
    name = "?"
    for n, f in funcs:
        if n <= line: name = f
        else: break
    return name
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cnt, infn, cur = collections.Counter(), False, ("?", 0)
for l in dis.split("\n"):
    if l.startswith(".text."):
        infn = kern in l; continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1), int(m.group(2))); continue
    if re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l): cnt[func_of(*cur)] += 1
tot = sum(cnt.values())
for f, n in cnt.most_common(40): print(f"{f:28s} {n:6d} {100*n/tot:5.1f}%")
print("total", tot, "instructions =", tot * 16 // 1024, "KB")
