import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d["precision"], d["case"], d["clips"], "ms %.2f" % d["ms"], "us/solve(longest) %.2f" % d["us_per_solve_of_longest_clip"], "Mframes/s %.2f" % (d["frames_per_s"] / 1e6))
