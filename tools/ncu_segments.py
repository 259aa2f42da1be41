#!/usr/bin/env python
"""Split the SASS of the first kernel in an .ncu-rep at barrier / vote / match / atomics and print the stall
samples and executed warp instructions of every segment, in program order.
usage: ncu_segments.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdrs = [i for i, r in enumerate(rows) if "# Samples" in r]
h, st = rows[hdrs[0]], hdrs[0] + 1
end = hdrs[1] if len(hdrs) > 1 else len(rows)
si, ji, ie = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
data = []
for r in rows[st:end]:
    if len(r) <= ji: continue
    try: data.append((float(r[ji] or 0), float(r[ie] or 0), r[si]))
    except ValueError: pass
tot = sum(d[0] for d in data) or 1
print("instructions", len(data), "samples", tot, "warp instructions", sum(d[1] for d in data))
acc = ex = n = 0; start = 0
for i, (s, e, t) in enumerate(data):
    acc += s; ex += e; n += 1
    w = t.split()
    op = w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "")
    if op.startswith(("BAR", "B2R", "SYNCS", "MATCH", "WARPSYNC", "UBLKCP", "VOTE", "ATOMS", "EXIT", "SHFL")) or i == len(data) - 1:
        if acc / tot >= 0.004 or ex > 0.01 * sum(d[1] for d in data):
            print("%5d-%5d  samples %5.1f%%  warp-instr %9d  | %s" % (start, i, 100 * acc / tot, ex, t[:70]))
        acc = ex = n = 0; start = i + 1
