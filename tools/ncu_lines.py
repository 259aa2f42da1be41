#!/usr/bin/env python
"""Stall samples of an .ncu-rep per CUDA source line (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep [min_percent]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; floor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for i, r in enumerate(rows[:80]):
    if "# Samples" in r:
        h, st = r, i + 1
        break
ln, src, ji, ie = h.index("Line No"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
samples, instr, text = collections.Counter(), collections.Counter(), {}
for r in rows[st:]:
    if len(r) <= ji: continue
    try: n = float(r[ji] or 0); e = float(r[ie] or 0)
    except ValueError: continue
    samples[r[ln]] += n; instr[r[ln]] += e; text[r[ln]] = r[src]
tot = sum(samples.values()) or 1
print("samples %d, warp instructions %d" % (tot, sum(instr.values())))
for k in sorted(samples, key=lambda x: int(x) if x.isdigit() else 0):
    if 100 * samples[k] / tot >= floor:
        print("%5s %5.1f%% %9d  %s" % (k, 100 * samples[k] / tot, instr[k], text[k].strip()[:120]))
