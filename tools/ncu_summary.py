#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + top stall lines of the source page) into text for profiles/."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "gpu__time_duration.sum", "smsp__cycles_elapsed.max", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_uniform.sum"]
for r in rows[2:]:
    print("== launch")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("  %-70s %s %s" % (w, r[i][:90], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = None
for i, r in enumerate(rows[:60]):
    if "Source" in r and "# Samples" in r:
        h, start = r, i + 1
        break
if h:
    si, ji, ie = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [k for k, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    seq, stalls, ops = [], collections.Counter(), collections.Counter()
    for r in rows[start:]:
        try:
            n = float(r[ji] or 0)
        except Exception:
            continue
        seq.append((n, float(r[ie] or 0), r[si], r))
        for k in stall_cols:
            try:
                stalls[h[k]] += float(r[k] or 0)
            except Exception:
                pass
        t = [x for x in r[si].split() if not x.startswith("@")]
        if t:
            ops[t[0].split(".")[0]] += float(r[ie] or 0)
    tot = sum(x[0] for x in seq) or 1
    print("== first captured launch: %d samples, %d warp instructions" % (tot, sum(x[1] for x in seq)))
    print("stall reasons (%% of samples):", {k: round(100 * v / (sum(stalls.values()) or 1), 1) for k, v in stalls.most_common(8)})
    print("warp instructions by opcode:", dict((k, int(v)) for k, v in ops.most_common(16)))
    print("top sampled instructions:")
    for d in sorted(seq, key=lambda x: -x[0])[:14]:
        st = {h[k]: d[3][k] for k in stall_cols if d[3][k] not in ("0", "")}
        top = sorted(st.items(), key=lambda kv: -float(kv[1]))[:2]
        print("  %5.1f%%  exec %9d  %-58s %s" % (100 * d[0] / tot, d[1], d[2][:58], top))
