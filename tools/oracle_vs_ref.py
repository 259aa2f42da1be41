#!/usr/bin/env python
"""Offline campaign (build container: needs oracle/_ref): the oracle against the unmodified reference on the same two input
families as tools/fuzz_campaign.py -- graph dump and output_edge text, and with the third argument 1 also every strain of
infer_strains (first family only).  Seeds 34 and 47 of the first family make the reference itself segfault in infer_strains
(every strain pruned / NaN abundances): run ranges around them.
usage: oracle_vs_ref.py lo hi with_strains"""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refpy
from rambl_b200 import synth
from helpers import fuzz_spec
lo, hi, with_strains = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = bad = 0
for seed in range(lo, hi):
    specs = [fuzz_spec(seed), dict(n_reads=300 + 7 * (seed % 50), read_len=60, n_strains=2 + seed % 4, seed=1000 + seed, window=(100, 400), sub_err=0.005,
                      indel_err=0.01 + 0.002 * (seed % 15), indel_frac=0.4, homopolymer_bias=True, divergence=(0.02, 0.06))]
    for which, spec in enumerate(specs):
        if seed == 66 and which == 1:
            continue  # the reference never finishes this one
        sg = synth.make_subgroup(**spec)
        if sg.n_unique == 0:
            continue
        o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="oracle")
        r = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn)
        n += 1
        if o.dump() != r.dump() or o.edges() != r.edges():
            print("GRAPH MISMATCH", seed, which); bad += 1; continue
        if with_strains and which == 0:
            so, _ = o.infer(sg.pair_off, sg.pair_val)
            sr, _ = r.infer(sg.pair_off, sg.pair_val)
            if so != sr:
                print("STRAIN MISMATCH", seed); bad += 1
print("seeds", lo, hi, "cases", n, "bad", bad, flush=True)
