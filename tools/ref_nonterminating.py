#!/usr/bin/env python
"""An input on which the unmodified reference (oracle/_ref) does not finish building its graph: 650 reads of 60 bp with
homopolymer indel errors (normal construction takes well under a second at this size).  Run under `timeout 120`; the library
refuses the same input with RAMBL_ERR_INVALID (tests/test_host_logic.py).  Build container only (needs oracle/_ref)."""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refpy
from rambl_b200 import synth
seed = 66
spec = dict(n_reads=300 + 7 * (seed % 50), read_len=60, n_strains=2 + seed % 4, seed=1000 + seed, window=(100, 400), sub_err=0.005, indel_err=0.01 + 0.002 * (seed % 15), indel_frac=0.4, homopolymer_bias=True, divergence=(0.02, 0.06))
sg = synth.make_subgroup(**spec)
t = time.time()
o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn)
print("reference built the graph in %.1f s, %d nodes" % (time.time() - t, o.num_nodes()))
