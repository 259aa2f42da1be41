#!/usr/bin/env python
"""Insertion-alignment kernel alone: many level-problems of the size deep homopolymer levels produce."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rambl_b200 import api
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
rnd = np.random.default_rng(1)
probs = []
for _ in range(P):
    n = int(rnd.integers(20, 400))          # reads with an insertion at this level (depth 800)
    hp = "ACGT"[int(rnd.integers(4))]
    seqs = []
    for _ in range(n):
        k = int(rnd.integers(1, 10))       # StrainCall keeps insertions shorter than 10
        s = [hp] * k                        # homopolymer run of varying length ...
        if rnd.random() < 0.1:              # ... with the odd sequencing error
            s[int(rnd.integers(k))] = "ACGT"[int(rnd.integers(4))]
        seqs.append("".join(s))
    probs.append(sorted(seqs, key=lambda s: -len(s)))
for rep in range(3):
    t = time.time(); rows, st = api.msa_align_batch(probs); dt = time.time() - t
    steps = sum(len(p) - 1 for p in probs)
    print("problems %d sequences %d dp_cells %d kernel %.3f ms -> %.3f GCUPS (cells), %.2f M alignment steps/s; call %.3f s" % (
        P, sum(len(p) for p in probs), st["dp_cells"], st["kernel_ms"], st["dp_cells"] / st["kernel_ms"] / 1e6,
        steps / st["kernel_ms"] / 1e3, dt), flush=True)
