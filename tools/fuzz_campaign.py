#!/usr/bin/env python
"""Offline fuzz campaign (CPU): the host graph builder (rows from the oracle's alignment) against the oracle on the fuzz
specs of the tests and on subgroups with homopolymer indel errors, seeds lo..hi-1.  Seed 66 of the second family is an input on
which the reference never ends (tools/ref_nonterminating.py): the builder refuses it, and so it is skipped here.
usage: fuzz_campaign.py lo hi"""
import sys, os, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refpy
from rambl_b200 import api, synth
from helpers import fuzz_spec, strip_sib
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0; n = 0; nprob = 0
t0 = time.time()
for seed in range(lo, hi):
    specs = [fuzz_spec(seed)]
    specs.append(dict(n_reads=300 + 7 * (seed % 50), read_len=60, n_strains=2 + seed % 4, seed=1000 + seed, window=(100, 400), sub_err=0.005,
                      indel_err=0.01 + 0.002 * (seed % 15), indel_frac=0.4, homopolymer_bias=True, divergence=(0.02, 0.06)))
    for which, spec in enumerate(specs):
        if seed == 66 and which == 1:
            continue
        try:
            sg = synth.make_subgroup(**spec)
        except Exception as e:
            continue
        if sg.n_unique == 0:
            continue
        b = api.StrainCallBatch(); b.add(sg); b.thread_reads()
        probs = b.msa_problems(); nprob += len(probs)
        try:
            b.finish_graphs_with_rows([refpy.msa_align(p, "oracle") for p in probs])
        except api.RamblError as e:
            # refused inputs must be refused by the oracle too
            try:
                refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="oracle")
                print("seed", seed, "refused here only:", e); bad += 1
            except Exception:
                pass
            continue
        o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="oracle")
        n += 1
        if strip_sib(b.graph_dump(0)) != strip_sib(o.dump()) or b.output_edge(0) != o.edges():
            print("MISMATCH seed", seed, spec); bad += 1
        b.close()
print("range", lo, hi, "graphs", n, "alignment problems", nprob, "bad", bad, "in %.0f s" % (time.time() - t0))
