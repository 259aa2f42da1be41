import os, sys, time
sys.path.insert(0, '/root/repo')
from rambl_b200 import api, synth
for (ie, frac, n) in [(0.001, 0.2, 5000), (0.0005, 0.2, 2500)]:
    sg = synth.make_subgroup(n, 250, 4, indel_err=ie, indel_frac=frac, homopolymer_bias=True, seed=0)
    b = api.StrainCallBatch(); b.add(sg)
    t = time.time(); b.build_graphs(); t1 = time.time(); b.infer(); t2 = time.time()
    st = b.stats()
    print("indel_err", ie, "frac", frac, "reads", sg.n_reads, "nodes", b.num_nodes(0), "status", b.status(0), "strains", len(b.strains(0)) if b.status(0)==0 else -1,
          "build %.2f infer %.2f" % (t1 - t, t2 - t1), "msa problems", st["msa_problems"], "cells", st["msa_dp_cells"], "ms %.3f" % st["msa_kernel_ms"], "levels", st["level_steps"], flush=True)
