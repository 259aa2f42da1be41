#!/usr/bin/env python
"""One configs[2] subgroup on its own through the device walk (development helper).  usage: one_subgroup.py k [cluster]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rambl_b200 import api, synth
k = int(sys.argv[1])
if len(sys.argv) > 2:
    os.environ["RAMBL_WALK_CLUSTER"] = sys.argv[2]
sg = synth.config2_subgroup(k)
b = api.StrainCallBatch()
b.add(sg)
b.build_graphs()
for _ in range(2):
    s0 = b.stats(); t = time.time(); b.infer(); dt = time.time() - t; s1 = b.stats()
    print("subgroup %d: %d reads, %d unique, %d nodes; infer %.3f s, walk kernel %.1f ms, draws %d, status %d, strains %d" % (
        k, sg.n_reads, sg.n_unique, b.num_nodes(0), dt, s1["dpm_kernel_ms"] - s0["dpm_kernel_ms"], s1["draws"] - s0["draws"], b.status(0),
        len(b.strains(0)) if b.status(0) == 0 else -1))
    print("   rounds %d passes %d (%.2f per round), level steps %d, loglik updates %d" % (
        s1["gibbs_rounds"] - s0["gibbs_rounds"], s1["gibbs_passes"] - s0["gibbs_passes"],
        (s1["gibbs_passes"] - s0["gibbs_passes"]) / max(1, s1["gibbs_rounds"] - s0["gibbs_rounds"]), s1["level_steps"] - s0["level_steps"],
        s1["loglik_updates"] - s0["loglik_updates"]))
