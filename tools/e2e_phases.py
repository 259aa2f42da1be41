#!/usr/bin/env python
"""Wall-clock phases of one configs[1] pass through the C ABI, with and without torch in the process."""
import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "torch":
    import torch
    torch.cuda.synchronize()
    if len(sys.argv) > 2:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
from rambl_b200 import api, synth
sg = synth.make_subgroup(20000, 150, 10, divergence=(0.01, 0.03), seed=0)
for rep in range(3):
    t0 = time.time(); b = api.StrainCallBatch(); b.add(sg); t1 = time.time(); b.build_graphs(); t2 = time.time(); b.infer(); t3 = time.time()
    f = b.fasta(0, "g", 1, len(sg.gene), 0.02); t4 = time.time(); st = b.stats(); b.close(); t6 = time.time()
    print("add %.3f build %.3f infer %.3f fasta %.3f close %.3f  (infer_gpu_ms %.0f)" % (t1-t0, t2-t1, t3-t2, t4-t3, t6-t4, st["infer_gpu_ms"]), flush=True)
