#!/usr/bin/env python
"""The static keys of the walk kernel's launch order for the 500 configs[2] subgroups, on the CPU (graphs built with the
oracle's alignment rows, rambl_batch_walk_plan): node counts, read-pool entries over the levels, which subgroups are handed
over at an early "$", and where subgroup 383 stands in the launch order (development helper; ~1.5 minutes on 8 cores)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from rambl_b200 import api
from oracle import refpy
def main():
    sgs = bench.load_subgroups(list(range(500)), "/tmp/rambl_b200_bench_cache", 8)
    out = []
    for k, sg in enumerate(sgs):
        b = api.StrainCallBatch(); b.add(sg); b.thread_reads()
        b.finish_graphs_with_rows([refpy.msa_align(p, "oracle") for p in b.msa_problems()])
        pl = b.walk_plan(0)
        out.append((k, b.num_nodes(0), pl["entries"], pl["handoff"], sg.n_reads))
        b.close()
    old = sorted(range(500), key=lambda i: -out[i][1])
    new = sorted(range(500), key=lambda i: -(out[i][1] * out[i][2]))
    print("383 in old order at", old.index(383), "in new order at", new.index(383))
    import numpy as np
    nodes = np.array([o[1] for o in out]); ent = np.array([o[2] for o in out])
    print("nodes min/med/max", nodes.min(), int(np.median(nodes)), nodes.max(), "entries min/med/max", ent.min(), int(np.median(ent)), ent.max())
    print("handoffs:", [o[0] for o in out if o[3]])
    ro = np.empty(500); ro[old] = np.arange(500); rn = np.empty(500); rn[new] = np.arange(500)
    print("rank correlation old/new", np.corrcoef(ro, rn)[0, 1])
    by_strains = {}
    for k in range(500):
        by_strains.setdefault(2 + k % 5, []).append(nodes[k])
    print({s: int(np.mean(v)) for s, v in by_strains.items()})
    

if __name__ == '__main__':
    main()
