#!/usr/bin/env python
"""Walk-kernel time of every rank's share of the 500 configs[2] subgroups at world size N, for cluster sizes 1 and 2
(development helper for the strong-scaling analysis).   usage: rank_walk.py [N=8]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from rambl_b200 import api, shard
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    parts = shard.assign([bench.config2_cost(k) for k in range(500)], N)
    sgs = bench.load_subgroups(list(range(500)), "/tmp/rambl_b200_bench_cache", os.cpu_count() or 1)
    for r, mine in enumerate(parts):
        b = api.StrainCallBatch()
        for k in mine:
            b.add(sgs[k])
        b.build_graphs()
        out = []
        for c in (1, 2):
            os.environ["RAMBL_WALK_CLUSTER"] = str(c)
            best = None
            for _ in range(2):
                s0 = b.stats()
                b.infer()
                s1 = b.stats()
                ms = s1["dpm_kernel_ms"] - s0["dpm_kernel_ms"]
                best = ms if best is None else min(best, ms)
            out.append(best)
        print("rank %d: %d subgroups, walk kernel %.1f ms with one CTA per subgroup, %.1f ms with clusters of two" % (
            r, len(mine), out[0], out[1]), flush=True)
        b.close()


if __name__ == "__main__":
    main()
