#!/usr/bin/env python
"""End-to-end time of rambl_batch_solve on the 500 configs[2] subgroups for chunk / driver counts (development helper).
usage: solve_sweep.py chunks:drivers ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from rambl_b200 import api
    sgs = bench.load_subgroups(list(range(500)), "/tmp/rambl_b200_bench_cache", os.cpu_count() or 1)
    for sg in sgs:
        sg.packed()
    for spec in sys.argv[1:]:
        c, d = spec.split(":")
        os.environ["RAMBL_SOLVE_CHUNKS"], os.environ["RAMBL_SOLVE_DRIVERS"] = c, d
        best = None
        for _ in range(2):
            b = api.StrainCallBatch()
            t = time.time()
            for sg in sgs:
                b.add(sg)
            b.solve()
            dt = time.time() - t
            b.close()
            best = dt if best is None else min(best, dt)
        print("chunks %s drivers %s: add + solve %.3f s (best of 2)" % (c, d, best), flush=True)


if __name__ == "__main__":
    main()
