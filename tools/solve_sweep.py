#!/usr/bin/env python
"""End-to-end time of rambl_batch_solve on the 500 configs[2] subgroups for chunk / driver counts (development helper).
usage: solve_sweep.py spec ...     spec = chunks:drivers (equal chunks) | first:N (first chunk of N subgroups, then the rest) |
default (the library's own layout: the first chunk is one wave of the walk kernel) | separate (build_graphs, then infer)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from rambl_b200 import api
    sgs = bench.load_subgroups(list(range(500)), "/tmp/rambl_b200_bench_cache", os.cpu_count() or 1)
    for sg in sgs:
        sg.packed()
    for spec in sys.argv[1:]:
        for k in ("RAMBL_SOLVE_CHUNKS", "RAMBL_SOLVE_DRIVERS", "RAMBL_SOLVE_FIRST"):
            os.environ.pop(k, None)
        if spec.startswith("first:"):
            os.environ["RAMBL_SOLVE_FIRST"] = spec.split(":")[1]
        elif spec not in ("default", "separate"):
            c, d = spec.split(":")
            os.environ["RAMBL_SOLVE_CHUNKS"], os.environ["RAMBL_SOLVE_DRIVERS"] = c, d
        best = None
        for _ in range(2):
            b = api.StrainCallBatch()
            t = time.time()
            for sg in sgs:
                b.add(sg)
            if spec == "separate":  # the two calls one after the other, no overlap
                b.build_graphs()
                t1 = time.time()
                b.infer()
                print("  build_graphs %.3f s, infer %.3f s" % (t1 - t, time.time() - t1), flush=True)
            else:
                b.solve()
            dt = time.time() - t
            b.close()
            best = dt if best is None else min(best, dt)
        print("layout %s: add + solve %.3f s (best of 2)" % (spec, best), flush=True)


if __name__ == "__main__":
    main()
