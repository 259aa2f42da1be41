#!/usr/bin/env python
"""Time rambl_batch_infer on the configs[2] batch for several CTA shapes of the device walk (development helper).
usage: walk_sweep.py [n_subgroups] [nb:tile ...]     e.g. walk_sweep.py 500 0:0 1:64 2:32 4:32 8:40"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from rambl_b200 import api
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    shapes = [tuple(int(x) for x in a.split(":")) for a in sys.argv[2:]] or [(0, 0)]
    t = time.time()
    sgs = bench.load_subgroups(list(range(n)), "/tmp/rambl_b200_bench_cache", os.cpu_count() or 1)
    print("workload: %d subgroups, %d reads, %.1f s" % (len(sgs), sum(s.n_reads for s in sgs), time.time() - t), flush=True)
    b = api.StrainCallBatch()
    for sg in sgs:
        b.add(sg)
    b.build_graphs()
    for nb, tile in shapes:
        os.environ["RAMBL_WALK_NB"] = str(nb)
        os.environ["RAMBL_WALK_TILE"] = str(tile)
        if nb < 0:
            os.environ["RAMBL_WALK"] = "0"
        else:
            os.environ.pop("RAMBL_WALK", None)
        s0 = b.stats()
        t = time.time()
        b.infer()
        dt = time.time() - t
        s1 = b.stats()
        print("nb %d tile %d: infer %.3f s wall, %.1f ms gpu, walk kernel %.1f ms, launches %d, ok %d" % (
            nb, tile, dt, s1["infer_gpu_ms"] - s0["infer_gpu_ms"], s1["dpm_kernel_ms"] - s0["dpm_kernel_ms"],
            s1["gpu_launches"] - s0["gpu_launches"], sum(1 for i in range(len(sgs)) if b.status(i) == 0)), flush=True)


if __name__ == "__main__":
    main()
