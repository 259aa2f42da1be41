#!/usr/bin/env python
"""Time one BASELINE config through the C ABI (development helper).
usage: time_config.py <config index> [scale] [repetitions]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _one(k):
    from rambl_b200 import synth
    return synth.make_subgroup(5000, 150, 2 + (k % 5), seed=k)


def main():
    from rambl_b200 import api, synth
    idx = int(sys.argv[1]); scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    if os.environ.get("RAMBL_GIBBS_NG"):
        assert api.lib().rambl_set_gibbs_blocks(int(os.environ["RAMBL_GIBBS_NG"])) == 0
    t = time.time()
    if idx == 2:  # 500 subgroups x 5k reads: generate on all host cores
        import multiprocessing as mp
        n = max(1, int(500 * scale))
        with mp.get_context("spawn").Pool(min(os.cpu_count() or 1, 32)) as pool:
            sgs = pool.map(_one, range(n), chunksize=4)
    else:
        sgs = synth.config_workload(idx, seed=0, scale=scale)
    print("synth %.2fs" % (time.time() - t), "subgroups", len(sgs), "reads", sum(s.n_reads for s in sgs),
          "unique", sum(s.n_unique for s in sgs), "raw", sum(s.n_raw_reads for s in sgs), flush=True)
    for rep in range(reps):
        b = api.StrainCallBatch()
        t0 = time.time()
        for sg in sgs: b.add(sg)
        t1 = time.time(); b.build_graphs(); t2 = time.time(); b.infer(); t3 = time.time()
        st = b.stats()
        print("add %.3f build %.3f infer %.3f total %.3f" % (t1 - t0, t2 - t1, t3 - t2, t3 - t0), flush=True)
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
        ok = sum(1 for i in range(len(sgs)) if b.status(i) == 0)
        print("status ok %d/%d" % (ok, len(sgs)), "strains", [len(b.strains(i)) if b.status(i) == 0 else -1 for i in range(min(8, len(sgs)))], "nodes", b.num_nodes(0))
        print("reads/s %.1f" % (sum(s.n_reads for s in sgs) / (t3 - t0)))
        b.close()


if __name__ == "__main__":
    main()
