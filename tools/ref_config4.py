#!/usr/bin/env python
"""What the UNMODIFIED reference (oracle/_ref) does on BASELINE configs[4] at full size (5 000 x 250 bp reads with
homopolymer indel errors, indel-rich strains), with a wall-clock limit.  Evidence for DESIGN.md section 6; run in
the build container only.   usage: ref_config4.py [limit seconds] [n_reads] [indel_err]"""
import multiprocessing as mp, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def work(n, ie, q):
    from oracle import refpy
    from rambl_b200 import synth
    sg = synth.make_subgroup(n, 250, 4, indel_err=ie, indel_frac=0.4, homopolymer_bias=True, seed=0)
    t = time.time()
    g = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn)
    q.put(("built", sg.n_reads, sg.n_unique, g.num_nodes(), time.time() - t))
    st, _ = g.infer(sg.pair_off, sg.pair_val, with_loglik=False)
    q.put(("done", len(st.get("final", [])), [s["abundance"] for s in st.get("final", [])][:8], time.time() - t))


if __name__ == "__main__":
    limit = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    ie = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
    q = mp.get_context("spawn").Queue()
    p = mp.get_context("spawn").Process(target=work, args=(n, ie, q))
    t0 = time.time()
    p.start()
    done = False
    while time.time() - t0 < limit:
        try:
            msg = q.get(timeout=5)
            print(n, ie, msg, flush=True)
            if msg[0] == "done":
                done = True
                break
        except Exception:
            if not p.is_alive():
                print(n, ie, "reference process died, exit code", p.exitcode, "after %.0f s" % (time.time() - t0), flush=True)
                done = True
                break
    if not done:
        print(n, ie, "reference still running after %.0f s: killed" % (time.time() - t0), flush=True)
        p.kill()
    p.join()
