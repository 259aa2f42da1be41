#!/usr/bin/env python
"""Quick GPU sanity run: CUDA path vs the oracle on a handful of small seeded cases."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refpy
from rambl_b200 import synth, api

def cmp_strains(a, b, tol=1e-6):
    """a: oracle stages, b: product stages -> list of problems"""
    bad = []
    for stage in a:
        if stage not in b: bad.append("missing stage " + stage); continue
        if len(a[stage]) != len(b[stage]): bad.append("%s: %d vs %d strains" % (stage, len(a[stage]), len(b[stage]))); continue
        for i, (x, y) in enumerate(zip(a[stage], b[stage])):
            if x["path"] != y["path"]: bad.append("%s[%d] path differs" % (stage, i))
            if x["plain"] != y["plain"]: bad.append("%s[%d] plain differs" % (stage, i))
            if abs(x["abundance"] - y["abundance"]) > tol * max(1.0, abs(x["abundance"])): bad.append("%s[%d] abundance %r vs %r" % (stage, i, x["abundance"], y["abundance"]))
            for q, (u, v) in enumerate(zip(x["sub"], y["sub"])):
                if abs(u - v) > tol * max(1.0, abs(u)): bad.append("%s[%d] sub[%d] %r vs %r" % (stage, i, q, u, v)); break
            if "loglik" in x and "loglik" in y:
                for rid, u in x["loglik"].items():
                    v = y["loglik"].get(rid)
                    if v is None or (abs(u - v) > tol * max(1.0, abs(u)) and not (u == v)):
                        bad.append("%s[%d] loglik[%d] %r vs %r" % (stage, i, rid, u, v)); break
    return bad

def main():
    print("devices", api.device_count())
    rnd = random.Random(7)
    nbad = 0
    # ---- MSA
    probs = []
    for it in range(300):
        n = rnd.randint(2, 12)
        alpha = "ACGT" if it % 7 else "ACGTNa-"
        seqs = ["".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 12))) for _ in range(n)]
        seqs.sort(key=lambda s: -len(s))
        probs.append(seqs)
    probs.append(["ACGTACGTAC" * 3] + ["ACG" * k for k in range(9, 0, -1)] * 3)
    t = time.time(); rows, st = api.msa_align_batch(probs); print("msa gpu %.3fs" % (time.time() - t), st)
    for p, r in zip(probs, rows):
        o = refpy.msa_align(p, "oracle")
        if o != r:
            nbad += 1
            if nbad < 5: print("MSA MISMATCH", p, o, r)
    print("msa mismatches", nbad)
    # ---- full path
    for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
        rng = random.Random(seed)
        L = rng.choice([60, 120, 200])
        kw = dict(n_reads=rng.choice([30, 80, 200]), read_len=rng.choice([30, 50]), n_strains=rng.choice([2, 3, 4]),
                  seed=seed, window=(100, 100 + L), sub_err=rng.choice([0.0, 0.005, 0.02]),
                  indel_err=rng.choice([0, 0, 0.01, 0.03]), indel_frac=rng.choice([0.1, 0.4]),
                  homopolymer_bias=rng.random() < 0.5, paired=rng.random() < 0.4, divergence=(0.02, 0.08))
        sg = synth.make_subgroup(**kw)
        if sg.n_unique == 0: continue
        og = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="oracle")
        t = time.time(); ost, _ = og.infer(sg.pair_off, sg.pair_val, do_assign=False); to = time.time() - t
        b = api.StrainCallBatch(); b.add(sg)
        t = time.time(); b.build_graphs()
        gok = b.graph_dump(0).replace(" | SIB", "") .split("\n")[0:1] is not None
        if len(ost["infer"]) == 0:
            b.infer(keep_loglik=True)
            print(seed, "oracle 0 strains; product status", b.status(0)); continue
        ost, _ = og.infer(sg.pair_off, sg.pair_val)
        b.infer(keep_loglik=True); tg = time.time() - t
        if b.status(0) != 0:
            print(seed, "PRODUCT STATUS", b.status(0)); nbad += 1; continue
        pst = refpy.parse_strain_dump(b.strains_text(0))
        bad = cmp_strains(ost, pst)
        print(seed, kw["n_reads"], L, "strains", len(ost["final"]), "OK" if not bad else "MISMATCH", "t_oracle %.2f t_gpu %.2f" % (to, tg), b.stats()["gpu_launches"], flush=True)
        for x in bad[:5]: print("   ", x)
        nbad += 1 if bad else 0
    print("TOTAL BAD", nbad)
    return 1 if nbad else 0

if __name__ == "__main__":
    sys.exit(main())
