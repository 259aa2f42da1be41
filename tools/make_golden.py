#!/usr/bin/env python
"""Generate tests/golden/*.json by running the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile
from /root/reference/StrainCall) on small seeded inputs.  Run in the build container only; the JSON
files are committed so that the GPU box (which has no /root/reference) can check against them."""
import json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refpy
from rambl_b200 import synth

OUT = os.path.join(ROOT, "tests", "golden")

def msa_cases():
    rnd = random.Random(20151003)
    cases = [["ACG", "AG", "ACGT", "A"], ["A", "AA", "AAA"][::-1], ["GATTACA", "GATACA", "ATTAC", "TT", "G"],
             ["ACGTN", "acg", "A-T"], ["TTTTTTTTT", "TTTTTTT", "TTTTT", "TTT", "T"]]
    for _ in range(60):
        n = rnd.randint(2, 10)
        seqs = ["".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 10))) for _ in range(n)]
        seqs.sort(key=lambda s: -len(s))
        cases.append(seqs)
    return [{"seqs": c, "rows": refpy.msa_align(c)} for c in cases]

POG_SPECS = [
    dict(n_reads=60, read_len=40, n_strains=2, seed=11, window=(300, 400), sub_err=0.005, divergence=(0.03, 0.06)),
    dict(n_reads=90, read_len=40, n_strains=3, seed=12, window=(100, 200), sub_err=0.003, indel_err=0.01,
         indel_frac=0.4, homopolymer_bias=True, divergence=(0.03, 0.08)),
    dict(n_reads=100, read_len=40, n_strains=3, seed=13, window=(700, 800), sub_err=0.003, paired=True,
         divergence=(0.02, 0.06)),
    dict(n_reads=300, read_len=30, n_strains=2, seed=14, window=(900, 960), sub_err=0.001, indel_err=0.01,
         indel_frac=0.5, homopolymer_bias=True, divergence=(0.02, 0.05)),
]

def pog_cases():
    out = []
    # the worked example in the reference's header (PartialOrderGraph.hpp:20-29)
    hdr = dict(gene="ACGTACGT", pos=[0, 0, 0], cigar=["8M", "3M1I5M", "3M1D4M"],
               seq=["ACCTACGT", "ACCCTACGT", "ACCACGT"], cn=[1, 1, 1], pair_off=[0, 1, 2, 3], pair_val=[-1, -1, -1])
    g = refpy.RefPog(hdr["gene"], hdr["pos"], hdr["cigar"], hdr["seq"], hdr["cn"])
    out.append(dict(name="header_example", input=hdr, dump=g.dump(), edges=g.edges(), strains=None))
    for spec in POG_SPECS:
        sg = synth.make_subgroup(**spec)
        g = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn)
        st, _ = g.infer(sg.pair_off, sg.pair_val)
        inp = dict(gene=sg.gene, pos=sg.pos, cigar=sg.cigar, seq=sg.seq, cn=sg.cn,
                   pair_off=[int(x) for x in sg.pair_off], pair_val=[int(x) for x in sg.pair_val])
        out.append(dict(name="synth_seed%d" % spec["seed"], spec={k: v for k, v in spec.items()}, input=inp,
                        dump=g.dump(), edges=g.edges(), strains=st))
    return out

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    json.dump(msa_cases(), open(os.path.join(OUT, "msa_golden.json"), "w"), indent=0)
    json.dump(pog_cases(), open(os.path.join(OUT, "pog_golden.json"), "w"), indent=0)
    for f in os.listdir(OUT):
        print(f, os.path.getsize(os.path.join(OUT, f)))
