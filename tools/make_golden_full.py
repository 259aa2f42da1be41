#!/usr/bin/env python
"""Generate tests/golden/full_*.json.gz: the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile from
/root/reference/StrainCall) on the FULL-SIZE workloads bench.py times -- BASELINE configs[0], configs[1] and
whole-gene subgroups of configs[2] -- plus how long the reference took on one core of the build container.
Run in the build container only (configs[1] alone is ~13 minutes); the files are committed so that the GPU box
(which has no /root/reference) can check the CUDA path against them at benchmark size.

usage: make_golden_full.py [case ...]        (default: every case, one process per case)
"""
import gzip, hashlib, json, multiprocessing as mp, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

OUT = os.path.join(ROOT, "tests", "golden")

# name -> synth.make_subgroup arguments (the same calls bench.py / synth.config_workload make)
CASES = {
    "config0_seed0": dict(n_reads=2000, read_len=100, n_strains=3, seed=0),
    "config1_seed0": dict(n_reads=20000, read_len=150, n_strains=10, divergence=(0.01, 0.03), seed=0),
    "config2_sub0": dict(n_reads=5000, read_len=150, n_strains=2, seed=0),
    "config2_sub3": dict(n_reads=5000, read_len=150, n_strains=5, seed=3),
    "config2_sub4": dict(n_reads=5000, read_len=150, n_strains=6, seed=4),
    # configs[3] ("one deep subgroup, 1M 150bp reads, 50 strains") at a tenth of the reads: StrainCall's -D 800 down-sampling
    # (rho = 800 / depth) leaves the same ~8 200 reads at any raw depth above 800, so the subgroup that reaches the graph
    # is of the size of the 1M case
    # configs[4] ("250bp MiSeq-like reads with indel-rich strains (homopolymer errors) stressing wide POA graphs") at full
    # size with the indels in the STRAINS: the largest form the reference finishes (tools/ref_config4.py: with read-level
    # indel errors at >= 0.0005 per base it is killed -- segfault or out of memory -- from 500 reads up) ...
    "config4_full": dict(n_reads=5000, read_len=250, n_strains=4, indel_err=0.0, indel_frac=0.4, homopolymer_bias=True, seed=0),
    # ... and with read-level homopolymer indel errors at the largest size the reference finishes
    "config4_1000_ie001": dict(n_reads=1000, read_len=250, n_strains=4, indel_err=0.001, indel_frac=0.4, homopolymer_bias=True, seed=0),
    "config3_100k": dict(n_reads=100000, read_len=150, n_strains=50, divergence=(0.01, 0.03), seed=0),
}


def run_case(name):
    from oracle import refpy
    from rambl_b200 import synth
    from helpers import strip_sib
    spec = CASES[name]
    sg = synth.make_subgroup(**spec)
    t0 = time.time()
    g = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn)
    t1 = time.time()
    st, _ = g.infer(sg.pair_off, sg.pair_val)
    t2 = time.time()
    dump, edges = g.dump(), g.edges()
    out = dict(name=name, spec=spec,
               input=dict(gene=sg.gene, pos=sg.pos, cigar=sg.cigar, seq=sg.seq, cn=sg.cn,
                          pair_off=[int(x) for x in sg.pair_off], pair_val=[int(x) for x in sg.pair_val]),
               n_raw_reads=sg.n_raw_reads, n_reads=sg.n_reads, n_nodes=g.num_nodes(),
               dump_nosib_sha256=hashlib.sha256(strip_sib(dump).encode()).hexdigest(),
               edges_sha256=hashlib.sha256(edges.encode()).hexdigest(),
               strains=st,
               reference_seconds=dict(build=t1 - t0, infer_and_assign=t2 - t1, cores=1,
                                      build_flags="oracle/_ref -O2 (see oracle/Makefile)"))
    path = os.path.join(OUT, "full_%s.json.gz" % name)
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump(out, f)
    print("%s: %d reads, %d nodes, %d final strains, build %.1f s, infer+assign %.1f s -> %s (%d bytes)" % (
        name, sg.n_reads, g.num_nodes(), len(st.get("final", [])), t1 - t0, t2 - t1, path, os.path.getsize(path)), flush=True)
    return name


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    os.makedirs(OUT, exist_ok=True)
    with mp.get_context("spawn").Pool(len(names)) as pool:
        for _ in pool.imap_unordered(run_case, names):
            pass
