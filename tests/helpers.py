"""Shared comparison helpers of the parity tests."""
import json
import os
import random

from oracle import refpy
from rambl_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# log-likelihoods / abundances: the north star asks for 1e-6 relative in FP64 against the reference's
# long double arithmetic; paths, graph edges, consensus sequences and assignments must be identical
REL_TOL = 1e-6


def load_golden(name):
    with open(os.path.join(ROOT, "tests", "golden", name)) as f:
        return json.load(f)


def load_golden_gz(name):
    import gzip
    with gzip.open(os.path.join(ROOT, "tests", "golden", name), "rt") as f:
        return json.load(f)


def strip_sib(txt):
    """Graph dump without the SIB field (sibling lists keep ids of deleted nodes in the reference)."""
    out = []
    for line in txt.split("\n"):
        if line.startswith("NODE "):
            parts = line.split(" | ")
            parts[3] = "SIB"
            line = " | ".join(parts)
        out.append(line.rstrip())
    return "\n".join(out).strip()


def close(u, v, tol=REL_TOL):
    if u == v:
        return True
    if u != u and v != v:  # both NaN
        return True
    return abs(u - v) <= tol * max(1.0, abs(u))


def compare_strains(want, got, tol=REL_TOL, stages=None):
    """want/got: parse_strain_dump() dicts.  Returns a list of human-readable differences."""
    bad = []
    for stage in (stages or want.keys()):
        if stage not in got:
            bad.append("missing stage " + stage)
            continue
        if len(want[stage]) != len(got[stage]):
            bad.append("%s: %d vs %d strains" % (stage, len(want[stage]), len(got[stage])))
            continue
        for i, (x, y) in enumerate(zip(want[stage], got[stage])):
            if x["path"] != y["path"]:
                bad.append("%s[%d] path differs" % (stage, i))
            if x["seq"] != y["seq"] or x["plain"] != y["plain"]:
                bad.append("%s[%d] sequence differs" % (stage, i))
            if not close(x["abundance"], y["abundance"], tol):
                bad.append("%s[%d] abundance %r vs %r" % (stage, i, x["abundance"], y["abundance"]))
            for q, (u, v) in enumerate(zip(x["sub"], y["sub"])):
                if not close(u, v, tol):
                    bad.append("%s[%d] sub[%d] %r vs %r" % (stage, i, q, u, v))
                    break
            if "loglik" in x and "loglik" in y:
                for rid, u in x["loglik"].items():
                    v = y["loglik"].get(rid, y["loglik"].get(str(rid)))
                    if v is None or not close(u, v, tol):
                        bad.append("%s[%d] loglik[%s] %r vs %r" % (stage, i, rid, u, v))
                        break
    return bad


def normalise_golden_strains(st):
    """JSON turns the integer keys of the loglik maps into strings."""
    for stage in st.values():
        for s in stage:
            if "loglik" in s:
                s["loglik"] = {int(k): v for k, v in s["loglik"].items()}
    return st


def fuzz_spec(seed):
    rng = random.Random(seed)
    L = rng.choice([60, 120, 200])
    return dict(n_reads=rng.choice([30, 80, 200]), read_len=rng.choice([30, 50]), n_strains=rng.choice([2, 3, 4]),
                seed=seed, window=(100, 100 + L), sub_err=rng.choice([0.0, 0.005, 0.02]),
                indel_err=rng.choice([0, 0, 0.01, 0.03]), indel_frac=rng.choice([0.1, 0.4]),
                homopolymer_bias=rng.random() < 0.5, paired=rng.random() < 0.4, divergence=(0.02, 0.08))


def msa_fuzz_problems(seed, count, max_n=10, max_len=10):
    rnd = random.Random(seed)
    probs = []
    for it in range(count):
        n = rnd.randint(2, max_n)
        alpha = "ACGT" if it % 7 else "ACGTNa-"
        seqs = ["".join(rnd.choice(alpha) for _ in range(rnd.randint(1, max_len))) for _ in range(n)]
        seqs.sort(key=lambda s: -len(s))
        probs.append(seqs)
    return probs


def subgroup_from_golden(inp):
    import numpy as np
    return synth.Subgroup(inp["gene"], inp["pos"], inp["cigar"], inp["seq"], inp["cn"],
                          np.asarray(inp["pair_off"], dtype=np.int32), np.asarray(inp["pair_val"], dtype=np.int32),
                          int(sum(inp["cn"])))
