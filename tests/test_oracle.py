"""CPU: the oracle (oracle/liboracle.so) against the committed golden vectors of the real reference and,
where oracle/_ref is present, against the reference itself on fresh seeded inputs."""
import pytest

from oracle import refpy
from rambl_b200 import synth

from helpers import compare_strains, fuzz_spec, load_golden, msa_fuzz_problems, normalise_golden_strains, subgroup_from_golden

HAVE_REF = refpy.available("")


def test_oracle_msa_matches_golden():
    for case in load_golden("msa_golden.json"):
        assert refpy.msa_align(case["seqs"], "oracle") == case["rows"], case["seqs"]


def test_oracle_graph_and_strains_match_golden():
    for case in load_golden("pog_golden.json"):
        sg = subgroup_from_golden(case["input"])
        g = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="oracle")
        assert g.dump() == case["dump"], case["name"]
        assert g.edges() == case["edges"], case["name"]
        if case["strains"] is None:
            continue
        got, _ = g.infer(sg.pair_off, sg.pair_val)
        want = normalise_golden_strains(case["strains"])
        # the oracle restates the reference in the reference's own precision: exact, not "close"
        assert compare_strains(want, got, tol=0.0) == [], case["name"]
        for stage in want:
            assert [s["abundance_ld"] for s in want[stage]] == [s["abundance_ld"] for s in got[stage]]


def test_msa_edge_cases_oracle():
    # single sequence, equal sequences, one-letter sequences, letters outside the scoring alphabet
    assert refpy.msa_align(["ACGT"], "oracle") == ["ACGT"]
    assert refpy.msa_align(["AC", "AC"], "oracle") == ["AC", "AC"]
    rows = refpy.msa_align(["ACGT", "T"], "oracle")
    assert len(rows[0]) == len(rows[1]) and rows[0].replace("-", "") == "ACGT" and rows[1].replace("-", "") == "T"
    rows = refpy.msa_align(["NNN", "N"], "oracle")
    assert len(rows[0]) == len(rows[1])


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_msa_matches_reference_fuzz():
    for seqs in msa_fuzz_problems(5, 150):
        assert refpy.msa_align(seqs, "oracle") == refpy.msa_align(seqs), seqs


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
def test_reference_build_variants_agree():
    """The -O2 build the CPU baseline is timed with and the -O0 build (how setup.py compiles) agree."""
    for seqs in msa_fuzz_problems(6, 40):
        assert refpy.msa_align(seqs, "O0") == refpy.msa_align(seqs)
    sg = synth.make_subgroup(**fuzz_spec(8))
    a = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="O0")
    b = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn)
    c = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="plain")
    assert a.dump() == b.dump() == c.dump()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [2, 6, 8, 12, 15, 22])
def test_oracle_matches_reference_on_fresh_inputs(seed):
    sg = synth.make_subgroup(**fuzz_spec(seed))
    o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="oracle")
    r = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn)
    assert o.dump() == r.dump()
    assert o.edges() == r.edges()
    so, _ = o.infer(sg.pair_off, sg.pair_val)
    assert len(so["infer"]) > 0
    sr, _ = r.infer(sg.pair_off, sg.pair_val)
    assert so == sr  # every printed digit of the long double values included
