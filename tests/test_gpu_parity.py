"""GPU: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs, against the
golden vectors of the real reference, and -- at full size -- through size-independent properties."""
import numpy as np
import pytest

from oracle import refpy
from rambl_b200 import api, synth

from helpers import (REL_TOL, compare_strains, fuzz_spec, load_golden, load_golden_gz, msa_fuzz_problems,
                     normalise_golden_strains, strip_sib, subgroup_from_golden)

pytestmark = pytest.mark.gpu


def _solve(sgs, **kw):
    b = api.StrainCallBatch()
    for sg in sgs:
        b.add(sg)
    b.build_graphs()
    b.infer(keep_loglik=True, **kw)
    return b


def test_device_present_and_library_native():
    assert api.device_count() >= 1
    st = _solve([synth.make_subgroup(60, 40, 2, seed=1, window=(0, 100))]).stats()
    assert st["gpu_launches"] > 0


# ---- alignment kernel: bit-exact rows -------------------------------------------------------------
def test_msa_kernel_matches_golden_reference_rows():
    cases = load_golden("msa_golden.json")
    rows, _ = api.msa_align_batch([c["seqs"] for c in cases])
    for c, r in zip(cases, rows):
        assert r == c["rows"], c["seqs"]


def test_msa_kernel_matches_oracle_fuzz():
    probs = msa_fuzz_problems(11, 400, max_n=14, max_len=13)
    rows, st = api.msa_align_batch(probs)
    assert st["dp_cells"] > 0
    for p, r in zip(probs, rows):
        assert r == refpy.msa_align(p, "oracle"), p


def test_msa_kernel_edge_cases():
    probs = [["ACGT"], ["AC", "AC"], ["ACGT", "T"], ["NNN", "N"], ["A" * 40, "A" * 3, "A"],
             ["ACGTACGTACGT", "ACGT", "GT", "G"] * 3, ["T" * 9] + ["T" * k for k in range(8, 0, -1)] * 6]
    probs = [sorted(p, key=lambda s: -len(s)) for p in probs]
    rows, _ = api.msa_align_batch(probs)
    for p, r in zip(probs, rows):
        assert r == refpy.msa_align(p, "oracle"), p
    assert api.msa_align_batch([])[0] == []


def test_msa_kernel_properties_at_scale():
    """Many wide problems (deep homopolymer levels): every row, gaps removed, is its input; all rows of a
    problem have the same width; aligning a problem alone or inside a big batch gives the same rows."""
    rnd = np.random.default_rng(5)
    probs = []
    for _ in range(2000):
        n = int(rnd.integers(2, 120))
        base = "".join(rnd.choice(list("ACGT"), size=12))
        seqs = []
        for _ in range(n):
            k = int(rnd.integers(1, 13))
            a = int(rnd.integers(0, 13 - k))
            seqs.append(base[a:a + k])
        probs.append(sorted(seqs, key=lambda s: -len(s)))
    rows, st = api.msa_align_batch(probs)
    for p, r in zip(probs, rows):
        assert len({len(x) for x in r}) == 1
        assert [x.replace("-", "") for x in r] == p
    for k in (0, 777, 1999):
        assert api.msa_align_batch([probs[k]])[0][0] == rows[k]
    for k in (3, 1234):
        assert rows[k] == refpy.msa_align(probs[k], "oracle")


def test_msa_size_classes_and_oversize_problems():
    """Problems are dealt into shared-memory size classes by their longest string; a profile that outgrows its class
    moves up, and anything beyond the largest class runs with its tables in global memory -- no compiled-in limit
    (the reference allocates on the heap).  Every class boundary and the oversize path, against the oracle."""
    rnd = np.random.default_rng(9)
    probs = [["A" * 100, "C" * 3], ["ACGT" * 40, "TTGA" * 30, "G" * 70, "AC"], ["A" * 64, "A" * 63, "C"]]
    for lm in (8, 9, 16, 17, 32, 33, 63, 64):
        for n in (2, 5, 40):
            seqs = ["".join(rnd.choice(list("ACGT"), size=int(rnd.integers(1, lm + 1)))) for _ in range(n)]
            seqs[0] = "".join(rnd.choice(list("ACGT"), size=lm))
            probs.append(sorted(seqs, key=lambda x: -len(x)))
    # many distinct short strings: the profile grows past the 16 columns of the smallest class
    probs.append(sorted(["".join(rnd.choice(list("ACGT"), size=int(rnd.integers(1, 9)))) for _ in range(60)], key=lambda x: -len(x)))
    rows, st = api.msa_align_batch(probs)
    for p, r in zip(probs, rows):
        assert r == refpy.msa_align(p, "oracle"), p


# ---- whole hot path ------------------------------------------------------------------------------
def test_graphs_built_on_device_match_golden_reference():
    for case in load_golden("pog_golden.json"):
        sg = subgroup_from_golden(case["input"])
        b = api.StrainCallBatch()
        b.add(sg)
        b.build_graphs()
        assert strip_sib(b.graph_dump(0)) == strip_sib(case["dump"]), case["name"]
        assert b.output_edge(0) == case["edges"], case["name"]


def test_strains_match_golden_reference():
    for case in load_golden("pog_golden.json"):
        if case["strains"] is None:
            continue
        sg = subgroup_from_golden(case["input"])
        b = _solve([sg])
        assert b.status(0) == api.RAMBL_OK
        got = refpy.parse_strain_dump(b.strains_text(0))
        assert compare_strains(normalise_golden_strains(case["strains"]), got) == [], case["name"]


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 5, 6, 8, 9, 12, 15, 18, 22])
def test_strains_match_oracle(seed):
    sg = synth.make_subgroup(**fuzz_spec(seed))
    o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="oracle")
    want, _ = o.infer(sg.pair_off, sg.pair_val, do_assign=False)
    b = _solve([sg])
    if len(want["infer"]) == 0:  # the reference is undefined here; the product reports it
        assert b.status(0) in (api.RAMBL_ERR_NO_STRAINS, api.RAMBL_ERR_CAPACITY)
        return
    want, _ = o.infer(sg.pair_off, sg.pair_val)
    assert b.status(0) == api.RAMBL_OK
    assert b.output_edge(0) == o.edges()
    got = refpy.parse_strain_dump(b.strains_text(0))
    assert compare_strains(want, got) == []
    # identical cluster assignments under the same seed: same strains, same order, same paths
    assert [s["path"] for s in want["final"]] == [s["path"] for s in got["final"]]


@pytest.mark.parametrize("seed", [1, 6, 15])
def test_strain_search_on_a_graph_built_elsewhere(seed):
    """the host keeps its own graph builder (here: the oracle's node list, or the
    real reference's when oracle/_ref is present) and only the strain search runs on the device."""
    sg = synth.make_subgroup(**fuzz_spec(seed))
    variant = "" if refpy.available("") else "oracle"
    o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant=variant)
    want, _ = o.infer(sg.pair_off, sg.pair_val)
    b = api.StrainCallBatch()
    b.add_graph(refpy.parse_graph_dump(o.dump()), sg.cn, sg.pair_off, sg.pair_val)
    b.infer(keep_loglik=True)
    assert b.status(0) == api.RAMBL_OK
    got = refpy.parse_strain_dump(b.strains_text(0))
    assert compare_strains(want, got) == []


def test_unusual_letters_and_clips():
    """N in the gene (the strain letter then copies the read letter, NonparametricClustering.cpp:358),
    soft clips and '='/'X' CIGAR
    operations (parse_cigar), duplicate reads (copy numbers > 1).  (Reads with N never reach the graph:
    StrainCall.cpp:547-550 drops them.)"""
    gene = "ACGTNACGTTGCATGCAAGGNTTACGATCGATTACGGATCCAT"
    reads = [(0, "43M", gene.replace("N", "A"), 3), (0, "43M", gene.replace("N", "C"), 2),
             (2, "10=1X9M", "GTAACGTTGCTTGCAAGGCT", 1), (5, "2S12M", "TTACGTTGCATGCA", 1),
             (8, "6M1I8M", "TTGCATAGCAAGGAT", 2), (8, "6M2D8M", "TTGCATCAAGGCTT", 1),
             (20, "23M", "CTTACGATCGATTACGGATCCAT", 1), (1, "12M", "CGTCACGTTGCA", 1)]
    pos, cig, seq, cn = zip(*reads)
    o = refpy.RefPog(gene, pos, cig, seq, cn, variant="oracle")
    pair_off = [0]
    for c in cn:
        pair_off.append(pair_off[-1] + c)
    pair_val = [-1] * pair_off[-1]
    want, _ = o.infer(pair_off, pair_val)
    b = api.StrainCallBatch()
    b.add_subgroup(gene, pos, cig, seq, cn)
    b.build_graphs()
    assert strip_sib(b.graph_dump(0)) == strip_sib(o.dump())
    b.infer(keep_loglik=True)
    if len(want["infer"]) == 0:
        assert b.status(0) != api.RAMBL_OK
        return
    assert b.status(0) == api.RAMBL_OK
    got = refpy.parse_strain_dump(b.strains_text(0))
    assert compare_strains(want, got) == []


def test_batch_with_an_empty_and_a_readless_subgroup():
    """A batch where one subgroup has no reads at all (StrainCall skips such windows, StrainCall.cpp:1009)
    next to normal ones: the others are unaffected."""
    sg = synth.make_subgroup(**fuzz_spec(2))
    b = api.StrainCallBatch()
    b.add(sg)
    b.add_subgroup("ACGTACGTAC", [], [], [], [])
    b.add(sg)
    b.build_graphs()
    b.infer()
    assert b.status(0) == api.RAMBL_OK and b.status(2) == api.RAMBL_OK
    assert b.strains_text(0) == b.strains_text(2)
    # the backbone-only graph has one path and no reads: one strain, abundance 1 after read_assign's normalise
    assert b.status(1) in (api.RAMBL_OK, api.RAMBL_ERR_NO_STRAINS)


def test_batch_equals_one_by_one():
    """Subgroups solved together (one launch set per level) give what they give alone."""
    sgs = [synth.make_subgroup(**fuzz_spec(s)) for s in (1, 2, 6, 9, 15)]
    together = _solve(sgs)
    for i, sg in enumerate(sgs):
        alone = _solve([sg])
        assert together.status(i) == alone.status(0)
        assert together.strains_text(i) == alone.strains_text(0)
        assert together.output_edge(i) == alone.output_edge(0)


def test_gibbs_block_count_does_not_change_the_chain():
    """The Gibbs kernels speculate over 1 to 8 blocks of 32 draws per round (positive: kernel chosen by strain
    count, negative: the four-warps-per-block kernel); each setting must settle on the same sequential chain,
    so every output is identical text."""
    sgs = [synth.make_subgroup(**fuzz_spec(s)) for s in (1, 2, 6, 9, 15)]
    sgs.append(synth.make_subgroup(n_reads=800, read_len=100, n_strains=4, seed=31, window=(300, 520)))
    texts = {}
    try:
        for blocks in (1, 2, 4, 8, -1, -2, -4):
            assert api.lib().rambl_set_gibbs_blocks(blocks) == api.RAMBL_OK
            b = _solve(sgs)
            texts[blocks] = [(b.status(i), b.strains_text(i)) for i in range(len(sgs))]
    finally:
        api.lib().rambl_set_gibbs_blocks(0)
    assert api.lib().rambl_set_gibbs_blocks(3) == api.RAMBL_ERR_INVALID
    for blocks in texts:
        assert texts[blocks] == texts[1], blocks
    assert any(st == api.RAMBL_OK for st, _ in texts[1])


def test_device_walk_equals_level_synchronous_path():
    """The strain search runs as one kernel that walks all levels on the device (default) or level by level with the
    host deciding in between (rambl_set_walk_mode(0)); with 1, 2, 4 or 8 warps per subgroup.  Same chains, same
    arithmetic in the same order: identical text, paired reads and collapsed nodes included."""
    sgs = [synth.make_subgroup(**fuzz_spec(s)) for s in (1, 2, 3, 5, 6, 9, 12, 15, 18, 22)]
    sgs.append(synth.make_subgroup(n_reads=800, read_len=100, n_strains=4, seed=31, window=(300, 520)))
    sgs.append(synth.make_subgroup(n_reads=300, read_len=30, n_strains=2, seed=4, window=(500, 580), sub_err=0.002,
                                   paired=True, divergence=(0.03, 0.06)))
    sgs.append(synth.make_subgroup(n_reads=60, read_len=60, n_strains=2, seed=7, window=(200, 420), sub_err=0.0,
                                   divergence=(0.01, 0.02)))  # clean reads: long collapsed nodes
    L = api.lib()
    texts = {}
    try:
        assert L.rambl_set_walk_mode(0) == api.RAMBL_OK
        b = _solve(sgs)
        texts["level-synchronous"] = [(b.status(i), b.strains_text(i)) for i in range(len(sgs))]
        assert b.stats()["dpm_launches"] == 0
        assert L.rambl_set_walk_mode(1) == api.RAMBL_OK
        for nb in (0, 1, 2, 4, 8):
            assert L.rambl_set_walk_blocks(nb) == api.RAMBL_OK
            b = _solve(sgs)
            texts["walk/%d" % nb] = [(b.status(i), b.strains_text(i)) for i in range(len(sgs))]
            assert b.stats()["dpm_launches"] == 1
        assert L.rambl_set_walk_blocks(0) == api.RAMBL_OK
        for ctas in (1, 2, 4, 8):  # a thread-block cluster per subgroup: the extra CTAs join the Gibbs chains
            assert L.rambl_set_walk_cluster(ctas) == api.RAMBL_OK
            b = _solve(sgs)
            texts["cluster/%d" % ctas] = [(b.status(i), b.strains_text(i)) for i in range(len(sgs))]
    finally:
        L.rambl_set_walk_mode(1)
        L.rambl_set_walk_blocks(0)
        L.rambl_set_walk_cluster(0)
    assert L.rambl_set_walk_blocks(3) == api.RAMBL_ERR_INVALID and L.rambl_set_walk_cluster(3) == api.RAMBL_ERR_INVALID
    for k in texts:
        for i in range(len(sgs)):
            assert texts[k][i] == texts["level-synchronous"][i], (k, i)
    assert sum(1 for st, _ in texts["walk/0"] if st == api.RAMBL_OK) >= 8


def test_paired_reads_and_copies():
    sg = synth.make_subgroup(n_reads=300, read_len=30, n_strains=2, seed=4, window=(500, 580), sub_err=0.002,
                             paired=True, divergence=(0.03, 0.06))
    assert max(sg.cn) > 1 and (sg.pair_val >= 0).any()
    o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="oracle")
    want, _ = o.infer(sg.pair_off, sg.pair_val)
    got = refpy.parse_strain_dump(_solve([sg]).strains_text(0))
    assert compare_strains(want, got) == []


@pytest.mark.parametrize("n,ie,W", [(200, 0.002, 450), (400, 0.001, 600)])
def test_config4_like_indel_rich_250bp_reads(n, ie, W):
    """BASELINE configs[4] at a size the oracle finishes: 250bp reads, homopolymer-biased indel errors and
    indel-bearing strains (wide graphs, insertion alignment on the device), against the oracle."""
    sg = synth.make_subgroup(n, 250, 3, indel_err=ie, indel_frac=0.3, homopolymer_bias=True, seed=4,
                             window=(300, 300 + W), divergence=(0.01, 0.03))
    o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="oracle")
    want, _ = o.infer(sg.pair_off, sg.pair_val)
    assert len(want["final"]) > 0
    b = _solve([sg])
    assert b.status(0) == api.RAMBL_OK
    assert strip_sib(b.graph_dump(0)) == strip_sib(o.dump())
    assert b.output_edge(0) == o.edges()
    got = refpy.parse_strain_dump(b.strains_text(0))
    assert compare_strains(want, got) == []


def test_depth800_sample_of_configs1_matches_reference():
    """configs[1] restricted to a 160 bp window (depth 800, 150 bp reads, 10 strains, ~840 reads, 50 sweeps of ~800
    draws per level) -- the deepest case the reference finishes in seconds.  Strain paths, order, consensus and
    abundances against the real reference when oracle/_ref is present (else the oracle)."""
    sg = synth.make_subgroup(int(20000 * 160 / 1542), 150, 10, divergence=(0.01, 0.03), seed=0, window=(600, 760))
    variant = "" if refpy.available("") else "oracle"
    o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant=variant)
    want, _ = o.infer(sg.pair_off, sg.pair_val)
    b = _solve([sg])
    assert b.status(0) == api.RAMBL_OK
    assert b.output_edge(0) == o.edges()
    got = refpy.parse_strain_dump(b.strains_text(0))
    assert compare_strains(want, got) == []
    assert b.stats()["draws"] > 1000000


def test_matched_sample_set_of_the_bench_matches_reference():
    """The sample set bench.py's reference arm times (configs[2] subgroups on a gene window) solved as ONE batch,
    subgroup by subgroup against the reference (oracle/_ref when present, else the oracle): the like-for-like pair
    of the bench line is also a parity case."""
    import bench
    w, sgs = bench.matched_sample_set(20)
    b = _solve(sgs)
    variant = "" if refpy.available("") else "oracle"
    for i, sg in enumerate(sgs[:6]):
        o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant=variant)
        want, _ = o.infer(sg.pair_off, sg.pair_val)
        assert b.status(i) == api.RAMBL_OK
        got = refpy.parse_strain_dump(b.strains_text(i))
        assert compare_strains(want, got) == [], i


FULL_GOLDEN = ["config0_seed0", "config1_seed0", "config2_sub0", "config2_sub3", "config2_sub4", "config4_full",
               "config4_1000_ie001", "config3_100k"]


@pytest.mark.parametrize("name", FULL_GOLDEN)
def test_full_size_matches_golden_reference(name):
    """BASELINE configs[0], configs[1] (the single-chain bench block), whole-gene subgroups of configs[2] (the bench
    workload), configs[4] at full size (5 000 x 250 bp reads, indel-rich strains: wide graphs, the insertion alignment
    on the device) and with read-level homopolymer indel errors at the largest size the reference finishes, and
    configs[3] (100k raw reads of 50 strains through the depth-800 down-sampling) against what the unmodified reference
    produced for the same inputs
    (tests/golden/full_*.json.gz, generated by tools/make_golden_full.py; the reference needs 1-13 minutes per case):
    graph (node dump and output_edge text by hash), strain paths, order, consensus, abundances, substitution tables
    and every per-read log-likelihood."""
    import hashlib
    case = load_golden_gz("full_%s.json.gz" % name)
    sg = subgroup_from_golden(case["input"])
    b = _solve([sg])
    assert b.status(0) == api.RAMBL_OK
    assert b.num_nodes(0) == case["n_nodes"]
    assert hashlib.sha256(strip_sib(b.graph_dump(0)).encode()).hexdigest() == case["dump_nosib_sha256"]
    assert hashlib.sha256(b.output_edge(0).encode()).hexdigest() == case["edges_sha256"]
    got = refpy.parse_strain_dump(b.strains_text(0))
    want = normalise_golden_strains(case["strains"])
    assert compare_strains(want, got) == []
    assert [s["path"] for s in want["final"]] == [s["path"] for s in got["final"]]
    # no level on which a one-letter strain label could meet a multi-letter read string (the one place
    # where the reference's std::map keys leave the 6x6 substitution table and this library does not follow)
    assert b.stats()["offtable_levels"] == 0


@pytest.mark.parametrize("cfg", [0, 1])
def test_full_size_properties(cfg):
    """BASELINE configs[0] (2k 100bp reads, 3 strains) and configs[1] (20k 150bp reads, 10 strains, the bench
    workload) at full size on the whole 16S gene: too slow for the oracle (the reference needs 13 minutes for
    configs[1]), so check what must hold at any size: the run is deterministic, abundances are normalised,
    every strain is a ^...$ path along edges of the graph whose letters are its consensus, FASTA = strains
    above tau in abundance order."""
    sg = synth.config_workload(cfg, seed=1)[0]
    b1, b2 = _solve([sg]), _solve([sg])
    assert b1.status(0) == api.RAMBL_OK
    assert b1.strains_text(0) == b2.strains_text(0)
    # ... and the chain does not depend on how many draws a round speculates over: the default (warp-per-block
    # kernel, up to 256 draws per round at ~48 candidate strains) against the four-warps-per-block kernel at
    # 32 draws per round (the configuration checked against the reference at depth 800) and against 64
    try:
        for blocks in (-1, 2):
            assert api.lib().rambl_set_gibbs_blocks(blocks) == api.RAMBL_OK
            assert _solve([sg]).strains_text(0) == b1.strains_text(0), blocks
    finally:
        api.lib().rambl_set_gibbs_blocks(0)
    st = b1.strains(0)
    assert abs(sum(s.abundance for s in st) - 1.0) < 1e-9
    nodes = refpy.parse_graph_dump(b1.graph_dump(0))
    for s in st:
        assert nodes[s.path[0]]["label"] == "^" and nodes[s.path[-1]]["label"] == "$"
        for u, v in zip(s.path, s.path[1:]):
            assert v in nodes[u]["out"]
        assert s.plain_seq() == "".join(nodes[u]["label"] for u in s.path
                                        if nodes[u]["label"] not in ("^", "$", "-", "="))
    order = b1.order(0)
    ab = [st[k].abundance for k in order]
    assert ab == sorted(ab, reverse=True)
    fasta = b1.fasta(0, "g", 1, len(sg.gene), 0.02).strip().split("\n")
    assert len(fasta) == 2 * sum(1 for a in ab if a >= np.float32(0.02))
    # consensus sequences are full-length 16S candidates
    for k in order:
        assert abs(len(st[k].plain_seq()) - len(sg.gene)) <= 0.05 * len(sg.gene)
    # every read's log-likelihood under every kept strain is a finite, non-positive number
    for s in b1.strains(0, with_loglik=True):
        assert np.all(np.isfinite(s.read_loglik)) and np.all(s.read_loglik <= 0)


@pytest.mark.parametrize("layout", [{}, {"RAMBL_SOLVE_FIRST": "7"}, {"RAMBL_SOLVE_CHUNKS": "3"},
                                    {"RAMBL_SOLVE_CHUNKS": "5", "RAMBL_SOLVE_DRIVERS": "3"}])
def test_overlapped_solve_equals_build_then_infer(layout, monkeypatch):
    """rambl_batch_solve (chunks on two streams, graph construction of one chunk under the strain search of the previous)
    gives, subgroup by subgroup, the text of rambl_batch_build_graphs + rambl_batch_infer -- for the default layout (a
    batch of up to one wave is one chunk) and for layouts that force several chunks on this small batch."""
    sgs = [synth.make_subgroup(**fuzz_spec(s)) for s in range(24)]
    a = _solve(sgs)
    for k, v in layout.items():
        monkeypatch.setenv(k, v)
    b = api.StrainCallBatch()
    for sg in sgs:
        b.add(sg)
    b.solve(keep_loglik=True)
    for i in range(len(sgs)):
        assert a.status(i) == b.status(i), i
        if a.status(i) == api.RAMBL_OK:
            assert a.strains_text(i) == b.strains_text(i), i
        assert a.output_edge(i) == b.output_edge(i), i
    assert b.stats()["gpu_launches"] > 0


def test_config2_slice_batched_equals_sharded():
    """A slice of configs[2] (subgroups x 5k reads): one batch of 6 subgroups gives, subgroup by subgroup, what
    two 'ranks' of 3 give -- the multi-GPU sharding changes nothing in the results."""
    from rambl_b200 import shard
    sgs = [synth.make_subgroup(5000, 150, 2 + (k % 5), seed=k, window=(0, 400)) for k in range(6)]
    whole = _solve(sgs)
    parts = shard.assign([shard.cost_proxy(s) for s in sgs], 2)
    assert sorted(parts[0] + parts[1]) == list(range(6))
    for mine in parts:
        b = _solve([sgs[i] for i in mine])
        for local, i in enumerate(mine):
            assert b.status(local) == whole.status(i)
            assert b.strains_text(local) == whole.strains_text(i)
