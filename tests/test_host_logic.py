"""CPU: the C-ABI library loads and exports every declared symbol, the host graph construction of the
product (phases A and C, with the device alignment step supplied by the test) equals the oracle's, and
the product fails loudly -- never silently computes on the CPU -- when there is no device."""
import ctypes
import os
import re

import pytest

from oracle import refpy
from rambl_b200 import api, synth

from helpers import ROOT, fuzz_spec, load_golden, strip_sib, subgroup_from_golden

try:
    HAVE_GPU = api.device_count() > 0
except Exception:
    HAVE_GPU = False


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rambl_b200.h")).read()
    declared = set(re.findall(r"\b(rambl_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = ctypes.CDLL(api.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export: " + name
    assert declared == set(api.SYMBOLS), declared ^ set(api.SYMBOLS)
    api.lib()


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "rambl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".cuh", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "refpy" not in txt and "import oracle" not in txt, f
                assert "from oracle" not in txt, f


def _build_with_supplied_rows(sg):
    b = api.StrainCallBatch()
    b.add(sg)
    b.thread_reads()
    probs = b.msa_problems()
    b.finish_graphs_with_rows([refpy.msa_align(p, "oracle") for p in probs])
    return b, len(probs)


@pytest.mark.parametrize("seed", list(range(0, 24)))
def test_graph_construction_matches_oracle(seed):
    spec = fuzz_spec(seed)
    spec["n_reads"] *= 2
    sg = synth.make_subgroup(**spec)
    if sg.n_unique == 0:
        pytest.skip("no reads survived the filters")
    b, _ = _build_with_supplied_rows(sg)
    o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="oracle")
    assert strip_sib(b.graph_dump(0)) == strip_sib(o.dump())
    assert b.output_edge(0) == o.edges()


def test_graph_construction_with_alignment_problems():
    """Indel-rich homopolymer reads: several levels need the sum-of-pairs alignment."""
    total = 0
    for seed in (21, 35, 38, 16):
        spec = dict(n_reads=600, read_len=60, n_strains=3, seed=seed, window=(100, 400), sub_err=0.005,
                    indel_err=0.03, indel_frac=0.4, homopolymer_bias=True, divergence=(0.02, 0.06))
        sg = synth.make_subgroup(**spec)
        b, nprob = _build_with_supplied_rows(sg)
        total += nprob
        o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="oracle")
        assert strip_sib(b.graph_dump(0)) == strip_sib(o.dump())
        assert b.output_edge(0) == o.edges()
    assert total > 0, "the cases were meant to exercise the alignment path"


def test_graph_matches_golden_reference_dumps():
    for case in load_golden("pog_golden.json"):
        sg = subgroup_from_golden(case["input"])
        b, _ = _build_with_supplied_rows(sg)
        assert strip_sib(b.graph_dump(0)) == strip_sib(case["dump"]), case["name"]
        assert b.output_edge(0) == case["edges"], case["name"]


def test_empty_and_degenerate_inputs():
    b = api.StrainCallBatch()
    b.add_subgroup("ACGTACGT", [], [], [], [])  # no reads at all: backbone only
    b.thread_reads()
    assert b.msa_problems() == []
    b.finish_graphs_with_rows([])
    o = refpy.RefPog("ACGTACGT", [], [], [], [], variant="oracle")
    assert b.output_edge(0) == o.edges()
    # malformed input is refused, not guessed at
    bad = api.StrainCallBatch()
    bad.add_subgroup("ACGT", [0], ["9M"], ["ACGTACGTA"], [1])
    with pytest.raises(api.RamblError) as ei:
        bad.thread_reads()
    assert ei.value.code == api.RAMBL_ERR_INVALID
    with pytest.raises(api.RamblError):
        api.StrainCallBatch().add_subgroup("ACGT", [0], ["4M"], ["ACGT"], [0])  # copy number 0


def test_calls_in_wrong_order_are_errors():
    b = api.StrainCallBatch()
    b.add_subgroup("ACGTACGT", [0], ["8M"], ["ACGTACGT"], [1])
    with pytest.raises(api.RamblError) as ei:
        b.infer()
    assert ei.value.code in (api.RAMBL_ERR_STATE, api.RAMBL_ERR_CUDA)


@pytest.mark.skipif(HAVE_GPU, reason="only meaningful on a box without a device")
def test_gibbs_block_hook_accepts_only_the_documented_values():
    """rambl_set_gibbs_blocks (a measurement / test hook, include/rambl_b200.h): 0 = automatic, 1/2/4/8 blocks of
    32 draws per round, -1/-2/-4 the same on the four-warps-per-block kernel; anything else is refused."""
    try:
        for ok in (1, 2, 4, 8, -1, -2, -4, 0):
            assert api.lib().rambl_set_gibbs_blocks(ok) == api.RAMBL_OK, ok
        for bad in (3, 5, 16, -3, -8, 100):
            assert api.lib().rambl_set_gibbs_blocks(bad) == api.RAMBL_ERR_INVALID, bad
    finally:
        api.lib().rambl_set_gibbs_blocks(0)


def test_no_device_is_a_loud_error():
    with pytest.raises(api.RamblError) as ei:
        api.msa_align_batch([["ACG", "A"]])
    assert ei.value.code == api.RAMBL_ERR_CUDA
    b = api.StrainCallBatch()
    b.add_subgroup("ACGTACGT", [0], ["8M"], ["ACGTACGT"], [1])
    with pytest.raises(api.RamblError) as ei:
        b.build_graphs()
    assert ei.value.code == api.RAMBL_ERR_CUDA
    b2 = api.StrainCallBatch()
    b2.add_subgroup("ACGTACGT", [0], ["8M"], ["ACGTACGT"], [1])
    b2.thread_reads()
    b2.finish_graphs_with_rows([])
    with pytest.raises(api.RamblError) as ei:
        b2.infer()
    assert ei.value.code == api.RAMBL_ERR_CUDA


def test_synthetic_downsampling_follows_std_mt19937():
    """synth reproduces std::mt19937(1234) + uniform_real_distribution (StrainCall.cpp:491-493)."""
    u = synth.StdMt19937(1234).canonical(3)
    # first raw outputs of std::mt19937 seeded with 1234 (the standard fixes the sequence)
    raw = synth.StdMt19937(1234).raw(2)
    assert int(raw[0]) == 822569775 and int(raw[1]) == 2137449171
    assert abs(u[0] - (822569775 + 2137449171 * 4294967296.0) / 18446744073709551616.0) < 1e-18


def test_graph_built_elsewhere_gets_the_reference_edge_counts():
    """rambl_batch_add_graph: a graph flattened from the oracle's node list; the reads-over-edge counts the
    library derives must be the ones output_edge prints."""
    for seed in (3, 9, 12):
        sg = synth.make_subgroup(**fuzz_spec(seed))
        o = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant="oracle")
        b = api.StrainCallBatch()
        b.add_graph(refpy.parse_graph_dump(o.dump()), sg.cn, sg.pair_off, sg.pair_val)
        assert b.num_nodes(0) == o.num_nodes()
        mine = [l for l in b.output_edge(0).split("\n") if l and not l.startswith("#")]
        want = [l for l in o.edges().split("\n") if l and not l.startswith("#")]
        assert mine == want
    bad = api.StrainCallBatch()
    with pytest.raises(api.RamblError):
        bad.add_graph([dict(st=0, label="A", out=[1], pool=[]), dict(st=0, label="$", out=[], pool=[])], [])


def _random_cigar_case(rnd):
    L = rnd.randint(8, 60)
    gene = "".join(rnd.choice("ACGT") for _ in range(L))
    reads = []
    for _ in range(rnd.randint(1, 25)):
        pos = rnd.randint(0, L - 2)
        ops, ref, seq = [], pos, []
        for k in range(rnd.randint(1, 6)):
            room = L - ref
            if room <= 0:
                break
            op = rnd.choice("MMMMIDS=X") if k else rnd.choice("MMMMMIS=")
            if op == "S" and k != 0:
                op = "M"
            n = rnd.randint(1, 5)
            if op in "M=X":
                n = min(n, room)
                seq += [gene[ref + q] if rnd.random() < 0.8 else rnd.choice("ACGT") for q in range(n)]
                ref += n
            elif op == "D":
                n = min(n, room)
                ref += n
            else:  # I, S
                seq += [rnd.choice("ACGT") for _ in range(n)]
            ops.append("%d%s" % (n, op))
        if ops:
            reads.append((pos, "".join(ops), "".join(seq), rnd.randint(1, 3)))
    return gene, reads


def test_random_cigars_match_oracle_or_are_refused():
    """Arbitrary CIGAR shapes (reads that begin with an insertion, end in a deletion, stacked indels, clips,
    =/X): the product builds the oracle's graph, or refuses the input with RAMBL_ERR_INVALID -- including the
    alignments on which the reference's construction never returns (a budget on the construction loops)."""
    import random
    import signal
    rnd = random.Random(2)
    refused = 0
    for _ in range(250):
        gene, reads = _random_cigar_case(rnd)
        if not reads:
            continue
        pos, cig, seq, cn = zip(*reads)
        b = api.StrainCallBatch()
        b.add_subgroup(gene, pos, cig, seq, cn)
        try:
            b.thread_reads()
            b.finish_graphs_with_rows([refpy.msa_align(p, "oracle") for p in b.msa_problems()])
        except api.RamblError as e:
            assert e.code == api.RAMBL_ERR_INVALID
            refused += 1
            continue
        signal.alarm(60)  # the oracle has no such budget
        o = refpy.RefPog(gene, pos, cig, seq, cn, variant="oracle")
        signal.alarm(0)
        assert strip_sib(b.graph_dump(0)) == strip_sib(o.dump()), (gene, reads)
        assert b.output_edge(0) == o.edges()
    assert refused < 125


def test_host_block_cache_keeps_and_reuses_the_flat_graph_arrays():
    """The per-entry arrays of a flat graph (>= 64 KB each) go to the host block cache when the batch is closed, the next
    batch takes them from there, and rambl_release_cached_memory() gives everything back."""
    sg = synth.make_subgroup(n_reads=1500, read_len=100, n_strains=3, seed=5)
    api.release_cached_memory()
    assert api.cached_host_bytes() == 0
    b, _ = _build_with_supplied_rows(sg)
    dump = b.graph_dump(0)
    assert api.cached_host_bytes() == 0  # in use, nothing released yet
    b.close()
    held = api.cached_host_bytes()
    assert held >= 3 * 64 * 1024  # read ids, copies, string offsets (4 bytes per entry each) at least
    b2, _ = _build_with_supplied_rows(sg)
    assert api.cached_host_bytes() < held  # the same sizes again: taken from the cache
    assert b2.graph_dump(0) == dump
    b2.close()
    assert api.cached_host_bytes() == held
    api.release_cached_memory()
    assert api.cached_host_bytes() == 0


def test_host_block_cache_respects_its_limit():
    """RAMBL_HOST_CACHE_MB bounds what is kept; the blocks released longest ago are dropped first."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from rambl_b200 import api, synth
        from test_host_logic import _build_with_supplied_rows
        held = []
        for n in (1500, 3000, 1500):
            b, _ = _build_with_supplied_rows(synth.make_subgroup(n_reads=n, read_len=100, n_strains=3, seed=5))
            b.close()
            held.append(api.cached_host_bytes())
        print(held)
        assert all(0 < h <= 1 << 20 for h in held), held
    """) % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, RAMBL_HOST_CACHE_MB="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_reads_as_strings_and_as_packed_arenas_give_the_same_graph():
    """rambl_batch_add_subgroup (one C string per read), rambl_batch_add_subgroup_packed, and the packed call on arenas
    whose first read does not start at offset 0 (a slice of a larger buffer) build identical graphs; offsets that
    decrease are refused."""
    import numpy as np
    sg = synth.make_subgroup(**fuzz_spec(3))
    pk = sg.packed()
    dumps = []
    for how in ("strings", "packed", "shifted"):
        b = api.StrainCallBatch()
        if how == "strings":
            b.add_subgroup(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, sg.pair_off, sg.pair_val)
        elif how == "packed":
            b.add(sg)
        else:
            b.add_subgroup_packed(sg.gene, pk["pos"], pk["cigar_off"] + 5, b"JUNK!" + pk["cigar_chars"], pk["seq_off"] + 3,
                                  b"xyz" + pk["seq_chars"], pk["cn"], sg.pair_off, sg.pair_val)
        b.thread_reads()
        b.finish_graphs_with_rows([refpy.msa_align(p, "oracle") for p in b.msa_problems()])
        dumps.append((b.graph_dump(0), b.output_edge(0)))
        b.close()
    assert dumps[0] == dumps[1] == dumps[2]
    bad = api.StrainCallBatch()
    off = np.array(pk["seq_off"], dtype=np.int64)
    off[2] = off[1] - 1
    with pytest.raises(api.RamblError):
        bad.add_subgroup_packed(sg.gene, pk["pos"], pk["cigar_off"], pk["cigar_chars"], off, pk["seq_chars"], pk["cn"])
    bad.close()


def test_input_on_which_the_reference_never_ends_is_refused():
    """Homopolymer indel errors can produce alignments whose canonised graph has a cycle; PartialOrderGraph::build (and the
    oracle's restatement of it) then loops in path_collapse for ever -- checked with oracle/_ref under a 120 s timeout,
    tools/ref_nonterminating.py.  The builder pays every loop into one work budget and refuses such a subgroup with
    RAMBL_ERR_INVALID instead of hanging the batch."""
    seed = 66
    spec = dict(n_reads=300 + 7 * (seed % 50), read_len=60, n_strains=2 + seed % 4, seed=1000 + seed, window=(100, 400),
                sub_err=0.005, indel_err=0.01 + 0.002 * (seed % 15), indel_frac=0.4, homopolymer_bias=True,
                divergence=(0.02, 0.06))
    sg = synth.make_subgroup(**spec)
    b = api.StrainCallBatch()
    b.add(sg)
    b.thread_reads()
    rows = [refpy.msa_align(p, "oracle") for p in b.msa_problems()]
    with pytest.raises(api.RamblError) as e:
        b.finish_graphs_with_rows(rows)
    assert e.value.code == api.RAMBL_ERR_INVALID and "does not terminate" in str(e.value)
    b.close()


def test_chunk_layout_of_the_overlapped_solve(monkeypatch):
    """rambl_solve_layout (no device): up to one wave of the walk kernel is one chunk; a larger batch starts with one wave
    minus 8 SMs and continues in chunks of at most eight waves; the two environment overrides; the bounds always start at
    0, end at N and do not decrease."""
    for k in ("RAMBL_SOLVE_FIRST", "RAMBL_SOLVE_CHUNKS"):
        monkeypatch.delenv(k, raising=False)
    assert api.solve_layout(0) == [0, 0]
    assert api.solve_layout(62) == [0, 62]
    assert api.solve_layout(148) == [0, 148]
    assert api.solve_layout(149) == [0, 140, 149]
    assert api.solve_layout(250) == [0, 140, 250]
    assert api.solve_layout(500) == [0, 140, 500]
    assert api.solve_layout(140 + 8 * 148) == [0, 140, 140 + 8 * 148]
    big = api.solve_layout(5000)
    assert big[:2] == [0, 140] and big[-1] == 5000 and len(big) == 2 + 5  # 4860 subgroups in five chunks of <= 1184
    assert all(b - a <= 8 * 148 for a, b in zip(big[1:], big[2:])) and big == sorted(big)
    assert api.solve_layout(500, sms=132) == [0, 124, 500]
    monkeypatch.setenv("RAMBL_SOLVE_FIRST", "7")
    assert api.solve_layout(24) == [0, 7, 24]
    assert api.solve_layout(5) == [0, 5]
    monkeypatch.delenv("RAMBL_SOLVE_FIRST")
    monkeypatch.setenv("RAMBL_SOLVE_CHUNKS", "3")
    assert api.solve_layout(24) == [0, 8, 16, 24]
    assert api.solve_layout(2) == [0, 1, 2]
    with pytest.raises(api.RamblError):
        api.solve_layout(-1)
