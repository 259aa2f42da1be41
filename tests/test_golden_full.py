"""CPU: the full-size golden fixtures (tests/golden/full_*.json.gz, the unmodified reference's output on the bench
workloads) are present, well-formed, and belong to the inputs bench.py generates today -- rambl_b200.synth must
reproduce every stored input bit for bit from the stored spec, else the GPU parity test at benchmark size would be
checking a different workload than the one that is timed."""
import os

import pytest

from helpers import ROOT, load_golden_gz
from rambl_b200 import synth

CASES = ["config0_seed0", "config1_seed0", "config2_sub0", "config2_sub3", "config2_sub4", "config4_full",
               "config4_1000_ie001", "config3_100k"]


@pytest.mark.parametrize("name", CASES)
def test_full_golden_belongs_to_todays_workload(name):
    path = os.path.join(ROOT, "tests", "golden", "full_%s.json.gz" % name)
    assert os.path.exists(path), "run tools/make_golden_full.py in the build container"
    case = load_golden_gz("full_%s.json.gz" % name)
    spec = dict(case["spec"])
    if "divergence" in spec:
        spec["divergence"] = tuple(spec["divergence"])
    sg = synth.make_subgroup(**spec)
    inp = case["input"]
    assert sg.gene == inp["gene"] and sg.pos == inp["pos"] and sg.cigar == inp["cigar"] and sg.seq == inp["seq"]
    assert sg.cn == inp["cn"]
    assert [int(x) for x in sg.pair_off] == inp["pair_off"] and [int(x) for x in sg.pair_val] == inp["pair_val"]
    assert case["n_reads"] == sg.n_reads and case["n_raw_reads"] == sg.n_raw_reads
    st = case["strains"]
    assert set(st) == {"infer", "assign", "final"} and len(st["final"]) >= 1
    assert abs(sum(s["abundance"] for s in st["final"]) - 1.0) < 1e-9
    assert case["reference_seconds"]["infer_and_assign"] > 10  # minutes of CPU: why these are fixtures, not live runs


def test_bench_workload_definitions_match_the_golden_specs():
    """bench.py's configs[2] subgroup k and configs[1] block are the calls the fixtures were generated with."""
    c = load_golden_gz("full_config2_sub3.json.gz")
    assert c["spec"] == dict(n_reads=5000, read_len=150, n_strains=2 + 3 % 5, seed=3)
    sg = synth.config2_subgroup(3)
    assert sg.seq == c["input"]["seq"] and sg.pos == c["input"]["pos"]


@pytest.mark.parametrize("name", CASES)
def test_host_graph_construction_at_full_size_equals_the_reference(name):
    """The host half of graph construction (pog.cpp) at benchmark size, with the insertion rows supplied by the oracle's
    alignment: node dump and output_edge text hash to what the unmodified reference produced (the GPU parity test repeats
    this with the rows from the device kernel)."""
    import hashlib
    from helpers import strip_sib
    from oracle import refpy
    from rambl_b200 import api
    case = load_golden_gz("full_%s.json.gz" % name)
    inp = case["input"]
    spec = dict(case["spec"])
    if "divergence" in spec:
        spec["divergence"] = tuple(spec["divergence"])
    sg = synth.make_subgroup(**spec)
    assert sg.seq == inp["seq"]
    b = api.StrainCallBatch()
    b.add(sg)
    b.thread_reads()
    b.finish_graphs_with_rows([refpy.msa_align(p, "oracle") for p in b.msa_problems()])
    assert b.num_nodes(0) == case["n_nodes"]
    assert hashlib.sha256(strip_sib(b.graph_dump(0)).encode()).hexdigest() == case["dump_nosib_sha256"]
    assert hashlib.sha256(b.output_edge(0).encode()).hexdigest() == case["edges_sha256"]
    # what the device walk will do with it is a property of the graph: every fixture goes to the walk kernel as a whole, and
    # none has a level on which a one-letter strain label could meet a multi-letter read string (deviation (i))
    plan = b.walk_plan(0)
    assert plan["eligible"] and not plan["handoff"] and plan["reason"] == 0 and plan["offtable_levels"] == 0, plan
    assert plan["levels"] > 100 and 0 < plan["max_draws"] <= 40000 * 8  # (collapsed paths shorten the walk, insertion columns lengthen it)
    b.close()


def test_subgroup_whose_graph_is_not_strictly_levelled_is_handed_over_at_the_early_end():
    """configs[2] subgroup 383: "$" shares a level with other nodes, every node sits on two levels (twice the read-pool entries
    of its neighbours -- why its chain is the longest of the workload); the walk takes it up to that level."""
    from oracle import refpy
    from rambl_b200 import api
    plans = {}
    for k in (382, 383):
        sg = synth.config2_subgroup(k)
        b = api.StrainCallBatch()
        b.add(sg)
        b.thread_reads()
        b.finish_graphs_with_rows([refpy.msa_align(p, "oracle") for p in b.msa_problems()])
        plans[k] = b.walk_plan(0)
        b.close()
    assert plans[382]["eligible"] and not plans[382]["handoff"]
    assert plans[383]["eligible"] and plans[383]["handoff"] and plans[383]["offtable_levels"] == 0
    assert plans[383]["entries"] > 1.8 * plans[382]["entries"]
