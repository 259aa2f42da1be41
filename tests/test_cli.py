"""The StrainCall drop-in command line (rambl_b200/StrainCall) against the reference CLI built from the
unmodified sources (oracle/_ref/StrainCall), both reading the same fixtures through tests/samtools_shim."""
import os
import subprocess

import pytest

from rambl_b200 import api, synth

from helpers import ROOT

CLI = os.path.join(ROOT, "rambl_b200", "StrainCall")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "StrainCall")
SHIM = os.path.join(ROOT, "tests", "samtools_shim")


def run_cli(binary, args, cwd):
    env = dict(os.environ)
    env["PATH"] = SHIM + os.pathsep + env.get("PATH", "")
    r = subprocess.run([binary] + args, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=600)
    return r.returncode, r.stdout, r.stderr


def test_cli_prints_the_reference_usage():
    code, out, err = run_cli(CLI, ["-h"], ROOT)
    assert code == 0 and "StrainCall marker_gene read_mapping" in err and "--max-depth" in err


def test_samtools_shim_roundtrip(tmp_path):
    gene, raw, _ = synth.simulate_raw_reads(50, 40, 2, seed=3, window=(0, 120), indel_err=0.02)
    fa, sam = synth.write_cli_fixture(str(tmp_path), "g1", gene, raw)
    code, out, _ = run_cli(os.path.join(SHIM, "samtools"), ["faidx", fa, "g1:11-30"], str(tmp_path))
    assert code == 0 and "".join(out.split("\n")[1:]) == gene[10:30]
    code, out, _ = run_cli(os.path.join(SHIM, "samtools"), ["view", sam, "-q", "3", "-F", "1804", "g1:1-120"], str(tmp_path))
    assert len(out.strip().split("\n")) == len(raw)
    code, out, _ = run_cli(os.path.join(SHIM, "samtools"), ["mpileup", "-q", "3", "-Q0", "-A", "-r", "g1:1-120", sam], str(tmp_path))
    assert out.count("\n") > 50


@pytest.mark.gpu
@pytest.mark.parametrize("spec", [
    dict(n_reads=150, read_len=50, n_strains=2, seed=21, window=(200, 330), sub_err=0.004, divergence=(0.03, 0.06)),
    dict(n_reads=200, read_len=60, n_strains=3, seed=22, window=(600, 760), sub_err=0.003, indel_err=0.01,
         homopolymer_bias=True, divergence=(0.03, 0.07)),
    dict(n_reads=160, read_len=40, n_strains=2, seed=23, window=(900, 1010), sub_err=0.003, paired=True,
         divergence=(0.03, 0.06)),
])
def test_cli_matches_reference_cli(tmp_path, spec):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/StrainCall not built")
    gene, raw, _ = synth.simulate_raw_reads(**spec)
    fa, sam = synth.write_cli_fixture(str(tmp_path), "gene7", gene, raw)
    args = ["-r", "gene7:1-%d" % len(gene), "-q", "0", "-D", "800", "-I", "13", "-l", "20", "-t", "0.02", "-d", "0.02",
            "-w", "5000", fa, sam]  # the option set scripts/rambl.py passes (rambl.py:181-187)
    code_r, out_r, err_r = run_cli(REF_CLI, args, str(tmp_path))
    code, out, err = run_cli(CLI, args, str(tmp_path))
    assert code == 0, err
    if code_r != 0:
        pytest.skip("the reference CLI crashed on this input (all strains pruned)")
    assert out == out_r
    assert out.startswith(">contiggene71")
    # -G: the graph text
    code_r, out_r, _ = run_cli(REF_CLI, args + ["-G"], str(tmp_path))
    code, out, err = run_cli(CLI, args + ["-G"], str(tmp_path))
    assert code == 0 and out == out_r
    assert os.listdir(str(tmp_path)).count("genes.fa") == 1 and len(os.listdir(str(tmp_path))) == 3  # temp files removed


def parse_dump(path):
    wins = []
    for line in open(path):
        f = line.split()
        if f[0] == "WINDOW":
            wins.append(dict(name=f[1], p0=int(f[2]), p1=int(f[3]), pos=[], cigar=[], seq=[], cn=[], mates=[]))
        elif f[0] == "GENE":
            wins[-1]["gene"] = f[1] if len(f) > 1 else ""
        elif f[0] == "READ":
            w = wins[-1]
            w["pos"].append(int(f[1])); w["cigar"].append(f[2]); w["seq"].append(f[3]); w["cn"].append(int(f[4]))
            w["mates"].append([int(x) for x in f[5:]])
    return wins


WINDOW_ARGS = ["-w", "120", "-o", "40", "-l", "30", "-q", "0"]


def window_fixture(tmp_path, indels):
    kw = dict(indel_err=0.01) if indels else dict(indel_frac=0.0)
    gene, raw, _ = synth.simulate_raw_reads(400, 60, 2, seed=31 if indels else 33, window=(100, 400), sub_err=0.003,
                                            divergence=(0.03, 0.06), **kw)
    return synth.write_cli_fixture(str(tmp_path), "w1", gene, raw)


def test_cli_io_glue_reproduces_reference_graphs(tmp_path):
    """CPU: the scan windows, cropped reads, filters, down-sampling and de-duplication of the drop-in CLI
    (--dump-inputs stops before the device is needed) give, window by window, the graphs the reference CLI
    prints with -G.  The graphs are built by the product's host phases with the alignment rows supplied by
    the test (no device here)."""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/StrainCall not built")
    from oracle import refpy
    fa, sam = window_fixture(tmp_path, indels=False)
    code, out_ref, _ = run_cli(REF_CLI, WINDOW_ARGS + [fa, sam, "-G"], str(tmp_path))
    assert code == 0
    dump = os.path.join(str(tmp_path), "dump.txt")
    code, _, err = run_cli(CLI, WINDOW_ARGS + [fa, sam, "--dump-inputs", dump], str(tmp_path))
    assert code == 0, err
    wins = parse_dump(dump)
    assert len(wins) >= 3
    text = ""
    for w in wins:
        if not w["pos"]:
            continue
        b = api.StrainCallBatch()
        b.add_subgroup(w["gene"], w["pos"], w["cigar"], w["seq"], w["cn"])
        b.thread_reads()
        b.finish_graphs_with_rows([refpy.msa_align(p, "oracle") for p in b.msa_problems()])
        text += b.output_edge(0)
    assert text == out_ref


def test_cli_inputs_equal_synth_subgroup(tmp_path):
    """CPU: with one window over the whole gene the CLI's reads are synth.make_subgroup's (same
    down-sampling stream, same AlignRead order, same ReadPairs)."""
    spec = dict(n_reads=3000, read_len=100, n_strains=3, seed=5, window=(0, 300), sub_err=0.004, paired=True,
                divergence=(0.02, 0.05))
    gene, raw, _ = synth.simulate_raw_reads(**spec)
    fa, sam = synth.write_cli_fixture(str(tmp_path), "gx", gene, raw)
    dump = os.path.join(str(tmp_path), "dump.txt")
    code, _, err = run_cli(CLI, ["-r", "gx:1-%d" % len(gene), "-w", "5000", "-q", "0", "-l", "20", fa, sam,
                                 "--dump-inputs", dump], str(tmp_path))
    assert code == 0, err
    (w,) = parse_dump(dump)
    sg = synth.make_subgroup(**spec)
    assert sg.n_reads < sg.n_raw_reads  # the depth cap was active
    assert (w["gene"], w["pos"], w["cigar"], w["seq"], w["cn"]) == (sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn)
    assert [m for ms in w["mates"] for m in ms] == [int(x) for x in sg.pair_val]


@pytest.mark.gpu
@pytest.mark.parametrize("indels", [False, True])
def test_cli_windows_match_reference_cli(tmp_path, indels):
    """Several overlapping scan windows (-w/-o) end to end.  With indel-bearing reads the reference's own
    cropping can leave a read that starts with an insertion at the window start; its graph construction
    then erases set::end() (PartialOrderGraph.cpp:795-796) and may never return -- the comparison is
    skipped when the reference does not finish, the drop-in must still finish."""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/StrainCall not built")
    fa, sam = window_fixture(tmp_path, indels)
    code, out, err = run_cli(CLI, WINDOW_ARGS + [fa, sam], str(tmp_path))
    assert code == 0, err
    assert out.count(">contigw1") >= 3
    try:
        env = dict(os.environ)
        env["PATH"] = SHIM + os.pathsep + env.get("PATH", "")
        r = subprocess.run([REF_CLI] + WINDOW_ARGS + [fa, sam], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, timeout=120)
    except subprocess.TimeoutExpired:
        pytest.skip("the reference CLI did not finish on this input")
    if r.returncode == 0:
        assert out == r.stdout


@pytest.mark.gpu
def test_cli_equals_library_path(tmp_path):
    spec = dict(n_reads=300, read_len=100, n_strains=3, seed=5, window=(0, 400), sub_err=0.004, divergence=(0.02, 0.05))
    gene, raw, _ = synth.simulate_raw_reads(**spec)
    fa, sam = synth.write_cli_fixture(str(tmp_path), "gx", gene, raw)
    code, out, err = run_cli(CLI, ["-r", "gx:1-%d" % len(gene), "-w", "5000", "-q", "0", fa, sam], str(tmp_path))
    assert code == 0, err
    sg = synth.make_subgroup(**spec)
    b = api.StrainCallBatch()
    b.add(sg)
    b.build_graphs()
    b.infer()
    assert out == b.fasta(0, "gx", 1, len(gene), 0.02)


def test_config3_reads_through_the_cli_glue_equal_the_golden_input(tmp_path):
    """CPU: BASELINE configs[3] at a tenth of its raw reads (100 000 x 150 bp of 50 strains, ~9 700x deep) through the
    drop-in CLI's window / filter / -D 800 down-sampling / de-duplication glue gives exactly the subgroup the unmodified
    reference was run on for tests/golden/full_config3_100k.json.gz (the GPU suite then checks the strains).  The depth
    cap makes the subgroup that reaches the graph independent of the raw depth, so this is the 1M-read case in all but
    the time the samtools stand-in needs to read the SAM text (tools/config3_glue.py runs the full million)."""
    from helpers import load_golden_gz
    path = os.path.join(ROOT, "tests", "golden", "full_config3_100k.json.gz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/full_config3_100k.json.gz not generated yet")
    case = load_golden_gz("full_config3_100k.json.gz")
    spec = dict(case["spec"])
    spec["divergence"] = tuple(spec["divergence"])
    gene, raw, _ = synth.simulate_raw_reads(**spec)
    fa, sam = synth.write_cli_fixture(str(tmp_path), "deep", gene, raw)
    dump = os.path.join(str(tmp_path), "dump.txt")
    code, _, err = run_cli(CLI, ["-r", "deep:1-%d" % len(gene), "-w", "5000", "-q", "0", "-D", "800", "-I", "13", "-l", "20",
                                 fa, sam, "--dump-inputs", dump], str(tmp_path))
    assert code == 0, err
    (w,) = parse_dump(dump)
    inp = case["input"]
    assert sum(w["cn"]) < 0.1 * len(raw)  # the depth cap was active
    assert (w["gene"], w["pos"], w["cigar"], w["seq"], w["cn"]) == (inp["gene"], inp["pos"], inp["cigar"], inp["seq"], inp["cn"])
    assert [m for ms in w["mates"] for m in ms] == inp["pair_val"]


def _two_gene_fixture(tmp_path):
    fa = os.path.join(str(tmp_path), "genes.fa")
    sam = os.path.join(str(tmp_path), "reads.sam")
    rois = []
    with open(fa, "w") as f, open(fa + ".fai", "w") as fi, open(sam, "w") as fs:
        off = 0
        for k, (seed, win) in enumerate([(41, (100, 300)), (42, (700, 880))]):
            gene, raw, _ = synth.simulate_raw_reads(220, 60, 2 + k, seed=seed, window=win, sub_err=0.003, divergence=(0.03, 0.06))
            name = "gene%d" % k
            rois.append("%s:1-%d" % (name, len(gene)))
            f.write(">%s\n" % name)
            off += len(name) + 2
            fi.write("%s\t%d\t%d\t60\t61\n" % (name, len(gene), off))
            for i in range(0, len(gene), 60):
                f.write(gene[i:i + 60] + "\n")
            off += len(gene) + (len(gene) + 59) // 60
            for (nm, p, cg, sq) in sorted(raw, key=lambda r: r[1]):
                fs.write("%s_%d\t0\t%s\t%d\t30\t%s\t*\t0\t0\t%s\t%s\n" % (nm, k, name, p + 1, cg, sq, "I" * len(sq)))
    return fa, sam, rois


@pytest.mark.gpu
def test_cli_several_regions_in_one_call(tmp_path):
    """`-r` may repeat and `--roi-file` lists regions: every window of every region goes into ONE batch, and stdout is
    what the single-region runs (the way scripts/rambl.py calls StrainCall, rambl.py:179-190) print one after the other."""
    fa, sam, rois = _two_gene_fixture(tmp_path)
    common = ["-q", "0", "-l", "20", "-w", "5000", fa, sam]
    singles = ""
    for r in rois:
        code, out, err = run_cli(CLI, ["-r", r] + common, str(tmp_path))
        assert code == 0, err
        singles += out
    code, out, err = run_cli(CLI, ["-r", rois[0], "-r", rois[1]] + common, str(tmp_path))
    assert code == 0, err
    assert out == singles and out.count(">contiggene0") >= 1 and out.count(">contiggene1") >= 1
    roi_file = os.path.join(str(tmp_path), "rois.txt")
    with open(roi_file, "w") as f:
        f.write("# seed genes of the sample\n" + "\n".join(rois) + "\n")
    code, out2, err = run_cli(CLI, ["--roi-file", roi_file, "--device", "0"] + common, str(tmp_path))
    assert code == 0, err
    assert out2 == singles


def test_cli_several_regions_dump_inputs(tmp_path):
    """CPU: the same, up to where the device takes over (--dump-inputs): the windows of all regions, in order."""
    fa, sam, rois = _two_gene_fixture(tmp_path)
    dump = os.path.join(str(tmp_path), "dump.txt")
    code, _, err = run_cli(CLI, ["-r", rois[0], "-r", rois[1], "-q", "0", "-l", "20", "-w", "5000", fa, sam, "--dump-inputs", dump],
                           str(tmp_path))
    assert code == 0, err
    wins = parse_dump(dump)
    assert [w["name"] for w in wins] == ["gene0", "gene1"] and all(len(w["pos"]) > 50 for w in wins)
