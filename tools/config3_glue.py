#!/usr/bin/env python
"""BASELINE configs[3] ("one deep subgroup, 1M 150bp reads, 50 strains") through the host glue of the drop-in CLI:
raw mapped reads -> window, filters, depth down-sampling (-D 800, std::mt19937(1234)), AlignRead de-duplication,
ReadPairs (StrainCall.cpp:480-670) -- `StrainCall --dump-inputs` stops where the device would take over.  Checks the
dump against rambl_b200.synth.make_subgroup (the in-memory restatement the golden fixtures were generated from) and
times the glue.  CPU only.   usage: config3_glue.py [n_reads] [workdir]"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from rambl_b200 import synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    work = sys.argv[2] if len(sys.argv) > 2 else tempfile.mkdtemp(prefix="config3_")
    spec = dict(n_reads=n, read_len=150, n_strains=50, divergence=(0.01, 0.03), seed=0)
    t = time.time()
    gene, raw, _ = synth.simulate_raw_reads(**spec)
    t_sim = time.time() - t
    fa, sam = synth.write_cli_fixture(work, "deep", gene, raw)
    env = dict(os.environ)
    env["PATH"] = os.path.join(ROOT, "tests", "samtools_shim") + os.pathsep + env.get("PATH", "")
    dump = os.path.join(work, "dump.txt")
    t = time.time()
    r = subprocess.run([os.path.join(ROOT, "rambl_b200", "StrainCall"), "-r", "deep:1-%d" % len(gene), "-w", "5000", "-q", "0",
                        "-D", "800", "-I", "13", "-l", "20", fa, sam, "--dump-inputs", dump], cwd=work, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    t_cli = time.time() - t
    assert r.returncode == 0, r.stderr
    from test_cli import parse_dump
    (w,) = parse_dump(dump)
    t = time.time()
    sg = synth.make_subgroup(**spec)
    t_sub = time.time() - t
    same = (w["gene"], w["pos"], w["cigar"], w["seq"], w["cn"]) == (sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn)
    same = same and [m for ms in w["mates"] for m in ms] == [int(x) for x in sg.pair_val]
    print("configs[3] host glue: %d raw reads (depth %.0fx) -> %d reads in %d unique AlignReads after -D 800; identical to "
          "synth.make_subgroup: %s" % (len(raw), len(raw) * 150.0 / len(gene), sum(w["cn"]), len(w["pos"]), same))
    print("  simulate %.1f s, StrainCall --dump-inputs %.1f s wall (of which the python samtools stand-in reads the %d-line SAM "
          "text), make_subgroup %.1f s" % (t_sim, t_cli, len(raw), t_sub))
    assert same


if __name__ == "__main__":
    main()
