#!/usr/bin/env python
"""What scripts/rambl.py's process pool (rambl.py:179-190: one StrainCall process per seed gene, `--cores` at a time)
costs on one GPU against ONE StrainCall call with every region (-r ... -r ... / --roi-file): G whole-gene subgroups of
BASELINE configs[2] in one FASTA + one SAM, (a) G concurrent single-region processes, (b) one process, G regions.
Both read through tests/samtools_shim (a python stand-in: its time is reported separately as the --dump-inputs run).
usage: cli_batch_vs_procs.py [G=16]"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from rambl_b200 import synth
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    work = tempfile.mkdtemp(prefix="cli_batch_")
    fa = os.path.join(work, "genes.fa")
    sam = os.path.join(work, "reads.sam")
    names = []
    with open(fa, "w") as f, open(fa + ".fai", "w") as fi, open(sam, "w") as fs:
        off = 0
        for k in range(G):
            gene, raw, _ = synth.simulate_raw_reads(5000, 150, 2 + (k % 5), seed=k)
            name = "gene%02d" % k
            names.append((name, len(gene)))
            f.write(">%s\n" % name)
            off += len(name) + 2
            fi.write("%s\t%d\t%d\t60\t61\n" % (name, len(gene), off))
            for i in range(0, len(gene), 60):
                f.write(gene[i:i + 60] + "\n")
            off += len(gene) + (len(gene) + 59) // 60
            for (nm, p, cg, sq) in sorted(raw, key=lambda r: r[1]):
                fs.write("%s_%d\t0\t%s\t%d\t30\t%s\t*\t0\t0\t%s\t%s\n" % (nm, k, name, p + 1, cg, sq, "I" * len(sq)))
    env = dict(os.environ)
    env["PATH"] = os.path.join(ROOT, "tests", "samtools_shim") + os.pathsep + env.get("PATH", "")
    cli = os.path.join(ROOT, "rambl_b200", "StrainCall")
    common = ["-q", "0", "-D", "800", "-I", "13", "-l", "20", "-t", "0.02", "-d", "0.02", "-w", "5000", fa, sam]
    rois = ["%s:1-%d" % (n, L) for n, L in names]

    def run_many(extra):
        t = time.time()
        procs = [subprocess.Popen([cli, "-r", r] + extra + common, cwd=work, env=env, stdout=subprocess.PIPE,
                                  stderr=subprocess.PIPE, text=True) for r in rois]
        outs = [p.communicate() for p in procs]
        assert all(p.returncode == 0 for p in procs), [o[1][-300:] for o in outs]
        return time.time() - t, "".join(o[0] for o in outs)

    def run_one(extra):
        t = time.time()
        args = []
        for r in rois:
            args += ["-r", r]
        r = subprocess.run([cli] + args + extra + common, cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           text=True)
        assert r.returncode == 0, r.stderr[-500:]
        return time.time() - t, r.stdout

    io_many, _ = run_many(["--dump-inputs", os.path.join(work, "dump_many.txt")])
    io_one, _ = run_one(["--dump-inputs", os.path.join(work, "dump_one.txt")])
    run_one([])  # warm the driver / caches once
    t_many, out_many = run_many([])
    t_one, out_one = run_one([])
    print("G = %d whole-gene subgroups of configs[2] (5000 x 150bp reads each)" % G)
    print("  %d concurrent single-region processes (rambl.py's pool): %.2f s wall (of which reading the inputs: %.2f s)" % (
        G, t_many, io_many))
    print("  one process, %d regions in one batch:                    %.2f s wall (of which reading the inputs: %.2f s)" % (
        G, t_one, io_one))
    print("  outputs identical: %s (%d FASTA bytes)" % (out_many == out_one, len(out_one)))


if __name__ == "__main__":
    main()
