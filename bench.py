#!/usr/bin/env python
"""bench.py -- StrainCall hot path on B200: reads/s through the C ABI, with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU code (oracle/_ref)

A "step" is one pass of the hot path (PartialOrderGraph construction with the device insertion
alignment, level-synchronous strain inference, read assignment) over one batch of synthetic input.
Workload at every N: each rank gets BASELINE.json configs[1] ("single subgroup, 20k 150bp reads,
10 strains at 1-3% divergence"), seeded by rank -- subgroups are independent, so ranks share nothing
on the data path ("scaling": "weak"); rank 0 gathers the FASTA records at the end of a step.
`value`  = reads entering the hot path per second with the graphs already built and flattened (the
           timed region is rambl_batch_infer, CUDA events, max over ranks);
`e2e`    = the same metric through the whole C-ABI call sequence from HOST buffers
           (add_subgroup -> build_graphs -> infer -> fasta), copies inside the timed region.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "straincall_reads_per_sec"
UNIT = "reads/s"
WORKLOAD = "configs[1]: single subgroup, 20k 150bp reads, 10 strains at 1-3% divergence, whole 16S gene window"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--reads", type=int, default=20000, help="raw reads per subgroup (configs[1]: 20000)")
    p.add_argument("--subgroups", type=int, default=1, help="subgroups per rank and step (configs[1]: 1)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-sample-window", type=int, default=160)
    p.add_argument("--ref-sample-window", type=int, default=200)
    return p.parse_args()


def make_workload(rank: int, reads: int, subgroups: int):
    from rambl_b200 import synth
    return [synth.make_subgroup(reads, 150, 10, divergence=(0.01, 0.03), seed=1000 * rank + k)
            for k in range(subgroups)]


def make_cpu_sample(window: int, seed: int = 0):
    """A bounded sample of the same workload: same depth, read length, strain count and divergence,
    on a `window`-bp slice of the gene (the reference needs minutes for the whole gene)."""
    from rambl_b200 import synth
    n = int(20000 * window / 1542)
    return synth.make_subgroup(n, min(150, window), 10, divergence=(0.01, 0.03), seed=seed, window=(600, 600 + window))


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def cpu_reference_once(sg):
    """The reference's own code (oracle/_ref, -O2 where its UB allows) on one subgroup; returns seconds."""
    from oracle import refpy
    variant = "" if refpy.available("") else "oracle"
    t = time.time()
    g = refpy.RefPog(sg.gene, sg.pos, sg.cigar, sg.seq, sg.cn, variant=variant)
    g.infer(sg.pair_off, sg.pair_val, with_loglik=False)
    g.close()
    return time.time() - t, ("reference" if variant == "" else "port")


def _cpu_worker(args):
    window, seed = args
    sg = make_cpu_sample(window, seed)
    dt, kind = cpu_reference_once(sg)
    return sg.n_reads, dt, kind


def cpu_baseline_block(window: int, cores: int = 1):
    sg = make_cpu_sample(window, 0)
    dt, kind = cpu_reference_once(sg)
    return {"value": sg.n_reads / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d reads (depth-800 down-sampled, 150bp, 10 strains, 1-3%% divergence) on a %dbp window of the "
                      "16S gene, graph build + infer_strains + read_assign, %.1f s; %s" % (sg.n_reads, window, dt, FULL_SIZE_NOTE)}


# dram__bytes_read.sum + dram__bytes_write.sum of one Gibbs-kernel launch of this workload, from the ncu --set full
# capture summarised in profiles/r01_k_gibbs_w_full.txt (mean of the two captured launches: 838 KB and 680 KB
# read, 0 written: the weight tiles once; the sweeps re-read them from L2 / shared memory)
GIBBS_DRAM_BYTES_PER_LAUNCH = 758912

FULL_SIZE_NOTE = ("the reference on the FULL configs[1] subgroup (8223 reads after down-sampling), measured once on one "
                  "core of the build container with oracle/_ref -O2: 805 s = 10.2 reads/s; windowed samples run faster "
                  "per read because fewer candidate strains accumulate")


def run_reference_arm(args, rank, world):
    """--impl reference: the reference CPU StrainCall path on all host cores (one independent sample
    subgroup per core and step, the way scripts/rambl.py spreads subgroups over a process pool).
    Timed steps use 200bp-window samples of configs[1] (about 50 s of CPU per core and step); the untimed
    warm-up steps use 60bp windows -- CPU code has nothing to warm up and the run must end in minutes."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, min(os.cpu_count() or 1, 64))
    # bounded: the largest sample window whose K timed steps fit in about 2.5 minutes (seconds per step per core,
    # measured with oracle/_ref -O2: 200bp ~50 s, 160bp ~16 s, 100bp ~11 s, 60bp ~5 s)
    window = 60
    for w, est in ((args.ref_sample_window, 50.0), (160, 16.0), (100, 11.0)):
        if w <= args.ref_sample_window and args.steps * est <= 150.0:
            window = w
            break
    ctx = mp.get_context("spawn")
    times = []
    reads = 0
    kind = "reference"
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            w = window if step >= args.warmup else 60
            t = time.time()
            res = pool.map(_cpu_worker, [(w, 100 * step + c) for c in range(cores)])
            dt = time.time() - t
            if step >= args.warmup:
                times.append(dt)
                reads += sum(r[0] for r in res)
            kind = res[0][2]
    total = sum(times)
    value = reads / total
    sample = ("%d independent %dbp-window samples of configs[1] per step, one per core (same depth, read length, strain "
              "count and divergence); %s" % (cores, window, FULL_SIZE_NOTE))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(1, args.steps), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f80", "data": "synthetic",
           "config": {"workload": WORKLOAD, "sample": sample},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def poa_block():
    """The other half of BASELINE.json's metric ("POA GCUPS"): the insertion-alignment kernel on a fixed batch of
    2000 level-problems as deep homopolymer levels produce them (20-400 insertion strings of 1-9 letters each).
    Outside the timed region; cells = profile columns x letters per alignment step, CUDA-event kernel time."""
    import numpy as np
    from rambl_b200 import api
    rnd = np.random.default_rng(1)
    probs = []
    for _ in range(2000):
        n = int(rnd.integers(20, 400))
        hp = "ACGT"[int(rnd.integers(4))]
        seqs = []
        for _ in range(n):
            k = int(rnd.integers(1, 10))
            s = [hp] * k
            if rnd.random() < 0.1:
                s[int(rnd.integers(k))] = "ACGT"[int(rnd.integers(4))]
            seqs.append("".join(s))
        probs.append(sorted(seqs, key=lambda x: -len(x)))
    api.msa_align_batch(probs)
    best = None
    for _ in range(3):
        _, st = api.msa_align_batch(probs)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    return {"gcups": best["dp_cells"] / best["kernel_ms"] / 1e6, "dp_cells": best["dp_cells"], "kernel_ms": best["kernel_ms"],
            "alignment_steps_per_s": sum(len(p) - 1 for p in probs) / (best["kernel_ms"] / 1e3),
            "workload": "2000 insertion-alignment problems, 423k homopolymer insertion strings of 1-9 letters"}


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return 0

    # one process per GPU: bind before CUDA comes up so that both torch and librambl_b200 see device 0
    if world > 1 and "CUDA_VISIBLE_DEVICES" not in os.environ:
        os.environ["CUDA_VISIBLE_DEVICES"] = str(local_rank)
    elif world > 1:
        vis = os.environ["CUDA_VISIBLE_DEVICES"].split(",")
        if len(vis) > local_rank:
            os.environ["CUDA_VISIBLE_DEVICES"] = vis[local_rank]
    import torch
    import torch.distributed as dist
    from rambl_b200 import api

    if not torch.cuda.is_available() or api.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(0)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sgs = make_workload(rank, args.reads, args.subgroups)
    reads_per_step = sum(s.n_reads for s in sgs)
    raw_per_step = sum(s.n_raw_reads for s in sgs)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def one_step():
        """Returns (e2e_ms, infer_ms, stats, fasta bytes) for one pass from host buffers."""
        flush.zero_()  # L2 flush: 256 MiB write, larger than the 126 MB L2
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b = api.StrainCallBatch()
        for sg in sgs:
            b.add(sg)
        b.build_graphs()
        b.infer()
        out = 0
        for i in range(len(sgs)):
            if b.status(i) == 0:
                out += len(b.fasta(i, "g", 1, len(sgs[i].gene), 0.02))
        e1.record()
        torch.cuda.synchronize()
        st = b.stats()
        b.close()
        return e0.elapsed_time(e1), st["infer_gpu_ms"], st, out

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(int(os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]) if world > 1 else 0)
    if rank == 0:
        sampler.start()
    t_begin = time.time()
    e2e_ms = infer_ms = 0.0
    launches = 0
    agg = {}
    out_bytes = 0
    for _ in range(args.steps):
        a, b_, st, ob = one_step()
        e2e_ms += a
        infer_ms += b_
        launches += st["gpu_launches"]
        out_bytes += ob
        for k, v in st.items():
            agg[k] = agg.get(k, 0) + v
    barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None

    t = torch.tensor([e2e_ms, infer_ms], dtype=torch.float64, device="cuda")
    r = torch.tensor([float(reads_per_step)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    e2e_max, infer_max = float(t[0]), float(t[1])
    total_reads_per_step = float(r[0])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    K = args.steps
    value = total_reads_per_step * K / (infer_max / 1000.0)
    e2e_value = total_reads_per_step * K / (e2e_max / 1000.0)
    peak, peak_src = measured_peak_gbs()
    gibbs_s = agg.get("gibbs_kernel_ms", 0.0) / 1000.0
    achieved = (agg.get("gibbs_alg_bytes", 0) / 1e9) / gibbs_s if gibbs_s > 0 else 0.0
    roofline = {"kernel": "k_gibbs_w (speculative block Gibbs sweeps, one warp per 32-draw block, up to 8 blocks per round)", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": GIBBS_DRAM_BYTES_PER_LAUNCH,
                "algorithmic_bytes_per_launch": (agg.get("gibbs_alg_bytes", 0) / agg["gibbs_launches"]) if agg.get("gibbs_launches") else None,
                "avg_launch_ms": (agg.get("gibbs_kernel_ms", 0.0) / agg["gibbs_launches"]) if agg.get("gibbs_launches") else None,
                "peak_source": peak_src,
                "share_of_infer_time": (agg.get("gibbs_kernel_ms", 0.0) / infer_ms) if infer_ms else None,
                "draws_per_s": agg.get("draws", 0) / gibbs_s if gibbs_s > 0 else None,
                "passes_per_round": (agg.get("gibbs_passes", 0) / agg["gibbs_rounds"]) if agg.get("gibbs_rounds") else None,
                "note": "a sequential Gibbs chain per subgroup: latency-bound by construction, its weights stay in L2/shared "
                        "memory; the algorithmic bytes are sweeps x draws x (S weights + 1 uniform) x 8"}
    poa = poa_block()
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline_block(args.cpu_sample_window)
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
           "ms_per_step": infer_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD, "subgroups_per_rank": len(sgs), "raw_reads_per_rank": raw_per_step,
                      "reads_after_depth800_downsampling_per_rank": reads_per_step, "sharding": "one subgroup set per rank, no collective",
                      "l2": "256 MiB device write between steps", "n": 5000, "e": 0.01, "tau": 0.02, "diff": 0.01},
           "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_max / K,
                   "h2d_bytes_per_step": int(agg.get("h2d_bytes", 0) / K), "d2h_bytes_per_step": int(agg.get("d2h_bytes", 0) / K),
                   "result_bytes_per_step": int(out_bytes / K)},
           "gpu_launches": int(launches),
           "roofline": roofline,
           "poa": dict(poa, msa_problems_in_step=agg.get("msa_problems", 0) / K,
                       msa_dp_cells_in_step=agg.get("msa_dp_cells", 0) / K),
           "level_steps_per_step": agg.get("level_steps", 0) / K, "draws_per_step": agg.get("draws", 0) / K}
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
