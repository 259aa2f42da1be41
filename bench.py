#!/usr/bin/env python
"""bench.py -- StrainCall hot path on B200: reads/s through the C ABI, with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU code (oracle/_ref)

Workload (every N): BASELINE.json configs[2], "500 synthetic taxonomic subgroups x 5k reads, sharded by subgroup
across 1/2/4/8 B200" -- the workload scripts/rambl.py produces (one StrainCall problem per seed gene,
rambl.py:165-194).  The 500 subgroups are FIXED; N ranks split them with rambl_b200.shard.assign (greedy LPT on a
cost proxy), so total work is constant: "scaling": "strong".  A "step" is one pass of the hot path over the
rank's share: add_subgroup -> build_graphs (host splice + device insertion alignment) -> infer (device strain
search + read assignment) -> FASTA text, then rank 0 gathers every rank's FASTA records; no collective on the
data path.
`value`  = reads of ALL 500 subgroups per second of the device-resident phase (rambl_batch_infer: graphs already
           flattened and uploaded, CUDA events, max over ranks);
`e2e`    = the same through the whole C-ABI call sequence from HOST buffers to the gathered FASTA on rank 0
           (copies and the gather inside the timed region), max over ranks.
Extra blocks on the same line (rank 0, outside the timed region):
`config1`  configs[1] (single subgroup, 20k 150bp reads, 10 strains): one chain on one GPU -- the latency case;
`matched`  the bounded sample set the reference arm times (same subgroups, same windows), solved as one batch here,
           so that `matched.e2e_reads_per_s` / the reference arm's value is a like-for-like ratio;
`poa`      the insertion-alignment kernel (GCUPS) on the alignment problems of a configs[4] subgroup;
`cpu_baseline`  the reference on one core on the first subgroups of the matched sample set.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import pickle
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "straincall_reads_per_sec"
UNIT = "reads/s"
N_SUBGROUPS = 500
WORKLOAD = ("configs[2]: 500 synthetic taxonomic subgroups x 5k 150bp reads (2-6 strains each, whole 16S gene), "
            "sharded by subgroup across the ranks")
MATCHED_SUBGROUPS = 16


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--subgroups", type=int, default=N_SUBGROUPS, help="subgroups of the whole job (configs[2]: 500)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extra", action="store_true", help="skip the config1 / matched / poa blocks")
    p.add_argument("--cache", default=os.environ.get("RAMBL_BENCH_CACHE", "/tmp/rambl_b200_bench_cache"))
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
# workloads
def matched_window(steps: int) -> int:
    """Width (bp) of the gene window of the matched sample set: the widest whose K timed steps keep the reference
    arm within ~3 minutes when every sample has a core (seconds per sample on one core of the build container with
    oracle/_ref -O2 -- see DESIGN.md section 6; the whole gene needs minutes per subgroup)."""
    per_step = 170.0 / max(1, steps)
    for w, est in MATCHED_WINDOW_SECONDS:
        if est <= per_step:
            return w
    return MATCHED_WINDOW_SECONDS[-1][0]


# (window bp, seconds of the slowest of the 16 samples -- the six-strain ones -- on one core of the build container)
MATCHED_WINDOW_SECONDS = [(300, 55.0), (200, 22.0), (150, 10.0), (100, 4.0), (60, 2.0)]


def matched_sample_set(steps: int):
    from rambl_b200 import synth
    w = matched_window(steps)
    return w, [synth.config2_subgroup(k, (600, 600 + w)) for k in range(MATCHED_SUBGROUPS)]


def matched_description(w: int) -> str:
    return ("%d subgroups of configs[2] (k=0..%d) restricted to the %dbp window [600,%d) of the 16S gene at the same depth; "
            "the whole-gene subgroups take the reference minutes each on one core (tests/golden/full_config2_*.json.gz "
            "record 1-core times)" % (MATCHED_SUBGROUPS, MATCHED_SUBGROUPS - 1, w, 600 + w))


def _gen_one(k):
    from rambl_b200 import synth
    return k, synth.config2_subgroup(k)


def load_subgroups(indices, cache_dir, procs):
    """configs[2] subgroups by index; generated on `procs` host processes and cached on local disk (the driver runs
    N=1,2,4,8 back to back on one box; generation is ~1.5 core-seconds per subgroup of pure Python)."""
    from rambl_b200 import synth
    with open(synth.__file__, "rb") as f:
        tag = hashlib.sha256(f.read()).hexdigest()[:12]
    d = os.path.join(cache_dir, tag)
    out, missing = {}, []
    for k in indices:
        p = os.path.join(d, "c2_%d.pkl" % k)
        try:
            with open(p, "rb") as f:
                out[k] = pickle.load(f)
        except Exception:
            missing.append(k)
    if missing:
        import multiprocessing as mp
        try:
            os.makedirs(d, exist_ok=True)
        except Exception:
            d = None
        if procs > 1 and len(missing) > 1:
            with mp.get_context("spawn").Pool(min(procs, len(missing))) as pool:
                res = pool.map(_gen_one, missing, chunksize=1)
        else:
            res = [_gen_one(k) for k in missing]
        for k, sg in res:
            out[k] = sg
            if d:
                try:
                    tmp = os.path.join(d, "c2_%d.pkl.%d" % (k, os.getpid()))
                    with open(tmp, "wb") as f:
                        pickle.dump(sg, f, protocol=4)
                    os.replace(tmp, os.path.join(d, "c2_%d.pkl" % k))
                except Exception:
                    pass
    return [out[k] for k in indices]


def config2_cost(k: int) -> float:
    """Cost proxy for the LPT split, computable without generating the subgroup: the strain count drives the
    candidate set (2 + k % 5 strains), every level draws <= 40000 times per candidate."""
    return 1.0 + 0.6 * (k % 5)


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


def ncu_traffic(kernel_key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed `ncu --set
    full` capture of this command (profiles/traffic.json, written by tools/ncu_summary.py); null when there is none."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# the reference on the host cores
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
    k, w = args
    from rambl_b200 import synth
    sg = synth.config2_subgroup(k, (600, 600 + w))
    dt, kind = cpu_reference_once(sg)
    return sg.n_reads, dt, kind


def common_config(args, w):
    """The `config` object both arms print (identical, so that the two lines describe one configuration)."""
    return {"workload": WORKLOAD, "subgroups": args.subgroups, "raw_reads_per_subgroup": 5000, "read_length": 150,
            "n": 5000, "e": 0.01, "tau": 0.02, "diff": 0.01,
            "reference_arm": "bounded: the reference needs minutes per whole-gene subgroup on one core, so it is timed on "
                             "the matched sample set; the CUDA arm times the full workload AND the same sample set "
                             "(`matched`)",
            "matched_sample": matched_description(w)}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference CPU StrainCall path (PartialOrderGraph(G,R) + infer_strains + read_assign,
    StrainCall.cpp:1014-1053 minus the samtools I/O) on all host cores, one subgroup per process the way
    scripts/rambl.py spreads seed genes over a Pool (rambl.py:168-194).  Every step solves the matched sample set."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, min(os.cpu_count() or 1, MATCHED_SUBGROUPS))
    w = matched_window(args.steps)
    ctx = mp.get_context("spawn")
    times, reads, kind = [], 0, "reference"
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            # CPU code has nothing to warm up: the untimed steps run the two cheapest samples only
            ks = list(range(MATCHED_SUBGROUPS)) if step >= args.warmup else [0, 1]
            t = time.time()
            res = pool.map(_cpu_worker, [(k, w) for k in ks], chunksize=1)
            dt = time.time() - t
            if step >= args.warmup:
                times.append(dt)
                reads += sum(r[0] for r in res)
            kind = res[0][2]
    total = sum(times)
    value = reads / total
    sample = "every step: " + matched_description(w) + "; %d worker processes" % cores
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(1, args.steps), "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f80", "data": "synthetic",
           "config": common_config(args, w),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "reads_per_step": reads // max(1, args.steps), "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def cpu_baseline_block(w: int):
    """One core, the first samples of the matched set, until ~15 s of CPU work are spent."""
    from rambl_b200 import synth
    reads, secs, n, kind = 0, 0.0, 0, "reference"
    for k in range(MATCHED_SUBGROUPS):
        sg = synth.config2_subgroup(k, (600, 600 + w))
        dt, kind = cpu_reference_once(sg)
        reads += sg.n_reads
        secs += dt
        n += 1
        if secs > 15.0:
            break
    return {"value": reads / secs, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "the first %d subgroups of the matched sample set (%s): %d reads, graph build + infer_strains + "
                      "read_assign, %.1f s on one core" % (n, matched_description(w), reads, secs)}


# ------------------------------------------------------------------------------------------------
def solve_batch(api, sgs, names=None):
    """One pass from host buffers through the public calls (add_subgroup_packed, then rambl_batch_solve = graph
    construction + strain search, overlapped chunk by chunk): returns (batch, [FASTA text per subgroup])."""
    b = api.StrainCallBatch()
    for sg in sgs:
        b.add(sg)
    b.solve()
    fasta = []
    for i, sg in enumerate(sgs):
        fasta.append(b.fasta(i, names[i] if names else "g", 1, len(sg.gene), 0.02) if b.status(i) == 0 else "")
    return b, fasta


def resident_pass(b):
    """`value`: the strain search alone on a batch whose graphs are built (rambl_batch_infer, CUDA events inside the
    library) -- the statistics of THIS pass only."""
    s0 = b.stats()
    b.infer()
    s1 = b.stats()
    return {k: s1[k] - s0[k] for k in s1}


def poa_block(api):
    """The other half of BASELINE.json's metric ("POA GCUPS"): the insertion-alignment kernel on the alignment problems
    of configs[4]-derived subgroups (250bp reads with homopolymer indel errors: every graph level whose insertions
    differ in length is one problem), replicated to fill the machine.  Outside the timed region; cells = profile
    columns x letters per alignment step, CUDA-event kernel time, best of 3."""
    from rambl_b200 import synth
    probs = []
    for seed in range(4):
        sg = synth.make_subgroup(5000, 250, 4, indel_err=0.004, indel_frac=0.4, homopolymer_bias=True, seed=seed)
        b = api.StrainCallBatch()
        b.add(sg)
        b.thread_reads()
        probs.extend(b.msa_problems())
        b.close()
    n_distinct = len(probs)
    while len(probs) < 60000 and n_distinct:  # ~ the alignment problems of 500 such subgroups in one batch
        probs.extend(probs[:n_distinct])
    if not probs:
        return None
    api.msa_align_batch(probs)
    best = None
    for _ in range(3):
        _, st = api.msa_align_batch(probs)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    return {"gcups": best["dp_cells"] / best["kernel_ms"] / 1e6, "dp_cells": best["dp_cells"], "kernel_ms": best["kernel_ms"],
            "alignment_steps_per_s": sum(len(p) - 1 for p in probs) / (best["kernel_ms"] / 1e3),
            "problems": len(probs), "distinct_problems": n_distinct,
            "workload": "insertion-alignment problems of 4 configs[4] subgroups (5000 x 250bp reads, homopolymer indel errors), "
                        "replicated to %d problems" % len(probs)}


def timed_passes(torch, api, sgs, steps, warmup, flush):
    """W untimed + K timed passes over one fixed batch on this process; returns per-pass sums."""
    agg, e2e_ms, infer_ms = {}, 0.0, 0.0
    for it in range(warmup + steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b, fasta = solve_batch(api, sgs)
        e1.record()
        torch.cuda.synchronize()
        st = resident_pass(b)
        b.close()
        if it >= warmup:
            e2e_ms += e0.elapsed_time(e1)
            infer_ms += st["infer_gpu_ms"]
            for k, v in st.items():
                agg[k] = agg.get(k, 0) + v
    return e2e_ms, infer_ms, agg


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
    # the ranks share the host: each takes its share of the cores for its worker threads (graph construction)
    cores = os.cpu_count() or 1
    my_cores = max(1, cores // world)
    os.environ.setdefault("RAMBL_HOST_THREADS", str(my_cores))
    import torch
    import torch.distributed as dist
    from rambl_b200 import api, shard

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

    # ---- the fixed job, split over the ranks
    parts = shard.assign([config2_cost(k) for k in range(args.subgroups)], world)
    mine = parts[rank]
    sgs = load_subgroups(mine, args.cache, my_cores)
    names = ["sg%03d" % k for k in mine]
    my_reads = sum(s.n_reads for s in sgs)
    my_raw = sum(s.n_raw_reads for s in sgs)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def one_step():
        """One pass from host buffers to the gathered FASTA; returns (e2e ms, infer ms, stats, FASTA bytes on rank 0)."""
        flush.zero_()  # L2 flush: 256 MiB write, larger than the 126 MB L2
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b, fasta = solve_batch(api, sgs, names)
        h2d, d2h = b.stats()["h2d_bytes"], b.stats()["d2h_bytes"]
        part = list(zip(mine, fasta))
        got = 0
        if world > 1:  # rank 0 gathers the FASTA records of every rank (the host-side gather of the north star)
            parts_all = [None] * world if rank == 0 else None
            dist.gather_object(part, parts_all, dst=0)
            if rank == 0:
                ordered = [None] * args.subgroups
                for p in parts_all:
                    for i, txt in p:
                        ordered[i] = txt
                got = sum(len(x) for x in ordered)
        else:
            got = sum(len(x) for x in fasta)
        e1.record()
        torch.cuda.synchronize()
        ok = sum(1 for i in range(len(sgs)) if b.status(i) == 0)
        msa = {k: b.stats()[k] for k in ("msa_problems", "msa_dp_cells")}
        st = resident_pass(b)  # the device-resident phase on its own: `value`
        st.update(msa)
        st["h2d_bytes"], st["d2h_bytes"] = h2d, d2h  # of the end-to-end pass
        b.close()
        return e0.elapsed_time(e1), st["infer_gpu_ms"], st, got, ok

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(int(os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]) if world > 1 else 0)
    if rank == 0:
        sampler.start()
    t_begin = time.time()
    e2e_ms = infer_ms = 0.0
    agg = {}
    out_bytes = 0
    n_ok = 0
    for _ in range(args.steps):
        a, b_, st, ob, ok = one_step()
        e2e_ms += a
        infer_ms += b_
        out_bytes += ob
        n_ok = ok
        for k, v in st.items():
            agg[k] = agg.get(k, 0) + v
    barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None

    tmax = torch.tensor([e2e_ms, infer_ms], dtype=torch.float64, device="cuda")
    keys = ["gpu_launches", "h2d_bytes", "d2h_bytes", "draws", "level_steps", "msa_problems", "msa_dp_cells",
            "dpm_kernel_ms", "dpm_alg_bytes", "dpm_launches", "gibbs_kernel_ms", "gibbs_alg_bytes", "gibbs_launches",
            "gibbs_rounds", "gibbs_passes", "loglik_updates"]
    tsum = torch.tensor([float(my_reads), float(my_raw), float(n_ok), float(infer_ms)] + [float(agg.get(k, 0)) for k in keys],
                        dtype=torch.float64, device="cuda")
    tmin = tmax.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    K = args.steps
    e2e_max, infer_max = float(tmax[0]), float(tmax[1])
    total_reads, total_raw, total_ok, infer_sum = float(tsum[0]), float(tsum[1]), int(tsum[2]), float(tsum[3])
    tot = {k: float(tsum[4 + i]) for i, k in enumerate(keys)}
    value = total_reads * K / (infer_max / 1000.0)
    e2e_value = total_reads * K / (e2e_max / 1000.0)
    peak, peak_src = measured_peak_gbs()

    # dominant kernel: the strain-search kernel(s) of rambl_batch_infer (device walk, or the per-level Gibbs kernel
    # of the level-synchronous path); algorithmic bytes are counted by the library
    if tot.get("dpm_kernel_ms", 0) > 0:
        kname, kms, kbytes, kl = "k_walk", tot["dpm_kernel_ms"], tot["dpm_alg_bytes"], tot["dpm_launches"]
    else:
        kname, kms, kbytes, kl = "k_gibbs_w", tot["gibbs_kernel_ms"], tot["gibbs_alg_bytes"], tot["gibbs_launches"]
    # kernel time is summed over ranks, and so are the bytes: achieved is the per-GPU rate
    achieved = (kbytes / 1e9) / (kms / 1000.0) if kms > 0 else 0.0
    roofline = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": ncu_traffic(kname),
                "algorithmic_bytes_per_launch": kbytes / kl if kl else None,
                "avg_launch_ms": kms / kl if kl else None, "peak_source": peak_src,
                "share_of_infer_time": kms / infer_sum if infer_sum else None,
                "draws_per_s_per_gpu": tot["draws"] / (kms / 1000.0) if kms > 0 else None,
                "passes_per_round": tot["gibbs_passes"] / tot["gibbs_rounds"] if tot.get("gibbs_rounds") else None,
                "note": "per-GPU rate of the strain-search kernel over all its launches (CUDA events); algorithmic bytes = "
                        "sweeps x draws x (S weights + 1 uniform) x 8 for the Gibbs chains + 16 B per log-likelihood update + "
                        "16 B per (draw, strain) weight; the chains re-read their weights from shared memory / L2, so DRAM "
                        "traffic is far below the algorithmic bytes"}
    w = matched_window(K)
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
           "ms_per_step": infer_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": common_config(args, w),
           "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_max / K,
                   "h2d_bytes_per_step": int(tot["h2d_bytes"] / K), "d2h_bytes_per_step": int(tot["d2h_bytes"] / K),
                   "result_bytes_per_step": int(out_bytes / K),
                   "includes": "host splice + device insertion alignment + strain search + FASTA + gather on rank 0"},
           "gpu_launches": int(tot["gpu_launches"]),
           "roofline": roofline,
           "job": {"subgroups": args.subgroups, "subgroups_ok": total_ok, "reads_per_step": int(total_reads),
                   "raw_reads_per_step": int(total_raw), "rank_imbalance": {"e2e_ms_min": float(tmin[0]) / K, "e2e_ms_max": e2e_max / K,
                                                                             "infer_ms_min": float(tmin[1]) / K, "infer_ms_max": infer_max / K},
                   "host_threads_per_rank": my_cores, "host_cores": cores,
                   "l2": "256 MiB device write between steps",
                   "level_steps_per_step": tot["level_steps"] / K, "draws_per_step": tot["draws"] / K,
                   "msa_problems_per_step": tot["msa_problems"] / K, "msa_dp_cells_per_step": tot["msa_dp_cells"] / K}}

    if not args.no_extra:
        from rambl_b200 import synth
        # ---- configs[1]: one chain on one GPU
        sg1 = synth.make_subgroup(20000, 150, 10, divergence=(0.01, 0.03), seed=0)
        e1, i1, a1 = timed_passes(torch, api, [sg1], min(K, 3), 1, flush)
        k1 = min(K, 3)
        gk = a1.get("dpm_kernel_ms", 0) or a1.get("gibbs_kernel_ms", 0)
        gb = a1.get("dpm_alg_bytes", 0) or a1.get("gibbs_alg_bytes", 0)
        out["config1"] = {"workload": "configs[1]: single subgroup, 20k 150bp reads (%d after depth-800 down-sampling), 10 strains at "
                                      "1-3%% divergence, whole gene" % sg1.n_reads,
                          "value": sg1.n_reads * k1 / (i1 / 1000.0), "e2e": sg1.n_reads * k1 / (e1 / 1000.0), "unit": UNIT,
                          "ms_per_step": i1 / k1, "e2e_ms_per_step": e1 / k1, "steps": k1,
                          "draws_per_s": a1.get("draws", 0) / (gk / 1000.0) if gk else None,
                          "chain_kernel_GBps": (gb / 1e9) / (gk / 1000.0) if gk else None,
                          "chain_kernel_frac_of_hbm_peak": ((gb / 1e9) / (gk / 1000.0)) / peak if gk else None,
                          "h2d_bytes_per_step": int(a1.get("h2d_bytes", 0) / k1), "gpu_launches_per_step": a1.get("gpu_launches", 0) / k1}
        # ---- the reference arm's sample set, as one batch
        _, msg = matched_sample_set(K)
        em, im, am = timed_passes(torch, api, msg, 3, 2, flush)
        mreads = sum(s.n_reads for s in msg)
        out["matched"] = {"sample": matched_description(w), "reads_per_step": mreads,
                          "e2e_reads_per_s": mreads * 3 / (em / 1000.0), "infer_reads_per_s": mreads * 3 / (im / 1000.0),
                          "e2e_ms_per_step": em / 3, "steps": 3,
                          "note": "divide by the value of `bench.py --impl reference` (same subgroups, same windows) for the "
                                  "like-for-like ratio"}
        out["poa"] = poa_block(api)
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline_block(w)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
