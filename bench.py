#!/usr/bin/env python
"""Benchmark of the pattern-partition DP on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N = 1  : BASELINE config 3 — one full 9-mer DP (`NNNNANNNN`, 2 562 890 625 patterns) + backtrack.
         A step = count expansion (K2) + wave-front DP with fused scoring (K3+K4) + backtrack (K5),
         k-mer count tables already resident in HBM.  metric = pattern-scores/sec.
N > 1  : BASELINE config 4 — the 3x3 (alpha, penalty) grid x 5 folds = 45 single-fold DP jobs of the
         same size, dealt to the ranks by job, one all_gather (NCCL) of the per-job losses.
         A step = fold count expansion + this rank's jobs + gather + selection, held-out fold tables
         already sampled.  metric = pattern-scores/sec over all jobs; the grid wall time is reported too.
--impl reference : the CPU oracle (C restatement of the reference's numba path, all host threads) on a
         bounded sample of the same workload (the 9-mer DP restricted to k-mers starting with A).

Prints ONE JSON line on rank 0.  Timing: CUDA events on the launching stream, barrier + synchronize on
both sides, max over ranks.  Tables (>= 12.8 GB) are far larger than the 126 MB L2, so no explicit flush.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GEN_PAT = "NNNNANNNN"
SAMPLE_GEN_PAT = "ANNNANNNN"     # CPU sample: the sub-problem of k-mers starting with A (1/15 of the patterns)
ALPHA, PENALTY = 1.0, 6.0
CV_ALPHAS, CV_PENALTIES, CV_FOLDS, CV_SEED = [0.5, 1.0, 10.0], [3.0, 5.0, 6.0], 5, 1
ALGO_BYTES_SINGLE = 41.0   # f32 best W+R (8) + int64 M,U W+R (32) + u8 split W (1)
ALGO_BYTES_CV = 48.0       # f32 train+test W+R (16) + read 4 int64 counts (32)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """Samples SM clocks and throttle reasons through NVML (in-process thread) while the timed region runs.
    NVML is initialised in __init__, well before the timed region: spawning nvidia-smi or initialising NVML
    next to the timed steps stalls the driver for tens of milliseconds and would distort the measurement."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period_s=0.025):
        self.rows, self.period, self.running, self.thread = [], period_s, False, None
        self.handle = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # NVML missing: report it, do not guess
            self.error = repr(e)

    def _loop(self):
        nv = self.nv
        while self.running:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((float(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.handle is None:
            return
        self.running = True
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.handle is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "error", "?")]}
        self.running = False
        self.thread.join(timeout=2)
        sm = [r[0] for r in self.rows]
        reasons = sorted({name for _, bits in self.rows for name, bit in self.REASONS.items() if bits & bit})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm)}


def dist_setup(ngpus):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def barrier_sync(world):
    import torch

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world, device):
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist

    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def pinned(arr):
    """Copy a numpy array into page-locked host memory (still a numpy view)."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t.numpy(), t


# ---------------------------------------------------------------------------------------------
# CPU legs (the oracle is only ever the baseline/checker here, never the product path)
# ---------------------------------------------------------------------------------------------
def cpu_sample_run(nthreads=0):
    from kmerpapa_b200 import synthetic
    from oracle import kp_oracle as O

    O.build()
    kmers, pos, neg = synthetic.negbin_counts(GEN_PAT, 9003)
    keep = [i for i, km in enumerate(kmers) if km[0] == "A"]          # k-mer index order is preserved
    M, U = pos[keep].astype(np.uint64), neg[keep].astype(np.uint64)
    mu = int(pos.sum()) / (int(pos.sum()) + int(neg.sum()))
    beta = (ALPHA * (1.0 - mu)) / mu
    npat, _, _ = O.plan_info(SAMPLE_GEN_PAT)
    threads = O.lib().kpo_max_threads() if nthreads == 0 else nthreads
    t0 = time.perf_counter()
    res = O.single_dp(SAMPLE_GEN_PAT, M, U, ALPHA, beta, PENALTY, nthreads=threads)
    n = len(O.backtrack(SAMPLE_GEN_PAT, res["split"]))
    dt = time.perf_counter() - t0
    return {"npat": npat, "seconds": dt, "threads": threads, "partition": n}


def run_reference(args, emit):
    """--impl reference: the reference's CPU path (oracle port, all host threads) on the bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        cpu_sample_run()
    times = []
    for _ in range(args.steps):
        r = cpu_sample_run()
        times.append(r["seconds"])
    ms = 1e3 * sum(times) / len(times)
    value = r["npat"] / (ms / 1e3)
    sample = f"9-mer DP restricted to k-mers starting with A ({SAMPLE_GEN_PAT}, {r['npat']} patterns), full DP + backtrack per step"
    line = {
        "impl": "reference", "metric": "pattern-scores/sec", "value": value, "unit": "patterns/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg3 synthetic 9-mer single DP (CPU arm on a bounded 1/15 sample)", "gen_pat": GEN_PAT,
                   "alpha": ALPHA, "penalty": PENALTY},
        "cpu_baseline": {"value": value, "unit": "patterns/s", "cores": r["threads"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "patterns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------
def load_traffic(kind):
    """DRAM bytes per DP from the committed ncu capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(kind)
    return None


def bench_single(args, rank, world, local):
    import torch

    from kmerpapa_b200 import synthetic
    from kmerpapa_b200.algorithms import bottum_up_array_w_numba as single
    from kmerpapa_b200.engine import get_plan

    dev = torch.device("cuda", local)
    kmers, pos, neg = synthetic.negbin_counts(GEN_PAT, 9003 + rank)
    codes, _c = pinned(synthetic.codes_of(kmers))
    pos_p, _p = pinned(pos)
    neg_p, _n = pinned(neg)
    mu = int(pos.sum()) / (int(pos.sum()) + int(neg.sum()))
    beta = (ALPHA * (1.0 - mu)) / mu
    max_count = int(pos.sum()) + int(neg.sum())
    plan = get_plan(GEN_PAT, local)
    npat = plan.npat
    kM, kU = plan.pack_counts(codes, pos_p, neg_p)

    def device_step(ev=None):
        eM, eU = plan.expand(kM, kU)
        if ev:
            ev[0].record()
        best, kept = plan.dp_single(eM, eU, max_count, ALPHA, beta, PENALTY)
        if ev:
            ev[1].record()
        patnums = plan.backtrack(best, kept)
        return plan.top_score(best), patnums

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        loss, patnums = device_step()
    barrier_sync(world)
    sampler.start()
    l0 = plan.launches
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start.record()
    for s in range(args.steps):
        loss, patnums = device_step(kev[s])
    end.record()
    barrier_sync(world)
    clocks = sampler.stop()
    launches = plan.launches - l0
    ms_total = max_over_ranks(start.elapsed_time(end), world, dev)
    ms_step = ms_total / args.steps
    dp_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps        # the DP wave kernels alone
    value = world * npat / (ms_step / 1e3)

    # end to end through the public array API: host buffers in, partition out
    def e2e_step():
        return single.partition_from_arrays(GEN_PAT, codes, pos_p, neg_p, ALPHA, beta, PENALTY, device=local)

    e2e_step()
    barrier_sync(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss2, pat2 = e2e_step()
    barrier_sync(world)
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0), world, dev) / args.steps
    assert loss2 == loss and np.array_equal(pat2, patnums)

    peak, peak_kind = measured_peak()
    achieved = ALGO_BYTES_SINGLE * npat / (dp_ms / 1e3) / 1e9
    line = {
        "metric": "pattern-scores/sec", "value": value, "unit": "patterns/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg3 synthetic 9-mer (neg-binomial), single penalty+pseudo, full DP + backtrack",
                   "gen_pat": GEN_PAT, "npat": npat, "alpha": ALPHA, "penalty": PENALTY, "partition_patterns": int(len(patnums)),
                   "loss": float(loss), "l2": "score table 11 GB >> 126 MB L2, no flush needed",
                   "replicas": world},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": load_traffic("single_dp_bytes"), "kernel": "kp_dp_rows_kernel (fused lazy score + min-plus), all 16 wave launches of one DP",
                     "kernel_ms": dp_ms, "algorithmic_bytes_per_pattern": ALGO_BYTES_SINGLE, "peak_kind": peak_kind,
                     "design_bytes_per_pattern": 4.0 * 3616 / 3375 * (1 + 16.7) + 2 * 226 / 3375.0},
        "e2e": {"value": world * npat / (e2e_ms / 1e3), "unit": "patterns/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(codes.nbytes + pos_p.nbytes + neg_p.nbytes),
                "d2h_bytes_per_step": int(4 + 8 * len(patnums) + 16)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    return line


def bench_cv(args, rank, world, local, steps=None, warmup=None):
    import torch

    from kmerpapa_b200 import CV_tools, synthetic
    from kmerpapa_b200.algorithms import bottum_up_array_penalty_plus_pseudo_CV as cv
    from kmerpapa_b200.engine import get_plan

    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    dev = torch.device("cuda", local)
    kmers, pos, neg = synthetic.negbin_counts(GEN_PAT, 9004)
    codes = synthetic.codes_of(kmers)
    plan = get_plan(GEN_PAT, local)
    npat = plan.npat
    prng = np.random.RandomState(CV_SEED)
    folds = CV_tools.sample_fold_counts(kmers, pos, neg, CV_FOLDS, prng)      # host sampler, outside the timed region
    njobs = CV_FOLDS * len(CV_ALPHAS) * len(CV_PENALTIES)

    def step():
        runner = cv.GpuFoldRunner(GEN_PAT, codes, pos, neg, device=local)
        res = cv.run_grid(GEN_PAT, kmers, codes, pos, neg, CV_ALPHAS, CV_PENALTIES, CV_FOLDS, 1, CV_SEED, runner=runner,
                          presampled=[folds])
        return cv.select_best(CV_ALPHAS, CV_PENALTIES, res, 1, CV_FOLDS, len(GEN_PAT))

    sampler = ClockSampler(local)
    for _ in range(warmup):
        best = step()
    barrier_sync(world)
    sampler.start()
    l0 = plan.launches
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    start.record()
    for _ in range(steps):
        best = step()
    end.record()
    barrier_sync(world)
    wall_ms = 1e3 * (time.perf_counter() - t0)     # host clock around the same steps, barrier included: the end-to-end time
    clocks = sampler.stop()
    launches = plan.launches - l0
    ms_step = max_over_ranks(start.elapsed_time(end), world, dev) / steps
    e2e_ms_step = max_over_ranks(wall_ms, world, dev) / steps
    value = njobs * npat / (ms_step / 1e3)
    peak, peak_kind = measured_peak()
    achieved = ALGO_BYTES_CV * njobs * npat / (ms_step / 1e3) / 1e9 / world
    out = {"jobs": njobs, "wall_s": ms_step / 1e3, "pattern_scores_per_s": value, "e2e_wall_s": e2e_ms_step / 1e3,
           "e2e_pattern_scores_per_s": njobs * npat / (e2e_ms_step / 1e3), "selected": [best[0], best[1], float(best[2])],
           "launches": int(launches), "clocks": clocks, "achieved_gbs_per_gpu": achieved, "frac_per_gpu": achieved / peak,
           "peak": peak, "peak_kind": peak_kind}
    return out


def bench_sharded_dp(rank, world, local, reps=6):
    """Secondary N>1 measurement: ONE 9-mer DP (config 3) sharded over the ranks (kmerpapa_b200/sharded.py, replicated
    mode: rows pushed to their readers over NVLink inside the DP kernel).  Device time, max over ranks, best of reps."""
    import torch
    import torch.distributed as dist

    from kmerpapa_b200 import sharded, synthetic
    from kmerpapa_b200.engine import get_plan

    kmers, pos, neg = synthetic.negbin_counts(GEN_PAT, 9003)
    plan = get_plan(GEN_PAT, local)
    kM, kU = plan.pack_counts(synthetic.codes_of(kmers), pos, neg)
    eM, eU = plan.expand(kM, kU)
    mc = int(pos.sum() + neg.sum())
    mu = int(pos.sum()) / mc
    alpha, penalty = 1.0, 6.0
    beta = alpha * (1.0 - mu) / mu
    sh = sharded.ShardedDP(plan, rank, world, replicate=True)
    sh.connect()
    ms = []
    for _ in range(reps):
        sh.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sh.run(eM, eU, mc, alpha, beta, penalty)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=plan.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(float(t.item()))
    part = sh.backtrack()
    top = float(sh.top_score())
    sh.close()
    best = min(ms[1:])
    return {"workload": "cfg3 single 9-mer DP sharded by the top high digit, replicated mode", "ms": best,
            "pattern_scores_per_s": plan.npat / (best / 1e3), "partition_patterns": int(len(part)), "loss": top,
            "reps_ms": [round(x, 3) for x in ms]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cv", action="store_true", help="N=1: skip the secondary CV-grid measurement")
    ap.add_argument("--no-sharded", action="store_true", help="N>1: skip the secondary sharded single-DP measurement")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything else that writes to file descriptor 1 (the NCCL banner,
    # stray prints of libraries) is sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, emit)
        return
    import __graft_entry__ as ge

    rank, world, local = dist_setup(args.gpus)
    if rank == 0:
        ge.build()
    barrier_sync(world)
    if world == 1:
        line = bench_single(args, rank, world, local)
        if not args.no_cv:
            cvres = bench_cv(args, rank, world, local, steps=1, warmup=1)
            line["cv_grid"] = cvres
    else:
        cvres = bench_cv(args, rank, world, local)
        line = {
            "metric": "pattern-scores/sec", "value": cvres["pattern_scores_per_s"], "unit": "patterns/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cvres["wall_s"] * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg4 synthetic 9-mer, 3x3 penalty x pseudo grid, 5-fold CV sharded by job",
                       "gen_pat": GEN_PAT, "jobs": cvres["jobs"], "alphas": CV_ALPHAS, "penalties": CV_PENALTIES,
                       "nfolds": CV_FOLDS, "l2": "train table 11 GB per job >> 126 MB L2, no flush needed"},
            "roofline": {"bound": "hbm", "achieved": cvres["achieved_gbs_per_gpu"], "peak": cvres["peak"], "unit": "GB/s",
                         "frac": cvres["frac_per_gpu"], "traffic": load_traffic("cv_job_bytes"),
                         "kernel": "kp_dp_rows_kernel on train counts + backtrack/leaf kernels, per GPU", "algorithmic_bytes_per_pattern": ALGO_BYTES_CV,
                         "peak_kind": cvres["peak_kind"]},
            "e2e": {"value": cvres["e2e_pattern_scores_per_s"], "unit": "patterns/s", "ms_per_step": cvres["e2e_wall_s"] * 1e3,
                    "h2d_bytes_per_step": int(65536 * 8 * 3 * 6), "d2h_bytes_per_step": int(8 * cvres["jobs"]),
                    "note": "host clock around the same steps (every step packs the host fold tables, H2D, runs its jobs, reads "
                            "every job's losses back, D2H, and gathers them over NCCL), max over ranks"},
            "cv_grid": cvres, "gpu_launches": cvres["launches"], "clocks": cvres["clocks"],
        }
        if not args.no_sharded and world <= 8:
            line["sharded_single_dp"] = bench_sharded_dp(rank, world, local)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_sample_run()
        line["cpu_baseline"] = {
            "value": r["npat"] / r["seconds"], "unit": "patterns/s", "cores": r["threads"], "kind": "port",
            "sample": f"{SAMPLE_GEN_PAT}: the 9-mer DP restricted to k-mers starting with A, {r['npat']} patterns, "
                      f"{r['seconds']:.1f} s, oracle C port of the numba path with OpenMP"}
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


if __name__ == "__main__":
    main()
