#!/usr/bin/env python
"""bench.py -- headline benchmark of the SI hot path (BASELINE.json):
    SI train iters/sec (n=5k DGP) + predict points/sec, FP64, at 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --steps K --warmup W    (CPU path on the box's host cores)

One "step" = one stochastic-EM iteration of `dgp.train` (I-step: ess_burn+1 = 11 ESS sweeps; M-step: L-BFGS-B
over every GP node) on BASELINE config 3: 3-layer DGP (8+8+2 squared-exponential nodes, global input
connections), n=5000, d=8, 2 outputs, synthetic data.  At N>1 every rank runs an independent replica of the
chain (the Markov chain itself does not shard, DESIGN.md "multi-GPU") and `value` is the aggregate.
The secondary metric `predict` is `emulator.predict(x, 'mean_var')` points/s with the test points sharded
over the ranks and the moments all-gathered over NCCL.

Prints ONE JSON line (see the task contract): metric/value/unit/..., `e2e`, `roofline`, `cpu_baseline`,
`gpu_launches`, `clocks`.
"""
from __future__ import annotations

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

SEED = 20261017 + 2
METRIC = "SI train iters/sec (n=5k DGP)"


# ------------------------------------------------------------------------------------------------ workload
def make_config3(n, rng):
    """BASELINE config 3: X~U[0,1]^{n x 8}; y1=sin(sum x)+x1 x2; y2=cos(2 pi x3) x4 + x5^2."""
    X = rng.uniform(0, 1, size=(n, 8))
    Y = np.stack([np.sin(X.sum(1)) + X[:, 0] * X[:, 1], np.cos(2 * np.pi * X[:, 2]) * X[:, 3] + X[:, 4] ** 2], 1)
    return X, Y


def layers_config3(make):
    l1 = [make(length=np.array([1.0]), name="sexp") for _ in range(8)]
    l2 = [make(length=np.array([1.0]), name="sexp", connect=np.arange(8)) for _ in range(8)]
    l3 = [make(length=np.array([1.0]), name="sexp", scale_est=True, connect=np.arange(8)) for _ in range(2)]
    return [l1, l2, l3]


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag = index, [], threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop_flag.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_counts_small(n_small=250):
    """Evaluation counts of one SEM iteration (they are what the reference executes per iteration: one
    Cholesky per prior draw, per threshold node and per proposal node, kernel_class.py:481-488,
    imputation.py:54-107; one `llik` per L-BFGS-B evaluation).  Measured on a small-n oracle run of the same
    model because they depend on the chain, not on n."""
    from oracle import dgp_oracle as O

    rng = np.random.default_rng(SEED)
    X, Y = make_config3(n_small, rng)
    layers = O.build_dgp(X, Y, layers_config3(lambda **kw: O.Node(**kw)))
    cnt = {"ll": 0, "grad": 0}
    orig_ll, orig_draw, orig_obj = O.Node.loglik, O.Node.prior_draw, O.Node.objective

    def ll(self):
        cnt["ll"] += 1
        return orig_ll(self)

    def draw(self, z):
        cnt["ll"] += 1  # one n^3/3 factorisation, same cost class as a likelihood
        return orig_draw(self, z)

    def obj(self, x):
        cnt["grad"] += 1
        return orig_obj(self, x)

    O.Node.loglik, O.Node.prior_draw, O.Node.objective = ll, draw, obj
    try:
        O.ess_sweeps(layers, 10, rng)  # burn-in sweep of the constructor (dgp.py:126)
        cnt["ll"] = cnt["grad"] = 0
        O.sem_iteration(layers, rng)
    finally:
        O.Node.loglik, O.Node.prior_draw, O.Node.objective = orig_ll, orig_draw, orig_obj
    return cnt


def cpu_step(n, rng):
    """One bounded CPU sample at the full n: one ESS likelihood (K build + Cholesky + solve) and one M-step
    objective with gradient, for a config-3 upper node (D = 8 local + 8 global)."""
    from oracle import dgp_oracle as O

    X = rng.uniform(0, 1, size=(n, 16))
    y = np.sin(X.sum(1))
    t0 = time.perf_counter()
    O.loglik_dense(X, y, np.array([1.0]), 1.0, 1e-6, "sexp")
    t1 = time.perf_counter()
    O.nllik_grad_dense(X, y, np.array([1.0]), 1.0, 1e-6, "sexp", False, False)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def cpu_cores():
    try:
        import psutil
        return psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:
        return os.cpu_count()


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core (the reference's default)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass


def blas_info():
    try:
        from threadpoolctl import threadpool_info
        return "; ".join(f"{d.get('internal_api')} {d.get('version')} x{d.get('num_threads')}" for d in threadpool_info())
    except Exception:
        return "unknown"


def run_reference(args):
    """`--impl reference`: the reference's CPU algorithm (oracle port: numpy + LAPACK on all host cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n
    rng = np.random.default_rng(SEED)
    use_all_host_threads()
    cnt = cpu_counts_small()
    for _ in range(args.warmup):
        cpu_step(min(n, 1500), rng)
    tl, tg = [], []
    t_start = time.perf_counter()
    for _ in range(args.steps):
        a, b = cpu_step(n, rng)
        tl.append(a)
        tg.append(b)
    wall = time.perf_counter() - t_start
    t_iter = cnt["ll"] * float(np.mean(tl)) + cnt["grad"] * float(np.mean(tg))
    value = 1.0 / t_iter
    sample = (f"per step: 1 ESS likelihood + 1 llik(grad) at n={n}, D=16 on the host "
              f"(mean {np.mean(tl):.2f}s / {np.mean(tg):.2f}s); extrapolated to one SEM iteration with the evaluation "
              f"counts of the oracle chain at n=250: {cnt['ll']} factorisations + {cnt['grad']} gradient evaluations; "
              f"BLAS: {blas_info()}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_iter, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": "iters/s", "cores": cpu_cores(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"BASELINE config 3: 3-layer DGP, 8+8+2 sexp nodes with global connections, n={args.n}, d=8, "
                        f"2 outputs; one step = one SEM iteration (11 ESS sweeps + M-step)",
            "n": args.n, "ess_burn": 10, "nodes": 18, "chains": 1, "gpus_sharing_the_chain": world,
            "cache": "working set per batched factorisation >> 126 MB L2 (8 x n^2 doubles); no L2 flush needed",
            "predict_points": args.predict_points, "predict_imputations": args.predict_imputations}


# ------------------------------------------------------------------------------------------------ GPU arm
def measure_fp64_peak(torch):
    """FP64 roofline denominator: cuBLAS DGEMM 8192^3 (MEASURED_PEAKS.json holds no FP64 figure)."""
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    best = 0.0
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        if i:
            best = max(best, 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    torch.cuda.empty_cache()
    return best


def run_gpu(args):
    import torch

    import dgp_b200 as D
    from dgp_b200 import _lib as L

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = L.load()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    peak = measure_fp64_peak(torch) if rank == 0 else 0.0

    # N > 1: ONE chain shared by the ranks (same data, same draws on every rank; ESS waves and M-step nodes are
    # dealt over the GPUs, dgp_b200/parallel.py) -- strong scaling of a single `train` call
    if dist is not None:
        from dgp_b200 import parallel
        parallel.enable(dist)
    rng = np.random.default_rng(SEED)
    np.random.seed(SEED)
    D.nb_seed(SEED)
    X, Y = make_config3(args.n, rng)
    model = D.dgp(X, Y, layers_config3(lambda **kw: D.kernel(**kw)))
    for _ in range(args.warmup):
        model.train(1, disable=True)

    h2d0, d2h0 = L.COUNTERS["h2d"], L.COUNTERS["d2h"]
    launches0 = lib.dgpb_launch_count()
    prop0 = model.imp.n_proposals
    lib.dgpb_profile(1)
    tim0 = dict(model.timing)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            model.train(1, disable=True)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    prof = (L.c_dbl * 4)()
    lib.dgpb_profile_read(prof)
    lib.dgpb_profile(0)
    launches = lib.dgpb_launch_count() - launches0
    h2d, d2h = L.COUNTERS["h2d"] - h2d0, L.COUNTERS["d2h"] - d2h0
    nprop = model.imp.n_proposals - prop0

    times = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = float(times[0]), float(times[1])

    # ---- secondary metric: predict points/s (test points sharded over ranks, NCCL all-gather) -------
    predict = None
    if args.predict_points > 0:
        from dgp_b200.parallel import predict_sharded
        emu = D.emulator(model.estimate(), N=args.predict_imputations)
        xt_all = np.random.default_rng(SEED + 99).uniform(0, 1, size=(args.predict_points * world, 8))
        predict_sharded(emu, xt_all, dist)  # warm-up at full size (see the Vecchia leg below)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tp0 = time.perf_counter()
        p0.record()
        mu, var = predict_sharded(emu, xt_all, dist)
        p1.record()
        barrier()
        tp = time.perf_counter() - tp0
        tt = torch.tensor([p0.elapsed_time(p1), tp * 1e3], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        predict = {"metric": "predict points/sec (mean_var)", "value": len(xt_all) / (float(tt[0]) * 1e-3),
                   "e2e_value": len(xt_all) / (float(tt[1]) * 1e-3), "unit": "points/s", "points": len(xt_all),
                   "imputations": args.predict_imputations, "warmup": "one full-size call",
                   "sharding": "test points / rank, all_gather of (mu, var)",
                   "finite": bool(np.all(np.isfinite(mu)) and np.all(np.isfinite(var)))}

    # ---- secondary metric 2: Vecchia DGP prediction, BASELINE config 4 shape (n=100k, d=10, m=25, 10+1 sexp
    #      nodes), test points sharded over the ranks; kNN included -----------------------------
    predict_v = None
    if args.vecchia_points > 0:
        from dgp_b200.parallel import predict_sharded
        seed4 = 20261017 + 3
        rng4 = np.random.default_rng(seed4)
        np.random.seed(seed4)
        D.nb_seed(seed4)
        n4, d4 = args.vecchia_n, 10
        X4 = rng4.uniform(0, 1, (n4, d4))
        f4 = lambda x: np.sin(2 * np.pi * x[:, 0] * x[:, 1]) + x[:, 2] ** 2 + np.cos(3 * x[:, 3:].sum(1))
        Y4 = (f4(X4) + 0.05 * rng4.standard_normal(n4)).reshape(-1, 1)
        l1 = [D.kernel(length=np.array([1.0]), name="sexp") for _ in range(10)]
        l2 = [D.kernel(length=np.array([1.0]), name="sexp", scale_est=True, nugget_est=True, nugget=1e-2,
                       connect=np.arange(10))]
        tv0 = time.perf_counter()
        m4 = D.dgp(X4, Y4, D.combine(l1, l2), vecchia=True, m=25)
        m4.train(args.vecchia_train_iters, disable=True)
        torch.cuda.synchronize()
        t_train = (time.perf_counter() - tv0)
        emu4 = D.emulator(m4.estimate(burnin=0), N=args.vecchia_imputations)
        xt4 = np.random.default_rng(seed4 + 99).uniform(0, 1, size=(args.vecchia_points, d4))
        # warm-up at full size: the first call at a new size re-sizes the library's scratch slots and torch's pool
        # (cudaMalloc / cudaFree, synchronising; 0.03-0.8 s here depending on what ran before), which is not throughput
        predict_sharded(emu4, xt4, dist, m=25)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tp0 = time.perf_counter()
        p0.record()
        mu4, var4 = predict_sharded(emu4, xt4, dist, m=25)
        p1.record()
        barrier()
        tp = time.perf_counter() - tp0
        tt = torch.tensor([p0.elapsed_time(p1), tp * 1e3], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        S4, M4 = args.vecchia_imputations, len(xt4)
        # algorithmic FLOPs: kNN 2D flop per (query, candidate) pair on the tensor path -- layer 1
        # once (one design matrix, shared length-scale), layer 2 once per imputation (D = 20); per block
        # b^3/3 + b^2 (3D + 25)/2 + 2 b^2 flop, b = 26, for the 10 first-layer nodes; ~70 kflop per linked block
        b = 26.0
        blk1 = b ** 3 / 3 + b * b * (3 * 10 + 25) / 2 + 2 * b * b
        flops = M4 * n4 * 2.0 * 10 + S4 * M4 * n4 * 2.0 * 20 + S4 * M4 * (10 * blk1 + 7.0e4)
        t_pred = float(tt[0]) * 1e-3
        predict_v = {"metric": "Vecchia DGP predict points/sec (mean_var, m=25)", "value": M4 / t_pred,
                     "e2e_value": M4 / (float(tt[1]) * 1e-3), "unit": "points/s", "points": M4, "n_train": n4,
                     "imputations": S4, "train_iters": args.vecchia_train_iters, "warmup": "one full-size call",
                     "train_s_per_iter_incl_construct": t_train / max(1, args.vecchia_train_iters),
                     "node_imputation_points_per_s": M4 * S4 * 11 / t_pred,
                     "algorithmic_tflops": flops / t_pred / 1e12,
                     "frac_of_fp64_peak": (flops / t_pred / 1e12 / peak) if peak else None,
                     "sharding": "test points / rank, all_gather of (mu, var)",
                     "rmse_vs_truth": float(np.sqrt(np.mean((mu4[:, 0] - f4(xt4)) ** 2))),
                     "finite": bool(np.all(np.isfinite(mu4)) and np.all(np.isfinite(var4)))}

    if rank == 0:
        value = args.steps / (dev_ms_max * 1e-3)     # one chain, whatever the number of GPUs
        e2e = args.steps / (wall_ms_max * 1e-3)
        upd_ms, upd_n, upd_flops = prof[0], prof[1], prof[2]
        achieved = (upd_flops / upd_n) / (upd_ms / upd_n * 1e-3) / 1e12 if upd_n > 0 else None
        roof = {"bound": "tensor", "kernel": "update_kernel, bulk launches (FP64 DMMA SYRK trailing update, K = 512 hyper-blocks / 128 tail)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                # DRAM bytes per launch: one `ncu --set full` capture of a K = 512 bulk launch (profiles/
                # r1s2_update_ncu_raw.txt: 770.1 MB read + written for 41.94 GFLOP; algorithmic 760 MB), scaled to
                # this run's average launch
                "traffic": (upd_flops / upd_n) * (770.08e6 / 41.94e9) if upd_n > 0 else None,
                "traffic_source": "profiles/r1s2_update_ncu_raw.txt (dram__bytes_read+write of one bulk launch, "
                                  "18.4 B per kFLOP) x this run's FLOPs per launch",
                "launches_timed": int(upd_n), "avg_launch_ms": upd_ms / upd_n if upd_n else None,
                "share_of_step": upd_ms / (dev_ms * 1.0) if dev_ms else None,
                "peak_source": "cuBLAS DGEMM 8192^3 via torch.matmul(float64), best of 5, measured in this run "
                               "(MEASURED_PEAKS.json has no FP64 entry)"}
        # whole-step view: every factorisation FLOP issued in the timed region / its duration.  The bulk update runs
        # on the lowest-priority stream and is pre-empted by the critical-path kernels, so its in-situ launch
        # durations (above) under-state the kernel; the isolated launch is re-measured here for context.
        probe = L.host_doubles(2)
        iso = None
        if lib.dgpb_probe_update(L.workspace(), 5000, 8, 16 << 8, 5, probe) == 0:
            iso = {"achieved": probe[1], "frac": probe[1] / peak if peak else None, "ms_per_launch": probe[0],
                   "what": "K = 512 bulk launch alone, window 4544, 8 matrices (dgpb_probe_update)"}
        roof["isolated"] = iso
        roof["note"] = ("bulk launches run on the lowest-priority stream and yield SM slots to the critical-path kernels "
                        "(panels, inner and look-ahead updates on a highest-priority stream), so their in-situ event "
                        "durations include time spent executing those kernels: `isolated` is the same launch alone, "
                        "`step` is every factorisation FLOP of the timed region over its duration")
        step_tf = prof[3] / (dev_ms * 1e-3) / 1e12 if dev_ms else None
        roof["step"] = {"achieved": step_tf, "frac": (step_tf / peak) if (step_tf and peak) else None,
                        "flops_per_step": prof[3] / args.steps,
                        "what": "all factorisation FLOPs issued in the timed region (n^3/3 per Cholesky incl. speculative "
                                "proposals, n^3 per gradient evaluation) / its duration"}
        cpu = None
        if not args.no_cpu_baseline:
            use_all_host_threads()
            cnt = cpu_counts_small()
            r2 = np.random.default_rng(1)
            cpu_step(1000, r2)
            a, b = cpu_step(args.n, r2)
            t_iter = cnt["ll"] * a + cnt["grad"] * b
            cpu = {"value": 1.0 / t_iter, "unit": "iters/s", "cores": cpu_cores(), "kind": "port",
                   "sample": f"1 ESS likelihood ({a:.2f}s) + 1 llik with gradient ({b:.2f}s) at n={args.n}, D=16 on "
                             f"the host, extrapolated with the oracle chain's counts at n=250 ({cnt['ll']} "
                             f"factorisations + {cnt['grad']} gradient evaluations per SEM iteration); BLAS: "
                             f"{blas_info()}"}
        line = {"metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, world),
                "e2e": {"value": e2e, "unit": "iters/s", "h2d_bytes_per_step": h2d // args.steps,
                        "d2h_bytes_per_step": d2h // args.steps},
                "gpu_launches": int(launches), "ess_proposals_per_step": nprop / args.steps,
                "phase_ms_per_step": {"i_step": 1e3 * (model.timing["i_step"] - tim0["i_step"]) / args.steps,
                                      "m_step": 1e3 * (model.timing["m_step"] - tim0["m_step"]) / args.steps,
                                      "m_step_in_library": 1e3 * (model.timing.get("m_batched_s", 0.0)
                                                                  - tim0.get("m_batched_s", 0.0)) / args.steps,
                                      "m_step_batched_calls": (model.timing.get("m_batched_calls", 0)
                                                               - tim0.get("m_batched_calls", 0)) / args.steps,
                                      "m_step_matrices": (model.timing.get("m_batched_matrices", 0)
                                                          - tim0.get("m_batched_matrices", 0)) / args.steps},
                "roofline": roof,
                "cpu_baseline": cpu, "clocks": clocks.summary(), "predict": predict, "predict_vecchia": predict_v}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=5000, help="training points (BASELINE config 3: 5000)")
    ap.add_argument("--predict-points", type=int, default=2048, help="test points PER GPU for the predict metric")
    ap.add_argument("--predict-imputations", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--vecchia-points", type=int, default=1000000,
                    help="test points of the Vecchia prediction metric (BASELINE config 4: 1M); 0 = skip")
    ap.add_argument("--vecchia-n", type=int, default=100000)
    ap.add_argument("--vecchia-imputations", type=int, default=2)
    ap.add_argument("--vecchia-train-iters", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
