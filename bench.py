#!/usr/bin/env python
"""bench.py -- headline benchmark of the SI hot path (BASELINE.json):
    SI train iters/sec (n=5k DGP) + predict points/sec, FP64, at 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --steps K --warmup W    (the UNMODIFIED reference on the box's host cores)

One "step" = one stochastic-EM iteration of `dgp.train` (I-step: ess_burn+1 = 11 ESS sweeps; M-step: L-BFGS-B
over every GP node) on BASELINE config 3: 3-layer DGP (8+8+2 squared-exponential nodes, global input
connections), n=5000, d=8, 2 outputs, synthetic data.  At N>1 the ranks share ONE chain (same data, same draws:
the candidate angles of every ESS wave and the GP nodes of the M-step are dealt over the GPUs,
dgp_b200/parallel.py), so `value` is the speed of a single `train` call and `scaling` is "strong".

Secondary legs (`legs` in the JSON line, `--legs` selects): the predict side of config 3 (10 000 points, N=10) and
BASELINE configs 4 (Vecchia, n=100k, 1M points, N=10), 2 (Matern, n=2000, 100k points, N=10), 5 (linked
GP -> DGP -> GP, 1M points, N=50) and 1 (step function, train(500)) at their stated sizes; test points are sharded
over the ranks.  At N=1 every leg also times the UNMODIFIED reference (baseline/_ref) on a bounded sample of the same
frozen inputs and compares the two results.

Prints ONE JSON line: metric/value/unit/..., `e2e`, `roofline`, `cpu_baseline`, `gpu_launches`, `clocks`, `legs`.
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

SEED0 = 20261017
SEED = SEED0 + 2
METRIC = "SI train iters/sec (n=5k DGP)"
DEFAULT_LEGS = "predict3,cfg4,cfg2,cfg5,cfg1"


# ------------------------------------------------------------------------------------------------ workloads
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


def f_config2(x):
    return np.sin(2 * np.pi * x[:, 0] * x[:, 1]) + (x[:, 2] - 0.5) ** 2 + x[:, 3] * np.exp(-x[:, 4])


def f_config4(x):
    return np.sin(2 * np.pi * x[:, 0] * x[:, 1]) + x[:, 2] ** 2 + np.cos(3 * x[:, 3:].sum(1))


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
def cpu_cores():
    try:
        import psutil
        return psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:
        return os.cpu_count()


def _headline_node_arrays(n, rng):
    """Inputs of one config-3 upper node: 8 latent + 8 connected global dimensions."""
    X = rng.uniform(0, 1, size=(n, 16))
    return X, np.sin(X.sum(1)).reshape(-1, 1)


class ReferenceArm:
    """The reference's CPU implementation of the config-3 hot path: the UNMODIFIED `dgpsi` from baseline/_ref when
    it is installed (kind "reference"), else the oracle port (kind "port").  A full SEM iteration at n = 5000 takes
    the reference about an hour, so a step is a BOUNDED SAMPLE of it -- one ESS likelihood (`log_likelihood_func`),
    one M-step objective with gradient (`llik`) and one prior draw (`fmvn`) on a real n = 5000, D = 16 node -- and
    the iteration time is those three costs times the number of calls the reference's own chain makes per
    iteration (counted by running the reference's `train(1)` on the same model at n = 200)."""

    def __init__(self, n):
        from baseline import ref_loader as R

        self.n, self.R = n, R
        R.use_all_host_threads()
        self.dgpsi = R.load()
        self.kind = "reference" if self.dgpsi is not None else "port"
        self.counts = self._counts_reference() if self.dgpsi is not None else self._counts_port()

    # -- how many of each call one SEM iteration makes (depends on the chain, not on n)
    def _counts_reference(self, n_small=200):
        d = self.dgpsi
        from dgpsi import imputation as IMP
        from dgpsi.kernel_class import kernel

        rng = np.random.default_rng(SEED)
        np.random.seed(SEED)
        d.nb_seed(SEED)
        X, Y = make_config3(n_small, rng)
        model = d.dgp(X, Y, d.combine(*layers_config3(lambda **kw: kernel(**kw))))
        cnt = {"loglik": 0, "llik": 0, "fmvn": 0}
        o_ll, o_llik, o_fmvn = kernel.log_likelihood_func, kernel.llik, IMP.fmvn

        def ll(self_):
            cnt["loglik"] += 1
            return o_ll(self_)

        def llik(self_, x):
            cnt["llik"] += 1
            return o_llik(self_, x)

        def fmvn(cov):
            cnt["fmvn"] += 1
            return o_fmvn(cov)

        kernel.log_likelihood_func, kernel.llik, IMP.fmvn = ll, llik, fmvn
        try:
            model.train(N=1, disable=True)
        finally:
            kernel.log_likelihood_func, kernel.llik, IMP.fmvn = o_ll, o_llik, o_fmvn
        return cnt

    def _counts_port(self, n_small=250):
        from oracle import dgp_oracle as O

        rng = np.random.default_rng(SEED)
        X, Y = make_config3(n_small, rng)
        layers = O.build_dgp(X, Y, layers_config3(lambda **kw: O.Node(**kw)))
        cnt = {"loglik": 0, "llik": 0, "fmvn": 0}
        o_ll, o_draw, o_obj = O.Node.loglik, O.Node.prior_draw, O.Node.objective

        def ll(self_):
            cnt["loglik"] += 1
            return o_ll(self_)

        def draw(self_, z):
            cnt["fmvn"] += 1
            return o_draw(self_, z)

        def obj(self_, x):
            cnt["llik"] += 1
            return o_obj(self_, x)

        O.Node.loglik, O.Node.prior_draw, O.Node.objective = ll, draw, obj
        try:
            O.ess_sweeps(layers, 10, rng)
            cnt["loglik"] = cnt["llik"] = cnt["fmvn"] = 0
            O.sem_iteration(layers, rng)
        finally:
            O.Node.loglik, O.Node.prior_draw, O.Node.objective = o_ll, o_draw, o_obj
        return cnt

    # -- one bounded sample: seconds of (loglik, llik, fmvn) at n points
    def sample(self, n, rng):
        X, y = _headline_node_arrays(n, rng)
        if self.dgpsi is not None:
            from dgpsi import imputation as IMP
            from dgpsi.kernel_class import kernel

            k = kernel(length=np.array([1.0]), name="sexp", connect=np.arange(8))
            k.input, k.global_input, k.input_dim = X[:, :8].copy(), X[:, 8:].copy(), np.arange(8)
            k.output, k.D = y, 16
            k.para_path = np.atleast_2d(np.concatenate((k.scale, k.length, k.nugget)))
            t0 = time.perf_counter()
            k.log_likelihood_func()
            t1 = time.perf_counter()
            k.llik(k.log_t().copy())
            t2 = time.perf_counter()
            IMP.fmvn(k.scale * k.k_matrix())     # the prior draw builds its covariance too (imputation.py:54-63)
            t3 = time.perf_counter()
            return t1 - t0, t2 - t1, t3 - t2
        from oracle import dgp_oracle as O

        t0 = time.perf_counter()
        O.loglik_dense(X, y, np.array([1.0]), 1.0, 1e-6, "sexp")
        t1 = time.perf_counter()
        O.nllik_grad_dense(X, y, np.array([1.0]), 1.0, 1e-6, "sexp", False, False)
        t2 = time.perf_counter()
        np.linalg.cholesky(O.k_matrix(X, np.array([1.0]), 1e-6, "sexp"))
        t3 = time.perf_counter()
        return t1 - t0, t2 - t1, t3 - t2

    def iteration_seconds(self, t_ll, t_llik, t_fmvn):
        c = self.counts
        return c["loglik"] * t_ll + c["llik"] * t_llik + c["fmvn"] * t_fmvn

    def describe(self, t_ll, t_llik, t_fmvn, steps):
        c = self.counts
        src = ("the unmodified reference (dgpsi 2.6.0 from baseline/_ref: kernel.log_likelihood_func, kernel.llik, "
               "imputation.fmvn)" if self.kind == "reference" else "the oracle port (reference not installed)")
        return (f"{steps} bounded sample(s) of {src} on one n={self.n}, D=16 node: log-likelihood {t_ll:.2f}s, objective + "
                f"gradient {t_llik:.2f}s, prior draw {t_fmvn:.2f}s; one SEM iteration = {c['loglik']} / {c['llik']} / "
                f"{c['fmvn']} such calls (counted on the reference's own train(1) of the same model at n=200); "
                f"host: {json.dumps(self.R.host_info())}")


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the headline workload on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = ReferenceArm(args.n)
    rng = np.random.default_rng(SEED)
    for _ in range(args.warmup):
        arm.sample(min(args.n, 1200), rng)     # JIT compilation and thread pools
    acc = np.zeros(3)
    t_start = time.perf_counter()
    for _ in range(args.steps):
        acc += arm.sample(args.n, rng)
    wall = time.perf_counter() - t_start
    t_ll, t_llik, t_fmvn = acc / args.steps
    t_iter = arm.iteration_seconds(t_ll, t_llik, t_fmvn)
    value = 1.0 / t_iter
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_iter, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": "iters/s", "cores": cpu_cores(), "kind": arm.kind,
                             "sample": arm.describe(t_ll, t_llik, t_fmvn, args.steps)},
            "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "sample_wall_s": wall,
            "note": "value is extrapolated from the bounded samples (a real iteration takes the reference ~1 h at n=5000)"}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"BASELINE config 3: 3-layer DGP, 8+8+2 sexp nodes with global connections, n={args.n}, d=8, "
                        f"2 outputs; one step = one SEM iteration (11 ESS sweeps + M-step)",
            "n": args.n, "ess_burn": 10, "nodes": 18, "chains": 1, "gpus_sharing_the_chain": world,
            "cache": "working set per batched factorisation >> 126 MB L2 (8 x n^2 doubles); no L2 flush needed",
            "legs": args.legs, "leg_seconds": args.leg_seconds}


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


def fit_points(stated, probe, t_probe, budget_s, granule):
    """Test points of a predict leg that fit `budget_s` seconds, from a timed probe of `probe` points: the stated
    number when it fits (or when there is no budget), else the largest multiple of `granule` that does (at least the
    probe).  Throughput of these legs is a steady-state rate (identical chunks), so fewer points change the duration,
    not the points/s."""
    if budget_s <= 0 or t_probe <= 0 or stated <= probe:
        return int(stated)
    if stated * t_probe / probe <= budget_s:
        return int(stated)
    fit = int(budget_s * probe / t_probe) // granule * granule
    return int(min(stated, max(probe, fit)))


class Ctx:
    """What every leg needs: torch, the package, the process group and a device-side max-over-ranks timer."""

    def __init__(self, args):
        import torch

        import dgp_b200 as D
        from dgp_b200 import _lib as L
        from dgp_b200 import parallel

        self.args, self.torch, self.D, self.L, self.parallel = args, torch, D, L, parallel
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.lib = L.load()
        self.peak = 0.0
        self.ref = None          # loaded lazily: the unmodified reference for the CPU samples (rank 0, N = 1)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def share_chain(self, on):
        if self.dist is None:
            return
        if on:
            if self.parallel.chain() is None:
                self.parallel.enable(self.dist)
        else:
            self.parallel.disable()

    def seed(self, seed):
        np.random.seed(seed)
        self.D.nb_seed(seed)
        return np.random.default_rng(seed)

    def timed(self, fn):
        """fn() between barriers: (result, device seconds, wall seconds), both maxima over the ranks."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = fn()
        e1.record()
        self.barrier()
        wall = time.perf_counter() - t0
        tt = torch.tensor([e0.elapsed_time(e1) * 1e-3, wall], dtype=torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return out, float(tt[0]), float(tt[1])

    def sum_over_ranks(self, values):
        tt = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(tt)
        return [float(v) for v in tt]

    def predict(self, emu, x, **kw):
        from dgp_b200.parallel import predict_sharded
        return predict_sharded(emu, x, self.dist, **kw)

    def bounded_points(self, emu, xt, stated, **kw):
        """`stated` test points, or as many as fit --leg-seconds (default run: the whole bench line within minutes on
        one GPU; `--leg-seconds 0` runs every leg at the size BASELINE.json states).  Decided from a timed probe whose
        duration is the maximum over the ranks, so every rank takes the same decision."""
        probe = min(stated, 1024 * self.world)
        _, t_probe, _ = self.timed(lambda: self.predict(emu, xt[:probe], **kw))
        return fit_points(stated, probe, t_probe, self.args.leg_seconds, 256 * self.world)

    def cpu_reference(self):
        """The unmodified reference for the bounded CPU samples of the legs: rank 0 of a single-GPU run only."""
        if self.args.no_cpu_baseline or self.world != 1:
            return None
        if self.ref is None:
            from baseline import ref_loader as R
            R.use_all_host_threads()
            self.ref = (R, R.load())
        return self.ref if self.ref[1] is not None else None


def _err(a, b):
    """GPU result against the reference's on the same frozen inputs: largest absolute difference and the scale."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return {"max_abs_diff": float(np.max(np.abs(a - b))), "max_abs_reference": float(np.max(np.abs(b)))}


def leg_train(ctx):
    """Headline: SEM iterations/s of config 3."""
    args, D, L, lib = ctx.args, ctx.D, ctx.L, ctx.lib
    ctx.share_chain(True)
    rng = ctx.seed(SEED)
    X, Y = make_config3(args.n, rng)
    model = D.dgp(X, Y, layers_config3(lambda **kw: D.kernel(**kw)))
    for _ in range(args.warmup):
        model.train(1, disable=True)

    h2d0, d2h0 = L.COUNTERS["h2d"], L.COUNTERS["d2h"]
    launches0 = lib.dgpb_launch_count()
    prop0 = model.imp.n_proposals
    lib.dgpb_profile(1)
    tim0 = dict(model.timing)
    with ClockSampler(ctx.local) as clocks:
        _, dev_s, wall_s = ctx.timed(lambda: [model.train(1, disable=True) for _ in range(args.steps)])
    prof = (L.c_dbl * 8)()
    lib.dgpb_profile_read(prof)
    lib.dgpb_profile(0)
    launches = lib.dgpb_launch_count() - launches0
    h2d, d2h = L.COUNTERS["h2d"] - h2d0, L.COUNTERS["d2h"] - d2h0
    nprop = model.imp.n_proposals - prop0
    upd_ms, upd_n, upd_flops, issued, wasted = ctx.sum_over_ranks(prof[:5])
    launches_all = ctx.sum_over_ranks([launches])[0]
    return {"model": model, "dev_s": dev_s, "wall_s": wall_s, "clocks": clocks.summary(), "launches": int(launches_all),
            "h2d": h2d, "d2h": d2h, "nprop": nprop, "timing0": tim0,
            "upd": (upd_ms, upd_n, upd_flops), "issued": issued, "wasted": wasted}


def leg_predict3(ctx, model):
    """Predict side of config 3: `emulator(N=10).predict` of 10 000 points (8 GP + 10 linked-GP nodes, n = 5000)."""
    args, D = ctx.args, ctx.D
    S, M = args.predict_imputations, args.predict_points
    n = args.n
    emu, t_emu, _ = ctx.timed(lambda: D.emulator(model.estimate(), N=S))
    xt = np.random.default_rng(SEED + 99).uniform(0, 1, size=(M, 8))
    ctx.predict(emu, xt[: min(M, 256 * ctx.world)])                     # sizes the scratch buffers
    M_stated, M = M, ctx.bounded_points(emu, xt, M)
    xt = xt[:M]
    (mu, var), t, t_wall = ctx.timed(lambda: ctx.predict(emu, xt))
    entries = float(M) * S * 10 * (n * (n + 1) / 2)                     # J entries of the 10 linked nodes
    leg = {"workload": f"config 3 predict: emulator(N={S}).predict, {M} points, n={n}, 8 gp + 10 link_gp nodes",
           "points": M, "points_stated": M_stated,
           "points_per_s": M / t, "e2e_points_per_s": M / t_wall, "seconds": t, "emulator_build_s": t_emu,
           "node_imputation_points_per_s": M * S * 18 / t,
           "J_entries_per_s": entries / t,
           "fp64_pipe_frac": entries * 98.0 / t / 1e12 / (ctx.peak * ctx.world) if ctx.peak else None,
           "fp64_pipe_frac_what": "J entries x 49 FMA (sexp, Dw=Dz=8) / time / (measured DGEMM peak x GPUs): DFMA and "
                                  "DMMA share the FP64 datapath",
           "sharding": "test points / rank, one all-gather of (mu, var) on the device",
           "finite": bool(np.all(np.isfinite(mu)) and np.all(np.isfinite(var)))}
    ref = ctx.cpu_reference()
    if ref is not None:
        R, dgpsi = ref
        # one linked node and one first-layer node of the frozen imputation 0, a handful of points, through the
        # reference's own kernel methods (its compute_stats at n = 5000 takes ~10 s per node)
        Mc = 4
        k2, k1 = emu.all_layer_set[0][1][0], emu.all_layer_set[0][0][0]
        r2, r1 = R.ref_kernel(dgpsi, k2), R.ref_kernel(dgpsi, k1)
        t0 = time.perf_counter()
        r2.compute_stats()
        r1.compute_stats()
        t_stats = time.perf_counter() - t0
        m_in = np.random.default_rng(5).uniform(0, 1, (Mc, 8))
        v_in = np.random.default_rng(6).uniform(1e-4, .05, (Mc, 8))
        r2.linkgp_prediction(m_in[:1], v_in[:1], xt[:1])        # JIT
        t0 = time.perf_counter()
        mr, vr = r2.linkgp_prediction(m_in, v_in, xt[:Mc])
        t_link = (time.perf_counter() - t0) / Mc
        r1.gp_prediction(xt[:1], None)
        t0 = time.perf_counter()
        m1r, v1r = r1.gp_prediction(xt[:64], None)
        t_gp = (time.perf_counter() - t0) / 64
        mg, vg = k2.linkgp_prediction(m_in, v_in, xt[:Mc])
        m1g, v1g = k1.gp_prediction(xt[:64], None)
        per_point = S * (10 * t_link + 8 * t_gp)
        leg["cpu_reference"] = {"points_per_s": 1.0 / per_point, "kind": "reference", "cores": cpu_cores(),
                                "sample": f"reference kernel.linkgp_prediction on {Mc} points ({t_link:.2f}s/point/node) and "
                                          f"gp_prediction on 64 points ({1e3 * t_gp:.2f} ms/point/node) of imputation 0 "
                                          f"(compute_stats {t_stats:.1f}s for 2 nodes, not counted); per point = "
                                          f"{S} imputations x (10 linked + 8 gp nodes)",
                                "gpu_vs_reference": {"link_mean": _err(mg, mr), "link_var": _err(vg, vr),
                                                        "gp_mean": _err(m1g, m1r), "gp_var": _err(v1g, v1r)}}
    return leg


def leg_cfg4(ctx):
    """BASELINE config 4: Vecchia DGP, n=100k, d=10, m=25, 10+1 nodes; `emulator(N=10).predict` of 1M points."""
    args, D = ctx.args, ctx.D
    ctx.share_chain(True)
    seed4 = SEED0 + 3
    rng = ctx.seed(seed4)
    n4, d4, S4, M4 = args.vecchia_n, 10, args.vecchia_imputations, args.vecchia_points
    X4 = rng.uniform(0, 1, (n4, d4))
    Y4 = (f_config4(X4) + 0.05 * rng.standard_normal(n4)).reshape(-1, 1)
    l1 = [D.kernel(length=np.array([1.0]), name="sexp") for _ in range(10)]
    l2 = [D.kernel(length=np.array([1.0]), name="sexp", scale_est=True, nugget_est=True, nugget=1e-2,
                   connect=np.arange(10))]
    m4, t_build, _ = ctx.timed(lambda: D.dgp(X4, Y4, D.combine(l1, l2), vecchia=True, m=25))
    _, t_train, _ = ctx.timed(lambda: m4.train(args.vecchia_train_iters, disable=True))
    emu4, t_emu, _ = ctx.timed(lambda: D.emulator(m4.estimate(burnin=0), N=S4))
    xt4 = np.random.default_rng(seed4 + 99).uniform(0, 1, size=(M4, d4))
    ctx.predict(emu4, xt4, m=25)       # first call at a new size re-sizes scratch slots (cudaMalloc, synchronising)
    (mu4, var4), t, t_wall = ctx.timed(lambda: ctx.predict(emu4, xt4, m=25))
    pairs = float(M4) * n4 * (1 + S4)   # kNN candidate pairs: layer 1 once (shared length-scale), layer 2 per imputation
    leg = {"workload": f"config 4: Vecchia DGP n={n4}, d=10, m=25, 10+1 sexp nodes; emulator(N={S4}).predict(m=25), "
                       f"{M4} points, kNN included",
           "points_per_s": M4 / t, "e2e_points_per_s": M4 / t_wall, "seconds": t,
           "node_imputation_points_per_s": M4 * S4 * 11 / t, "knn_pairs_per_s_lower_bound": pairs / t,
           "train_s_per_iter": t_train / max(1, args.vecchia_train_iters), "train_iters": args.vecchia_train_iters,
           "construct_s": t_build, "emulator_build_s": t_emu,
           "sharding": "test points / rank, one all-gather of (mu, var) on the device",
           "rmse_vs_truth": float(np.sqrt(np.mean((mu4[:, 0] - f_config4(xt4)) ** 2))),
           "finite": bool(np.all(np.isfinite(mu4)) and np.all(np.isfinite(var4)))}
    ref = ctx.cpu_reference()
    if ref is not None:
        R, dgpsi = ref
        Mc, Sc = 2000, 1
        remu = R.ref_emulator(dgpsi, emu4, Sc)
        remu.predict(xt4[:50], m=25)     # JIT
        t0 = time.perf_counter()
        mr, vr = remu.predict(xt4[:Mc], m=25)
        tc = time.perf_counter() - t0
        sub = D.emulator.__new__(D.emulator)
        sub.all_layer, sub.n_layer, sub.vecch = emu4.all_layer_set[0], emu4.n_layer, True
        sub.all_layer_set = emu4.all_layer_set[:Sc]
        mg, vg = sub.predict(xt4[:Mc], m=25)
        leg["cpu_reference"] = {"points_per_s": Mc / tc * Sc / S4, "kind": "reference", "cores": cpu_cores(),
                                "sample": f"reference emulator.predict(m=25) on {Mc} points x {Sc} of the {S4} frozen "
                                          f"imputations: {tc:.1f}s (sklearn kd-tree kNN; scaled linearly to {S4} imputations)",
                                "gpu_vs_reference": {"mean": _err(mg, mr), "var": _err(vg, vr)}}
    return leg


def leg_cfg2(ctx):
    """BASELINE config 2: 2-layer Matern-2.5 DGP, n=2000, d=5; `emulator(N=10).predict` of 100k points."""
    args, D = ctx.args, ctx.D
    ctx.share_chain(True)
    rng = ctx.seed(SEED0 + 1)
    n, d, S, M = 2000, 5, args.cfg2_imputations, args.cfg2_points
    X = rng.uniform(0, 1, (n, d))
    Y = f_config2(X).reshape(-1, 1)
    l1 = [D.kernel(length=np.array([1.]), name='matern2.5') for _ in range(5)]
    l2 = [D.kernel(length=np.array([1.]), name='matern2.5', scale_est=True, connect=np.arange(5))]
    model = D.dgp(X, Y, D.combine(l1, l2))
    model.train(3, disable=True)
    iters = args.cfg2_train_iters
    _, t_train, _ = ctx.timed(lambda: model.train(iters, disable=True))
    emu, t_emu, _ = ctx.timed(lambda: D.emulator(model.estimate(), N=S))
    xt = rng.uniform(0, 1, (M, d))
    ctx.predict(emu, xt[: min(M, 256 * ctx.world)])
    M_stated, M = M, ctx.bounded_points(emu, xt, M)
    xt = xt[:M]
    (mu, var), t, t_wall = ctx.timed(lambda: ctx.predict(emu, xt))
    evals = float(M) * S * (n * (n + 1) / 2) * 5          # Jd evaluations: pairs x Dw
    leg = {"workload": f"config 2: 2-layer Matern-2.5 DGP (5 + 1 nodes, global connection), n={n}, d=5; train({iters}) then "
                       f"emulator(N={S}).predict, {M} points",
           "points": M, "points_stated": M_stated,
           "points_per_s": M / t, "e2e_points_per_s": M / t_wall, "seconds": t,
           "train_iters_per_s": iters / t_train, "emulator_build_s": t_emu,
           "node_imputation_points_per_s": M * S * 6 / t,
           "J_entries_per_s": float(M) * S * (n * (n + 1) / 2) / t, "Jd_evaluations_per_s": evals / t,
           "fp64_pipe_frac": evals * 220.0 / t / 1e12 / (ctx.peak * ctx.world) if ctx.peak else None,
           "fp64_pipe_frac_what": "Jd evaluations x ~110 FMA (tabulated Matern kernel) / time / (measured DGEMM peak x GPUs)",
           "sharding": "test points / rank, one all-gather of (mu, var) on the device",
           "rmse_vs_truth": float(np.sqrt(np.mean((mu[:, 0] - f_config2(xt)) ** 2))),
           "finite": bool(np.all(np.isfinite(mu)) and np.all(np.isfinite(var)))}
    ref = ctx.cpu_reference()
    if ref is not None:
        R, dgpsi = ref
        Mc, Sc = 24, 1
        t0 = time.perf_counter()
        remu = R.ref_emulator(dgpsi, emu, Sc)
        t_stats = time.perf_counter() - t0
        remu.predict(xt[:2])
        t0 = time.perf_counter()
        mr, vr = remu.predict(xt[:Mc])
        tc = time.perf_counter() - t0
        sub = D.emulator.__new__(D.emulator)
        sub.all_layer, sub.n_layer, sub.vecch = emu.all_layer_set[0], emu.n_layer, False
        sub.all_layer_set = emu.all_layer_set[:Sc]
        mg, vg = sub.predict(xt[:Mc])
        leg["cpu_reference"] = {"points_per_s": Mc / tc * Sc / S, "kind": "reference", "cores": cpu_cores(),
                                "sample": f"reference emulator.predict on {Mc} points x {Sc} of the {S} frozen imputations: "
                                          f"{tc:.1f}s (scaled linearly to {S} imputations; its compute_stats for 6 nodes "
                                          f"{t_stats:.1f}s not counted)",
                                "gpu_vs_reference": {"mean": _err(mg, mr), "var": _err(vg, vr)}}
    return leg


def leg_cfg5(ctx):
    """BASELINE config 5: linked GP -> 2-layer DGP -> GP (n=500 each), `lgp(N=50).predict` of 1M points."""
    args, D = ctx.args, ctx.D
    ctx.share_chain(False)     # n = 500: a wave is microseconds of work, sharing the chain would only add latency
    rng = ctx.seed(SEED0 + 4)
    n, S, M = 500, args.cfg5_imputations, args.cfg5_points
    X1 = rng.uniform(0, 1, (n, 2))
    Y1 = (np.sin(3 * X1[:, 0]) + X1[:, 1] ** 2).reshape(-1, 1)
    g1 = D.gp(X1, Y1, D.kernel(length=np.array([1., 1.]), name='matern2.5', scale_est=True))
    g1.train()
    X2 = rng.uniform(-0.2, 2.0, (n, 1))
    Y2 = np.tanh(2 * (X2 - 0.9))
    d2 = D.dgp(X2, Y2, D.combine([D.kernel(length=np.array([1.]), name='matern2.5')],
                                 [D.kernel(length=np.array([1.]), name='matern2.5', scale_est=True, connect=np.arange(1))]))
    d2.train(20, disable=True)
    X3 = rng.uniform(-1.1, 1.1, (n, 1))
    Y3 = X3 ** 2 - 0.3 * X3
    g3 = D.gp(X3, Y3, D.kernel(length=np.array([1.]), name='sexp', scale_est=True))
    g3.train()
    system, t_build, _ = ctx.timed(lambda: D.lgp(D.combine([D.container(g1.export(), np.array([0, 1]))],
                                                           [D.container(d2.estimate(), np.array([0]))],
                                                           [D.container(g3.export(), np.array([0]))]), N=S))
    xt = rng.uniform(0, 1, (M, 2))
    ctx.predict(system, xt[: min(M, 256 * ctx.world)])
    M_stated, M = M, ctx.bounded_points(system, xt, M)
    xt = xt[:M]
    (mu, var), t, t_wall = ctx.timed(lambda: ctx.predict(system, xt))
    truth = np.tanh(2 * ((np.sin(3 * xt[:, 0]) + xt[:, 1] ** 2) - 0.9))
    truth = truth ** 2 - 0.3 * truth
    pairs = n * (n + 1) / 2
    leg = {"workload": f"config 5: linked GP(2-D, Matern) -> 2-layer DGP (Matern) -> GP (sexp), n={n} each; "
                       f"lgp(N={S}).predict, {M} points",
           "points": M, "points_stated": M_stated,
           "points_per_s": M / t, "e2e_points_per_s": M / t_wall, "seconds": t, "lgp_build_s": t_build,
           "emulator_imputation_points_per_s": M * S * 3 / t,
           "J_entries_per_s": float(M) * S * 3 * pairs / t,
           "fp64_pipe_frac": float(M) * S * pairs * (3 * 220.0 + 98.0) / t / 1e12 / (ctx.peak * ctx.world) if ctx.peak else None,
           "fp64_pipe_frac_what": "per point and imputation: 3 Matern Jd dimensions x ~110 FMA + one sexp J x 49 FMA per "
                                  "pair / time / (measured DGEMM peak x GPUs)",
           "sharding": "test points / rank, one all-gather of (mu, var) on the device",
           "rmse_vs_truth": float(np.sqrt(np.mean((mu[0][:, 0] - truth) ** 2))),
           "finite": bool(np.all(np.isfinite(mu[0])) and np.all(np.isfinite(var[0])))}
    ref = ctx.cpu_reference()
    if ref is not None:
        R, dgpsi = ref
        Mc, Sc = 200, 2
        rs = R.ref_lgp(dgpsi, system, Sc)
        rs.predict(xt[:4])
        t0 = time.perf_counter()
        mr, vr = rs.predict(xt[:Mc])
        tc = time.perf_counter() - t0
        sub = D.lgp.__new__(D.lgp)
        sub.L, sub.num_model = system.L, system.num_model
        sub.all_layer_set = system.all_layer_set[:Sc]
        sub.all_layer = sub.all_layer_set[0]
        mg, vg = sub.predict(xt[:Mc])
        leg["cpu_reference"] = {"points_per_s": Mc / tc * Sc / S, "kind": "reference", "cores": cpu_cores(),
                                "sample": f"reference lgp.predict on {Mc} points x {Sc} of the {S} frozen imputation sets: "
                                          f"{tc:.1f}s (scaled linearly to {S})",
                                "gpu_vs_reference": {"mean": _err(mg[0], mr[0]), "var": _err(vg[0], vr[0])}}
    return leg


def leg_cfg1(ctx):
    """BASELINE config 1 (demo/step_fct.ipynb): 3 GP layers of one sexp node on the 1-D step function, n=10,
    train(500) and `emulator(N=10).predict` of 300 points."""
    args, D = ctx.args, ctx.D
    ctx.share_chain(False)
    iters = args.cfg1_iters

    def build(mod, kernel):
        X = np.linspace(0, 1, 10)[:, None]
        Y = np.array([[-1.0] if i < 0.5 else [1.0] for i in X[:, 0]])
        layers = [[kernel(length=np.array([1.0]), name='sexp')], [kernel(length=np.array([1.0]), name='sexp')],
                  [kernel(length=np.array([1.0]), name='sexp', scale_est=True)]]
        return mod.dgp(X, Y, mod.combine(*layers))

    ctx.seed(SEED0)
    model = build(D, D.kernel)
    model.train(10, disable=True)
    _, t_train, t_wall = ctx.timed(lambda: model.train(iters, disable=True))
    emu = D.emulator(model.estimate(), N=10)
    xt = np.linspace(0, 1, 300)[:, None]
    emu.predict(xt)
    (mu, var), t_pred, _ = ctx.timed(lambda: emu.predict(xt))
    leg = {"workload": f"config 1: demo/step_fct.ipynb, 3 GP layers x 1 sexp node, n=10; train({iters}) and "
                       f"emulator(N=10).predict(300 points); every rank runs it (n = 10 does not shard)",
           "train_iters_per_s": iters / t_train, "e2e_train_iters_per_s": iters / t_wall,
           "predict_points_per_s": 300 / t_pred, "notebook_iters_per_s_unknown_hardware": 24.75,
           "finite": bool(np.all(np.isfinite(mu)) and np.all(np.isfinite(var)))}
    ref = ctx.cpu_reference()
    if ref is not None:
        R, dgpsi = ref
        from dgpsi.kernel_class import kernel as rkernel
        np.random.seed(SEED0)
        dgpsi.nb_seed(SEED0)
        rm = build(dgpsi, rkernel)
        rm.train(N=10, disable=True)
        t0 = time.perf_counter()
        rm.train(N=iters, disable=True)
        tc = time.perf_counter() - t0
        remu = dgpsi.emulator(rm.estimate(), N=10)
        remu.predict(xt)
        t0 = time.perf_counter()
        remu.predict(xt)
        tp = time.perf_counter() - t0
        leg["cpu_reference"] = {"train_iters_per_s": iters / tc, "predict_points_per_s": 300 / tp, "kind": "reference",
                                "cores": cpu_cores(), "sample": f"the full run: reference train({iters}) {tc:.1f}s, "
                                                                f"predict 300 points {tp:.3f}s"}
    return leg


LEGS = {"predict3": None, "cfg4": leg_cfg4, "cfg2": leg_cfg2, "cfg5": leg_cfg5, "cfg1": leg_cfg1}


def run_gpu(args):
    ctx = Ctx(args)
    torch, L, lib, rank, world = ctx.torch, ctx.L, ctx.lib, ctx.rank, ctx.world
    ctx.peak = measure_fp64_peak(torch)
    if ctx.dist is not None:   # every rank uses rank 0's figure
        pk = torch.tensor([ctx.peak], dtype=torch.float64, device="cuda")
        ctx.dist.broadcast(pk, 0)
        ctx.peak = float(pk[0])
    peak = ctx.peak

    tr = leg_train(ctx)
    model = tr["model"]
    legs = {}
    wanted = [w for w in args.legs.split(",") if w]
    for name in wanted:
        if name not in LEGS:
            raise SystemExit(f"unknown leg {name!r}; choose from {sorted(LEGS)}")
        t0 = time.perf_counter()
        try:
            legs[name] = leg_predict3(ctx, model) if name == "predict3" else LEGS[name](ctx)
        except Exception as exc:  # a secondary leg must not take the headline down with it
            if world > 1:
                raise
            import traceback
            legs[name] = {"error": f"{type(exc).__name__}: {exc}", "traceback": traceback.format_exc()[-1500:]}
        legs[name]["leg_wall_s"] = time.perf_counter() - t0
        torch.cuda.empty_cache()

    if rank == 0:
        steps = args.steps
        dev_s, wall_s = tr["dev_s"], tr["wall_s"]
        value = steps / dev_s                      # one chain, whatever the number of GPUs
        e2e = steps / wall_s
        upd_ms, upd_n, upd_flops = tr["upd"]
        achieved = (upd_flops / upd_n) / (upd_ms / upd_n * 1e-3) / 1e12 if upd_n > 0 else None
        roof = {"bound": "tensor",
                "kernel": "update_kernel, bulk launches (FP64 DMMA SYRK trailing update, K = 512 hyper-blocks / 128 tail)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                # DRAM bytes per launch: one `ncu --set full` capture of a K = 512 bulk launch (profiles/
                # r1s2_update_ncu_raw.txt: 770.1 MB read + written for 41.94 GFLOP; algorithmic 760 MB), scaled to
                # this run's average launch
                "traffic": (upd_flops / upd_n) * (770.08e6 / 41.94e9) if upd_n > 0 else None,
                "traffic_source": "profiles/r1s2_update_ncu_raw.txt (dram__bytes_read+write of one bulk launch, "
                                  "18.4 B per kFLOP) x this run's FLOPs per launch",
                "launches_timed": int(upd_n), "avg_launch_ms": upd_ms / upd_n if upd_n else None,
                "share_of_step": upd_ms / (dev_s * 1e3 * world) if dev_s else None,
                "peak_source": "cuBLAS DGEMM 8192^3 via torch.matmul(float64), best of 5, measured in this run "
                               "(MEASURED_PEAKS.json has no FP64 entry)"}
        probe = L.host_doubles(2)
        iso = None
        if lib.dgpb_probe_update(L.workspace(), 5000, 8, 16 << 8, 5, probe) == 0:
            iso = {"achieved": probe[1], "frac": probe[1] / peak if peak else None, "ms_per_launch": probe[0],
                   "what": "K = 512 bulk launch alone, window 4544, 8 matrices (dgpb_probe_update)"}
        roof["isolated"] = iso
        roof["note"] = ("bulk launches run on the lowest-priority stream and yield SM slots to the critical-path kernels "
                        "(panels, inner and look-ahead updates on a highest-priority stream), so their in-situ event "
                        "durations include time spent executing those kernels: `isolated` is the same launch alone, "
                        "`step` is the factorisation FLOPs of the timed region over its duration")
        issued, wasted = tr["issued"], tr["wasted"]
        useful = issued - wasted
        agg_peak = peak * world
        roof["step"] = {"achieved": useful / dev_s / 1e12, "frac": useful / dev_s / 1e12 / agg_peak if agg_peak else None,
                        "flops_useful_per_step": useful / steps, "flops_issued_per_step": issued / steps,
                        "issued_frac": issued / dev_s / 1e12 / agg_peak if agg_peak else None,
                        "what": "USEFUL factorisation FLOPs (n^3/3 per Cholesky the reference's schedule needs -- prior "
                                "draws, thresholds, every candidate up to the accepted one -- and n^3 per gradient "
                                "evaluation) / duration / (peak x GPUs); `issued` also counts the speculative candidates "
                                "of a wave that lay behind the accepted one"}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            arm = ReferenceArm(args.n)
            r2 = np.random.default_rng(1)
            arm.sample(1000, r2)
            a, b, c = arm.sample(args.n, r2)
            cpu = {"value": 1.0 / arm.iteration_seconds(a, b, c), "unit": "iters/s", "cores": cpu_cores(),
                   "kind": arm.kind, "sample": arm.describe(a, b, c, 1)}
        tim, tim0 = model.timing, tr["timing0"]
        line = {"metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": world, "steps": steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / steps, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, world),
                "e2e": {"value": e2e, "unit": "iters/s", "h2d_bytes_per_step": tr["h2d"] // steps,
                        "d2h_bytes_per_step": tr["d2h"] // steps},
                "gpu_launches": tr["launches"], "ess_proposals_per_step": tr["nprop"] / steps,
                "phase_ms_per_step": {"i_step": 1e3 * (tim["i_step"] - tim0["i_step"]) / steps,
                                      "m_step": 1e3 * (tim["m_step"] - tim0["m_step"]) / steps,
                                      "m_step_in_library": 1e3 * (tim.get("m_batched_s", 0.0) - tim0.get("m_batched_s", 0.0)) / steps,
                                      "m_step_between_rounds": 1e3 * (tim.get("m_round_gap_s", 0.0) - tim0.get("m_round_gap_s", 0.0)) / steps,
                                      "m_step_before_first_round": 1e3 * (tim.get("m_first_round_s", 0.0) - tim0.get("m_first_round_s", 0.0)) / steps,
                                      "m_step_rounds": 1e3 * (tim.get("m_rounds_s", 0.0) - tim0.get("m_rounds_s", 0.0)) / steps,
                                      "m_step_after_last_round": 1e3 * (tim.get("m_tail_s", 0.0) - tim0.get("m_tail_s", 0.0)) / steps,
                                      "m_step_batched_calls": (tim.get("m_batched_calls", 0) - tim0.get("m_batched_calls", 0)) / steps,
                                      "m_step_matrices": (tim.get("m_batched_matrices", 0) - tim0.get("m_batched_matrices", 0)) / steps},
                "multi_gpu": None if world == 1 else {
                    "what": "one chain shared by the ranks: ESS wave candidates dealt over the GPUs (all-gather of <= 1 KB "
                            "per rank and wave, broadcast of the prior draws), M-step nodes dealt over the GPUs (one "
                            "all-reduce of 36 doubles per node)",
                    "limiting_collective": "ncclAllGather of the per-matrix results, once per ESS wave (latency-bound)"},
                "roofline": roof, "cpu_baseline": cpu, "clocks": tr["clocks"], "legs": legs}
        print(json.dumps(line), flush=True)
    if ctx.dist is not None:
        ctx.share_chain(False)
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=5000, help="training points (BASELINE config 3: 5000)")
    ap.add_argument("--legs", default=DEFAULT_LEGS,
                    help="secondary legs, comma separated (predict3, cfg4, cfg2, cfg5, cfg1); '' = headline only")
    ap.add_argument("--predict-points", type=int, default=10000, help="config 3 predict: test points (10 000)")
    ap.add_argument("--predict-imputations", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--vecchia-points", type=int, default=1000000, help="config 4: test points (1M)")
    ap.add_argument("--vecchia-n", type=int, default=100000)
    ap.add_argument("--vecchia-imputations", type=int, default=10)
    ap.add_argument("--vecchia-train-iters", type=int, default=2)
    ap.add_argument("--cfg2-points", type=int, default=100000)
    ap.add_argument("--cfg2-imputations", type=int, default=10)
    ap.add_argument("--cfg2-train-iters", type=int, default=10)
    ap.add_argument("--cfg5-points", type=int, default=1000000)
    ap.add_argument("--cfg5-imputations", type=int, default=50)
    ap.add_argument("--cfg1-iters", type=int, default=500)
    ap.add_argument("--leg-seconds", type=float, default=75.0,
                    help="time budget of the timed predict call of the predict3 / cfg2 / cfg5 legs: when the stated "
                         "number of test points would take longer, as many as fit are used (same model, same "
                         "imputations; reported as `points` next to `points_stated`); 0 = always the stated sizes")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
