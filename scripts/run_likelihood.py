"""Likelihood-layer workloads at a size where the device kernels matter:
  python scripts/run_likelihood.py
(1) Hetero, dense, n = 1500: node-wise sweeps with the exact conditional draw of the mean (one shifted n^3
factorisation per sweep);  (2) Poisson under Vecchia, n = 20000, m = 25: ESS over a likelihood node with sparse prior
draws, prediction of 50k points."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgp_b200 as D

rng = np.random.default_rng(3); np.random.seed(3); D.nb_seed(3)
n, d = 1500, 2
X = rng.uniform(0, 1, (n, d))
mean = np.sin(4 * X[:, 0]) + X[:, 1]
sd = np.exp(-2.0 + 1.5 * X[:, 0])
Y = (mean + sd * rng.standard_normal(n)).reshape(-1, 1)
l1 = [D.kernel(length=np.array([0.5]), name='sexp') for _ in range(d)]
l2 = [D.kernel(length=np.array([0.5]), name='sexp', scale_est=True, connect=np.arange(d)) for _ in range(2)]
t = time.perf_counter()
m = D.dgp(X, Y, D.combine(l1, l2, [D.Hetero()]))
t1 = time.perf_counter()
m.train(3, disable=True)
torch.cuda.synchronize(); t2 = time.perf_counter()
emu = D.emulator(m.estimate(), N=2)
xt = rng.uniform(0, 1, (2000, d))
t3 = time.perf_counter()
mu, var = emu.predict(xt)
t4 = time.perf_counter()
truth = np.sin(4 * xt[:, 0]) + xt[:, 1]
print('Hetero n=%d: construct %.2fs, %.2f s/iter, predict 2000 pts %.2fs; rmse(mean) %.4f, mean predicted sd %.3f '
      '(true %.3f); finite %s' % (n, t1 - t, (t2 - t1) / 3, t4 - t3, np.sqrt(np.mean((mu[:, 0] - truth) ** 2)),
                                  np.mean(np.sqrt(var)), np.mean(np.exp(-2.0 + 1.5 * xt[:, 0])),
                                  bool(np.isfinite(mu).all() and np.isfinite(var).all())), flush=True)

n = 20000
X = rng.uniform(0, 1, (n, d))
rate = lambda x: np.exp(1.0 + np.sin(3 * x[:, 0]) + x[:, 1])
Y = rng.poisson(rate(X)).astype(float).reshape(-1, 1)
l1 = [D.kernel(length=np.array([0.5]), name='sexp') for _ in range(d)]
l2 = [D.kernel(length=np.array([0.5]), name='sexp', scale_est=True, connect=np.arange(d))]
t = time.perf_counter()
m = D.dgp(X, Y, D.combine(l1, l2, [D.Poisson()]), vecchia=True, m=25)
t1 = time.perf_counter()
m.train(2, disable=True)
torch.cuda.synchronize(); t2 = time.perf_counter()
emu = D.emulator(m.estimate(burnin=0), N=2)
xt = rng.uniform(0, 1, (50000, d))
t3 = time.perf_counter()
mu, var = emu.predict(xt, m=25)
t4 = time.perf_counter()
print('Poisson Vecchia n=%d: construct %.2fs, %.2f s/iter, predict 50k pts %.2fs; relative rmse(rate) %.3f; finite %s'
      % (n, t1 - t, (t2 - t1) / 2, t4 - t3, np.sqrt(np.mean((mu[:, 0] / rate(xt) - 1) ** 2)),
         bool(np.isfinite(mu).all() and np.isfinite(var).all())), flush=True)
