"""Vecchia DGP prediction workload (BASELINE config 4 shape, reduced M / S) for launch-list profiling:
  python scripts/prof_vecchia_predict.py [M] [S] [n]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgp_b200 as D
M = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
d = 10
seed = 20261017 + 3
rng = np.random.default_rng(seed); np.random.seed(seed); D.nb_seed(seed)
X = rng.uniform(0, 1, (n, d))
f = lambda x: np.sin(2*np.pi*x[:,0]*x[:,1]) + x[:,2]**2 + np.cos(3*x[:,3:].sum(1))
Y = (f(X) + 0.05*rng.standard_normal(n)).reshape(-1, 1)
l1 = [D.kernel(length=np.array([1.]), name='sexp') for _ in range(10)]
l2 = [D.kernel(length=np.array([1.]), name='sexp', scale_est=True, nugget_est=True, nugget=1e-2, connect=np.arange(10))]
m = D.dgp(X, Y, D.combine(l1, l2), vecchia=True, m=25)
m.train(1, disable=True)
emu = D.emulator(m.estimate(burnin=0), N=S)
xt = rng.uniform(0, 1, (M, d))
torch.cuda.synchronize(); t = time.perf_counter()
mu, var = emu.predict(xt, m=25)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print('predict %d pts x %d imputations (m=25): %.3fs -> %.1f pts/s' % (M, S, dt, M/dt), 'finite', np.isfinite(mu).all() and np.isfinite(var).all())
