"""Vecchia DGP training (BASELINE config 4 shape) phase split: python scripts/prof_vecchia_train.py [iters] [n]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgp_b200 as D
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
d = 10
seed = 20261017 + 3
rng = np.random.default_rng(seed); np.random.seed(seed); D.nb_seed(seed)
X = rng.uniform(0, 1, (n, d))
f = lambda x: np.sin(2*np.pi*x[:,0]*x[:,1]) + x[:,2]**2 + np.cos(3*x[:,3:].sum(1))
Y = (f(X) + 0.05*rng.standard_normal(n)).reshape(-1, 1)
l1 = [D.kernel(length=np.array([1.]), name='sexp') for _ in range(10)]
l2 = [D.kernel(length=np.array([1.]), name='sexp', scale_est=True, nugget_est=True, nugget=1e-2, connect=np.arange(10))]
t = time.perf_counter(); m = D.dgp(X, Y, D.combine(l1, l2), vecchia=True, m=25); print('construct %.2fs' % (time.perf_counter() - t))
m.train(1, disable=True)
t0 = dict(m.timing); t = time.perf_counter()
m.train(iters, disable=True)
dt = time.perf_counter() - t
print('train %d iters: %.3f s/iter; I-step %.3f s, M-step %.3f s per iter; proposals/iter %.1f' % (
    iters, dt / iters, (m.timing['i_step'] - t0['i_step']) / iters, (m.timing['m_step'] - t0['m_step']) / iters,
    m.imp.n_proposals / (iters + 1 + 1)))
