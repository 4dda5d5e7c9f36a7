"""Repeat the Vecchia DGP prediction of bench.py's secondary metric and print each call's duration:
  python scripts/time_vecchia_predict.py [M] [S] [reps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgp_b200 as D
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n, d = 100000, 10
seed = 20261017 + 3
rng = np.random.default_rng(seed); np.random.seed(seed); D.nb_seed(seed)
X = rng.uniform(0, 1, (n, d))
f = lambda x: np.sin(2*np.pi*x[:,0]*x[:,1]) + x[:,2]**2 + np.cos(3*x[:,3:].sum(1))
Y = (f(X) + 0.05*rng.standard_normal(n)).reshape(-1, 1)
l1 = [D.kernel(length=np.array([1.]), name='sexp') for _ in range(10)]
l2 = [D.kernel(length=np.array([1.]), name='sexp', scale_est=True, nugget_est=True, nugget=1e-2, connect=np.arange(10))]
t = time.perf_counter()
m = D.dgp(X, Y, D.combine(l1, l2), vecchia=True, m=25)
m.train(2, disable=True)
torch.cuda.synchronize()
print('construct + 2 iterations: %.2fs' % (time.perf_counter() - t))
emu = D.emulator(m.estimate(burnin=0), N=S)
xt = np.random.default_rng(seed + 99).uniform(0, 1, (M, d))
emu.predict(xt[:256], m=25)
for r in range(reps):
    torch.cuda.synchronize(); t = time.perf_counter()
    mu, var = emu.predict(xt, m=25)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print('call %d: %d pts x %d imputations: %.3fs -> %.0f pts/s' % (r, M, S, dt, M / dt), flush=True)
for key in ('knn_mma=0', 'vecchia_small=0'):
    os.environ['DGPB_TUNE'] = key
    from dgp_b200 import _lib as L
    k, v = key.split('=')
    L.check(L.load().dgpb_tune(k.encode(), int(v)))
    torch.cuda.synchronize(); t = time.perf_counter()
    emu.predict(xt, m=25)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print('%s: %.3fs' % (key, dt), flush=True)
    L.check(L.load().dgpb_tune(k.encode(), 3 if k == 'knn_mma' else 1))
