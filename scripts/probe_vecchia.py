"""Component timings of the Vecchia prediction path at BASELINE config-4 size (n=100k, d=10, m=25)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgp_b200 as D
from dgp_b200 import _lib as L, vecchia as V
rng = np.random.default_rng(0)
n, d, M = 100000, 10, int(sys.argv[1]) if len(sys.argv) > 1 else 100000
X = rng.uniform(0, 1, (n, d)); y = np.sin(X.sum(1, keepdims=True))
def timed(name, f, reps=2):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): r = f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
    print(f"{name:34s} {dt*1e3:9.2f} ms"); return r
Xd = L.to_dev(X); xq = L.to_dev(rng.uniform(0, 1, (M, d)))
timed("nn ordered n=100k m=25", lambda: V.nn(X, 25), 1)
NN = timed(f"knn M={M} n=100k m=25", lambda: V.get_pred_nn_dev(xq, Xd, 25))
NN50 = timed(f"knn M={M} n=100k m=50", lambda: V.get_pred_nn_dev(xq, Xd, 50))
k = D.kernel(length=np.array([0.8]), name='sexp', nugget=1e-4)
k.input, k.output, k.vecch, k.m = X, y, True, 25
k.ord_nn()
for pm in (25, 50):
    k.pred_m = pm
    timed(f"gp_prediction vecch M={M} m={pm}", lambda: k._gp_prediction_dev(xq, None))
k2 = D.kernel(length=np.array([0.8]), name='sexp', nugget=1e-2, scale_est=True, nugget_est=True, connect=np.arange(d))
k2.input, k2.global_input, k2.output, k2.vecch, k2.m = X.copy(), X, y, True, 25
k2.ord_nn()
mq, vq = L.to_dev(rng.uniform(0, 1, (M, d))), L.to_dev(rng.uniform(1e-3, 1e-2, (M, d)))
for pm in (25, 50):
    k2.pred_m = pm
    timed(f"linkgp_prediction vecch M={M} m={pm} (Dw=10,Dz=10)", lambda: k2._linkgp_prediction_dev(mq, vq, xq))
timed("vecchia_llik n=100k", lambda: k.log_likelihood_func_vecch())
timed("llik_vecch (grad) n=100k", lambda: k2.llik_vecch(k2.log_t()))
timed("fmvn_sp n=100k", lambda: V.fmvn_sp(X[k.ord], k.NNarray, 1.0, k.length, 1e-4, 'sexp', z=rng.standard_normal(n)))
