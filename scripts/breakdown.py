"""Wall-clock split of one SEM iteration (I-step vs M-step per layer) at BASELINE config 3."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgp_b200 as D
from bench import SEED, layers_config3, make_config3
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rng = np.random.default_rng(SEED); np.random.seed(SEED); D.nb_seed(SEED)
X, Y = make_config3(n, rng)
t0 = time.perf_counter(); model = D.dgp(X, Y, layers_config3(lambda **kw: D.kernel(**kw))); torch.cuda.synchronize()
print("construct (11 sweeps): %.2fs" % (time.perf_counter() - t0))
model.train(1, disable=True)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); p0 = model.imp.n_proposals
    model.imp.sample(burnin=10); torch.cuda.synchronize(); t1 = time.perf_counter()
    evals = []
    tl = []
    for l, layer in enumerate(model.all_layer):
        ts = time.perf_counter(); ne = 0
        for k in layer:
            if l: k.r2()
            c0 = D._lib.load().dgpb_launch_count()
            k.maximise()
        torch.cuda.synchronize(); tl.append(time.perf_counter() - ts)
    print("iter %d: I-step %.2fs (%d proposals)  M-step layers (one node at a time) %s" % (it, t1 - t0, model.imp.n_proposals - p0, ["%.2f" % t for t in tl]))
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model.imp.sample(burnin=10); torch.cuda.synchronize(); t1 = time.perf_counter()
    model._m_step(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("iter %d: I-step %.2fs, threaded M-step %.2fs" % (it, t1 - t0, t2 - t1))
