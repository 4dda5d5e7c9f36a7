"""Run the BASELINE.json configurations other than the bench workload at (scaled) full size and print timings.
  python scripts/run_configs.py 2|4|5 [scale]
scale < 1 shrinks the number of test points / iterations, never the training size."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgp_b200 as D

cfg = int(sys.argv[1]); scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
seed = 20261017 + cfg - 1
rng = np.random.default_rng(seed); np.random.seed(seed); D.nb_seed(seed)

def tic(): torch.cuda.synchronize(); return time.perf_counter()

if cfg == 1:
    X = np.linspace(0, 1, 10)[:, None]
    Y = np.array([[-1.0] if i < 0.5 else [1.0] for i in X[:, 0]])
    layers = [[D.kernel(length=np.array([1.0]), name='sexp')], [D.kernel(length=np.array([1.0]), name='sexp')],
              [D.kernel(length=np.array([1.0]), name='sexp', scale_est=True)]]
    t = tic(); m = D.dgp(X, Y, D.combine(*layers)); print('construct %.2fs' % (tic()-t))
    iters = max(10, int(500*scale)); t = tic(); m.train(iters, disable=True); dt = tic()-t
    print('train %d iters: %.2f s = %.2f it/s (%d proposals, %d launches)' % (iters, dt, iters/dt, m.imp.n_proposals, D._lib.load().dgpb_launch_count()))
    t = tic(); emu = D.emulator(m.estimate(), N=10); print('emulator(N=10) %.2fs' % (tic()-t))
    xt = np.linspace(0, 1, 300)[:, None]
    t = tic(); mu, var = emu.predict(xt); dt = tic()-t
    print('predict 300 pts: %.4fs' % dt)
elif cfg == 2:
    n, d = 2000, 5
    X = rng.uniform(0, 1, (n, d))
    Y = (np.sin(2*np.pi*X[:,0]*X[:,1]) + (X[:,2]-0.5)**2 + X[:,3]*np.exp(-X[:,4])).reshape(-1, 1)
    l1 = [D.kernel(length=np.array([1.]), name='matern2.5') for _ in range(5)]
    l2 = [D.kernel(length=np.array([1.]), name='matern2.5', scale_est=True, connect=np.arange(5))]
    t = tic(); m = D.dgp(X, Y, D.combine(l1, l2)); print('construct %.2fs' % (tic()-t))
    iters = max(2, int(10*scale)); t = tic(); m.train(iters, disable=True); dt = tic()-t
    print('train %d iters: %.3f s/iter (%d proposals)' % (iters, dt/iters, m.imp.n_proposals))
    S = max(2, int(10*scale)); t = tic(); emu = D.emulator(m.estimate(), N=S); print('emulator(N=%d) %.2fs' % (S, tic()-t))
    M = int(100000*scale); xt = rng.uniform(0, 1, (M, d))
    t = tic(); mu, var = emu.predict(xt); dt = tic()-t
    print('predict %d pts x %d imputations: %.2fs -> %.1f pts/s' % (M, S, dt, M/dt), 'finite', np.isfinite(mu).all() and np.isfinite(var).all())
    rmse = np.sqrt(np.mean((mu[:,0] - (np.sin(2*np.pi*xt[:,0]*xt[:,1]) + (xt[:,2]-0.5)**2 + xt[:,3]*np.exp(-xt[:,4])))**2)); print('rmse', rmse)
elif cfg == 4:
    n, d = int(100000), 10
    X = rng.uniform(0, 1, (n, d))
    f = lambda x: np.sin(2*np.pi*x[:,0]*x[:,1]) + x[:,2]**2 + np.cos(3*x[:,3:].sum(1))
    Y = (f(X) + 0.05*rng.standard_normal(n)).reshape(-1, 1)
    l1 = [D.kernel(length=np.array([1.]), name='sexp') for _ in range(10)]
    l2 = [D.kernel(length=np.array([1.]), name='sexp', scale_est=True, nugget_est=True, nugget=1e-2, connect=np.arange(10))]
    t = tic(); m = D.dgp(X, Y, D.combine(l1, l2), vecchia=True, m=25); print('construct %.2fs' % (tic()-t))
    iters = max(2, int(20*scale)); t = tic(); m.train(iters, disable=True); dt = tic()-t
    print('train %d iters: %.3f s/iter' % (iters, dt/iters))
    S = max(2, int(10*scale)); t = tic(); emu = D.emulator(m.estimate(), N=S); print('emulator(N=%d) %.2fs' % (S, tic()-t))
    M = int(1000000*scale); xt = rng.uniform(0, 1, (M, d))
    t = tic(); mu, var = emu.predict(xt, m=25); dt = tic()-t
    print('predict %d pts x %d imputations (m=25): %.2fs -> %.1f pts/s' % (M, S, dt, M/dt), 'finite', np.isfinite(mu).all() and np.isfinite(var).all())
    print('rmse', np.sqrt(np.mean((mu[:,0]-f(xt))**2)))
elif cfg == 5:
    n = 500
    X1 = rng.uniform(0, 1, (n, 2)); Y1 = (np.sin(3*X1[:,0]) + X1[:,1]**2).reshape(-1, 1)
    g1 = D.gp(X1, Y1, D.kernel(length=np.array([1., 1.]), name='matern2.5', scale_est=True)); g1.train()
    X2 = rng.uniform(-0.2, 2.0, (n, 1)); Y2 = np.tanh(2*(X2-0.9))
    d2 = D.dgp(X2, Y2, D.combine([D.kernel(length=np.array([1.]), name='matern2.5')], [D.kernel(length=np.array([1.]), name='matern2.5', scale_est=True, connect=np.arange(1))]))
    d2.train(max(2, int(20*scale)), disable=True)
    X3 = rng.uniform(-1.1, 1.1, (n, 1)); Y3 = X3**2 - 0.3*X3
    g3 = D.gp(X3, Y3, D.kernel(length=np.array([1.]), name='sexp', scale_est=True)); g3.train()
    S = max(2, int(50*scale))
    t = tic(); sys_ = D.lgp(D.combine([D.container(g1.export(), np.array([0, 1]))], [D.container(d2.estimate(), np.array([0]))], [D.container(g3.export(), np.array([0]))]), N=S); print('lgp(N=%d) %.2fs' % (S, tic()-t))
    M = int(1000000*scale); xt = rng.uniform(0, 1, (M, 2))
    t = tic(); mu, var = sys_.predict(xt); dt = tic()-t
    print('lgp predict %d pts x %d imputations: %.2fs -> %.1f pts/s' % (M, S, dt, M/dt), 'finite', np.isfinite(mu[0]).all())
    truth = np.tanh(2*((np.sin(3*xt[:,0]) + xt[:,1]**2)-0.9)); truth = truth**2 - 0.3*truth
    print('rmse', np.sqrt(np.mean((mu[0][:,0]-truth)**2)))
