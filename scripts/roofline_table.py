"""Per-subsystem throughput at BASELINE sizes against the roofline that bounds each, timed with CUDA
events through the public node API / C-ABI.  Prints a markdown table:  python scripts/roofline_table.py > profiles/x.md"""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgp_b200 as D
from dgp_b200 import _lib as L
from dgp_b200 import vecchia as V

lib = L.load(); L.device()
rng = np.random.default_rng(7)
HBM = 6537.3  # GB/s, MEASURED_PEAKS.json


def fp64_peak():
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    best = 0.0
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        if i: best = max(best, 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


PEAK = fp64_peak()
rows = []
def row(name, work, unit, t, bound, achieved, peak, punit):
    rows.append((name, work, f"{t*1e3:.2f} ms", bound, f"{achieved:.1f} {punit}", f"{100*achieved/peak:.0f} %"))

def node(n, Dl, Dg, name="sexp", ard=False, **kw):
    k = D.kernel(length=np.full(Dl + Dg if ard else 1, 0.8), name=name, nugget=1e-4, scale=1.2,
                 connect=np.arange(Dg) if Dg else None, **kw)
    k.input = rng.uniform(0, 1, (n, Dl)); k.input_dim = np.arange(Dl)
    if Dg: k.global_input = rng.uniform(0, 1, (n, Dg))
    k.output = np.sin(3 * k.input.sum(1, keepdims=True)); k.D = Dl + Dg
    k.para_path = np.atleast_2d(np.concatenate((k.scale, k.length, k.nugget)))
    return k

# 1. kernel matrix (device buffers; HBM-write bound)
n, Dd = 5000, 16
Xd = L.to_dev(rng.uniform(0, 1, (n, Dd))); Kd = L.empty((n, n)); larr, lptr = L.length_host(np.array([0.8]))
t = timeit(lambda: L.check(lib.dgpb_kmatrix(L.ptr(Xd), n, Dd, lptr, 1, 1e-6, None, 0, 0, L.ptr(Kd), None, L.stream())), 10)
row("k_matrix sexp, n=5000, D=16 (full symmetric K materialised)", "8n^2 B written", "", t, "HBM", 8 * n * n / t / 1e9, HBM, "GB/s")
# 2. dense likelihood / gradient / wave
out = L.host_doubles(2)
for B, aug, label in ((1, 0, "log-likelihood, 1 node"), (8, 0, "ESS wave, 8 matrices"), (16, 0, "ESS wave, 16 matrices"), (1, 1, "nllik + gradient (K^-1 by sliding window)")):
    L.check(lib.dgpb_probe_factorize(L.workspace(), 5000, B, aug, 3, out))
    row(f"{label}, n=5000 (assemble + factorise + reduce)", "n^3/3" if not aug else "n^3", "", out[0] * 1e-3, "FP64 tensor", out[1], PEAK, "TFLOP/s")
k = node(5000, 8, 8)
t = timeit(lambda: k.llik(k.log_t().copy()), 2)
row("kernel.llik end to end (upload, K, K^-1, fused gradient, readback), n=5000, D=16", "n^3", "", t, "FP64 tensor", 5000 ** 3 / t / 1e12, PEAK, "TFLOP/s")
# 3. kNN
Mq, nn_, Dk = 200000, 100000, 10
xq, xw = L.to_dev(rng.uniform(0, 1, (Mq, Dk))), L.to_dev(rng.uniform(0, 1, (nn_, Dk)))
t = timeit(lambda: V.get_pred_nn_dev(xq, xw, 25), 2)
row("get_pred_nn, 200k queries x 100k points, D=10, m=25 (screen + exact ranking)", "2D flop/pair", "", t, "FP64 tensor", Mq * nn_ * 2 * Dk / t / 1e12, PEAK, "TFLOP/s")
xw20, xq20 = L.to_dev(rng.uniform(0, 1, (nn_, 20))), L.to_dev(rng.uniform(0, 1, (Mq, 20)))
t = timeit(lambda: V.get_pred_nn_dev(xq20, xw20, 25), 2)
row("get_pred_nn, 200k x 100k, D=20, m=25", "2D flop/pair", "", t, "FP64 tensor", Mq * nn_ * 2 * 20 / t / 1e12, PEAK, "TFLOP/s")
# 4. Vecchia blocks
Xv = rng.uniform(0, 1, (nn_, Dk)); yv = np.sin(3 * Xv.sum(1))
NNv = V.nn(Xv / 0.8, 25)
Xvd, yvd, NNd, _ = V._prep(Xv, yv, NNv, None)
o1 = L.host_doubles(1)
b = 26.0; blk = b ** 3 / 3 + b * b * (3 * Dk + 25) / 2 + 2 * b * b
t = timeit(lambda: L.check(lib.dgpb_vecchia_llik(L.ptr(Xvd), L.ptr(yvd), L.ptr(NNd), nn_, Dk, 26, lptr, 1, 1.0, 1e-4, None, 0, o1, L.stream())), 5)
row("vecchia_llik, n=100k, m=25, D=10", "b^3/3 + b^2(3D+25)/2 + 2b^2 per block", "", t, "FP64 vector", nn_ * blk / t / 1e12, PEAK, "TFLOP/s")
kv = node(nn_, Dk, 0); kv.vecch, kv.pred_m = True, 25
xt = rng.uniform(0, 1, (Mq, Dk))
with L.predict_cache():
    xtd = L.to_dev(xt); kv._gp_prediction_dev(xtd, None)
    t = timeit(lambda: kv._gp_prediction_dev(xtd, None), 3)   # neighbour search cached: block kernel only
row("gp_vecch, 200k test points, m=25, D=10 (blocks only)", "same per block", "", t, "FP64 vector", Mq * blk / t / 1e12, PEAK, "TFLOP/s")
# 5. dense predictions
kd = node(5000, 8, 0); kd.compute_stats()
xt5 = rng.uniform(0, 1, (10000, 8))
t = timeit(lambda: kd.gp_prediction(xt5, None), 2)
row("gp, 10k test points, n=5000, D=8 (R^-1 form, reference's)", "2 n^2 flop/pt", "", t, "FP64 tensor", 10000 * 2 * 5000.0 ** 2 / t / 1e12, PEAK, "TFLOP/s")
kl = node(5000, 8, 8); kl.compute_stats()
Ml = 512
m_in, v_in, z = rng.uniform(0, 1, (Ml, 8)), rng.uniform(1e-4, 0.02, (Ml, 8)), rng.uniform(0, 1, (Ml, 8))
t = timeit(lambda: kl.linkgp_prediction(m_in, v_in, z), 2)
ent = Ml * 5000 * 5001 / 2
row("link_gp sexp, 512 test points, n=5000, Dw=8, Dz=8", "n(n+1)/2 J entries/pt, 49 FP64 FMA each (28 tensor + exp + 2)", "", t, "FP64 FMA", ent * 49 * 2 / t / 1e12, PEAK, "TFLOP/s")
km = node(2000, 5, 0, name="matern2.5"); km.compute_stats()
Mm = 128
m_in, v_in = rng.uniform(0, 1, (Mm, 5)), rng.uniform(1e-4, 0.02, (Mm, 5))
t = timeit(lambda: km.linkgp_prediction(m_in, v_in, None), 2)
ent = Mm * 2000 * 2001 / 2 * 5
row("link_gp Matern-2.5, 128 test points, n=2000, Dw=5", "n(n+1)/2 x Dw Jd evaluations/pt (~400 flop each)", "", t, "FP64 vector", ent * 400 / t / 1e12, PEAK, "TFLOP/s")

print(f"# Sub-system throughput vs roofline (1 x B200; FP64 peak = cuBLAS DGEMM 8192^3 = {PEAK:.1f} TFLOP/s measured in this run; HBM = {HBM} GB/s from MEASURED_PEAKS.json)\n")
print("| sub-system / call | algorithmic work | time | bound | achieved | of peak |")
print("|---|---|---|---|---|---|")
for r in rows: print("| " + " | ".join(r) + " |")
