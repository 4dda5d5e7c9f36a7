"""kNN screen throughput with and without list maintenance: python scripts/probe_knn.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgp_b200 import _lib as L
from dgp_b200 import vecchia as V
lib = L.load(); L.device()
rng = np.random.default_rng(3)
for D, m in ((10, 25), (20, 25), (10, 50)):
    M, n = 200000, 100000
    xq, xw = L.to_dev(rng.uniform(0, 1, (M, D))), L.to_dev(rng.uniform(0, 1, (n, D)))
    for mode, name in ((5, "tcgen05 TF32 screen + rank"), (3, "split-TF32 screen + rank"), (1, "DMMA screen + lists + rank"), (2, "DMMA screen only (probe)"), (0, "scalar exact kernel")):
        if mode == 0 and D == 20: continue
        L.check(lib.dgpb_tune(b"knn_mma", mode))
        V.get_pred_nn_dev(xq, xw, m); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); V.get_pred_nn_dev(xq, xw, m); e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3
        print(f"D={D} m={m} {name:24s} {t*1e3:8.2f} ms  {M*n/t/1e9:7.1f} G pairs/s  {M*n*2*D/t/1e12:5.1f} TFLOP/s (2D flop/pair)", flush=True)
L.check(lib.dgpb_tune(b"knn_mma", 5))
