"""One neighbour search for profiling: python scripts/probe_knn_one.py [D] [mode] [m]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgp_b200 import _lib as L
from dgp_b200 import vecchia as V
D = int(sys.argv[1]) if len(sys.argv) > 1 else 10
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 5
m = int(sys.argv[3]) if len(sys.argv) > 3 else 25
lib = L.load(); L.device()
rng = np.random.default_rng(3)
M, n = 200000, 100000
xq, xw = L.to_dev(rng.uniform(0, 1, (M, D))), L.to_dev(rng.uniform(0, 1, (n, D)))
L.check(lib.dgpb_tune(b"knn_mma", mode))
V.get_pred_nn_dev(xq, xw, m); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); V.get_pred_nn_dev(xq, xw, m); e1.record(); torch.cuda.synchronize()
print("mode", mode, "D", D, "m", m, "%.2f ms" % e0.elapsed_time(e1))
