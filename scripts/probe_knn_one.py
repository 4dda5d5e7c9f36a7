import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgp_b200 import _lib as L
from dgp_b200 import vecchia as V
lib = L.load(); L.device()
rng = np.random.default_rng(3)
D = int(sys.argv[1]) if len(sys.argv) > 1 else 10
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 3
xq, xw = L.to_dev(rng.uniform(0, 1, (200000, D))), L.to_dev(rng.uniform(0, 1, (100000, D)))
L.check(lib.dgpb_tune(b"knn_mma", mode))
V.get_pred_nn_dev(xq, xw, 25); torch.cuda.synchronize()
