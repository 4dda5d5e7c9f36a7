"""Short, representative slice of one SEM iteration at BASELINE config-3 size for ncu captures:
one blocked ESS update of layer 1 (8 target nodes, 8 upper nodes), one M-step gradient evaluation,
one compute_stats and a 256-point 2-layer prediction.  Usage: python scripts/prof_workload.py [n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dgp_b200 as D  # noqa: E402
from bench import SEED, layers_config3, make_config3  # noqa: E402
from dgp_b200.imputation import _DeviceLayers  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rng = np.random.default_rng(SEED)
X, Y = make_config3(n, rng)
layers = layers_config3(lambda **kw: D.kernel(**kw))
In = X
for l, layer in enumerate(layers):
    for k, node in enumerate(layer):
        node.input_dim = np.arange(In.shape[1])
        node.input = In[:, node.input_dim].copy()
        if node.connect is not None:
            node.global_input = X[:, node.connect]
        node.output = Y[:, [k]].copy() if l == 2 else In[:, [k]].copy()
        node.D = node.input.shape[1] + (0 if node.connect is None else len(node.connect))
        node.para_path = np.atleast_2d(np.concatenate((node.scale, node.length, node.nugget)))
        node.vecch = False
dev = _DeviceLayers(layers)
nprop, _ = dev.ess_call(0, list(range(8)), list(range(8)), rng.standard_normal((8, n)), rng.uniform(size=64))
dev.write_back()
top = layers[2][0]
f, g = top.llik(top.log_t().copy())
for node in layers[0][:2] + [top]:
    node.compute_stats()
m, v = layers[0][0].gp_prediction(rng.uniform(0, 1, (256, 8)), None)
m2, v2 = top.linkgp_prediction(rng.uniform(0, 1, (256, 8)), rng.uniform(1e-3, 1e-2, (256, 8)), rng.uniform(0, 1, (256, 8)))
print("proposals", nprop, "nllik", f, "pred", float(m.mean()), float(m2.mean()))
