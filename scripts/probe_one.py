"""One batched likelihood pipeline for an ncu launch list: python scripts/probe_one.py n B aug hb min_w"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgp_b200 import _lib as L
lib = L.load(); L.device()
n, B, aug, hb, minw = (int(v) for v in sys.argv[1:6])
L.check(lib.dgpb_tune(b"hb", hb)); L.check(lib.dgpb_tune(b"hb_min_w", minw))
out = L.host_doubles(2)
L.check(lib.dgpb_probe_factorize(L.workspace(), n, B, aug, 1, out))
print(out[0], out[1])
