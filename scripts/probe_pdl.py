"""Factorisation time with and without programmatic dependent launch on the critical path (dgpb_tune 'pdl')."""
import os, sys
sys.path.insert(0, os.getcwd())
from dgp_b200 import _lib as L
lib = L.load(); L.device()
out = L.host_doubles(2)
for pdl in (0, 1, 8):
    L.check(lib.dgpb_tune(b"pdl", pdl))
    for n, aug, B in ((5000, 0, 1), (5000, 0, 4), (5000, 0, 8), (5000, 1, 1), (2000, 0, 1), (1000, 0, 1), (500, 0, 8)):
        L.check(lib.dgpb_probe_factorize(L.workspace(), n, B, aug, 5, out))
        print(f"pdl={pdl} n={n} aug={aug} B={B}: {out[0]:.3f} ms  {out[1]:.2f} TFLOP/s", flush=True)
