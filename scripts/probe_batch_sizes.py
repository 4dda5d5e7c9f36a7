import os, sys
sys.path.insert(0, os.getcwd())
from dgp_b200 import _lib as L
lib = L.load(); L.device()
out = L.host_doubles(2)
for n in (5000, 2000):
    for aug in (0, 1):
        for B in ((1, 2, 3, 4, 8, 16) if not aug else (1,)):
            L.check(lib.dgpb_probe_factorize(L.workspace(), n, B, aug, 5, out))
            print(f"n={n} aug={aug} B={B}: {out[0]:.3f} ms  {out[1]:.2f} TFLOP/s", flush=True)
