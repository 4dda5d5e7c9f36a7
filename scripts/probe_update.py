import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgp_b200 import _lib as L
lib = L.load(); L.device()
for n, B in ((5000, 8), (5000, 1), (2500, 8)):
    for flags, name in ((0, "full"), (1, "no C load"), (2, "no C store"), (3, "no C load/store"), (8, "no panel loads"), (11, "math only"), (4, "no math"), (15, "empty")):
        out = L.host_doubles(2)
        L.check(lib.dgpb_probe_update(L.workspace(), n, B, flags, 10, out))
        print(f"n={n} B={B} {name:16s} {out[0]*1e3:9.1f} us  {out[1]:6.2f} TFLOP/s")
