import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgp_b200 import _lib as L
lib = L.load(); L.device()
for n, B in ((5500, 8), (3000, 8), (1700, 8), (5500, 1)):
    for kc in (4, 8, 16):
        for flags, name in ((0, "full"), (3, "no C load/store"), (8, "no panel loads"), (11, "math only"), (4, "no math")):
            out = L.host_doubles(2)
            L.check(lib.dgpb_probe_update(L.workspace(), n, B, flags | (kc << 8), 5, out))
            print(f"n={n} B={B} K={32*kc:4d} {name:16s} {out[0]*1e3:9.1f} us  {out[1]:6.2f} TFLOP/s", flush=True)
