"""Hyper-block width / threshold sweep of the factorisation at small batch sizes (dgpb_tune 'hb', 'hb_min_w')."""
import os, sys
sys.path.insert(0, os.getcwd())
from dgp_b200 import _lib as L
lib = L.load(); L.device()
out = L.host_doubles(2)
for hb, minw, graded in ((512, 2560, 1), (256, 2560, 1), (128, 2560, 1), (256, 3840, 1), (256, 5120, 1)):
    L.check(lib.dgpb_tune(b"hb", hb)); L.check(lib.dgpb_tune(b"hb_min_w", minw)); L.check(lib.dgpb_tune(b"hb_graded", graded))
    row = []
    for n, aug, B in ((5000, 0, 1), (5000, 0, 2), (5000, 0, 3), (5000, 0, 4), (5000, 0, 6), (5000, 1, 1), (2000, 0, 1), (2000, 0, 8)):
        L.check(lib.dgpb_probe_factorize(L.workspace(), n, B, aug, 5, out))
        row.append(f"{n}{'aug' if aug else ''} B={B}: {out[0]:.3f}")
    print(f"hb={hb} min_w={minw} graded={graded} | " + " | ".join(row), flush=True)
