"""Time the complete batched likelihood pipeline (assemble + blocked factorisation + reduce) on synthetic
inputs for several hyper-block settings:  python scripts/probe_factorize.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgp_b200 import _lib as L
lib = L.load(); L.device()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
def run(B, aug, reps=3):
    out = L.host_doubles(2)
    L.check(lib.dgpb_probe_factorize(L.workspace(), n, B, aug, reps, out))
    return out[0], out[1]
for hb, minw in ((128, 0), (256, 0), (512, 0), (512, 1536), (512, 2560), (512, 3584), (256, 1536), (1024, 2560)):
    L.check(lib.dgpb_tune(b"hb", hb)); L.check(lib.dgpb_tune(b"hb_min_w", minw))
    row = []
    for B, aug in ((8, 0), (2, 0), (1, 0), (16, 0), (1, 1)):
        ms, tf = run(B, aug)
        row.append(f"B={B}{'a' if aug else ''}: {ms:7.2f} ms {tf:5.1f} TF")
    print(f"hb={hb:4d} min_w={minw:4d} | " + " | ".join(row), flush=True)
