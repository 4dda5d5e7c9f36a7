import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgp_b200 import _lib as L
lib = L.load(); L.device()
def run(B, aug, n=5000, reps=3):
    out = L.host_doubles(2)
    L.check(lib.dgpb_probe_factorize(L.workspace(), n, B, aug, reps, out))
    return out[0], out[1]
for crit in (1, 0, 1):
    L.check(lib.dgpb_tune(b"crit_stream", crit))
    row = []
    for B, aug in ((8, 0), (2, 0), (1, 0), (16, 0), (1, 1)):
        ms, tf = run(B, aug)
        row.append(f"B={B}{'a' if aug else ''}: {ms:7.2f} ms {tf:5.1f} TF")
    print(f"crit_stream={crit} | " + " | ".join(row), flush=True)
