"""Import shim for the UNMODIFIED reference (`/root/reference/dgpsi`) in the build container.

Only used by `make_golden.py` (fixture generation) -- never at test/bench run time, because
`/root/reference` does not exist on the GPU box.  Nothing in the reference is modified:

* `matplotlib` / `pathos` are not installed and are not on the serial arithmetic path -> stub modules.
* `dgpsi/functions.py:13` and `dgpsi/vecchia.py:17` force `numba.config.THREADING_LAYER='tbb'`; TBB is
  absent here, so writes of that attribute are dropped and numba falls back to its `omp` layer.
* `faiss` is absent -> the reference itself selects sklearn's kd-tree kNN (`dgpsi/vecchia.py:6-11`).
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


def import_reference():
    if "dgpsi" in sys.modules:
        return sys.modules["dgpsi"]
    for name in ("matplotlib", "matplotlib.pyplot", "pathos", "pathos.multiprocessing"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pathos"].multiprocessing = sys.modules["pathos.multiprocessing"]
    sys.modules["pathos.multiprocessing"].ProcessingPool = object

    import numba.core.config as nbcfg

    class _Cfg(types.ModuleType):
        def __setattr__(self, key, value):
            if key == "THREADING_LAYER":
                return
            super().__setattr__(key, value)

    nbcfg.__class__ = _Cfg
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import dgpsi  # noqa: E402

    return dgpsi
