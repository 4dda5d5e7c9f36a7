"""Loader for the UNMODIFIED reference (`dgpsi` 2.6.0) that `__graft_entry__.build()` installs into
`baseline/_ref` with

    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
        --target baseline/_ref /root/reference

(`baseline/_ref` is git-ignored but travels to the GPU box).  Only bench.py's CPU legs (`--impl reference`,
`cpu_baseline`) import this module; the product (`dgp_b200/`) never does.  Nothing in the reference is modified:

* `matplotlib` / `pathos` are not installed and are not on the serial arithmetic path -> stub modules;
* `dgpsi/functions.py:13` and `dgpsi/vecchia.py:17` force numba's TBB threading layer, which is absent here: writes
  of that one attribute are dropped and numba uses its `omp` layer;
* `faiss` is absent -> the reference itself selects sklearn's kd-tree kNN (`dgpsi/vecchia.py:6-11`).

The helpers below rebuild reference objects around the SAME arrays a dgp_b200 model holds (training inputs, imputed
latent layers, hyper-parameters), so the reference and the GPU path are timed -- and compared -- on identical inputs.
"""
import copy
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "dgpsi"))


def load():
    """Import `dgpsi` from baseline/_ref; returns the module or None when it is not installed."""
    if "dgpsi" in sys.modules:
        return sys.modules["dgpsi"]
    if not available():
        return None
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
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import dgpsi  # noqa: E402

    return dgpsi


def host_info():
    """Cores, BLAS and numba threading actually in use (reported beside every CPU number)."""
    info = {"cores_logical": os.cpu_count()}
    try:
        import psutil
        info["cores_physical"] = psutil.cpu_count(logical=False)
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_info
        info["blas"] = "; ".join(f"{d.get('internal_api')} {d.get('version')} x{d.get('num_threads')}"
                                 for d in threadpool_info())
    except Exception:
        pass
    try:
        import numba
        info["numba_threads"] = numba.get_num_threads()
    except Exception:
        pass
    return info


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the reference's default is every core (dgpsi/functions.py:10-14)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    try:
        import numba
        numba.set_num_threads(min(os.cpu_count(), numba.config.NUMBA_NUM_THREADS))
    except Exception:
        pass


# ---- reference objects around the arrays of a dgp_b200 model --------------------------------------------------------
def ref_kernel(dgpsi, k):
    """A reference `kernel` with the data and hyper-parameters of the dgp_b200 kernel `k` (GP nodes only)."""
    from dgpsi.kernel_class import kernel

    r = kernel(length=np.array(k.length, dtype=np.float64), scale=float(k.scale[0]), nugget=float(k.nugget[0]),
               name=k.name, scale_est=bool(k.scale_est), nugget_est=bool(k.nugget_est),
               input_dim=None if k.input_dim is None else np.array(k.input_dim),
               connect=None if k.connect is None else np.array(k.connect))
    r.input = np.array(k.input, dtype=np.float64)
    r.output = np.array(k.output, dtype=np.float64)
    r.global_input = None if k.global_input is None else np.array(k.global_input, dtype=np.float64)
    r.D = r.input.shape[1] + (0 if r.global_input is None else r.global_input.shape[1])
    r.para_path = np.atleast_2d(np.concatenate((r.scale, r.length, r.nugget)))
    r.vecch = bool(k.vecch)
    if k.vecch:
        r.m, r.ord, r.NNarray = k.m, np.array(k.ord), np.array(k.NNarray)
        r.rev_ord = np.argsort(r.ord)
    return r


def ref_layers(dgpsi, all_layer, stats=True):
    layers = []
    for layer in all_layer:
        out = []
        for k in layer:
            r = ref_kernel(dgpsi, k)
            if stats and not r.vecch:
                r.compute_stats()
            out.append(r)
        layers.append(out)
    return layers


def ref_emulator(dgpsi, emu, imputations):
    """A reference `emulator` holding the first `imputations` frozen imputations of the dgp_b200 emulator `emu`
    (K^-1 recomputed by the reference's own `compute_stats`)."""
    e = dgpsi.emulator.__new__(dgpsi.emulator)
    e.all_layer_set = [ref_layers(dgpsi, al) for al in emu.all_layer_set[:imputations]]
    e.all_layer = e.all_layer_set[0]
    e.n_layer = len(e.all_layer)
    e.vecch = bool(emu.vecch)
    return e


def ref_lgp(dgpsi, system, imputations):
    """A reference `lgp` around the first `imputations` imputation sets of the dgp_b200 linked system."""
    sets = []
    for one in system.all_layer_set[:imputations]:
        layers = []
        for layer in one:
            conts = []
            for c in layer:
                rc = dgpsi.container.__new__(dgpsi.container)
                rc.type = c.type
                rc.structure = ref_layers(dgpsi, [[c.structure]])[0][0] if c.type == "gp" else ref_layers(dgpsi, c.structure)
                rc.vecch = bool(c.vecch)
                rc.local_input_idx = copy.copy(c.local_input_idx)
                conts.append(rc)
            layers.append(conts)
        sets.append(layers)
    s = dgpsi.lgp.__new__(dgpsi.lgp)
    s.L, s.all_layer, s.all_layer_set, s.num_model = system.L, sets[0], sets, list(system.num_model)
    return s
