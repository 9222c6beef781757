"""Import the UNMODIFIED reference modules from /root/reference under shims.

TEST INFRASTRUCTURE ONLY (see oracle/george_shim.py header).  This loader only
works in the build container, where /root/reference exists; it is used by
``oracle/gen_golden.py`` to produce the committed fixtures in ``tests/golden/``
and by CPU-only tests that are skipped when the reference tree is absent.
Nothing here runs on the GPU box.

Shims injected through ``sys.modules`` (none of these packages is installed and
there is no network; each is a restatement of the package's published
behaviour, PARITY UNPINNED by the reference):

* ``george``            -> oracle.george_shim          (gpModel.py:8)
* ``pyDOE.lhs``         -> classic LHS on the legacy global numpy RNG
                           (util.py:9, barefoot.py:26, sample_code.py:13)
* ``sklearn_extra.cluster.KMedoids`` -> oracle.kmedoids_shim (util.py:13)
* ``ray.util.multiprocessing.Pool``  -> serial or fork pool (barefoot.py:32)
* ``matplotlib`` / ``matplotlib.pyplot`` -> empty stubs
                           (reificationFusion.py:10)
* ``scipy.random = numpy.random`` (acquisitionFunc.py:10) and
  ``numpy.float = float`` (barefoot.py:1671,1850,...) for the NumPy-2/SciPy-1.18
  image.
"""
import importlib
import multiprocessing
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("BAREFOOT_REFERENCE_DIR", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "reificationFusion.py"))


def lhs(n, samples=None, criterion=None, iterations=None):
    """pyDOE.lhs with criterion=None ("classic"): one uniform draw per stratum,
    then an independent random permutation per column."""
    if samples is None:
        samples = n
    if criterion is not None:
        raise NotImplementedError("only the classic criterion is restated")
    cut = np.linspace(0, 1, samples + 1)
    u = np.random.rand(samples, n)
    a = cut[:samples]
    b = cut[1:samples + 1]
    rdpoints = np.zeros_like(u)
    for j in range(n):
        rdpoints[:, j] = u[:, j] * (b - a) + a
    H = np.zeros_like(rdpoints)
    for j in range(n):
        order = np.random.permutation(range(samples))
        H[:, j] = rdpoints[order, j]
    return H


class SerialPool:
    """Stand-in for ray.util.multiprocessing.Pool: deterministic, in-process."""

    def __init__(self, *a, **k):
        pass

    def map(self, fn, it):
        return [fn(x) for x in it]

    def terminate(self):
        pass

    def close(self):
        pass

    def join(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _fork_pool(*a, **k):
    return multiprocessing.get_context("fork").Pool(*a, **k)


def install_shims(parallel_pool=False):
    from . import george_shim, kmedoids_shim

    if not hasattr(np, "float"):
        np.float = float
    import scipy
    if not hasattr(scipy, "random"):
        scipy.random = np.random

    g = types.ModuleType("george")
    g.kernels = george_shim.kernels
    g.GP = george_shim.GP
    sys.modules["george"] = g
    gk = types.ModuleType("george.kernels")
    for name in ("ExpSquaredKernel", "Matern32Kernel", "Matern52Kernel"):
        setattr(gk, name, getattr(george_shim, name))
    sys.modules["george.kernels"] = gk

    p = types.ModuleType("pyDOE")
    p.lhs = lhs
    sys.modules["pyDOE"] = p

    se = types.ModuleType("sklearn_extra")
    sec = types.ModuleType("sklearn_extra.cluster")
    sec.KMedoids = kmedoids_shim.KMedoids
    se.cluster = sec
    sys.modules["sklearn_extra"] = se
    sys.modules["sklearn_extra.cluster"] = sec

    ray = types.ModuleType("ray")
    ray_util = types.ModuleType("ray.util")
    ray_mp = types.ModuleType("ray.util.multiprocessing")
    ray_mp.Pool = _fork_pool if parallel_pool else SerialPool
    ray.util = ray_util
    ray_util.multiprocessing = ray_mp
    sys.modules["ray"] = ray
    sys.modules["ray.util"] = ray_util
    sys.modules["ray.util.multiprocessing"] = ray_mp

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


_REF_MODULES = ("gpModel", "reificationFusion", "acquisitionFunc", "kmedoids", "util",
                "barefoot")


def load_reference(names=("gpModel", "reificationFusion", "acquisitionFunc", "kmedoids",
                          "util"), parallel_pool=False):
    """Return {name: module} of the verbatim reference modules."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
    install_shims(parallel_pool=parallel_pool)
    # the product package mirrors the reference's module names; make sure the
    # bare names resolve to /root/reference and nothing else
    for n in _REF_MODULES:
        m = sys.modules.get(n)
        if m is not None and not getattr(m, "__file__", "").startswith(REFERENCE_DIR):
            del sys.modules[n]
    sys.path.insert(0, REFERENCE_DIR)
    sys.dont_write_bytecode = True
    try:
        out = {n: importlib.import_module(n) for n in names}
    finally:
        sys.path.remove(REFERENCE_DIR)
    return out
