"""In-memory loader for the REAL reference (linpde-gp + its vendored probnum fork).

TEST INFRASTRUCTURE ONLY.  Used in the build container (where /root/reference exists) to
(a) validate the numpy restatement in ``oracle/`` and (b) generate the frozen golden vectors
under ``tests/golden/`` (see ``oracle/make_golden.py``).  Nothing on the GPU box may import this
module: /root/reference does not exist there.

The reference pins python<3.12 / numpy<2 and imports jax + pykeops at module import time; neither
is installed here.  The recipe below patches three numpy-2 aliases and injects stub
modules so that the reference's *numpy/scipy path* -- the one parity is demanded against -- runs
unmodified.  Nothing is written under /root/reference.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LINPDE_GP_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "linpde_gp"))


def load():
    """Return the tuple ``(probnum, linpde_gp)`` of the real reference modules."""
    if "linpde_gp" in sys.modules and "probnum" in sys.modules:
        return sys.modules["probnum"], sys.modules["linpde_gp"]
    if not available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")

    import numpy as np

    sys.path[:0] = [
        os.path.join(REFERENCE_ROOT, "probnum", "src"),
        os.path.join(REFERENCE_ROOT, "src"),
    ]

    ver = types.ModuleType("probnum._version")
    ver.version = "0.1.25.dev0"
    ver.__version__ = ver.version
    sys.modules["probnum._version"] = ver

    # numpy-2 aliases the reference still uses
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    if not hasattr(np, "find_common_type"):
        np.find_common_type = lambda a, b: np.result_type(*a, *b)
    if not hasattr(np, "AxisError"):
        np.AxisError = np.exceptions.AxisError

    import probnum  # noqa: F401  (first: its _USE_KEOPS flags must resolve to False)

    def _jit(fn=None, **_kw):
        if fn is None:
            return lambda f: f
        return fn

    def _no_autodiff(*_a, **_k):
        raise NotImplementedError("jax stub: autodiff is not available in the oracle shim")

    jax = types.ModuleType("jax")
    jax.jit = _jit
    jax.hessian = _no_autodiff
    jax.jvp = _no_autodiff
    jax.grad = _no_autodiff
    jax.vmap = _no_autodiff
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        if not name.startswith("__"):
            setattr(jnp, name, getattr(np, name))
    jax.numpy = jnp
    jsp = types.ModuleType("jax.scipy")
    import scipy.special

    jsp.special = scipy.special
    jax.scipy = jsp
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.scipy": jsp, "jax.scipy.special": scipy.special})

    class LazyTensor:  # pylint: disable=too-few-public-methods
        def __init__(self, *a, **k):
            raise NotImplementedError("pykeops stub")

    pk = types.ModuleType("pykeops")
    pkn = types.ModuleType("pykeops.numpy")
    pkn.LazyTensor = LazyTensor
    pkn.Pm = pkn.Vi = pkn.Vj = _no_autodiff
    pk.numpy = pkn
    sys.modules.update({"pykeops": pk, "pykeops.numpy": pkn})

    import linpde_gp  # noqa: F401

    return sys.modules["probnum"], sys.modules["linpde_gp"]
