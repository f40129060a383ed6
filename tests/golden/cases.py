"""Declarative kernel x operator cases shared by the golden generator (``oracle/make_golden.py``, which
runs the REAL reference), the oracle tests and the GPU parity tests.

The families, RNG seeds and the Sobol-128 input set mirror the reference's own test-suite
(tests/linpde_gp/randprocs/kernels/linfuncops/diffops/test_diffops.py:15-20 and
cases/cases_{expquad,matern,tensor_product}.py), plus the north-star configurations of BASELINE.json.

Spec format (all JSON-able):
    kernel: {"scale": float|None, "base": {...}}   (see oracle/covfuncs.py)
    L0/L1 : None | [[scalar, [kind, payload]], ...] with kind in {"wl","dd","pd"}; "pd" payload is a list of
            [multi_index_list, coeff] pairs.
"""
from __future__ import annotations

import functools
import operator

import numpy as np


def sobol_points(input_shape) -> np.ndarray:
    """The reference's test inputs: 128 Sobol points in [-3, 3]^d (test_diffops.py:15-20)."""
    import scipy.stats

    d = functools.reduce(operator.mul, input_shape, 1)
    sampler = scipy.stats.qmc.Sobol(d, seed=109134809 + d)
    xs01 = sampler.random_base2(7)
    xs = scipy.stats.qmc.scale(xs01, -3.0, 3.0)
    return xs.reshape((-1,) + tuple(input_shape))


def _wl(weights, scalar=1.0):
    return [[float(scalar), ["wl", np.asarray(weights, dtype=float).tolist()]]]


def _dd(direction, scalar=1.0):
    return [[float(scalar), ["dd", np.asarray(direction, dtype=float).tolist()]]]


def _heat(d, alpha):
    w = [0.0] + [-float(alpha)] * (d - 1)
    return [[1.0, ["pd", [[[1] + [0] * (d - 1), 1.0]]]], [1.0, ["wl", w]]]


def _base(kind, input_shape, **kw):
    return {"kind": kind, "input_shape": list(input_shape), **kw}


def _tp(*factors):
    return {"kind": "tensor_product", "factors": list(factors)}


def _m(nu, ell=1.0):
    return {"kind": "matern", "input_shape": [], "nu": float(nu), "lengthscales": float(ell)}


def _e(ell=1.0):
    return {"kind": "expquad", "input_shape": [], "lengthscales": float(ell)}


def build_cases():
    cases = []

    def add(name, base, L0, L1, scale=None):
        cases.append({"name": name, "kernel": {"scale": scale, "base": base}, "L0": L0, "L1": L1})

    shapes = ((), (1,), (3,))
    # --- ExpQuad (cases_expquad.py) ---------------------------------------------------
    for shp in shapes:
        tag = "x".join(map(str, shp)) or "s"
        k = _base("expquad", shp, lengthscales=1.0)
        rng = np.random.default_rng(390852098)
        add(f"expquad_{tag}_id_dd", k, None, _dd(2.0 * rng.standard_normal(size=shp)))
        rng = np.random.default_rng(4158976)
        add(f"expquad_{tag}_dd_id", k, _dd(2.0 * rng.standard_normal(size=shp)), None)
        rng = np.random.default_rng(52469753628)
        d0 = rng.standard_normal(size=shp)
        d1 = rng.standard_normal(size=shp)
        add(f"expquad_{tag}_dd_dd", k, _dd(d0), _dd(d1))
        rng = np.random.default_rng(524390)
        add(f"expquad_{tag}_id_wl", k, None, _wl(2.0 * rng.standard_normal(size=shp)))
        rng = np.random.default_rng(2309823372)
        add(f"expquad_{tag}_wl_id", k, _wl(2.0 * rng.standard_normal(size=shp)), None)
        rng = np.random.default_rng(235890)
        w0 = 2.0 * rng.standard_normal(size=shp)
        w1 = 2.0 * rng.standard_normal(size=shp)
        add(f"expquad_{tag}_wl_wl", k, _wl(w0), _wl(w1))
        rng = np.random.default_rng(4158976)
        dr = 2.0 * rng.standard_normal(size=shp)
        ww = 2.0 * rng.standard_normal(size=shp)
        add(f"expquad_{tag}_dd_wl", k, _dd(dr), _wl(ww))
        rng = np.random.default_rng(654890)
        dr = 2.0 * rng.standard_normal(size=shp)
        ww = 2.0 * rng.standard_normal(size=shp)
        add(f"expquad_{tag}_wl_dd", k, _wl(ww), _dd(dr))
        add(f"expquad_{tag}_plain", k, None, None)
    # ARD lengthscales (north-star "ExpQuad ARD" closed forms)
    k = _base("expquad", (2,), lengthscales=[0.7, 1.3])
    add("expquad_ard2_neglap_neglap", k, _wl([1.0, 1.0], -1.0), _wl([1.0, 1.0], -1.0), scale=2.25)
    add("expquad_ard2_id_neglap", k, None, _wl([1.0, 1.0], -1.0), scale=2.25)
    add("expquad_ard2_plain", k, None, None, scale=2.25)

    # --- Matern (cases_matern.py) -----------------------------------------------------
    for shp in shapes:
        tag = "x".join(map(str, shp)) or "s"
        for nu in (1.5, 2.5, 3.5, 4.5):
            k = _base("matern", shp, nu=nu, lengthscales=1.0)
            rng = np.random.default_rng(390852098)
            add(f"matern{nu}_{tag}_id_dd", k, None, _dd(2.0 * rng.standard_normal(size=shp)))
            rng = np.random.default_rng(4158976)
            add(f"matern{nu}_{tag}_dd_id", k, _dd(2.0 * rng.standard_normal(size=shp)), None)
            rng = np.random.default_rng(413598)
            d0 = rng.standard_normal(size=shp)
            d1 = rng.standard_normal(size=shp)
            if not (nu == 1.5 and shp == (3,)):  # reference raises: Matern-3/2 multi-d DD/DD has no closed form
                add(f"matern{nu}_{tag}_dd_dd", k, _dd(d0), _dd(d1))
            add(f"matern{nu}_{tag}_plain", k, None, None)
    for nu in (2.5, 3.5, 4.5):
        k = _base("matern", (), nu=nu, lengthscales=1.0)
        rng = np.random.default_rng(5468907)
        add(f"matern{nu}_s_id_wl", k, None, _wl(2.0 * rng.standard_normal(size=())))
        rng = np.random.default_rng(87905642)
        add(f"matern{nu}_s_wl_id", k, _wl(2.0 * rng.standard_normal(size=())), None)
        rng = np.random.default_rng(257834)
        w0 = 2.0 * rng.standard_normal(size=())
        w1 = 2.0 * rng.standard_normal(size=())
        add(f"matern{nu}_s_wl_wl", k, _wl(w0), _wl(w1))
        rng = np.random.default_rng(4158976)
        dr = 2.0 * rng.standard_normal(size=())
        ww = 2.0 * rng.standard_normal(size=())
        add(f"matern{nu}_s_dd_wl", k, _dd(dr), _wl(ww))
        rng = np.random.default_rng(654890)
        dr = 2.0 * rng.standard_normal(size=())
        ww = 2.0 * rng.standard_normal(size=())
        add(f"matern{nu}_s_wl_dd", k, _wl(ww), _dd(dr))

    # --- TensorProduct (cases_tensor_product.py) -----------------------------------------
    rng = np.random.default_rng(390852098)
    dr = rng.standard_normal(size=(2,))
    dr /= np.sqrt(np.sum(dr**2))
    add("tp_m15m15_id_dd", _tp(_m(1.5), _m(1.5)), None, _dd(dr))
    add("tp_m15m15_dd_id", _tp(_m(1.5), _m(1.5)), _dd(dr), None)
    rng = np.random.default_rng(390852098)
    d0 = rng.standard_normal(size=(2,))
    d0 /= np.sqrt(np.sum(d0**2))
    d1 = rng.standard_normal(size=(2,))
    d1 /= np.sqrt(np.sum(d1**2))
    add("tp_m15m15_dd_dd", _tp(_m(1.5), _m(1.5)), _dd(d0), _dd(d1))
    rng = np.random.default_rng(67835487)
    w = 2.0 * rng.standard_normal(size=(2,))
    add("tp_m25m25_id_wl", _tp(_m(2.5), _m(2.5)), None, _wl(w))
    add("tp_m25m25_wl_id", _tp(_m(2.5), _m(2.5)), _wl(w), None)
    rng = np.random.default_rng(67835487)
    w0 = 2.0 * rng.standard_normal(size=(2,))
    w1 = 2.0 * rng.standard_normal(size=(2,))
    add("tp_m25m25_wl_wl", _tp(_m(2.5), _m(2.5)), _wl(w0), _wl(w1))
    rng = np.random.default_rng(89012645)
    dr = rng.standard_normal(size=(2,))
    ww = 2.0 * rng.standard_normal(size=(2,))
    add("tp_m25m25_dd_wl", _tp(_m(2.5), _m(2.5)), _dd(dr), _wl(ww))
    add("tp_m25m25_wl_dd", _tp(_m(2.5), _m(2.5)), _wl(ww), _dd(dr))
    add("tp_m15m25_id_heat", _tp(_m(1.5), _m(2.5)), None, _heat(2, 0.1))
    add("tp_m15m25_heat_heat", _tp(_m(1.5), _m(2.5)), _heat(2, 0.2), _heat(2, 0.1))
    add("tp_m15m25_plain", _tp(_m(1.5), _m(2.5)), None, None)
    # product of 1-D ExpQuads == ARD ExpQuad (tests/…/kernels/test_tensor_product.py:39-47)
    add("tp_e_e_e_plain", _tp(_e(0.8), _e(1.1), _e(1.7)), None, None)
    add("tp_e_e_neglap_neglap", _tp(_e(0.7), _e(1.3)), _wl([1.0, 1.0], -1.0), _wl([1.0, 1.0], -1.0), scale=2.25)
    add("tp_m25e_heat_heat", _tp(_m(2.5, 1.4), _e(0.9)), _heat(2, 0.3), _heat(2, 0.3))
    add("tp_m35m35m35_neglap_neglap", _tp(_m(3.5, 1.1), _m(3.5, 0.9), _m(3.5, 1.3)), _wl([1.0] * 3, -1.0), _wl([1.0] * 3, -1.0), scale=0.5)

    # --- north-star configurations (BASELINE.json configs[1..4]) -----------------
    ns = _tp(_m(2.5, 0.6), _m(2.5, 0.6))
    add("ns_poisson2d_LkL", ns, _wl([1.0, 1.0], -1.0), _wl([1.0, 1.0], -1.0), scale=4.0)
    add("ns_poisson2d_kL", ns, None, _wl([1.0, 1.0], -1.0), scale=4.0)
    add("ns_poisson2d_Lk", ns, _wl([1.0, 1.0], -1.0), None, scale=4.0)
    add("ns_poisson2d_k", ns, None, None, scale=4.0)
    hs = _tp(_m(1.5, 2.5), _m(2.5, 2.0))
    add("ns_heat_LkL", hs, _heat(2, 0.1), _heat(2, 0.1))
    add("ns_heat_kL", hs, None, _heat(2, 0.1))
    add("ns_heat_k", hs, None, None)
    add("ns_poisson1d_expquad_LkL", _base("expquad", (), lengthscales=1.0), _wl(1.0, -1.0), _wl(1.0, -1.0), scale=4.0)
    add("ns_poisson1d_matern25_LkL", _base("matern", (), nu=2.5, lengthscales=1.0), _wl(1.0, -1.0), _wl(1.0, -1.0), scale=4.0)
    add("ns_poisson1d_matern35_kL", _base("matern", (), nu=3.5, lengthscales=1.0), None, _wl(1.0, -1.0), scale=4.0)
    return cases


def spec_to_oracle_op(L):
    """JSON op -> oracle op (``oracle/covfuncs.py``)."""
    if L is None:
        return None
    out = []
    for scalar, (kind, payload) in L:
        if kind == "pd":
            out.append((scalar, ("pd", {tuple(mi): c for mi, c in payload})))
        else:
            out.append((scalar, (kind, np.asarray(payload, dtype=float))))
    return out


def kernel_input_shape(kernel):
    base = kernel["base"]
    if base["kind"] == "tensor_product":
        return (len(base["factors"]),)
    return tuple(base["input_shape"])


# ---- tensor-grid (Kronecker) structure path: SURVEY.md section 8f item 3 ------------------------------------------
def build_kron_cases():
    """Tensor-product kernels on TensorProductGrids (src/linpde_gp/randprocs/covfuncs/_tensor_product.py:64-82,
    .../linfuncops/diffops/_tensor_product.py:140-156).  ``factors1=None`` -> symmetric block (``x1=None``)."""
    g7, g5 = np.linspace(0.0, 1.0, 7).tolist(), np.linspace(-1.0, 1.0, 5).tolist()
    h4, h6 = np.linspace(0.1, 0.9, 4).tolist(), np.linspace(-0.7, 1.3, 6).tolist()
    tp2 = _tp(_m(2.5, 0.4), _m(2.5, 0.7))
    tpm = _tp(_m(1.5, 0.9), _m(2.5, 0.5))
    tp3 = _tp(_e(0.8), _m(2.5, 0.6), _m(1.5, 1.1))
    lap2, lap3 = _wl([1.0, 1.0], -1.0), _wl([1.0, 0.5, 2.0])
    cases = [
        dict(name="kron_tp2_k", kernel=dict(scale=None, base=tp2), L0=None, L1=None, factors0=[g7, g5], factors1=None),
        dict(name="kron_tp2_k_cross", kernel=dict(scale=3.0, base=tp2), L0=None, L1=None, factors0=[g7, g5], factors1=[h4, h6]),
        dict(name="kron_tp2_kL", kernel=dict(scale=2.0, base=tp2), L0=None, L1=lap2, factors0=[g7, g5], factors1=[h4, h6]),
        dict(name="kron_tp2_LkL", kernel=dict(scale=2.0, base=tp2), L0=lap2, L1=lap2, factors0=[g7, g5], factors1=None),
        dict(name="kron_tp2_LkL_edge", kernel=dict(scale=2.0, base=tp2), L0=lap2, L1=None, factors0=[g7, g5], factors1=[[0.0], h6]),
        dict(name="kron_heat_LkL", kernel=dict(scale=None, base=tpm), L0=_heat(2, 0.1), L1=_heat(2, 0.1), factors0=[g7, g5], factors1=None),
        dict(name="kron_heat_Lk", kernel=dict(scale=None, base=tpm), L0=_heat(2, 0.2), L1=None, factors0=[h4, h6], factors1=[g7, g5]),
        dict(name="kron_tp3_k", kernel=dict(scale=None, base=tp3), L0=None, L1=None, factors0=[h4, [0.0, 0.5, 1.5], g5], factors1=None),
        dict(name="kron_tp3_LkL", kernel=dict(scale=1.5, base=tp3), L0=lap3, L1=lap3, factors0=[h4, [0.0, 0.5, 1.5], g5], factors1=None),
    ]
    return cases


# kernels / operators of the seam goldens (tests/golden/seams.npz, oracle/make_golden.py::make_seams)
SEAM_KERNELS = {
    "matern_tp": {"scale": 2.5, "base": {"kind": "tensor_product", "factors": [
        {"kind": "matern", "nu": 2.5, "lengthscales": 0.6, "input_shape": []},
        {"kind": "matern", "nu": 3.5, "lengthscales": 0.9, "input_shape": []}]}},
    "expquad": {"scale": 1.7, "base": {"kind": "expquad", "lengthscales": [0.7, 1.1], "input_shape": [2]}},
}
SEAM_OPS = {"id": None, "neglap": [(-1.0, ("wl", [1.0, 1.0]))], "dd": [(1.0, ("dd", [0.3, -1.2]))]}
