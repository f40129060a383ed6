"""Test helpers: JSON case spec -> product descriptor, and a slow numpy evaluation of a descriptor (used on CPU
to check the host-side lowering without a GPU; never used by the product)."""
from __future__ import annotations

import itertools

import numpy as np


def flatten_op(L, d):
    """JSON op (list of scaled summands) -> {multi_index: coeff} (sum of all summands)."""
    if L is None:
        return None
    out = {}
    for scalar, (kind, payload) in L:
        if kind == "pd":
            items = [(tuple(mi), c) for mi, c in payload]
        else:
            arr = np.atleast_1d(np.asarray(payload, dtype=float)).reshape(-1)
            order = 2 if kind == "wl" else 1
            items = []
            for i, c in enumerate(arr):
                if c != 0.0:
                    mi = [0] * d
                    mi[i] = order
                    items.append((tuple(mi), c))
        for mi, c in items:
            out[mi] = out.get(mi, 0.0) + scalar * c
    return out


def factors_from_base(base):
    from linpde_gp_b200._lowering import Factor1D

    kind = base["kind"]
    if kind == "tensor_product":
        fs = []
        for f in base["factors"]:
            fs.extend(factors_from_base(f))
        return fs
    shape = tuple(base.get("input_shape", ()))
    d = int(np.prod(shape)) if shape else 1
    ls = np.broadcast_to(np.asarray(base["lengthscales"], dtype=float), (d,))
    if kind == "expquad":
        return [Factor1D("expquad", l) for l in ls]
    if kind == "matern":
        if d != 1:
            raise NotImplementedError("isotropic multi-d Matern is not of product form")
        return [Factor1D("matern", ls[0], nu=base["nu"])]
    raise ValueError(kind)


def desc_from_spec(spec):
    from linpde_gp_b200._lowering import lower

    factors = factors_from_base(spec["kernel"]["base"])
    d = len(factors)
    scale = spec["kernel"].get("scale")
    return lower(factors, flatten_op(spec["L0"], d), flatten_op(spec["L1"], d), 1.0 if scale is None else scale)


def eval_desc_numpy(desc, X0, X1):
    """Reference semantics of ``lpgp_kernel_desc`` (include/lpgp.h) in numpy; O(N0*N1*ncoef)."""
    d = desc.d
    X0 = np.asarray(X0, dtype=float).reshape(len(X0), d)
    X1 = np.asarray(X1, dtype=float).reshape(len(X1), d)
    nbt = [desc.nbasis[i] * (2 if desc.has_odd[i] else 1) for i in range(d)]
    coef = np.array(desc.coef[: int(np.prod(nbt))]).reshape(nbt)
    g = np.zeros((len(X0), len(X1)))
    basis = []
    for i in range(d):
        u = (X0[:, None, i] - X1[None, :, i]) * desc.scale[i]
        if desc.dim_type[i] == 1:
            v = u
            g += 0.5 * u * u
        else:
            v = np.abs(u)
            g += v
        b = [v**e for e in range(desc.nbasis[i])]
        if desc.has_odd[i]:
            b += [u * v**e for e in range(desc.nbasis[i])]
        basis.append(b)
    out = np.zeros_like(g)
    for idx in itertools.product(*[range(n) for n in nbt]):
        c = coef[idx]
        if c == 0.0:
            continue
        term = c
        for i, bi in enumerate(idx):
            term = term * basis[i][bi]
        out += term
    return out * np.exp(-g)
