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


def _direction(L, d):
    """JSON op -> direction vector of a first-order operator (None = identity)."""
    terms = flatten_op(L, d)
    if terms is None:
        return None
    vec = np.zeros(d)
    for mi, c in terms.items():
        assert sum(mi) == 1
        vec[mi.index(1)] += c
    return vec


def desc_from_spec(spec):
    from linpde_gp_b200._lowering import lower, lower_radial

    base = spec["kernel"]["base"]
    shape = tuple(base.get("input_shape", ()))
    if base["kind"] == "matern" and shape and int(np.prod(shape)) > 1:  # isotropic multi-d Matern: radial family
        d = int(np.prod(shape))
        ls = np.broadcast_to(np.asarray(base["lengthscales"], dtype=float), (d,))
        scale = spec["kernel"].get("scale")
        return lower_radial(base["nu"], np.sqrt(2 * base["nu"]) / ls, _direction(spec["L0"], d), _direction(spec["L1"], d),
                            1.0 if scale is None else scale)
    factors = factors_from_base(spec["kernel"]["base"])
    d = len(factors)
    scale = spec["kernel"].get("scale")
    return lower(factors, flatten_op(spec["L0"], d), flatten_op(spec["L1"], d), 1.0 if scale is None else scale)


def eval_desc_numpy(desc, X0, X1):
    """Reference semantics of ``lpgp_kernel_desc`` (include/lpgp.h) in numpy; O(N0*N1*ncoef)."""
    d = desc.d
    X0 = np.asarray(X0, dtype=float).reshape(len(X0), d)
    X1 = np.asarray(X1, dtype=float).reshape(len(X1), d)
    if desc.dim_type[0] == 2:  # LPGP_DIM_RADIAL: exp(-r) (Q0(r) + <a,u> Q1(r) + <a,u><b,u> Q2(r))
        nq = 6
        coef = np.array(desc.coef[: 3 * nq + 2 * d])
        s = np.array(desc.scale[:d])
        u = (X0[:, None, :] - X1[None, :, :]) * s
        r = np.sqrt(np.sum(u * u, axis=-1))
        a, b = coef[3 * nq : 3 * nq + d], coef[3 * nq + d : 3 * nq + 2 * d]
        pa, pb = u @ a, u @ b
        q = [sum(coef[t * nq + i] * r**i for i in range(nq)) for t in range(3)]
        return (q[0] + pa * q[1] + pa * pb * q[2]) * np.exp(-r)
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


# ---- JSON spec -> objects of the product's reference-style API ---------------------------------------------
def api_base(base):
    from linpde_gp_b200.randprocs import covfuncs

    kind = base["kind"]
    if kind == "tensor_product":
        return covfuncs.TensorProduct(*(api_base(f) for f in base["factors"]))
    shape = tuple(base.get("input_shape", ()))
    ls = base["lengthscales"]
    ls = np.asarray(ls, dtype=float) if isinstance(ls, (list, tuple)) else float(ls)
    if kind == "matern":
        return covfuncs.Matern(shape, nu=base["nu"], lengthscales=ls)
    return covfuncs.ExpQuad(shape, lengthscales=ls)


def api_kernel(kernel):
    k = api_base(kernel["base"])
    if kernel.get("scale") is not None:
        k = kernel["scale"] * k
    return k


def api_op(L):
    from linpde_gp_b200.linfuncops import SumLinearFunctionOperator, diffops

    if L is None:
        return None
    summands = []
    for scalar, (kind, payload) in L:
        if kind == "wl":
            op = diffops.WeightedLaplacian(np.asarray(payload, dtype=float))
        elif kind == "dd":
            op = diffops.DirectionalDerivative(np.asarray(payload, dtype=float))
        else:
            assert len(payload) == 1 and payload[0][1] == 1.0
            op = diffops.PartialDerivative(diffops.MultiIndex(payload[0][0]))
        if scalar != 1.0:
            op = scalar * op
        summands.append(op)
    return summands[0] if len(summands) == 1 else SumLinearFunctionOperator(*summands)


def api_L0kL1(spec):
    k = api_kernel(spec["kernel"])
    L0, L1 = api_op(spec["L0"]), api_op(spec["L1"])
    kk = L1(k, argnum=1) if L1 is not None else k
    return L0(kk, argnum=0) if L0 is not None else kk


def api_solve(problem):
    """Run a golden GP problem through the product API, block by block (like oracle/make_golden.py does with the
    real reference)."""
    import linpde_gp_b200 as lg
    from tests.golden import cases as gcases

    kernel = problem["kernel"]
    shape = gcases.kernel_input_shape(kernel)
    prior = lg.GaussianProcess(lg.functions.Zero(input_shape=shape), api_kernel(kernel))
    post = prior
    for blk in problem["blocks"]:
        X = np.asarray(blk["X"], dtype=float)
        Y = np.asarray(blk["Y"], dtype=float)
        if blk.get("grid") is not None:  # gridded batch: hand the product the TensorProductGrid itself
            X = lg.randprocs.covfuncs.TensorProductGrid(*[np.asarray(f, dtype=float) for f in blk["grid"]])
            Y = Y.reshape(X.shape[:-1])
        b = None
        if blk.get("noise_var") is not None:
            nv = np.broadcast_to(np.asarray(blk["noise_var"], dtype=float), Y.shape).copy()
            b = lg.randvars.Normal(np.zeros_like(Y), lg.linops.Scaling(nv))
        post = post.condition_on_observations(Y, X=X, L=api_op(blk["L"]), b=b)
    Xt = np.asarray(problem["Xt"], dtype=float)
    Xc = Xt[: problem.get("n_cov", 16)]
    return post, {
        "w": post.representer_weights,
        "mean": post.mean(Xt),
        "var": post.cov(Xt, None),
        "cov": post.cov.matrix(Xc),
        "gram": post.gram.todense(),
    }


def api_solve_multi_output(problem, one_shot: bool = False):
    """Run a multi-output golden problem (oracle/multi_output.py spec) through the product API: independent-output
    prior, observation operators ``sum_t c_t * (D_t @ SelectOutput(o_t))``, posterior of every selected output."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200 import linfuncops
    from linpde_gp_b200.randprocs import covfuncs
    from tests.golden import cases as gcases

    shape = gcases.kernel_input_shape(problem["kernels"][0])
    nout = len(problem["kernels"])
    prior = lg.GaussianProcess(
        lg.functions.StackedFunction(*(lg.functions.Constant(shape, m) for m in problem["means"])),
        covfuncs.IndependentMultiOutputCovarianceFunction(*(api_kernel(k) for k in problem["kernels"])),
    )
    sel = [linfuncops.SelectOutput((shape, (nout,)), idx=j) for j in range(nout)]
    batches = []
    for blk in problem["blocks"]:
        if "functional" in blk:  # scalar observation: sum of Lebesgue integrals and point evaluations (X=None)
            L = None
            for a in blk["functional"]:
                if a[0] == "int":
                    t = a[2] * lg.linfunctls.LebesgueIntegral((a[3], a[4])) @ sel[a[1]]
                else:
                    t = a[2] * sel[a[1]].to_linfunctl(a[3])
                L = t if L is None else L + t
            batches.append((float(blk["Y"][0]), None, L, None))
            continue
        X = np.asarray(blk["X"], dtype=float)
        Y = np.asarray(blk["Y"], dtype=float)
        L = None
        for o, c, op in blk["Ls"]:
            t = sel[o] if op is None else api_op(op) @ sel[o]
            if c != 1.0:
                t = c * t
            L = t if L is None else L + t
        b = None
        if blk.get("noise_var") is not None:
            nv = np.broadcast_to(np.asarray(blk["noise_var"], dtype=float), Y.shape).copy()
            b = lg.randvars.Normal(np.zeros_like(Y), lg.linops.Scaling(nv))
        batches.append((Y, X, L, b))
    if one_shot:
        post = lg.ConditionalGaussianProcess.from_observation_batches(prior, batches)
    else:
        post = prior
        for Y, X, L, b in batches:
            post = post.condition_on_observations(Y, X=X, L=L, b=b)
    Xt = np.asarray(problem["Xt"], dtype=float)
    Xc = Xt[: problem.get("n_cov", 8)]
    outs = [s(post) for s in sel]
    return post, {
        "w": post.representer_weights,
        "gram": post.gram.todense(),
        "mean": np.stack([o.mean(Xt) for o in outs]),
        "var": np.stack([o.cov(Xt, None) for o in outs]),
        "cov": np.stack([o.cov.matrix(Xc) for o in outs]),
    }
