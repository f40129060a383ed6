"""Generate the frozen golden vectors under ``tests/golden/`` by running the REAL reference.

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (needs /root/reference, loaded through
``oracle/refshim.py``); the resulting ``.npz`` files are committed so that the GPU box -- which has no
reference tree -- can check both the oracle and the CUDA path against outputs of the reference itself.

    python -m oracle.make_golden            # writes tests/golden/kernels.npz and tests/golden/gp_*.npz

While generating, every vector is also compared against the numpy restatement in ``oracle/`` and the
largest deviation is printed (the restatement must agree to a few ulps).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import covfuncs as ocf  # noqa: E402
from oracle import gp as ogp  # noqa: E402
from oracle import refshim  # noqa: E402
from tests.golden import cases as gcases  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_projections():
    """L2 projections onto piecewise-linear bases evaluated by the REAL reference (tests/golden/projections.npz):
    the reference's own test case (tests/linpde_gp/randprocs/crosscov/linfunctls/projections/test_matern_l2_projection.py:
    Matern-3/2, 7 nodes on [-1, 1], 50 points) in closed form AND through the generic quadrature class, other kernels /
    grids through the quadrature class, the covariance of two projections (dblquad), and a GP conditioned on projection
    observations followed by point observations."""
    pn, lg = refshim.load()
    from linpde_gp.randprocs.crosscov.linfunctls import projections as ref_proj
    from oracle import projections as oproj

    out, worst = {}, 0.0
    cases = {
        "m32_ref": dict(kernel={"kind": "matern", "input_shape": [], "nu": 1.5, "lengthscales": 1.0}, grid=np.linspace(-1.0, 1.0, 7),
                        zero_boundary=False, normalized=True, xs=np.linspace(-1.0, 1.0, 50)),
        "m32_zb": dict(kernel={"kind": "matern", "input_shape": [], "nu": 1.5, "lengthscales": 0.3},
                       grid=np.array([-0.5, -0.2, 0.0, 0.35, 0.6, 1.1]), zero_boundary=True, normalized=False,
                       xs=np.linspace(-0.8, 1.4, 23)),
        "m52": dict(kernel={"kind": "matern", "input_shape": [], "nu": 2.5, "lengthscales": 0.7}, grid=np.linspace(0.0, 2.0, 6),
                    zero_boundary=False, normalized=True, xs=np.linspace(-0.3, 2.3, 14)),
        "m12": dict(kernel={"kind": "matern", "input_shape": [], "nu": 0.5, "lengthscales": 0.5}, grid=np.linspace(0.0, 1.0, 5),
                    zero_boundary=True, normalized=True, xs=np.linspace(-0.2, 1.2, 15)),
        "eq": dict(kernel={"kind": "expquad", "input_shape": [], "lengthscales": 0.4}, grid=np.linspace(-1.0, 1.0, 6),
                   zero_boundary=False, normalized=True, xs=np.linspace(-1.2, 1.2, 13)),
    }
    for name, c in cases.items():
        k = ref_base(c["kernel"])
        basis = lg.functions.bases.UnivariateLinearInterpolationBasis(c["grid"], zero_boundary=c["zero_boundary"])
        proj = basis.l2_projection(normalized=c["normalized"])
        kPa = proj(k, argnum=1)
        val = np.asarray(kPa(c["xs"]))
        if c["kernel"]["kind"] == "matern" and c["kernel"]["nu"] == 1.5:
            assert isinstance(kPa, ref_proj.Matern32_L2Projection_UnivariateLinearInterpolationBasis)
            gen = ref_proj.CovarianceFunction_L2Projection_UnivariateLinearInterpolationBasis(k, proj, reverse=False)
            dev = np.max(np.abs(np.asarray(gen(c["xs"])) - val))
            print(f"projections {name}: closed form vs the reference's quadrature class {dev:.2e}")
        rev = np.asarray(proj(k, argnum=0)(c["xs"]))
        assert np.array_equal(rev, np.moveaxis(val, -1, 0))
        ob = oproj.Basis(c["grid"], c["zero_boundary"])
        oval = (oproj.crosscov_matern32(c["kernel"]["lengthscales"], ob, c["xs"], c["normalized"])
                if c["kernel"].get("nu") == 1.5 else oproj.crosscov_quad(c["kernel"], ob, c["xs"], c["normalized"]))
        worst = max(worst, np.max(np.abs(oval - val)) / np.max(np.abs(val)))
        out[f"{name}_kPa"] = val
        out[f"{name}_spec"] = np.frombuffer(json.dumps({**c, "grid": list(map(float, c["grid"])), "xs": list(map(float, c["xs"]))}).encode(), dtype=np.uint8)
    # covariance of two projections (dblquad in the reference): small bases
    for name, kernel, grid in (("m32", {"kind": "matern", "input_shape": [], "nu": 1.5, "lengthscales": 0.8}, np.linspace(-1.0, 1.0, 5)),
                               ("eq", {"kind": "expquad", "input_shape": [], "lengthscales": 0.6}, np.linspace(0.0, 1.0, 4))):
        k = ref_base(kernel)
        proj = lg.functions.bases.UnivariateLinearInterpolationBasis(grid, zero_boundary=False).l2_projection()
        C = np.asarray(proj(proj(k, argnum=1)).array)
        ob = oproj.Basis(grid, False)
        worst = max(worst, np.max(np.abs(oproj.covariance_dblquad(kernel, ob, ob) - C)) / np.max(np.abs(C)))
        out[f"PkP_{name}"] = C
        out[f"PkP_{name}_grid"] = grid
        out[f"PkP_{name}_kernel"] = np.frombuffer(json.dumps(kernel).encode(), dtype=np.uint8)
    # conditioning: projection observations first, point observations second (the order the reference supports)
    kernel = {"kind": "matern", "input_shape": [], "nu": 1.5, "lengthscales": 0.6}
    grid = np.linspace(-1.0, 1.0, 6)
    k = ref_base(kernel)
    prior = pn.randprocs.GaussianProcess(lg.functions.Constant(input_shape=(), value=0.3), 2.0 * k)
    proj = lg.functions.bases.UnivariateLinearInterpolationBasis(grid, zero_boundary=False).l2_projection()
    rng = np.random.default_rng(11)
    Yp = np.sin(2.0 * grid) + 0.3
    Xo = np.array([-0.75, 0.1, 0.55])
    Yo = np.sin(2.0 * Xo) + 0.3
    post1 = prior.condition_on_observations(Yp, L=proj)
    post2 = post1.condition_on_observations(Yo, X=Xo)
    Xt = np.linspace(-1.1, 1.1, 21)
    out.update(gp_kernel=np.frombuffer(json.dumps(kernel).encode(), dtype=np.uint8), gp_grid=grid, gp_Yp=Yp, gp_Xo=Xo, gp_Yo=Yo,
               gp_Xt=Xt, gp_mean1=np.asarray(post1.mean(Xt)), gp_var1=np.asarray(post1.cov(Xt, None)),
               gp_mean2=np.asarray(post2.mean(Xt)), gp_var2=np.asarray(post2.cov(Xt, None)),
               gp_gram2=np.asarray(post2.gram.todense()), gp_w2=np.asarray(post2.representer_weights))
    del rng
    np.savez(os.path.join(GOLDEN, "projections.npz"), **out)
    print(f"projections.npz: {len(out)} arrays, worst oracle-vs-reference deviation {worst:.2e}")
    return worst


# ----------------------------------------------------------------------------------------
# spec -> reference objects
# ----------------------------------------------------------------------------------------
def ref_base(base):
    from linpde_gp.randprocs import covfuncs

    kind = base["kind"]
    if kind == "tensor_product":
        return covfuncs.TensorProduct(*(ref_base(f) for f in base["factors"]))
    shape = tuple(base["input_shape"])
    ls = base["lengthscales"]
    ls = np.asarray(ls, dtype=float) if isinstance(ls, (list, tuple)) else float(ls)
    if kind == "matern":
        return covfuncs.Matern(shape, nu=base["nu"], lengthscales=ls)
    if kind == "expquad":
        return covfuncs.ExpQuad(shape, lengthscales=ls)
    raise ValueError(kind)


def ref_kernel(kernel):
    k = ref_base(kernel["base"])
    if kernel.get("scale") is not None:
        k = kernel["scale"] * k
    return k


def ref_op(L, input_shape):
    from linpde_gp.linfuncops import SumLinearFunctionOperator, diffops

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
    if len(summands) == 1:
        return summands[0]
    return SumLinearFunctionOperator(*summands)


def ref_L0kL1(spec):
    k = ref_kernel(spec["kernel"])
    shape = gcases.kernel_input_shape(spec["kernel"])
    L0, L1 = ref_op(spec["L0"], shape), ref_op(spec["L1"], shape)
    kk = L1(k, argnum=1) if L1 is not None else k
    return L0(kk, argnum=0) if L0 is not None else kk


# ----------------------------------------------------------------------------------------
def make_kernels():
    out = {}
    specs = gcases.build_cases()
    worst = 0.0
    for spec in specs:
        name = spec["name"]
        shape = gcases.kernel_input_shape(spec["kernel"])
        X = gcases.sobol_points(shape)
        X0, X1 = X[:32], X
        kref = ref_L0kL1(spec)
        K = np.asarray(kref.matrix(X0, X1))
        diag = np.asarray(kref(X0, None))
        # heat test-case style check of the un-flattened call too
        assert K.shape == (32, 128)
        o_op0, o_op1 = gcases.spec_to_oracle_op(spec["L0"]), gcases.spec_to_oracle_op(spec["L1"])
        Ko = ocf.matrix(spec["kernel"], o_op0, o_op1, X0, X1)
        do = ocf.diagonal(spec["kernel"], o_op0, o_op1, X0)
        scale = max(np.max(np.abs(K)), 1e-300)
        err = max(np.max(np.abs(K - Ko)) / scale, np.max(np.abs(diag - do)) / scale)
        worst = max(worst, err)
        print(f"{name:42s} type={type(kref).__name__:48s} max|K|={scale:10.3e} oracle-ref rel {err:.2e}")
        out[f"{name}__K"] = K
        out[f"{name}__diag"] = np.broadcast_to(diag, (32,)).copy()
    out["__specs__"] = np.frombuffer(json.dumps(specs).encode(), dtype=np.uint8)
    np.savez(os.path.join(GOLDEN, "kernels.npz"), **out)
    print(f"kernels.npz: {len(specs)} cases, worst oracle-vs-reference deviation {worst:.2e}")
    return worst


def make_kron():
    """Tensor-grid structure path: the reference's Kronecker linops on TensorProductGrids, densified and applied to a
    fixed block of vectors.  Derivative kernels with an explicit second grid go through ``.matrix`` (the reference's
    derivative ``linop(x0, x1)`` pairs x0's factors with themselves, diffops/_tensor_product.py:147-148)."""
    from linpde_gp.randprocs.covfuncs import TensorProductGrid

    from oracle import kron as okron

    out, worst = {}, 0.0
    specs = gcases.build_kron_cases()
    for spec in specs:
        kref = ref_L0kL1(spec)
        g0 = TensorProductGrid(*[np.asarray(f) for f in spec["factors0"]])
        g1 = None if spec["factors1"] is None else TensorProductGrid(*[np.asarray(f) for f in spec["factors1"]])
        plain = spec["L0"] is None and spec["L1"] is None
        if g1 is None or plain:
            op = kref.linop(g0, g1)
            K = np.asarray(op.todense())
        else:
            op = None
            K = np.asarray(kref.matrix(np.asarray(g0), np.asarray(g1)))
        V = np.random.default_rng(len(spec["name"])).standard_normal((K.shape[1], 3))
        KV = np.asarray(op @ V) if op is not None else K @ V
        o0, o1 = gcases.spec_to_oracle_op(spec["L0"]), gcases.spec_to_oracle_op(spec["L1"])
        terms = okron.kronecker_terms(spec["kernel"], o0, o1, spec["factors0"], spec["factors1"])
        Ko, KVo = okron.dense(terms), okron.matvec(terms, V)
        sc = np.max(np.abs(K))
        err = max(np.max(np.abs(K - Ko)) / sc, np.max(np.abs(KV - KVo)) / np.max(np.abs(KV)))
        worst = max(worst, err)
        print(f"{spec['name']:24s} linop={type(op).__name__:28s} shape={K.shape} terms={len(terms)} oracle-ref rel {err:.2e}")
        out[spec["name"] + "__K"], out[spec["name"] + "__V"], out[spec["name"] + "__KV"] = K, V, KV
    out["__specs__"] = np.frombuffer(json.dumps(specs).encode(), dtype=np.uint8)
    np.savez(os.path.join(GOLDEN, "kron.npz"), **out)
    print(f"kron.npz: {len(specs)} cases, worst oracle-vs-reference deviation {worst:.2e}")
    return worst


def make_gp():
    worst = 0.0
    for name, problem in ogp.golden_problems().items():
        res = run_reference_gp(problem)
        ores = ogp.solve(problem)
        for key in ("w", "mean", "var", "cov"):
            sc = max(np.max(np.abs(res[key])), 1e-300)
            err = np.max(np.abs(res[key] - ores[key])) / sc
            worst = max(worst, err)
            print(f"{name:28s} {key:5s} oracle-ref rel {err:.2e}")
        np.savez(
            os.path.join(GOLDEN, f"gp_{name}.npz"),
            problem=np.frombuffer(json.dumps(problem).encode(), dtype=np.uint8),
            **res,
        )
    return worst


def run_reference_gp(problem):
    """Run the unmodified reference: prior.condition_on_observations(...) block by block."""
    pn, lg = refshim.load()
    from linpde_gp import functions

    kernel = problem["kernel"]
    shape = gcases.kernel_input_shape(kernel)
    k = ref_kernel(kernel)
    prior = pn.randprocs.GaussianProcess(functions.Zero(input_shape=shape), k)
    post = prior
    for blk in problem["blocks"]:
        X = np.asarray(blk["X"], dtype=float)
        Y = np.asarray(blk["Y"], dtype=float)
        L = ref_op(blk["L"], shape)
        b = None
        if blk.get("noise_var") is not None:
            nv = np.asarray(blk["noise_var"], dtype=float)
            b = pn.randvars.Normal(np.zeros_like(Y), pn.linops.Scaling(np.broadcast_to(nv, Y.shape).copy()))
        post = post.condition_on_observations(Y, X=X, L=L, b=b)
    Xt = np.asarray(problem["Xt"], dtype=float)
    mean = np.asarray(post.mean(Xt))
    var = np.asarray(post.cov(Xt, None))
    Xc = Xt[: problem.get("n_cov", 16)]
    cov = np.asarray(post.cov.matrix(Xc))  # the path GaussianProcess.__call__ takes (pn _gaussian_process.py:75-79)
    w = np.asarray(post.representer_weights)
    gram = np.asarray(post.gram.todense())
    return {"w": w, "mean": mean, "var": var, "cov": cov, "gram": gram}


def run_reference_multi_output(problem):
    """The unmodified reference on a multi-output problem: IndependentMultiOutputCovarianceFunction prior,
    ``L = sum_t c_t * (D_t @ SelectOutput(o_t))`` observations, ``SelectOutput(j)(posterior)`` evaluated."""
    pn, lg = refshim.load()
    from linpde_gp import functions, linfuncops
    from linpde_gp.randprocs import covfuncs

    ks = [ref_kernel(k) for k in problem["kernels"]]
    shape = gcases.kernel_input_shape(problem["kernels"][0])
    nout = len(ks)
    prior = pn.randprocs.GaussianProcess(
        functions.StackedFunction(*(functions.Constant(shape, m) for m in problem["means"])),
        covfuncs.IndependentMultiOutputCovarianceFunction(*ks),
    )
    sel = [linfuncops.SelectOutput((shape, (nout,)), idx=j) for j in range(nout)]
    post = prior
    for blk in problem["blocks"]:
        if "functional" in blk:  # scalar observation: sum of Lebesgue integrals and point evaluations
            from linpde_gp import linfunctls

            L = None
            for a in blk["functional"]:
                if a[0] == "int":
                    t = a[2] * linfunctls.LebesgueIntegral((a[3], a[4])) @ sel[a[1]]
                else:
                    t = a[2] * sel[a[1]].to_linfunctl(a[3])
                L = t if L is None else L + t
            post = post.condition_on_observations(Y=float(blk["Y"][0]), L=L)
            continue
        X = np.asarray(blk["X"], dtype=float)
        Y = np.asarray(blk["Y"], dtype=float)
        L = None
        for o, c, op in blk["Ls"]:
            t = sel[o] if op is None else ref_op(op, shape) @ sel[o]
            if c != 1.0:
                t = c * t
            L = t if L is None else L + t
        b = None
        if blk.get("noise_var") is not None:
            nv = np.asarray(blk["noise_var"], dtype=float)
            b = pn.randvars.Normal(np.zeros_like(Y), pn.linops.Scaling(np.broadcast_to(nv, Y.shape).copy()))
        post = post.condition_on_observations(Y, X=X, L=L, b=b)
    Xt = np.asarray(problem["Xt"], dtype=float)
    Xc = Xt[: problem.get("n_cov", 8)]
    outs = [s(post) for s in sel]
    return {
        "w": np.asarray(post.representer_weights),
        "gram": np.asarray(post.gram.todense()),
        "mean": np.stack([np.asarray(o.mean(Xt)) for o in outs]),
        "var": np.stack([np.asarray(o.cov(Xt, None)) for o in outs]),
        "cov": np.stack([np.asarray(o.cov.matrix(Xc)) for o in outs]),
    }


def make_multi_output():
    from oracle import multi_output as omo

    worst = 0.0
    for name, problem in omo.golden_problems().items():
        res = run_reference_multi_output(problem)
        ores = omo.solve(problem)
        for key in ("w", "gram", "mean", "var", "cov"):
            sc = max(np.max(np.abs(res[key])), 1e-300)
            err = np.max(np.abs(res[key] - ores[key])) / sc
            worst = max(worst, err)
            print(f"mo_{name:25s} {key:5s} oracle-ref rel {err:.2e}")
        np.savez(
            os.path.join(GOLDEN, f"mo_{name}.npz"),
            problem=np.frombuffer(json.dumps(problem).encode(), dtype=np.uint8),
            **res,
        )
    return worst


INTEGRAL_NUS = (0.5, 1.5, 2.5, 3.5, 4.5)
INTEGRAL_LENGTHSCALES = (0.8, 1.1, 2.1)
INTEGRAL_DOMAINS = ((-2.2, -1.8), (-0.8, 0.5))
INTEGRAL_DOMAIN_PAIRS = (((-2.2, -1.8), (-0.8, 0.5)), ((-1.3, 0.0), (-0.2, 0.1)), ((0.25, 0.75), (0.3, 0.6)))


def make_integrals():
    """Lebesgue integrals of univariate half-integer Matern kernels on the reference's own test cases
    (tests/linpde_gp/randprocs/crosscov/linfunctls/cases/cases_integral_matern.py:11-44 and
    tests/linpde_gp/randprocs/cov/linfunctls/cases/cases_integral_matern.py:11-48)."""
    pn, lg = refshim.load()
    from linpde_gp import linfunctls
    from linpde_gp.randprocs.crosscov.linfunctls import integrals as ref_integrals
    from oracle import integrals as oint

    Lk = np.zeros((len(INTEGRAL_NUS), len(INTEGRAL_LENGTHSCALES), len(INTEGRAL_DOMAINS), 10))
    Xs = np.zeros((len(INTEGRAL_DOMAINS), 10))
    LkL = np.zeros((len(INTEGRAL_NUS), len(INTEGRAL_LENGTHSCALES), len(INTEGRAL_DOMAIN_PAIRS)))
    worst = 0.0
    for i, nu in enumerate(INTEGRAL_NUS):
        for j, ell in enumerate(INTEGRAL_LENGTHSCALES):
            k = pn.randprocs.covfuncs.Matern(input_shape=(), nu=nu, lengthscales=ell)
            for d, (a, b) in enumerate(INTEGRAL_DOMAINS):
                hw = (b - a) / 2
                Xs[d] = np.linspace(a - hw, b + hw, 10)
                L = linfunctls.LebesgueIntegral((a, b))
                kL, Lk_ = L(k, argnum=1), L(k, argnum=0)
                assert isinstance(kL, ref_integrals.UnivariateHalfIntegerMaternLebesgueIntegral)
                Lk[i, j, d] = np.asarray(kL(Xs[d]))
                assert np.array_equal(Lk[i, j, d], np.asarray(Lk_(Xs[d])))
                worst = max(worst, np.max(np.abs(Lk[i, j, d] - oint.matern_lebesgue_integral(int(nu - 0.5), ell, a, b, Xs[d]))))
            for d, (d0, d1) in enumerate(INTEGRAL_DOMAIN_PAIRS):
                L0, L1 = linfunctls.LebesgueIntegral(d0), linfunctls.LebesgueIntegral(d1)
                LkL[i, j, d] = float(np.asarray(L0(L1(k, argnum=1)).array if hasattr(L0(L1(k, argnum=1)), "array") else L0(L1(k, argnum=1))))
                worst = max(worst, abs(LkL[i, j, d] - oint.matern_lebesgue_integral_lebesgue_integral(int(nu - 0.5), ell, d0, d1)))
    np.savez(os.path.join(GOLDEN, "integrals.npz"), nus=np.array(INTEGRAL_NUS), lengthscales=np.array(INTEGRAL_LENGTHSCALES),
             domains=np.array(INTEGRAL_DOMAINS), domain_pairs=np.array(INTEGRAL_DOMAIN_PAIRS), X=Xs, Lk=Lk, LkL=LkL)
    print(f"integrals.npz: {Lk.size + LkL.size} values, worst oracle-vs-reference deviation {worst:.2e}")
    return worst


# ----------------------------------------------------------------------------------------
def make_seams():
    """Objects of the reference's Python seams evaluated by the REAL reference (tests/golden/seams.npz):

    * process-vector cross-covariances ``linfunctl(k, argnum)`` -- ``pv(x)`` for both ``argnum`` (crosscov/_pv_crosscov.py,
      crosscov/linfunctls/_evaluation.py:21-328) and the ``Covariance`` of two functionals ``L0(L1(k, argnum=1))``
      (:11-18; randvars/_covariance.py);
    * ``BlockMatrix2x2`` quantities (linops/_block.py:191-292): bordered Cholesky factor, Schur complement, ``L_A_inv_B``,
      ``schur_update``, solve, determinant, for a 2 + 1 nest of ExpQuad blocks (test_symmetric_block.py) and a Matern
      Gram matrix cut at 129 of 200."""
    pn, lg = refshim.load()
    from linpde_gp import linfunctls
    from linpde_gp.linops import BlockMatrix2x2

    rng = np.random.default_rng(20240)
    out = {"X": rng.uniform(0, 1, (37, 2)), "Xt": rng.uniform(0, 1, (5, 4, 2)), "X0": rng.uniform(0, 1, (3, 6, 2))}
    worst = 0.0
    for kname, kspec in gcases.SEAM_KERNELS.items():
        k = ref_kernel(kspec)
        for oname, ospec in gcases.SEAM_OPS.items():
            L = ref_op(ospec, (2,))
            fctl = (linfunctls._EvaluationFunctional((2,), (), out["X"]) if L is None  # pylint: disable=protected-access
                    else L.to_linfunctl(out["X"]))
            # (argnum=0 cannot be generated: the reference's CovarianceFunction_Evaluation_Identity reads a
            # non-existent `covfunc.output_shape`, crosscov/linfunctls/_evaluation.py:185, and raises AttributeError)
            pv1 = fctl(k, argnum=1)
            assert not pv1.reverse
            v1 = np.asarray(pv1(out["Xt"]))
            assert v1.shape == (5, 4, 37)
            out[f"pv__{kname}__{oname}__argnum1"] = v1
            out[f"pvlinop__{kname}__{oname}__argnum1"] = np.asarray(pv1.evaluate_linop(out["Xt"]).todense())
            o_op = gcases.spec_to_oracle_op(ospec)
            Ko = ocf.matrix(kspec, None, o_op, out["Xt"].reshape(-1, 2), out["X"])
            worst = max(worst, np.max(np.abs(v1.reshape(20, 37) - Ko)) / np.max(np.abs(Ko)))
            f0 = linfunctls._EvaluationFunctional((2,), (), out["X0"])  # pylint: disable=protected-access
            cov = f0(pv1)
            assert cov.shape0 == (3, 6) and cov.shape1 == (37,)
            out[f"cov__{kname}__{oname}"] = np.asarray(cov.array)
            out[f"covmat__{kname}__{oname}"] = np.asarray(cov.matrix)

    def spd(M):
        op = pn.linops.Matrix(M)
        op.is_symmetric = True
        op.is_positive_definite = True
        return op

    def block_case(name, K, cuts):
        """nested BlockMatrix2x2 over the index cuts (e.g. (2, 3) -> ((2 + 1) ...)), quantities of the outermost block"""
        A = spd(K[: cuts[0], : cuts[0]])
        lo = cuts[0]
        for hi in list(cuts[1:]) + [len(K)]:
            sbm = BlockMatrix2x2(A, pn.linops.Matrix(K[:lo, lo:hi]), None, spd(K[lo:hi, lo:hi]), is_spd=True)
            A, lo_prev, lo = sbm, lo, hi
        r = np.random.default_rng(len(K))
        u, v, B = r.standard_normal(lo_prev), r.standard_normal(len(K) - lo_prev), r.standard_normal((len(K), 3))
        out[f"blk__{name}__K"] = K
        out[f"blk__{name}__cuts"] = np.asarray(cuts)
        out[f"blk__{name}__chol"] = np.asarray(sbm.cholesky(True).todense())
        out[f"blk__{name}__schur"] = np.asarray(sbm.schur.todense())
        out[f"blk__{name}__LAinvB"] = np.asarray(sbm.L_A_inv_B.todense())
        out[f"blk__{name}__u"], out[f"blk__{name}__v"], out[f"blk__{name}__B"] = u, v, B
        out[f"blk__{name}__schur_update"] = np.asarray(sbm.schur_update(sbm.A.inv() @ u, v))
        out[f"blk__{name}__solve"] = np.asarray(sbm.solve(B))
        out[f"blk__{name}__inv_u"] = np.asarray(sbm.inv() @ np.concatenate([u, v]))
        out[f"blk__{name}__det"] = np.asarray(sbm.det())  # (underflows to 0 for the 200 x 200 case, as in the reference)
        return float(np.max(np.abs(out[f"blk__{name}__chol"] - np.linalg.cholesky(K))))

    x5 = np.arange(1.0, 6.0)
    e1 = block_case("expquad_nested5", np.exp(-0.5 * (x5[:, None] - x5[None, :]) ** 2), (2, 4))
    x3 = np.arange(1.0, 4.0)
    e2 = block_case("expquad3", np.exp(-0.5 * (x3[:, None] - x3[None, :]) ** 2), (2,))
    xs = np.sort(np.random.default_rng(7).uniform(0, 4, 200))
    r_ = np.abs(xs[:, None] - xs[None, :]) * (np.sqrt(5.0) / 0.5)
    e3 = block_case("matern200", (1 + r_ + r_ * r_ / 3) * np.exp(-r_) + 1e-6 * np.eye(200), (129,))
    np.savez(os.path.join(GOLDEN, "seams.npz"), **out)
    print(f"seams.npz: {len(out)} arrays, worst crosscov oracle-vs-reference deviation {worst:.2e}, "
          f"block Cholesky vs numpy {max(e1, e2, e3):.2e}")
    return worst


if __name__ == "__main__":
    refshim.load()
    if "--seams-only" in sys.argv:
        print(f"worst deviation: seams {make_seams():.2e}")
        sys.exit(0)
    if "--projections-only" in sys.argv:
        print(f"worst deviation: projections {make_projections():.2e}")
        sys.exit(0)
    if "--multi-output-only" in sys.argv:
        print(f"worst deviation: multi-output {make_multi_output():.2e}")
        sys.exit(0)
    if "--integrals-only" in sys.argv:
        print(f"worst deviation: integrals {make_integrals():.2e}, multi-output {make_multi_output():.2e}")
        sys.exit(0)
    w1 = make_kernels()
    w2 = make_gp()
    w3 = make_kron()
    w4 = make_multi_output()
    w5 = make_integrals()
    w6 = make_seams()
    w7 = make_projections()
    print(f"worst deviations: kernels {w1:.2e}, gp {w2:.2e}, kron {w3:.2e}, multi-output {w4:.2e}, integrals {w5:.2e}, "
          f"seams {w6:.2e}, projections {w7:.2e}")
