"""``ConditionalGaussianProcess``: a GP conditioned on batches of linear (PDE-collocation, boundary, plain
evaluation) observations, with a device-resident cached Cholesky factor of the Gram matrix that is EXTENDED
when another batch is added.

Mirrors src/linpde_gp/randprocs/_gaussian_process/_conditional.py (``from_observations`` :27-54,
``condition_on_observations`` :253-294, ``_preprocess_observations`` :296-399, ``Mean`` :177-197,
``CovarianceFunction`` :199-251, push-forwards :432-467).  Numerics: Gram blocks by the pairwise CUDA kernel,
bordered FP64 Cholesky + triangular solves on the DMMA path, matrix-free posterior mean, chunked TRSM posterior
variance -- all in ``liblpgp.so``.  Objects are immutable after construction (new conditionings copy the factor),
like the reference's functional API.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import _lowering, backend, functions, linfunctls, linops, randvars
from ..linfuncops import LinearFunctionOperator
from . import covfuncs
from ._gaussian_process import GaussianProcess

VAR_CHUNK_BYTES = 16 << 30  # upper bound of the cross-covariance workspace per chunk of test points (and <= free / 4)


def _descs(k: covfuncs.CovarianceFunction):
    """Device descriptors of a (possibly scaled sum of) kernel(s), one per summand without a common product form; the
    zero kernel has none.  Scalars distribute over sums (``sigma^2 * (k1 + k2)``, what ``apply_linfuncop`` and
    ``atom.coef * kk`` produce for sum priors -- pn/randprocs/covfuncs/_arithmetic_fallbacks.py:21-118)."""
    return covfuncs.device_descriptors(k)


def _gram_into(k: covfuncs.CovarianceFunction, X0, X1, out: "torch.Tensor", lower: bool = False, *,
               accumulate: bool = False, alpha: float = 1.0) -> None:
    """``out (+)= alpha * k(X0, X1)`` (``X1=None``: symmetric block, lower triangle when ``lower``): one pairwise-kernel
    launch per device descriptor, accumulating from the second on; blocks of the zero kernel are cleared."""
    descs = _descs(k)
    if not descs and not accumulate:
        out.zero_()
    for t, dsc in enumerate(descs):
        backend.gram(dsc, X0, X1, out=out, lower=lower, accumulate=accumulate or t > 0, alpha=alpha)


def _integral_terms(k: covfuncs.CovarianceFunction):
    """``[(scale, nu, lengthscale), ...]`` such that ``k = sum scale * Matern_nu(lengthscale)`` on scalar inputs -- the
    kernels whose Lebesgue integrals have closed forms (dispatch of
    src/linpde_gp/randprocs/covfuncs/linfunctls/_registry.py:25-41, 175-193 on Scaled / Sum / Matern kernels).  The
    reference integrates everything else numerically with scipy.integrate.quad (_covfunc_lebesgue.py:45-51); there
    is no host fallback here."""
    if isinstance(k, covfuncs.Zero):
        return []
    if isinstance(k, covfuncs.ScaledCovarianceFunction):
        return [(float(k.scalar) * s, nu, ell) for s, nu, ell in _integral_terms(k.covfunc)]
    if isinstance(k, covfuncs.SumCovarianceFunction):
        return [t for summand in k.summands for t in _integral_terms(summand)]
    if isinstance(k, covfuncs.Matern) and k.input_size == 1 and k.is_half_integer and k.output_shape_0 == () \
            and k.output_shape_1 == ():
        return [(1.0, k.nu, float(np.broadcast_to(k.lengthscales, (1,))[0]))]
    raise NotImplementedError(
        f"Lebesgue integral of {type(k).__name__}: closed forms exist for (sums of scaled) univariate half-integer "
        "Matern kernels only; the reference's scipy.integrate.quad fallback is host code and not provided"
    )


def _is_smooth(k: covfuncs.CovarianceFunction) -> bool:
    """Kernels that are analytic in both arguments (ExpQuad and its products / sums / scalings): Gauss-Legendre
    quadrature of their sections converges spectrally.  Matern kernels are not (kink / finite smoothness at x = x')."""
    if isinstance(k, covfuncs.Zero):
        return True
    if isinstance(k, covfuncs.ScaledCovarianceFunction):
        return _is_smooth(k.covfunc)
    if isinstance(k, covfuncs.SumCovarianceFunction):
        return all(_is_smooth(s) for s in k.summands)
    if isinstance(k, covfuncs.TensorProduct):
        return all(_is_smooth(f) for f in k.factors)
    return isinstance(k, covfuncs.ExpQuad)


def _proj_pts_block(k: covfuncs.CovarianceFunction, proj, X: "torch.Tensor", alpha: float = 1.0, op_pts=None,
                    op_proj=None) -> "torch.Tensor":
    """Device matrix ``alpha * Cov((op_pts f)(X_i), P[op_proj f]_j)`` (n x m) for ``f ~ GP(., k)`` and an L2 projection
    ``P`` onto hat functions, normaliser included.  Half-integer Matern kernels without operators: closed form
    (``lpgp_matern_hat_integral``; crosscov/linfunctls/projections.py:129-170 for every nu); kernels that are smooth
    inside the elements: Gauss-Legendre nodes ``T`` per element, ``K(X, T)`` by the Gram kernel, times the
    (basis x nodes) weight matrix on the DMMA path (the reference: scipy.integrate.quad per point and basis
    function, projections.py:47-66)."""
    m = len(proj.basis)
    out = backend.alloc_matrix(X.shape[0], m)
    terms = None
    if op_pts is None and op_proj is None:
        try:
            terms = _integral_terms(k)
        except NotImplementedError:
            terms = None
    if terms is not None:
        if not terms:
            out.zero_()
        for t, (scale, nu, ell) in enumerate(terms):
            backend.matern_hat_integral(_lowering.matern_integral_desc(nu, ell), proj.device_grid(), m,
                                        not proj.basis.zero_boundary, X, out, alpha=alpha * scale, accumulate=t > 0)
    else:
        if not _is_smooth(k):
            raise NotImplementedError(
                f"L2 projection of {type(k).__name__} composed with an operator: closed forms exist for plain half-integer "
                "Matern kernels, quadrature only for kernels that are smooth inside the elements")
        kk = k if op_proj is None else op_proj(k, argnum=1)
        kk = kk if op_pts is None else op_pts(kk, argnum=0)
        T, W, _ = proj.device_quadrature()
        K = backend.alloc_matrix(X.shape[0], T.shape[0])
        _gram_into(kk, X, T, K)
        backend.gemm_nt(K, W, out, alpha, 0.0)
    return proj.normalize_rows(out)


def _proj_proj_block(k: covfuncs.CovarianceFunction, projA, projB, alpha: float = 1.0, opA=None, opB=None) -> "torch.Tensor":
    """``alpha * Cov(P_A[op_A f], P_B[op_B f])`` (m_A x m_B): the inner projection in closed form / by quadrature at the
    Gauss-Legendre nodes of P_A's elements (the section ``s -> Cov(f(s), P_B f)`` is smooth INSIDE the elements when
    both projections share their nodes), then the outer quadrature as a DMMA GEMM (the reference: scipy dblquad per
    entry, projections.py:77-108)."""
    if not np.all(np.isin(projA.basis.elements(), projB.basis.grid)) and not _is_smooth(k):
        raise NotImplementedError("covariance of two L2 projections of a non-smooth kernel on different grids")
    T, W, _ = projA.device_quadrature()
    inner = _proj_pts_block(k, projB, T, alpha, op_pts=opA, op_proj=opB)  # (Q_A x m_B)
    C = backend.alloc_matrix(W.shape[0], inner.shape[1])
    backend.gemm_nn(W, inner, C, 1.0, 0.0)
    if projA.normalized:
        Ct = backend.alloc_matrix(C.shape[1], C.shape[0])
        Ct.copy_(C.T)
        projA.normalize_rows(Ct)
        C.copy_(Ct.T)
    return C


class _Atom:
    """One summand of an observation functional: ``coef * (op f)(X)`` (``kind == "pts"``),
    ``coef * int_a^b (op f)(t) dt`` (``kind == "int"``, one row) or ``coef * P[op f]`` for an L2 projection onto m hat
    functions (``kind == "proj"``, m rows)."""

    def __init__(self, coef, kind, op, payload, d: int):
        self.coef, self.kind, self.op = float(coef), kind, op
        self.proj = None
        if kind == "proj":
            if d != 1:
                raise NotImplementedError("L2 projections need a univariate input domain")
            self.X_host = self.X = self.dom = None
            self.proj = payload
            self.n = len(payload.basis)
        elif kind == "pts":
            self.X_host = np.asarray(payload, dtype=np.double)
            self.X = backend.points(self.X_host, d)
            self.n = self.X.shape[0]
            self.dom = None
        else:
            if d != 1:
                raise NotImplementedError("integral observations need a univariate input domain")
            self.X_host = self.X = None
            self.n = 1
            self.dom = (float(payload[0]), float(payload[1]))

    def apply(self, k, argnum: int):
        return k if self.op is None else self.op(k, argnum=argnum)


def _atom_cov_into(k: covfuncs.CovarianceFunction, A: _Atom, B: _Atom, out: "torch.Tensor", accumulate: bool) -> None:
    """``out (+)= coef_A coef_B cov(atom_A f, atom_B f)`` for ``f ~ GP(., k)`` (out: n_A x n_B view of a row-major
    device matrix)."""
    alpha = A.coef * B.coef
    if A.kind == "proj" or B.kind == "proj":
        if "int" in (A.kind, B.kind):
            raise NotImplementedError("covariance between an integral and an L2-projection observation")
        if A.kind == "proj" and B.kind == "proj":
            blk = _proj_proj_block(k, A.proj, B.proj, alpha, A.op, B.op)
        elif B.kind == "proj":
            blk = _proj_pts_block(k, B.proj, A.X, alpha, op_pts=A.op, op_proj=B.op)
        else:
            blk = _proj_pts_block(k, A.proj, B.X, alpha, op_pts=B.op, op_proj=A.op).T
        if accumulate:
            out.add_(blk)
        else:
            out.copy_(blk)
        return
    kk = A.apply(B.apply(k, 1), 0)
    if A.kind == "pts" and B.kind == "pts":
        _gram_into(kk, A.X, B.X, out, accumulate=accumulate, alpha=alpha)
        return
    terms = _integral_terms(kk)
    if not terms and not accumulate:
        out.zero_()
    ld = out.stride(0) if out.shape[0] > 1 else 1
    for t, (scale, nu, ell) in enumerate(terms):
        dsc = _lowering.matern_integral_desc(nu, ell)
        acc = accumulate or t > 0
        if A.kind == "int" and B.kind == "int":
            backend.matern_integral2(dsc, A.dom, B.dom, out, alpha=alpha * scale, accumulate=acc)
        elif A.kind == "int":  # one row: int_a^b k(t, X_B) dt
            backend.matern_integral(dsc, A.dom[0], A.dom[1], B.X, out, out_stride=1, alpha=alpha * scale, accumulate=acc)
        else:  # one column
            backend.matern_integral(dsc, B.dom[0], B.dom[1], A.X, out, out_stride=ld, alpha=alpha * scale, accumulate=acc)


def _block_cov_into(k, blk: "_Block", pb: "_Block", out: "torch.Tensor") -> None:
    """``out <- cov(block, pb)`` as the sum over all atom pairs (n_blk x n_pb)."""
    first = True
    for A in blk.atoms:
        for B in pb.atoms:
            _atom_cov_into(k, A, B, out, accumulate=not first)
            first = False


def _assemble_kronecker(k: covfuncs.CovarianceFunction, grid0, grid1, out: "torch.Tensor", lower: bool) -> bool:
    """Tensor-grid structure path: if both point sets are intact TensorProductGrids and
    ``k.linop`` yields a sum of Kronecker products, densify it with the Kronecker assembly kernel (one multiply-add
    per term and entry, HBM-write bound) instead of evaluating the kernel pair by pair.  ``grid1=None`` = symmetric
    block.  Returns False when the structure is not available (the caller then uses the pairwise Gram kernel)."""
    if grid0 is None or (grid1 is None and not lower):  # off-diagonal blocks need BOTH grids
        return False
    try:
        op = k.linop(grid0, grid1)
    except NotImplementedError:
        return False
    terms = op.kron_terms()
    if terms is None:
        return False
    backend.kron_sum(terms, out=out, lower=lower)
    return True


class _Block:
    """One observation batch: points, operator, logical / physical (even-padded) size, offset in the factor.

    ``atoms`` (``LinearFunctional._atoms()``) generalises the batch to a SUM of point-evaluation and integral atoms
    (``simple`` is False then and ``X`` / ``op`` are unset); plain ``L[f](X)`` batches keep the single-atom fast paths
    (Kronecker assembly, one-shot / multi-GPU assembly)."""

    def __init__(self, X_host, op: Optional[LinearFunctionOperator], d: int, col_off: int, atoms=None):
        self.col_off = col_off
        if atoms is not None and not (len(atoms) == 1 and atoms[0][1] == "pts" and atoms[0][0] == 1.0):
            self.simple = False
            self.grid = self.X_host = self.X = self.op = None
            self.atoms = [_Atom(c, kind, aop, payload, d) for c, kind, aop, payload in atoms]
            self.n = self.atoms[0].n
            if any(a.n != self.n for a in self.atoms):
                raise ValueError("all summands of an observation functional must produce the same number of rows")
            self.n_phys = self.n + (self.n % 2)
            return
        if atoms is not None:
            _, _, op, X_host = atoms[0]
        self.simple = True
        # intact TensorProductGrid: Gram blocks against other gridded batches are sums of Kronecker products
        self.grid = X_host if covfuncs._grid_factors(X_host) is not None else None  # pylint: disable=protected-access
        X_host = np.asarray(X_host, dtype=np.double)
        self.X_host = X_host
        self.op = op
        self.X = backend.points(X_host, d)
        self.n = self.X.shape[0]
        self.n_phys = self.n + (self.n % 2)
        atom = _Atom.__new__(_Atom)
        atom.coef, atom.kind, atom.op, atom.X_host, atom.X, atom.n, atom.dom = 1.0, "pts", op, X_host, self.X, self.n, None
        atom.proj = None
        self.atoms = [atom]


STRUCTURED_MIN_N = int(__import__("os").environ.get("LPGP_STRUCTURED_MIN_N", "20000"))  # below: dense factor (general)


class KroneckerFactor:
    """Cached factor of a Gram matrix that IS a Kronecker product, ``G = alpha * K_1 (x) K_2`` -- plain observations of a
    (scaled) two-factor ``TensorProduct`` kernel on an intact ``TensorProductGrid`` (the reference keeps
    this Gram matrix a lazy ``pn.linops.Kronecker``, covfuncs/_tensor_product.py:64-82, whose ``solve`` / ``cholesky``
    act on the factors, pn/linops/_kronecker.py:122-140, 233-242).  Nothing of size N x N is ever formed:

    * representer weights: two multi-right-hand-side Cholesky solves with the n_1 x n_1 and n_2 x n_2 factors;
    * pointwise variance: ``k(X, x) = alpha k_1(X_1, x_1) (x) k_2(X_2, x_2)`` and ``L = sqrt(alpha) L_1 (x) L_2`` give
      ``|L^{-1} k(X, x)|^2 = alpha q_1(x_1) q_2(x_2)`` with ``q_i = |L_i^{-1} k_i(X_i, x_i)|^2`` -- O(M (n_1^2 + n_2^2))
      instead of O(M N^2);
    * covariance blocks: Hadamard product of the per-dimension ``V_i V_i'^T``.

    N = n_1 n_2 can exceed what a dense factor could hold (1024 x 1024 grid: a dense FP64 Gram matrix would be 8.8 TB)."""

    structured = True

    def __init__(self, alpha: float, kernels, grids):
        from .. import linops as _linops  # pylint: disable=import-outside-toplevel

        self.alpha = float(alpha)
        self.kernels = tuple(kernels)
        self.grids = tuple(backend.points(np.asarray(g, dtype=np.double), 1) for g in grids)
        self.sizes = tuple(int(g.shape[0]) for g in self.grids)
        self.n = int(np.prod(self.sizes))
        ops = []
        for k, g in zip(self.kernels, self.grids):
            Ki = backend.alloc_matrix(g.shape[0], g.shape[0])
            _gram_into(k, g, None, Ki, lower=True)
            backend.symmetrize_lower(Ki)
            op = _linops._Device(Ki)  # pylint: disable=protected-access
            op.is_symmetric = True
            ops.append(op)
        self.op = _linops.Kronecker(ops[0], ops[1])
        self.chol = [o.cholesky(True) for o in ops]  # small dense factors (LinAlgError if a factor is not SPD)

    def potrs(self, B: "torch.Tensor") -> "torch.Tensor":
        if B.dim() == 1:
            B = B.reshape(1, -1)
        B.copy_(self.op._solve_rows_device(B))  # pylint: disable=protected-access
        return B.mul_(1.0 / self.alpha)

    def logdet(self) -> float:
        return self.op.logabsdet() + self.n * float(np.log(self.alpha))

    def _dim_rows(self, i: int, x: "torch.Tensor") -> "torch.Tensor":
        """``V_i = k_i(x, X_i) L_i^{-T}`` for the i-th coordinates ``x`` (M,) of the test points: (M x n_i)."""
        Ki = backend.alloc_matrix(x.shape[0], self.sizes[i])
        _gram_into(self.kernels[i], x.reshape(-1, 1).contiguous(), self.grids[i], Ki)
        # (odd n_i: the small factor is padded by an identity row and the solve returns a gathered copy -> re-align for TMA)
        return linops._aligned(self.chol[i]._solve_rows_device(Ki))  # pylint: disable=protected-access

    def post_var(self, Xt: "torch.Tensor", prior_diag: float) -> "torch.Tensor":
        q = None
        for i in range(2):
            qi = backend.row_sumsq(self._dim_rows(i, Xt[:, i].contiguous()))
            q = qi if q is None else q * qi
        return prior_diag - self.alpha * q

    def post_cov_sub(self, X0: "torch.Tensor", X1, C: "torch.Tensor") -> None:
        """``C -= alpha * prod_i V_i(X0) V_i(X1)^T`` (Hadamard product over the dimensions)."""
        H = None
        for i in range(2):
            V0 = self._dim_rows(i, X0[:, i].contiguous())
            V1 = V0 if X1 is None else self._dim_rows(i, X1[:, i].contiguous())
            Gi = backend.alloc_matrix(*C.shape)
            backend.gemm_nt(V0, V1, Gi, 1.0, 0.0)
            H = Gi if H is None else H.mul_(Gi)
        C.sub_(H.mul_(self.alpha))

    def extended(self, new_size: int):
        raise NotImplementedError(
            "a posterior with a Kronecker-structured Gram factor cannot be extended by bordering (the bordered matrix is no "
            "Kronecker product); condition on the gridded batch LAST, or set LPGP_STRUCTURED_MIN_N above its size to "
            "use the dense appendable factor")


def _kronecker_structure(prior, atoms, noise):
    """``(alpha, [k_1, k_2], [grid_1, grid_2])`` if the Gram matrix of this single batch is ``alpha K_1 (x) K_2``."""
    if noise is not None or len(atoms) != 1:
        return None
    coef, kind, op, payload = atoms[0]
    if kind != "pts" or op is not None or coef != 1.0:
        return None
    grids = covfuncs._grid_factors(payload)  # pylint: disable=protected-access
    k, alpha = prior.cov, 1.0
    while isinstance(k, covfuncs.ScaledCovarianceFunction):
        alpha *= float(k.scalar)
        k = k.covfunc
    if grids is None or len(grids) != 2 or not isinstance(k, covfuncs.TensorProduct) or len(k.factors) != 2:
        return None
    if int(np.prod([len(g) for g in grids])) < STRUCTURED_MIN_N or alpha <= 0.0:
        return None
    return alpha, list(k.factors), list(grids)


class _PosteriorState:
    """Everything a conditioned process needs to evaluate itself: prior, observation blocks, the device-resident
    factor, residuals and representer weights.  ``ConditionalGaussianProcess``, its ``Mean`` and its
    ``CovarianceFunction`` all point HERE and the state points back at none of them, so the object graph has no
    reference cycle: the device buffers (34 GB at N = 64k) are released by reference counting the moment the last
    posterior object using them goes away, not at some later pass of the cyclic garbage collector (which used to
    leave several dead factors resident between conditioning steps)."""

    def __init__(self, *, prior, base_prior, Ys, Ls, bs, blocks, factor, resid, weights, test_op):
        self._prior = prior
        self._base_prior = base_prior  # the process the observations refer to
        self._Ys = tuple(Ys)
        self._Ls = tuple(Ls)
        self._bs = tuple(bs)
        self._blocks = tuple(blocks)
        self._factor = factor
        self._resid = resid
        self._w = weights
        self._test_op = test_op
        self._gram_op = None

    # -- state ---------------------------------------------------------------------------------------------
    @property
    def _logical_index(self) -> np.ndarray:
        return np.concatenate([np.arange(b.col_off, b.col_off + b.n) for b in self._blocks])

    @property
    def gram(self) -> "GramFactorOperator":
        if getattr(self._factor, "distributed", False):
            raise NotImplementedError("the Gram factor is distributed over several GPUs (replicate=False)")
        if getattr(self._factor, "structured", False):  # the lazy Kronecker product itself (solve / cholesky on the factors)
            return self._factor.alpha * self._factor.op
        if self._gram_op is None:
            self._gram_op = GramFactorOperator(self._factor, self._logical_index)
        return self._gram_op

    # -- kernels between the test side and one atom of an observation block ----------------------------------------
    def _k_test_obs(self, atom: _Atom):
        kk = atom.apply(self._base_prior.cov, 1)
        return kk if self._test_op is None else self._test_op(kk, argnum=0)

    def _obs_blocks(self) -> backend.ObsBlocks:
        """Device view of ``k(x_test, observations)``: one entry per (point atom, kernel descriptor) -- entries on the
        same columns accumulate -- plus the integral atoms (``extras``), whose single column is the closed-form
        ``int_a^b k(x_test, t) dt`` evaluated by ``lpgp_matern_integral``."""
        descs, Xs, offs, extras = [], [], [], []
        for blk in self._blocks:
            for atom in blk.atoms:
                if atom.kind == "proj":  # dense block of m columns: Cov((test_op f)(x), P f)
                    extras.append((blk.col_off, atom.n, self._proj_fill(atom)))
                    continue
                kk = self._k_test_obs(atom)
                if atom.kind == "int":
                    terms = [(atom.coef * sc, _lowering.matern_integral_desc(nu, ell)) for sc, nu, ell in _integral_terms(kk)]
                    if terms:
                        extras.append((blk.col_off, atom.dom, terms))
                    continue
                if atom.coef != 1.0:
                    kk = atom.coef * kk
                for dsc in _descs(kk):
                    descs.append(dsc)
                    Xs.append(atom.X)
                    offs.append(blk.col_off)
        return backend.ObsBlocks(descs, Xs, offs, extras=extras)

    def _proj_fill(self, atom: _Atom):
        k, test_op = self._base_prior.cov, self._test_op
        return lambda Xt: _proj_pts_block(k, atom.proj, Xt, atom.coef, op_pts=test_op, op_proj=atom.op)

    def _obs_blocks_unique(self) -> backend.ObsBlocks:
        """Entries for the cross-covariance workspace ``k(x_test, X_obs)``: consecutive entries on the same columns
        (sum kernels) are accumulated by ``lpgp_crosscov``, columns nobody covers (zero kernels) are cleared."""
        return self._obs_blocks()

    @property
    def _is_multi_output(self) -> bool:
        return self._prior.output_shape != ()

    def _select(self, j: int) -> "ConditionalGaussianProcess":
        from ..linfuncops import SelectOutput

        return self._apply_linfuncop(SelectOutput((self._prior.input_shape, self._prior.output_shape), idx=j))

    def _apply_linfuncop(self, L: LinearFunctionOperator) -> "ConditionalGaussianProcess":
        if self._test_op is not None:
            raise NotImplementedError("composition of two operators on a conditioned process")
        return ConditionalGaussianProcess(
            prior=L(self._prior), Ys=self._Ys, Ls=self._Ls, bs=self._bs, blocks=self._blocks, factor=self._factor,
            resid=self._resid, weights=self._w, test_op=L, base_prior=self._base_prior,
        )

    def _var_distributed(self, Xt: "torch.Tensor", prior_diag: float) -> "torch.Tensor":
        """Pointwise variance of THIS rank's test points with a distributed factor (collective)."""
        return self._factor.post_var(self._obs_blocks_unique(), Xt, prior_diag, min_chunk_bytes=4 << 30)


class ConditionalGaussianProcess(GaussianProcess):
    @classmethod
    def from_observations(cls, prior: GaussianProcess, Y, X=None, *, L=None, b=None):
        Y, Lf, b, atoms, resid, noise = cls._preprocess_observations(prior=prior, Y=Y, X=X, L=L, b=b)
        blk = _Block(None, None, prior.cov.input_size, 0, atoms=atoms)
        structure = _kronecker_structure(prior, atoms, noise)
        if structure is not None:  # Gram matrix = alpha K_1 (x) K_2: factor the two small matrices only
            with backend.phase("factor"):
                factor = KroneckerFactor(*structure)
            y = backend.to_device(resid).clone()
            with backend.phase("solve"):
                w = factor.potrs(y.clone().reshape(1, -1)).reshape(-1)
            return cls(prior=prior, Ys=(Y,), Ls=(Lf,), bs=(b,), blocks=(blk,), factor=factor, resid=y, weights=w)
        factor = backend.DeviceFactor([blk.n_phys])
        with backend.phase("assemble"):
            cls._assemble_rows(prior, [], blk, factor, noise)
        with backend.phase("factor"):
            factor.potrf()
        y = torch.zeros(factor.n, dtype=torch.float64, device=factor.L.device)
        y[: blk.n].copy_(backend.to_device(resid))
        with backend.phase("solve"):
            w = factor.potrs(y.clone().reshape(1, -1)).reshape(-1)
        return cls(prior=prior, Ys=(Y,), Ls=(Lf,), bs=(b,), blocks=(blk,), factor=factor, resid=y, weights=w)

    @classmethod
    def from_observation_batches(cls, prior: GaussianProcess, batches, *, process_group=None, nb: int = 1024,
                                 replicate: bool = True):
        """Condition on several observation batches AT ONCE: ``batches`` is a sequence of
        ``(Y, X, L, b)`` tuples (same meaning as the arguments of :meth:`condition_on_observations`).

        The posterior is identical to conditioning batch by batch (the bordered factor of a block matrix IS the
        Cholesky factor of the whole matrix), but the Gram matrix is assembled and factorised in one go -- and,
        when ``torch.distributed`` is initialised with more than one rank, across all GPUs of the process group
        (block-row cyclic layout, NCCL panel exchange, see ``linpde_gp_b200/distributed.py``).  Every rank must
        call this with the same arguments.  ``replicate=True``: every rank ends up with the full factor, so that
        posterior evaluation can shard test points freely and without further communication.  ``replicate=False``:
        the factor stays distributed (N beyond one GPU's memory); ``mean`` works as usual, ``var`` / ``cov`` become
        COLLECTIVE calls in which every rank passes its own shard of test points and block rows of the factor are
        streamed over NVLink (``DistributedCholesky.solve_rows``)."""
        import torch.distributed as dist

        from .. import distributed

        pre = [cls._preprocess_observations(prior=prior, Y=t[0], X=t[1] if len(t) > 1 else None,
                                            L=t[2] if len(t) > 2 else None, b=t[3] if len(t) > 3 else None)
               for t in batches]
        d = prior.cov.input_size
        blocks, off = [], 0
        for (_, _, _, atoms, _, _) in pre:
            blk = _Block(None, None, d, off, atoms=atoms)
            if not blk.simple:
                raise NotImplementedError(
                    "one-shot / multi-GPU conditioning takes plain `L[f](X)` batches; add functional (integral, sum) "
                    "observations afterwards with `condition_on_observations`")
            blocks.append(blk)
            off += blk.n_phys
        n = off
        noises = [p[5] for p in pre]
        world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        if world > 1 or not replicate:
            ch = distributed.DistributedCholesky(n, nb=nb, group=process_group)
            with backend.phase("assemble"):
                for i in ch.layout.local_blocks(ch.rank):
                    g0, g1 = ch.layout.block_bounds(i)
                    cls._assemble_range(prior, blocks, noises, ch.local_block_rows(i), g0, g1)
            with backend.phase("factor"):
                if replicate:
                    factor = backend.DeviceFactor([n], reserve_rows=0)
                    ch.factor(factor.L)  # the gathered panels are the block columns of L: replicated on the fly
                    factor.dinv[: ch.dinv.numel() - 8].copy_(ch.dinv[:-8])
                    del ch
                else:
                    ch.factor()
                    factor = distributed.DistributedFactor(ch)
        else:
            factor = backend.DeviceFactor([n])
            with backend.phase("assemble"):
                cls._assemble_range(prior, blocks, noises, factor.L, 0, n)
            with backend.phase("factor"):
                factor.potrf()
        factor.factored_segments = 1
        y = torch.zeros(n, dtype=torch.float64, device=backend._require_cuda())  # pylint: disable=protected-access
        for blk, p in zip(blocks, pre):
            y[blk.col_off : blk.col_off + blk.n].copy_(backend.to_device(p[4]))
        with backend.phase("solve"):
            w = factor.potrs(y.clone().reshape(1, -1)).reshape(-1)
        return cls(prior=prior, Ys=tuple(p[0] for p in pre), Ls=tuple(p[1] for p in pre), bs=tuple(p[2] for p in pre),
                   blocks=tuple(blocks), factor=factor, resid=y, weights=w)

    @staticmethod
    def _assemble_range(prior, blocks, noises, out: "torch.Tensor", g0: int, g1: int) -> None:
        """Fill ``out`` (rows g0..g1 of the global Gram matrix, all n columns addressable) with the lower part
        G[g0:g1, 0:g1] of the block-structured Gram matrix of all batches (identity padding rows included)."""
        k = prior.cov
        for bi, blk in enumerate(blocks):
            r_lo, r_hi = max(g0, blk.col_off), min(g1, blk.col_off + blk.n_phys)
            if r_lo >= r_hi:
                continue
            rows = out[r_lo - g0 : r_hi - g0]
            l_lo, l_hi = r_lo - blk.col_off, min(r_hi - blk.col_off, blk.n)  # logical rows of this batch
            Xi = blk.X[l_lo:l_hi]
            for bj, pb in enumerate(blocks[: bi + 1]):
                kj = k if pb.op is None else pb.op(k, argnum=1)
                kij = kj if blk.op is None else blk.op(kj, argnum=0)
                c_hi = pb.n if bj < bi else min(pb.n, l_hi)  # own batch: columns up to the last row's diagonal
                if l_hi > l_lo and c_hi > 0:
                    _gram_into(kij, Xi, pb.X[:c_hi], rows[: l_hi - l_lo, pb.col_off : pb.col_off + c_hi])
                if pb.n_phys != pb.n and pb.col_off + pb.n < r_hi:
                    rows[:, pb.col_off + pb.n] = 0.0  # padding column of batch bj
            if noises[bi] is not None and l_hi > l_lo:
                kind, val = noises[bi]
                diag = rows[: l_hi - l_lo, blk.col_off + l_lo : blk.col_off + l_hi]
                if kind == "diag":
                    backend.add_diag(diag, backend.to_device(val[l_lo:l_hi]), 1.0)
                elif kind == "kernel":
                    Xn = backend.points(val._x0, val.covfunc.input_size)  # pylint: disable=protected-access
                    val.assemble_into(rows[: l_hi - l_lo, blk.col_off : blk.col_off + l_hi], X0=Xn[l_lo:l_hi], X1=Xn[:l_hi],
                                      accumulate=True)
                else:
                    diag.add_(backend.to_device(val[l_lo:l_hi, l_lo:l_hi]))
                    if l_lo > 0:
                        rows[: l_hi - l_lo, blk.col_off : blk.col_off + l_lo].add_(backend.to_device(val[l_lo:l_hi, :l_lo]))
            if blk.n_phys != blk.n and r_hi == blk.col_off + blk.n_phys:  # the identity padding row is in range
                pr = rows[blk.n - l_lo]
                pr[: blk.col_off + blk.n_phys] = 0.0
                pr[blk.col_off + blk.n] = 1.0

    def __init__(self, *, prior, Ys, Ls, bs, blocks, factor, resid, weights, test_op=None, base_prior=None):
        self._state = _PosteriorState(prior=prior, base_prior=prior if base_prior is None else base_prior, Ys=Ys, Ls=Ls,
                                      bs=bs, blocks=blocks, factor=factor, resid=resid, weights=weights, test_op=test_op)
        super().__init__(
            mean=ConditionalGaussianProcess.Mean(self._state),
            cov=ConditionalGaussianProcess.CovarianceFunction(self._state),
        )

    # -- state (lives in ``_PosteriorState``; see there why) ------------------------------------------------------
    _prior = property(lambda self: self._state._prior)
    _base_prior = property(lambda self: self._state._base_prior)
    _Ys = property(lambda self: self._state._Ys)
    _Ls = property(lambda self: self._state._Ls)
    _bs = property(lambda self: self._state._bs)
    _blocks = property(lambda self: self._state._blocks)
    _factor = property(lambda self: self._state._factor)
    _resid = property(lambda self: self._state._resid)
    _w = property(lambda self: self._state._w)
    _test_op = property(lambda self: self._state._test_op)
    _logical_index = property(lambda self: self._state._logical_index)
    _is_multi_output = property(lambda self: self._state._is_multi_output)

    @property
    def gram(self) -> "GramFactorOperator":
        return self._state.gram

    @property
    def representer_weights(self) -> np.ndarray:
        return self._w.cpu().numpy()[self._logical_index]

    def _obs_blocks(self) -> backend.ObsBlocks:
        return self._state._obs_blocks()

    def _obs_blocks_unique(self) -> backend.ObsBlocks:
        return self._state._obs_blocks_unique()

    def _select(self, j: int) -> "ConditionalGaussianProcess":
        return self._state._select(j)

    # -- posterior mean / covariance -----------------------------------------------------------------------------
    class Mean(functions.Function):
        def __init__(self, post: "_PosteriorState"):
            self._post = post
            super().__init__(input_shape=post._prior.mean.input_shape, output_shape=post._prior.mean.output_shape)

        def _evaluate(self, x: np.ndarray) -> np.ndarray:
            post = self._post
            if post._is_multi_output:  # one matrix-free pass per output (outputs share the representer weights)
                return np.stack([post._select(j).mean(x) for j in range(post._prior.output_shape[0])], axis=-1)
            batch = x.shape[: x.ndim - self.input_ndim]
            m_x = post._prior.mean(x)
            blocks = post._obs_blocks()
            if blocks.empty:  # no observation is correlated with this (output of the) process
                return m_x
            Xt = backend.points(x, post._base_prior.cov.input_size)
            with backend.phase("mean"):
                upd = backend.post_mean(blocks, post._w, Xt)
            return m_x + upd.cpu().numpy().reshape(batch)

    class CovarianceFunction(covfuncs.CovarianceFunction):
        def __init__(self, post: "_PosteriorState"):
            self._post = post
            super().__init__(post._prior.cov.input_shape)

        @property
        def output_shape_0(self):
            return self._post._prior.cov.output_shape_0

        @property
        def output_shape_1(self):
            return self._post._prior.cov.output_shape_1

        def _require_scalar(self):
            if self._post._is_multi_output:
                raise NotImplementedError(
                    "posterior covariance of a multi-output process: apply `SelectOutput` to the posterior first "
                    "(the reference's un-selected posterior covariance is not evaluable either)"
                )

        def _evaluate(self, x0, x1, batch):
            post = self._post
            self._require_scalar()
            d = self.input_size
            if x1 is None:  # pointwise variance
                Xt = backend.points(x0, d)
                prior_diag = _descs(post._prior.cov)
                diag = sum(dsc.diag_value for dsc in prior_diag)
                n = post._factor.n
                if post._obs_blocks().empty:
                    return np.full(batch, diag)
                if getattr(post._factor, "distributed", False):
                    with backend.phase("var"):
                        var = post._var_distributed(Xt, diag)
                    return var.cpu().numpy().reshape(batch)
                if getattr(post._factor, "structured", False):
                    if post._test_op is not None:
                        raise NotImplementedError("operator push-forwards of a posterior with a Kronecker-structured factor")
                    with backend.phase("var"):
                        var = post._factor.post_var(Xt, diag)
                    return var.cpu().numpy().reshape(batch)
                chunk = backend.var_chunk_rows(n, Xt.shape[0], max_bytes=VAR_CHUNK_BYTES)
                with backend.phase("var"):
                    var = backend.post_var(post._obs_blocks_unique(), post._factor, Xt, diag, chunk=chunk)
                return var.cpu().numpy().reshape(batch)
            nd = self.input_ndim
            b0, b1 = x0.shape[: x0.ndim - nd], x1.shape[: x1.ndim - nd]
            nb = len(batch)
            b0 = (1,) * (nb - len(b0)) + tuple(b0)
            b1 = (1,) * (nb - len(b1)) + tuple(b1)
            if not all(s0 == 1 or s1 == 1 for s0, s1 in zip(b0, b1)):
                raise NotImplementedError("posterior covariance: only outer-product broadcasting is supported")
            K = self._dense(x0.reshape((-1,) + self.input_shape), x1.reshape((-1,) + self.input_shape))
            K = K.cpu().numpy().reshape(tuple(b0) + tuple(b1))
            perm = [ax for pair in zip(range(nb), range(nb, 2 * nb)) for ax in pair]
            return np.ascontiguousarray(np.transpose(K, perm)).reshape(batch)

        def _dense(self, x0: np.ndarray, x1: Optional[np.ndarray]) -> torch.Tensor:
            """k(x0,x1) - K_0 G^{-1} K_1^T = k(x0,x1) - (K_0 L^{-T})(K_1 L^{-T})^T  (_conditional.py:245-251)."""
            post = self._post
            self._require_scalar()
            d = self.input_size
            X0 = backend.points(x0, d)
            X1 = None if x1 is None else backend.points(x1, d)
            C = backend.alloc_matrix(X0.shape[0], X0.shape[0] if X1 is None else X1.shape[0])
            _gram_into(post._prior.cov, X0, X1, C)
            blocks = post._obs_blocks_unique()
            if blocks.empty:
                return C
            if getattr(post._factor, "structured", False):
                if post._test_op is not None:
                    raise NotImplementedError("operator push-forwards of a posterior with a Kronecker-structured factor")
                post._factor.post_cov_sub(X0, X1, C)
                return C
            V0 = backend.crosscov(blocks, post._factor.n, X0)
            post._factor.trsm_rlt(V0)
            if x1 is None:
                V1 = V0
            else:
                V1 = backend.crosscov(blocks, post._factor.n, X1)
                post._factor.trsm_rlt(V1)
            backend.gemm_nt(V0, V1, C, -1.0, 1.0)
            return C

        def linop(self, x0, x1=None):
            x0 = self._preprocess_linop_input(x0, 0)
            x1 = None if x1 is None else self._preprocess_linop_input(x1, 1)
            C = self._dense(x0, x1)
            op = _DeviceMatrix(C)
            if x1 is None:
                op.is_symmetric = True
            return op

    # -- adding observations ----------------------------------------------------------------------------------
    def condition_on_observations(self, Y, X=None, *, L=None, b=None):
        if self._test_op is not None:
            raise NotImplementedError("condition the original process, then apply the operator")
        prior = self._base_prior
        Y, Lf, b, atoms, resid, noise = self._preprocess_observations(prior=prior, Y=Y, X=X, L=L, b=b)
        blk = _Block(None, None, prior.cov.input_size, self._factor.n, atoms=atoms)
        with backend.phase("extend"):
            factor = self._factor.extended(blk.n_phys)
        with backend.phase("assemble"):
            self._assemble_rows(prior, self._blocks, blk, factor, noise)
        with backend.phase("factor"):
            factor.append_last()
        y = torch.zeros(factor.n, dtype=torch.float64, device=factor.L.device)
        y[: self._factor.n].copy_(self._resid)
        y[blk.col_off : blk.col_off + blk.n].copy_(backend.to_device(resid))
        # representer weights of the extended system.  The reference updates them with the Schur-complement
        # formulas of BlockMatrix2x2.schur_update (_block.py:226-231); solving with the extended factor is the same
        # linear system and costs the same O(N^2).
        with backend.phase("solve"):
            w = factor.potrs(y.clone().reshape(1, -1)).reshape(-1)
        return ConditionalGaussianProcess(
            prior=prior, Ys=self._Ys + (Y,), Ls=self._Ls + (Lf,), bs=self._bs + (b,),
            blocks=self._blocks + (blk,), factor=factor, resid=y, weights=w,
        )

    @staticmethod
    def _assemble_rows(prior, prev_blocks, blk: _Block, factor: backend.DeviceFactor, noise) -> None:
        """Fill rows [col_off, col_off + n_phys) of the factor buffer with [L_new k L_j^*(X_new, X_j) ... | D]."""
        k = prior.cov
        r0, n = blk.col_off, blk.n
        rows = factor.L[r0 : r0 + blk.n_phys]
        for pb in prev_blocks:
            out = rows[:n, pb.col_off : pb.col_off + pb.n]
            if blk.simple and pb.simple:
                kj = k if pb.op is None else pb.op(k, argnum=1)
                kij = kj if blk.op is None else blk.op(kj, argnum=0)
                if not _assemble_kronecker(kij, blk.grid, pb.grid, out, lower=False):
                    _gram_into(kij, blk.X, pb.X, out)
            else:  # functional observations: sum over all pairs of atoms
                _block_cov_into(k, blk, pb, out)
            if pb.n_phys != pb.n:
                rows[:, pb.col_off + pb.n] = 0.0
        D = rows[:n, r0 : r0 + n]
        if blk.simple:
            kj = k if blk.op is None else blk.op(k, argnum=1)
            kii = kj if blk.op is None else blk.op(kj, argnum=0)
            if not _assemble_kronecker(kii, blk.grid, None, D, lower=True):
                _gram_into(kii, blk.X, None, D, lower=True)
        else:
            _block_cov_into(k, blk, blk, D)
        if noise is not None:
            kind, val = noise
            if kind == "diag":
                backend.add_diag(D, backend.to_device(val), 1.0)
            elif kind == "kernel":
                val.assemble_into(D, lower=True, accumulate=True)
            else:
                D.add_(backend.to_device(val))
        if blk.n_phys != n:  # identity padding row keeps segment offsets even (16-byte aligned TMA rows)
            rows[n, : r0 + n + 1] = 0.0
            rows[n, r0 + n] = 1.0

    @classmethod
    def _preprocess_observations(cls, *, prior, Y, X, L, b):
        """Argument handling and error behaviour of _conditional.py:296-399."""
        if isinstance(L, linfunctls.LinearFunctional):
            if X is not None:
                raise TypeError("If `L` is a `LinearFunctional`, `X` must be `None`.")
            Lf = L
        elif isinstance(L, LinearFunctionOperator):
            if X is None:
                raise ValueError("`X` must not be omitted if `L` is a `LinearFunctionOperator`.")
            Lf = L.to_linfunctl(X)
        elif L is None:
            if X is None:
                raise ValueError("`X` and `L` can not be omitted at the same time.")
            Lf = linfunctls._EvaluationFunctional(  # pylint: disable=protected-access
                input_domain_shape=prior.input_shape, input_codomain_shape=prior.output_shape, X=X
            )
        else:
            raise TypeError("`L` must be a `LinearFunctional`, a `LinearFunctionOperator` or `None`.")
        atoms = Lf._atoms()  # pylint: disable=protected-access
        if prior.output_shape != () and any(op is None or tuple(op.output_codomain_shape) != () for _, _, op, _ in atoms):
            raise NotImplementedError(
                "observations of a multi-output process must be scalar-valued: compose the operator with `SelectOutput`"
            )

        if b is not None:
            b = randvars.asrandvar(b)
            if not isinstance(b, (randvars.Constant, randvars.Normal)):
                raise TypeError(f"`b` must be a `Normal` or a `Constant` `RandomVariable` ({type(b)=})")
            if tuple(b.shape) != tuple(Lf.output_shape):
                raise ValueError(f"{b.shape=} must be equal to {Lf.output_shape}")

        pred_mean = np.asarray(Lf(prior.mean), dtype=np.double).reshape(-1, order="C")
        Y = np.asarray(Y, dtype=np.double)
        if Y.shape != tuple(Lf.output_shape):
            raise ValueError(f"Expected Y to have shape {Lf.output_shape}, got shape {Y.shape}.")
        Y = Y.reshape(-1, order="C")

        noise = None
        if b is not None:
            pred_mean = pred_mean + np.asarray(b.mean, dtype=np.double).reshape(-1, order="C")
            cov = b.cov
            if isinstance(cov, linops.Scaling):
                noise = ("diag", cov.factors)
            elif isinstance(cov, linops.CovarianceLinearOperator) and cov._x1 is None:  # pylint: disable=protected-access
                # b is the marginal of another GP at the observation points (uncertain right-hand side, ``b=-f(X)``):
                # its covariance block is assembled on the device, straight into the Gram rows (accumulate mode)
                noise = ("kernel", cov)
            elif isinstance(cov, linops.LinearOperator):
                noise = ("dense", cov.todense())
            else:
                noise = ("dense", np.asarray(cov, dtype=np.double))
        return Y, Lf, b, atoms, Y - pred_mean, noise

    # -- push-forwards L(posterior) (_conditional.py:432-467) --------------------------------------------------------
    def _apply_linfuncop(self, L: LinearFunctionOperator) -> "ConditionalGaussianProcess":
        return self._state._apply_linfuncop(L)

    def _apply_linfunctl(self, Lf) -> randvars.Normal:
        op, X = Lf._as_observation()  # pylint: disable=protected-access
        gp = self if op is None else self._apply_linfuncop(op)
        X = np.asarray(X, dtype=np.double)
        return randvars.Normal(mean=np.asarray(gp.mean(X)).reshape(-1), cov=gp.cov.linop(X))


class _DeviceMatrix(linops.LinearOperator):
    def __init__(self, dev: torch.Tensor):
        super().__init__(dev.shape)
        self._dev = dev

    def device_dense(self):
        return self._dev


class GramFactorOperator(linops.LinearOperator):
    """The (factored) Gram matrix of all observation batches -- the reference's nested ``BlockMatrix2x2``
    (src/linpde_gp/linops/_block.py:84-292) flattened into one device-resident bordered Cholesky factor."""

    def __init__(self, factor: backend.DeviceFactor, logical_index: np.ndarray):
        n = len(logical_index)
        super().__init__((n, n))
        self.factor = factor
        self._idx = logical_index
        self.is_symmetric = True
        self.is_positive_definite = True

    def device_dense(self) -> torch.Tensor:
        f = self.factor
        Lt = backend.alloc_matrix(f.n, f.n)
        Lt.copy_(torch.tril(f.L))
        G = backend.alloc_matrix(f.n, f.n)
        backend.gemm_nt(Lt, Lt, G, 1.0, 0.0)
        idx = torch.as_tensor(self._idx, device=G.device)
        return G[idx][:, idx].contiguous()

    def cholesky(self, lower: bool = True):
        fac = _LogicalCholesky(self.factor, self._idx)
        return fac if lower else fac.T

    def solve(self, B):
        B = np.asarray(B, dtype=np.double)
        n, f = self.shape[0], self.factor
        vec = B.ndim == 1
        cols = B[:, None] if vec else B
        if cols.shape[0] != n:
            raise ValueError("`b` must be a vector or a (stack of) matrices.")
        dev = torch.zeros((cols.shape[1], f.n), dtype=torch.float64, device=f.L.device)
        idx = torch.as_tensor(self._idx, device=dev.device)
        dev[:, idx] = backend.to_device(np.ascontiguousarray(cols.T))
        f.potrs(dev)
        res = dev[:, idx].cpu().numpy().T
        return res[:, 0] if vec else res

    def logdet(self) -> float:
        return self.factor.logdet()


class _LogicalCholesky(linops.LinearOperator):
    def __init__(self, factor, idx):
        super().__init__((len(idx), len(idx)))
        self.factor = factor
        self._idx = idx
        self.is_lower_triangular = True

    def device_dense(self):
        idx = torch.as_tensor(self._idx, device=self.factor.L.device)
        return torch.tril(self.factor.L)[idx][:, idx].contiguous()
