"""CPU test double of the C ABI (``include/lpgp.h``) -- TEST INFRASTRUCTURE ONLY.

``install_c_abi(monkeypatch)`` replaces the ctypes handle ``backend.lib`` by :class:`_HostLib`, whose methods are numpy
statements of the entry points the symbolic and conditioning layers call (assembly from a kernel descriptor via
``tests/helpers.eval_desc_numpy``, Cholesky / bordered append / triangular solves via ``numpy.linalg``, posterior mean /
cross-covariance / variance, Kronecker sums, the closed-form Matern integrals via ``oracle/integrals.py``), operating on
the SAME pointers and structs -- the buffers are CPU torch tensors.  With it the real HOST code of the product runs on a
CPU-only machine: ``DeviceFactor`` (storage, in-place extension, copy-on-append), ``ObsBlocks``, the block assembly order,
noise handling, chunking of the posterior evaluation, the seam classes (cross-covariances, ``BlockMatrix2x2``, ...), and is
checked by ``tests/test_host_path.py`` against the real-reference goldens with the test bodies of the GPU suite.

It is installed by monkeypatching inside a test and never imported by the product, which has no CPU path
(``backend._require_cuda``); numbers obtained through it say nothing about the CUDA kernels -- those are checked by
``pytest -m gpu`` on the B200."""
import ctypes

import numpy as np
import torch

from tests import helpers


def _addr(p):
    if isinstance(p, ctypes.c_void_p):
        return p.value or 0
    if hasattr(p, "value"):
        return p.value or 0
    return int(p)


def _vec(p, n):
    a = _addr(p)
    if a == 0 or n == 0:
        return None if a == 0 else np.zeros(0)
    return np.ctypeslib.as_array((ctypes.c_double * int(n)).from_address(a))


def _mat(p, rows, cols, ld):
    rows, cols, ld = int(rows), int(cols), int(ld)
    if rows == 0 or cols == 0:
        return np.zeros((rows, cols))
    flat = _vec(p, (rows - 1) * ld + cols)
    return np.lib.stride_tricks.as_strided(flat, shape=(rows, cols), strides=(8 * ld, 8))


def _obj(ref):
    return ref._obj if hasattr(ref, "_obj") else ref  # ctypes.byref(struct) -> struct


def _desc_eval(desc_ref, X0, X1):
    desc = _obj(desc_ref)
    if hasattr(desc, "contents"):
        desc = desc.contents
    return helpers.eval_desc_numpy(desc, X0, X1)


class _HostLib:
    """Stands in for the ctypes handle ``backend.lib``: listed entry points are numpy restatements (include/lpgp.h), pure
    host entry points (sizes, error strings, options) go to the real library, anything else raises."""

    _PASS = {"lpgp_factor_dinv_bytes", "lpgp_error_string", "lpgp_version", "lpgp_build_arch", "lpgp_set_option",
             "lpgp_launch_count"}

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        if name in self._PASS:
            return getattr(self._real, name)
        raise NotImplementedError(f"tests/host_double.py has no double of {name}")

    # ---- (1) assembly
    def lpgp_gram(self, desc, X0, n0, X1, n1, out, ld, mode, accumulate, alpha, stream):
        d = _obj(desc).d
        A0 = _mat(X0, n0, d, d)
        A1 = A0 if _addr(X1) == 0 else _mat(X1, n1, d, d)
        K = float(alpha) * _desc_eval(desc, A0, A1)
        O = _mat(out, n0, n1, ld)
        if int(mode) == 1:  # LOWER: only the lower triangle of the symmetric block is written
            idx = np.tril_indices(int(n0))
            O[idx] = (O[idx] if accumulate else 0.0) + K[idx]
        elif accumulate:
            O += K
        else:
            O[...] = K
        return 0

    def lpgp_gram_diag(self, desc, n0, out, alpha, stream):
        d = _obj(desc).d
        z = np.zeros((1, d))
        _vec(out, n0)[...] = float(alpha) * _desc_eval(desc, z, z)[0, 0]
        return 0

    def lpgp_gram_pairs(self, desc, X0, X1, n, out, alpha, stream):
        d = _obj(desc).d
        A0, A1, o = _mat(X0, n, d, d), _mat(X1, n, d, d), _vec(out, n)
        for i in range(int(n)):
            o[i] = float(alpha) * _desc_eval(desc, A0[i : i + 1], A1[i : i + 1])[0, 0]
        return 0

    def lpgp_add_diag(self, A, n, ld, v, scalar, stream):
        M = _mat(A, n, n, ld)
        vv = _vec(v, n)
        M[np.diag_indices(int(n))] += float(scalar) * (1.0 if vv is None else vv)
        return 0

    def lpgp_symmetrize_lower(self, A, n, ld, stream):
        M = _mat(A, n, n, ld)
        iu = np.triu_indices(int(n), 1)
        M[iu] = M.T[iu]
        return 0

    def lpgp_kron_sum(self, nterms, A, lda, B, ldb, alpha, n1, m1, n2, m2, out, ld, mode, accumulate, stream):
        K = np.zeros((int(n1) * int(n2), int(m1) * int(m2)))
        for t in range(int(nterms)):
            K += float(alpha[t]) * np.kron(_mat(A[t], n1, m1, lda[t]), _mat(B[t], n2, m2, ldb[t]))
        O = _mat(out, K.shape[0], K.shape[1], ld)
        if int(mode) == 1:
            idx = np.tril_indices(K.shape[0])
            O[idx] = (O[idx] if accumulate else 0.0) + K[idx]
        elif accumulate:
            O += K
        else:
            O[...] = K
        return 0

    def lpgp_gemm_nt(self, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower, stream):
        P = float(alpha) * (_mat(A, m, k, lda) @ _mat(B, n, k, ldb).T)
        Cm = _mat(C, m, n, ldc)
        Cm[...] = P if float(beta) == 0.0 else float(beta) * Cm + P
        return 0

    def lpgp_gemm_nn(self, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stream):
        P = float(alpha) * (_mat(A, m, k, lda) @ _mat(B, k, n, ldb))
        Cm = _mat(C, m, n, ldc)
        Cm[...] = P if float(beta) == 0.0 else float(beta) * Cm + P
        return 0

    def lpgp_gemv(self, trans, m, n, alpha, A, lda, x, y, stream):
        M = _mat(A, m, n, lda)
        if int(trans):
            _vec(y, n)[...] += float(alpha) * (M.T @ _vec(x, m))
        else:
            _vec(y, m)[...] += float(alpha) * (M @ _vec(x, n))
        return 0

    # ---- (4) closed-form Matern integrals: the oracle's restatement (oracle/integrals.py) of the same formulas
    @staticmethod
    def _matern(desc):
        d = _obj(desc)
        p = int(d.ncoef) - 1
        return p, float(np.sqrt(2.0 * (p + 0.5)) / d.scale)

    def lpgp_matern_integral(self, desc, a, b, x, n, alpha, w, out, out_stride, accumulate, stream):
        from oracle import integrals as oint  # pylint: disable=import-outside-toplevel

        p, ell = self._matern(desc)
        ww = _vec(w, 1)
        v = float(alpha) * (1.0 if ww is None else float(ww[0])) * oint.matern_lebesgue_integral(p, ell, float(a), float(b), _vec(x, n))
        o = np.lib.stride_tricks.as_strided(_vec(out, (int(n) - 1) * int(out_stride) + 1), shape=(int(n),),
                                            strides=(8 * int(out_stride),))
        o[...] = (o if accumulate else 0.0) + v
        return 0

    def lpgp_matern_integral2(self, desc, a, b, c, d, alpha, out, accumulate, stream):
        from oracle import integrals as oint  # pylint: disable=import-outside-toplevel

        p, ell = self._matern(desc)
        o = _vec(out, 1)
        o[0] = (o[0] if accumulate else 0.0) + float(alpha) * oint.matern_lebesgue_integral_lebesgue_integral(
            p, ell, (float(a), float(b)), (float(c), float(d)))
        return 0

    def lpgp_matern_hat_integral(self, desc, grid, m, half_ends, x, n, alpha, out, ld, accumulate, stream):
        """int phi_j(t) k(x_i, t) dt by Gauss-Legendre on the pieces where the integrand is smooth (the elements, split at
        the kink t = x_i of the half-integer Matern kernel) -- 48 nodes per piece: machine precision for these analytic pieces."""
        import math  # pylint: disable=import-outside-toplevel

        p, ell = self._matern(desc)
        s = math.sqrt(2.0 * (p + 0.5)) / ell
        cf = [math.factorial(p) / math.factorial(2 * p) * math.factorial(p + i) / (math.factorial(i) * math.factorial(p - i))
              for i in range(p + 1)]

        def k(r):
            u = s * np.abs(r)
            return np.exp(-u) * sum(c * (2.0 * u) ** (p - i) for i, c in enumerate(cf))

        z, wq = np.polynomial.legendre.leggauss(48)
        g, xs = _vec(grid, int(m) + 2), _vec(x, n)
        O = _mat(out, n, m, ld)
        for i, xi in enumerate(xs):
            for j in range(int(m)):
                val = 0.0
                for a, b, rising in ((g[j], g[j + 1], True), (g[j + 1], g[j + 2], False)):
                    if int(half_ends) and ((j == 0 and rising) or (j == int(m) - 1 and not rising)):
                        continue
                    cuts = [a] + ([xi] if a < xi < b else []) + [b]
                    for lo, hi in zip(cuts[:-1], cuts[1:]):
                        t = 0.5 * (hi - lo) * z + 0.5 * (hi + lo)
                        phi = (t - a) / (b - a) if rising else (b - t) / (b - a)
                        val += 0.5 * (hi - lo) * float(np.sum(wq * phi * k(t - xi)))
                O[i, j] = (O[i, j] if accumulate else 0.0) + float(alpha) * val
        return 0

    # ---- (2) factor
    @staticmethod
    def _L(f):
        f = _obj(f)
        return f, _mat(f.L, f.n, f.n, f.ld)

    def lpgp_potrf(self, f, stream):
        f, L = self._L(f)
        G = np.tril(L) + np.tril(L, -1).T
        try:
            C = np.linalg.cholesky(G)
        except np.linalg.LinAlgError:
            for j in range(1, f.n + 1):  # LAPACK info: order of the first leading minor that is not positive definite
                try:
                    np.linalg.cholesky(G[:j, :j])
                except np.linalg.LinAlgError:
                    return j
            return int(f.n)
        L[np.tril_indices(f.n)] = C[np.tril_indices(f.n)]
        return 0

    def lpgp_chol_append(self, f, stream):
        f, L = self._L(f)
        n0 = int(f.seg_off[f.nseg - 1])
        if n0 == 0:
            return self.lpgp_potrf(f, stream)
        L11 = np.tril(L[:n0, :n0])
        L21 = np.linalg.solve(L11, L[n0:, :n0].T).T  # B^T L11^{-T}
        D = np.tril(L[n0:, n0:]) + np.tril(L[n0:, n0:], -1).T
        S = D - L21 @ L21.T
        try:
            C = np.linalg.cholesky(S)
        except np.linalg.LinAlgError:
            for j in range(1, S.shape[0] + 1):
                try:
                    np.linalg.cholesky(S[:j, :j])
                except np.linalg.LinAlgError:
                    return n0 + j
            return int(f.n)
        L[n0:, :n0] = L21
        idx = np.tril_indices(f.n - n0)
        L[n0:, n0:][idx] = C[idx]
        return 0

    def lpgp_trsm_rlt(self, f, nlead, X, m, ldx, stream):
        f, L = self._L(f)
        Xm = _mat(X, m, nlead, ldx)
        Xm[...] = np.linalg.solve(np.tril(L[: int(nlead), : int(nlead)]), Xm.T).T
        return 0

    def lpgp_trsm_rln(self, f, X, m, ldx, stream):
        f, L = self._L(f)
        Xm = _mat(X, m, f.n, ldx)
        Xm[...] = np.linalg.solve(np.tril(L).T, Xm.T).T
        return 0

    def lpgp_potrs(self, f, B, nrhs, ldb, stream):
        f, L = self._L(f)
        Bm = _mat(B, nrhs, f.n, ldb)
        Lt = np.tril(L)
        Bm[...] = np.linalg.solve(Lt.T, np.linalg.solve(Lt, Bm.T)).T
        return 0

    def lpgp_trsv(self, f, trans, b, stream):
        f, L = self._L(f)
        v = _vec(b, f.n)
        Lt = np.tril(L)
        v[...] = np.linalg.solve(Lt.T if int(trans) else Lt, v)
        return 0

    def lpgp_logdet(self, f, out, stream):
        f, L = self._L(f)
        _vec(out, 1)[0] = float(np.sum(np.log(np.diag(L) ** 2)))
        return 0

    # ---- (3) posterior
    @staticmethod
    def _blocks(blocks, nblocks):
        for i in range(int(nblocks)):
            b = blocks[i]
            d = b.desc.contents.d
            yield b.desc.contents, _mat(b.X, b.n, d, d), int(b.n), int(b.col_off)

    def lpgp_post_mean(self, blocks, nblocks, w, Xt, m, out, accumulate, stream):
        o = _vec(out, m)
        acc = o.copy() if accumulate else np.zeros(int(m))
        for desc, X, n, off in self._blocks(blocks, nblocks):
            T = _mat(Xt, m, desc.d, desc.d)
            acc += helpers.eval_desc_numpy(desc, T, X) @ _vec(_addr(w) + 8 * off, n)
        o[...] = acc
        return 0

    def lpgp_crosscov(self, blocks, nblocks, n, Xt, m, K, ldk, stream):
        Km = _mat(K, m, n, ldk)
        Km[...] = 0.0
        for desc, X, nb, off in self._blocks(blocks, nblocks):
            T = _mat(Xt, m, desc.d, desc.d)
            Km[:, off : off + nb] += helpers.eval_desc_numpy(desc, T, X)
        return 0

    def lpgp_row_sumsq(self, A, m, n, ld, scale, offset, out, stream):
        _vec(out, m)[...] = float(offset) + float(scale) * np.sum(_mat(A, m, n, ld) ** 2, axis=1)
        return 0

    def lpgp_post_var(self, blocks, nblocks, f, Xt, m, prior_diag, K, ldk, out, stream):
        fo = _obj(f)
        self.lpgp_crosscov(blocks, nblocks, fo.n, Xt, m, K, ldk, stream)
        self.lpgp_trsm_rlt(f, fo.n, K, m, ldk, stream)
        return self.lpgp_row_sumsq(K, m, fo.n, ldk, -1.0, prior_diag, out, stream)


def install_c_abi(monkeypatch):
    """Patch ``backend`` so that every C-ABI call of the conditioning path lands in :class:`_HostLib` and all buffers
    are CPU tensors (undone by pytest's ``monkeypatch``)."""
    from linpde_gp_b200 import backend

    cpu = torch.device("cpu")

    def to_device(x, *, pinned=False):  # pylint: disable=unused-argument
        if isinstance(x, torch.Tensor):
            return x.to(dtype=torch.float64).contiguous()
        return torch.from_numpy(np.array(x, dtype=np.float64, order="C", copy=True))

    monkeypatch.setattr(backend, "_require_cuda", lambda: cpu)
    monkeypatch.setattr(backend, "_stream", lambda: ctypes.c_void_p(0))
    monkeypatch.setattr(backend, "to_device", to_device)
    monkeypatch.setattr(backend, "points", lambda X, d: to_device(X).reshape(-1, d))
    monkeypatch.setattr(backend, "lib", _HostLib(backend.lib))
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a, **k: (32 << 30, 64 << 30))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
