"""Thin device layer: torch CUDA tensors as buffers, hand-written sm_100a kernels through ``liblpgp.so``.

torch is used for device memory, streams and host<->device copies only; every numerical operation of the hot
path goes through the C ABI (``include/lpgp.h``).  There is no CPU fallback: all functions require a CUDA
device and raise otherwise.
"""
from __future__ import annotations

import contextlib
import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

F64 = torch.float64


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "linpde_gp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback for the hot path"
        )
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ------------------------------------------------------------------------------------------------------------
# CUDA-event phase timers (the reference has no tracing; this is the build's).  The host code
# of the path brackets its device phases with ``phase(name)``; nothing is recorded (and nothing costs anything)
# unless a ``PhaseTimer`` is active.  Events are recorded on torch's current stream -- the stream every C-ABI call
# of this module is enqueued on (the library's internal lookahead stream is forked from and joined back into it).
class PhaseTimer:
    """``with PhaseTimer() as t: ...; t.totals_ms()`` -> device milliseconds per phase name (one host
    synchronisation, when the totals are read)."""

    def __init__(self):
        self._spans = {}
        self._prev = None

    def __enter__(self):
        global _ACTIVE_TIMER  # pylint: disable=global-statement
        self._prev, _ACTIVE_TIMER = _ACTIVE_TIMER, self
        return self

    def __exit__(self, *exc):
        global _ACTIVE_TIMER  # pylint: disable=global-statement
        _ACTIVE_TIMER = self._prev
        return False

    def add(self, name: str, e0, e1) -> None:
        self._spans.setdefault(name, []).append((e0, e1))

    def totals_ms(self) -> dict:
        torch.cuda.synchronize()
        return {k: float(sum(e0.elapsed_time(e1) for e0, e1 in v)) for k, v in self._spans.items()}

    def reset(self) -> None:
        self._spans = {}


_ACTIVE_TIMER: Optional[PhaseTimer] = None


@contextlib.contextmanager
def phase(name: str):
    t = _ACTIVE_TIMER
    if t is None:
        yield
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    try:
        yield
    finally:
        e1.record()
        t.add(name, e0, e1)


def round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def to_device(x, *, pinned: bool = False) -> torch.Tensor:
    """Host array (anything ``np.asarray`` accepts) -> contiguous float64 CUDA tensor."""
    dev = _require_cuda()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=F64).contiguous()
    a = np.array(x, dtype=np.float64, order="C", copy=True)  # private, writable host staging copy
    t = torch.from_numpy(a)
    if pinned:
        t = t.pin_memory()
    return t.to(dev, non_blocking=pinned)


def points(X, d: int) -> torch.Tensor:
    """Flatten the batch shape C-order (pn ``_preprocess_linop_input``, _covariance_function.py:676-693)."""
    t = to_device(X, pinned=True)
    return t.reshape(-1, d) if d > 0 else t.reshape(-1, 1)


def alloc_matrix(rows: int, cols: int) -> torch.Tensor:
    """(rows x cols) view into a buffer whose leading dimension is a multiple of 16 doubles (128-byte rows)."""
    dev = _require_cuda()
    ld = max(round_up(cols, 16), 16)
    return torch.empty((max(rows, 1), ld), dtype=F64, device=dev)[:rows, :cols]


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and (t.shape[1] <= 1 or t.stride(1) == 1), "row-major matrix expected"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


# ------------------------------------------------------------------------------------------------------------
def gram(desc: _lib.KernelDesc, X0: torch.Tensor, X1: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         *, lower: bool = False, accumulate: bool = False, alpha: float = 1.0) -> torch.Tensor:
    """K[i, j] (+)= alpha * (L0 k L1*)(X0[i], X1[j]);  X1=None -> symmetric block on X0."""
    _require_cuda()
    n0 = X0.shape[0]
    n1 = n0 if X1 is None else X1.shape[0]
    if out is None:
        out = alloc_matrix(n0, n1)
    assert out.shape == (n0, n1)
    rc = lib.lpgp_gram(ctypes.byref(desc), _ptr(X0), n0, _ptr(X1), n1, _ptr(out), _ld(out),
                       _lib.GRAM_LOWER if lower else _lib.GRAM_FULL, int(accumulate), float(alpha), _stream())
    check(rc, "lpgp_gram")
    return out


def gram_pairs(desc: _lib.KernelDesc, X0: torch.Tensor, X1: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """out[i] = alpha * value(X0[i], X1[i]) (explicit pairs)."""
    n = X0.shape[0]
    assert X1.shape == X0.shape
    out = torch.empty(n, dtype=F64, device=_require_cuda())
    check(lib.lpgp_gram_pairs(ctypes.byref(desc), _ptr(X0), _ptr(X1), n, _ptr(out), float(alpha), _stream()), "lpgp_gram_pairs")
    return out


def gram_diag(desc: _lib.KernelDesc, n: int, alpha: float = 1.0) -> torch.Tensor:
    out = torch.empty(n, dtype=F64, device=_require_cuda())
    check(lib.lpgp_gram_diag(ctypes.byref(desc), n, _ptr(out), float(alpha), _stream()), "lpgp_gram_diag")
    return out


def add_diag(A: torch.Tensor, v: Optional[torch.Tensor] = None, scalar: float = 1.0) -> None:
    check(lib.lpgp_add_diag(_ptr(A), A.shape[0], _ld(A), _ptr(v), float(scalar), _stream()), "lpgp_add_diag")


def symmetrize_lower(A: torch.Tensor) -> None:
    check(lib.lpgp_symmetrize_lower(_ptr(A), A.shape[0], _ld(A), _stream()), "lpgp_symmetrize_lower")


def kron_sum(terms, out: Optional[torch.Tensor] = None, *, lower: bool = False, accumulate: bool = False) -> torch.Tensor:
    """out (+)= sum_t alpha_t * kron(A_t, B_t) for ``terms = [(alpha, A, B), ...]`` (device matrices, all A_t of one
    shape and all B_t of one shape).  One multiply-add per term and entry: HBM-write bound."""
    _require_cuda()
    assert len(terms) >= 1
    (n1, m1), (n2, m2) = terms[0][1].shape, terms[0][2].shape
    assert all(A.shape == (n1, m1) and B.shape == (n2, m2) for _, A, B in terms)
    if out is None:
        out = alloc_matrix(n1 * n2, m1 * m2)
    assert out.shape == (n1 * n2, m1 * m2)
    nt = len(terms)
    vp, i64, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double
    A = (vp * nt)(*[t[1].data_ptr() for t in terms])
    B = (vp * nt)(*[t[2].data_ptr() for t in terms])
    lda = (i64 * nt)(*[_ld(t[1]) for t in terms])
    ldb = (i64 * nt)(*[_ld(t[2]) for t in terms])
    al = (dbl * nt)(*[float(t[0]) for t in terms])
    rc = lib.lpgp_kron_sum(nt, A, lda, B, ldb, al, n1, m1, n2, m2, _ptr(out), _ld(out),
                           _lib.GRAM_LOWER if lower else _lib.GRAM_FULL, int(accumulate), _stream())
    check(rc, "lpgp_kron_sum")
    return out


def gemm_nt(A: torch.Tensor, B: torch.Tensor, C: torch.Tensor, alpha: float = 1.0, beta: float = 0.0,
            lower: bool = False) -> torch.Tensor:
    """C = beta*C + alpha * A @ B.T on the DMMA path."""
    m, k = A.shape
    n = B.shape[0]
    assert B.shape[1] == k and C.shape == (m, n)
    rc = lib.lpgp_gemm_nt(m, n, k, float(alpha), _ptr(A), _ld(A), _ptr(B), _ld(B), float(beta), _ptr(C), _ld(C),
                          int(lower), _stream())
    check(rc, "lpgp_gemm_nt")
    return C


def gemm_nn(A: torch.Tensor, B: torch.Tensor, C: torch.Tensor, alpha: float = 1.0, beta: float = 0.0) -> torch.Tensor:
    """C = beta*C + alpha * A @ B on the DMMA path (B row-major k x n)."""
    m, k = A.shape
    n = B.shape[1]
    assert B.shape[0] == k and C.shape == (m, n)
    rc = lib.lpgp_gemm_nn(m, n, k, float(alpha), _ptr(A), _ld(A), _ptr(B), _ld(B), float(beta), _ptr(C), _ld(C), _stream())
    check(rc, "lpgp_gemm_nn")
    return C


def gemm_nt_limited(A: torch.Tensor, B: torch.Tensor, C: torch.Tensor, col_limit: torch.Tensor, alpha: float = 1.0,
                    beta: float = 0.0, col_base: int = 0) -> torch.Tensor:
    """C = beta*C + alpha * A @ B.T, each block of 128 rows restricted to the columns j with
    col_base + j < col_limit[block] (int32, device)."""
    m, k = A.shape
    n = B.shape[0]
    assert B.shape[1] == k and C.shape == (m, n) and col_limit.dtype == torch.int32 and col_limit.numel() >= (m + 127) // 128
    rc = lib.lpgp_gemm_nt_limited(m, n, k, float(alpha), _ptr(A), _ld(A), _ptr(B), _ld(B), float(beta), _ptr(C), _ld(C),
                                  _ptr(col_limit), int(col_base), _stream())
    check(rc, "lpgp_gemm_nt_limited")
    return C


def gemv(A: torch.Tensor, x: torch.Tensor, y: torch.Tensor, alpha: float = 1.0, trans: bool = False) -> torch.Tensor:
    """y += alpha * A @ x  (or A.T @ x) in place; A row-major (view with unit column stride)."""
    m, n = A.shape
    assert x.numel() == (m if trans else n) and y.numel() == (n if trans else m)
    assert x.is_contiguous() and y.is_contiguous()
    rc = lib.lpgp_gemv(int(trans), m, n, float(alpha), _ptr(A), _ld(A), _ptr(x), _ptr(y), _stream())
    check(rc, "lpgp_gemv")
    return y


def row_sumsq(A: torch.Tensor, scale: float = 1.0, offset: float = 0.0) -> torch.Tensor:
    out = torch.empty(A.shape[0], dtype=F64, device=A.device)
    check(lib.lpgp_row_sumsq(_ptr(A), A.shape[0], A.shape[1], _ld(A), float(scale), float(offset), _ptr(out), _stream()),
          "lpgp_row_sumsq")
    return out


def matern_integral(desc: _lib.MaternIntegralDesc, a: float, b: float, x: torch.Tensor, out: torch.Tensor, *,
                    out_stride: int = 1, alpha: float = 1.0, w: Optional[torch.Tensor] = None,
                    accumulate: bool = False) -> torch.Tensor:
    """out[i * out_stride] (+)= alpha * (w or 1) * int_a^b k(x[i], t) dt  (``out``: any view whose first element is
    the first target; ``w``: one-element device tensor)."""
    _require_cuda()
    x = x.reshape(-1)
    assert x.is_contiguous() and x.dtype == F64
    rc = lib.lpgp_matern_integral(ctypes.byref(desc), float(a), float(b), _ptr(x), x.numel(), float(alpha), _ptr(w),
                                  _ptr(out), int(out_stride), int(accumulate), _stream())
    check(rc, "lpgp_matern_integral")
    return out


def matern_hat_integral(desc: _lib.MaternIntegralDesc, grid: torch.Tensor, m: int, half_ends: bool, x: torch.Tensor,
                        out: torch.Tensor, *, alpha: float = 1.0, accumulate: bool = False) -> torch.Tensor:
    """out[i, j] (+)= alpha * int phi_j(t) k(x[i], t) dt for the m hat functions on the nodes ``grid`` (m + 2 entries)."""
    _require_cuda()
    x = x.reshape(-1)
    assert x.is_contiguous() and x.dtype == F64 and grid.is_contiguous() and grid.numel() == m + 2
    assert out.shape == (x.numel(), m)
    rc = lib.lpgp_matern_hat_integral(ctypes.byref(desc), _ptr(grid), int(m), int(half_ends), _ptr(x), x.numel(),
                                      float(alpha), _ptr(out), _ld(out), int(accumulate), _stream())
    check(rc, "lpgp_matern_hat_integral")
    return out


def matern_integral2(desc: _lib.MaternIntegralDesc, dom0, dom1, out: torch.Tensor, *, alpha: float = 1.0,
                     accumulate: bool = False) -> torch.Tensor:
    """out[0] (+)= alpha * int_dom0 int_dom1 k(s, t) dt ds."""
    _require_cuda()
    rc = lib.lpgp_matern_integral2(ctypes.byref(desc), float(dom0[0]), float(dom0[1]), float(dom1[0]), float(dom1[1]),
                                   float(alpha), _ptr(out), int(accumulate), _stream())
    check(rc, "lpgp_matern_integral2")
    return out


# ------------------------------------------------------------------------------------------------------------
# growth policy of appendable factors: a fresh allocation leaves room for `fraction` * n (at least `min_rows`) more
# rows, provided the padded square stays below `max_free_fraction` of the free device memory
FACTOR_RESERVE = {"fraction": 0.125, "min_rows": 1024, "max_free_fraction": 0.35}
_LOWER_COPY_ROWS = 2048


class _FactorStorage:
    """One device allocation shared by a chain of factors that extend each other: ``L`` (cap x cap, row-major,
    padded leading dimension) and ``dinv`` (inverted leaf blocks + status area).  ``claimed`` is the number of rows
    handed out so far: only the factor whose ``n == claimed`` (the tip of the chain) may grow in place.  Rows and
    leaf blocks below a factor's own extent are never written again once factored, so every older factor (and the
    posterior object holding it) stays valid -- the versioned extent of SURVEY.md Appendix B."""

    def __init__(self, rows: int, leaves: int, reserve_rows: Optional[int]):
        dev = _require_cuda()
        if reserve_rows is None:
            reserve_rows = max(int(FACTOR_RESERVE["min_rows"]), int(rows * FACTOR_RESERVE["fraction"]))
            free, _ = torch.cuda.mem_get_info()
            if 8.0 * float(round_up(rows + reserve_rows, 16)) ** 2 > FACTOR_RESERVE["max_free_fraction"] * free:
                reserve_rows = 0
        self.cap_rows = rows + max(0, int(reserve_rows))
        # every later segment may end in one partial leaf
        self.cap_leaves = leaves + (self.cap_rows - rows + _lib.LEAF - 1) // _lib.LEAF + (_lib.MAX_SEG if reserve_rows else 0)
        self.L = alloc_matrix(self.cap_rows, self.cap_rows)
        self.dinv = torch.empty(self.cap_leaves * _lib.LEAF * _lib.LEAF + 8, dtype=F64, device=dev)
        self.claimed = rows


class DeviceFactor:
    """Device-resident, appendable lower Cholesky factor (``lpgp_factor``).

    ``seg_sizes`` are the (even) physical sizes of the observation batches.  ``L`` is the n x n view (row-major,
    padded leading dimension) of the storage, ``dinv`` the inverted 128x128 diagonal blocks.  Appending never
    mutates an existing factor: :meth:`extended` returns a new object (the reference's conditioning API is
    functional) that SHARES the storage when the reserved capacity suffices -- no allocation,
    no copy -- and otherwise moves the lower triangle (only) into a larger allocation."""

    def __init__(self, seg_sizes: Sequence[int], reserve_rows: Optional[int] = None, _storage: Optional[_FactorStorage] = None):
        _require_cuda()
        seg_sizes = [int(s) for s in seg_sizes]
        if any(s <= 0 or s % 2 for s in seg_sizes):
            raise ValueError("segment sizes must be positive and even (pad odd batches)")
        if len(seg_sizes) > _lib.MAX_SEG:
            raise ValueError(f"at most {_lib.MAX_SEG} observation batches per factor")
        self.seg_off = [0]
        for s in seg_sizes:
            self.seg_off.append(self.seg_off[-1] + s)
        self.n = self.seg_off[-1]
        self.nleaves = sum((s + _lib.LEAF - 1) // _lib.LEAF for s in seg_sizes)
        arr = (ctypes.c_int64 * len(self.seg_off))(*self.seg_off)
        if lib.lpgp_factor_dinv_bytes(arr, len(seg_sizes)) == 0:
            raise ValueError("invalid segment layout")
        self._storage = _FactorStorage(self.n, self.nleaves, reserve_rows) if _storage is None else _storage
        assert self.n <= self._storage.cap_rows and self.nleaves <= self._storage.cap_leaves
        self.L = self._storage.L[: self.n, : self.n]
        self.dinv = self._storage.dinv
        self.factored_segments = 0
        self._claimed_from = None  # extent of the factor this one grew from in place (rolled back when it dies unused)

    def __del__(self):
        # an in-place extension that dies while still the tip of its storage (a failed or discarded conditioning step)
        # hands its rows back, so that the factor it grew from can grow in place again
        try:
            st = self._storage
            if self._claimed_from is not None and st.claimed == self.n:
                st.claimed = self._claimed_from
        except Exception:  # pylint: disable=broad-except  (interpreter shutdown)
            pass

    @property
    def ld(self) -> int:
        return _ld(self._storage.L)

    @property
    def capacity(self) -> int:
        return self._storage.cap_rows

    def _struct(self, nseg: Optional[int] = None) -> _lib.Factor:
        nseg = len(self.seg_off) - 1 if nseg is None else nseg
        f = _lib.Factor()
        f.L = self.L.data_ptr()
        f.n = self.seg_off[nseg]
        f.ld = self.ld
        f.dinv = self.dinv.data_ptr()
        f.nseg = nseg
        for i in range(nseg + 1):
            f.seg_off[i] = self.seg_off[i]
        return f

    def segment_rows(self, s: int) -> torch.Tensor:
        return self.L[self.seg_off[s] : self.seg_off[s + 1]]

    def potrf(self) -> None:
        """Factor everything from scratch (all segments as one matrix)."""
        if self._storage.claimed != self.n:
            raise RuntimeError("this factor has been extended in place: its rows are shared and immutable")
        f = self._struct()
        check(lib.lpgp_potrf(ctypes.byref(f), _stream()), "lpgp_potrf")
        self.factored_segments = len(self.seg_off) - 1

    def append_last(self) -> None:
        """Segments 0..nseg-2 are factored; the last segment's rows hold the new Gram rows."""
        f = self._struct()
        check(lib.lpgp_chol_append(ctypes.byref(f), _stream()), "lpgp_chol_append")
        self.factored_segments = len(self.seg_off) - 1

    def extended(self, new_size: int, reserve_rows: Optional[int] = None) -> "DeviceFactor":
        """New factor object with one more (unfactored) segment.  In place when this factor is the tip of its
        storage and the capacity suffices; else the lower triangle and the leaf inverses move to a new allocation."""
        new_size = int(new_size)
        sizes = [self.seg_off[i + 1] - self.seg_off[i] for i in range(len(self.seg_off) - 1)] + [new_size]
        st = self._storage
        n_new = self.n + new_size
        leaves_new = self.nleaves + (new_size + _lib.LEAF - 1) // _lib.LEAF
        if (new_size > 0 and new_size % 2 == 0 and st.claimed == self.n and n_new <= st.cap_rows
                and leaves_new <= st.cap_leaves and len(sizes) <= _lib.MAX_SEG):
            st.claimed = n_new
            new = DeviceFactor(sizes, _storage=st)
            new._claimed_from = self.n  # pylint: disable=protected-access
        else:
            new = DeviceFactor(sizes, reserve_rows=reserve_rows)
            for r0 in range(0, self.n, _LOWER_COPY_ROWS):  # lower triangle only, in row slabs
                r1 = min(self.n, r0 + _LOWER_COPY_ROWS)
                new.L[r0:r1, :r1].copy_(self.L[r0:r1, :r1])
            nold = self.nleaves * _lib.LEAF * _lib.LEAF  # all leaf blocks, without the status area
            new.dinv[:nold].copy_(self.dinv[:nold])
        new.factored_segments = self.factored_segments
        return new

    def trsm_rlt(self, X: torch.Tensor, nlead: Optional[int] = None) -> torch.Tensor:
        """X <- X L^{-T} in place (rows of X are right-hand sides)."""
        nlead = self.n if nlead is None else nlead
        assert X.shape[1] == nlead
        f = self._struct()
        check(lib.lpgp_trsm_rlt(ctypes.byref(f), nlead, _ptr(X), X.shape[0], _ld(X), _stream()), "lpgp_trsm_rlt")
        return X

    def trsm_rln(self, X: torch.Tensor) -> torch.Tensor:
        """X <- X L^{-1} in place (L^{-T} b for every row b of X)."""
        assert X.shape[1] == self.n
        f = self._struct()
        check(lib.lpgp_trsm_rln(ctypes.byref(f), _ptr(X), X.shape[0], _ld(X), _stream()), "lpgp_trsm_rln")
        return X

    def ozaki_eligible(self, kblock: int) -> bool:
        """The emulated solve cuts the factor into column blocks of ``kblock`` columns that must start on leaf boundaries:
        every segment but the last a multiple of 128 rows, and at least two blocks."""
        return all(o % _lib.LEAF == 0 for o in self.seg_off[:-1]) and self.n >= 2 * kblock

    def ozaki_planes(self, nslices: int, kblock: int) -> OzakiPlanes:
        """Digit planes of the factor (strictly-lower K-blocks only), computed once and cached."""
        key = (int(nslices), int(kblock))
        cache = self.__dict__.setdefault("_ozaki_cache", {})
        if key not in cache:
            P = OzakiPlanes(self.n, self.n, nslices, kblock)
            nfull = (self.n // kblock) * kblock  # K never extends into a partial last block
            if nfull:
                P.split(self.L[:, :nfull], lower_blocks=True)
            cache.clear()
            cache[key] = P
        return cache[key]

    def trsm_rlt_ozaki(self, X: torch.Tensor, nslices: int, kblock: int = 1024, XP: Optional[OzakiPlanes] = None) -> torch.Tensor:
        """X <- X L^{-T} in place with the O(m n^2) part emulated on the INT8 tensor cores (``lpgp_trsm_rlt_ozaki``)."""
        assert X.shape[1] == self.n
        LP = self.ozaki_planes(nslices, kblock)
        if XP is None or XP.rows < X.shape[0] or XP.nslices != nslices or XP.kblock != kblock or XP.cols < self.n:
            XP = OzakiPlanes(X.shape[0], self.n, nslices, kblock)
        f = self._struct()
        rc = lib.lpgp_trsm_rlt_ozaki(ctypes.byref(f), _ptr(X), X.shape[0], _ld(X), ctypes.byref(LP.struct),
                                     ctypes.byref(XP.struct), _stream())
        check(rc, "lpgp_trsm_rlt_ozaki")
        return X

    def trsv(self, b: torch.Tensor, trans: bool = False) -> torch.Tensor:
        """b <- L^{-1} b (``trans``: L^{-T} b) in place; one contiguous right-hand side of length n."""
        assert b.numel() == self.n and b.is_contiguous()
        f = self._struct()
        check(lib.lpgp_trsv(ctypes.byref(f), int(trans), _ptr(b), _stream()), "lpgp_trsv")
        return b

    def potrs(self, B: torch.Tensor) -> torch.Tensor:
        """Rows of B <- G^{-1} rows of B, in place."""
        if B.dim() == 1:
            B = B.reshape(1, -1)
        assert B.shape[1] == self.n
        f = self._struct()
        check(lib.lpgp_potrs(ctypes.byref(f), _ptr(B), B.shape[0], _ld(B), _stream()), "lpgp_potrs")
        return B

    def logdet(self) -> float:
        out = torch.empty(1, dtype=F64, device=self.L.device)
        f = self._struct()
        check(lib.lpgp_logdet(ctypes.byref(f), _ptr(out), _stream()), "lpgp_logdet")
        return float(out.item())


# ------------------------------------------------------------------------------------------------------------
# FP64 emulated on the INT8 tensor cores (Ozaki scheme, csrc/ozaki.cu): opt-in solver of the posterior-variance TRSM
class OzakiPlanes:
    """Digit planes of an FP64 matrix (``lpgp_ozaki_planes``): ``nslices`` byte planes of ``rows x cols`` plus one int32
    exponent per (row, K-block of ``kblock`` columns).  Owns its buffers."""

    def __init__(self, rows: int, cols: int, nslices: int, kblock: int = 1024):
        dev = _require_cuda()
        if not 1 <= nslices <= _lib.OZAKI_MAX_SLICES:
            raise ValueError(f"nslices must be in 1..{_lib.OZAKI_MAX_SLICES}")
        if kblock % 128 or not 128 <= kblock <= 4096:
            raise ValueError("kblock must be a multiple of 128 in 128..4096")
        self.rows, self.cols, self.nslices, self.kblock = int(rows), int(cols), int(nslices), int(kblock)
        self.pitch = round_up(max(self.cols, 1), 128)
        self.buf = torch.empty((self.nslices, max(self.rows, 1), self.pitch), dtype=torch.uint8, device=dev)
        self.nkb = (self.cols + kblock - 1) // kblock
        self.exps = torch.zeros((max(self.nkb, 1), max(self.rows, 1)), dtype=torch.int32, device=dev)
        st = _lib.OzakiPlanes()
        st.planes, st.exps = self.buf.data_ptr(), self.exps.data_ptr()
        st.rows, st.cols, st.pitch = self.rows, self.cols, self.pitch
        st.plane_stride = max(self.rows, 1) * self.pitch
        st.lde = max(self.rows, 1)
        st.nslices, st.kblock = self.nslices, self.kblock
        self.struct = st

    def split(self, A: torch.Tensor, row_off: int = 0, col0: int = 0, lower_blocks: bool = False) -> None:
        """Digits of ``A`` (rows x ncols, ncols a multiple of kblock) -> rows ``row_off..``, columns ``col0..``."""
        rows, ncols = A.shape
        check(lib.lpgp_ozaki_split(_ptr(A), _ld(A), rows, int(row_off), int(col0), ncols, ctypes.byref(self.struct),
                                   int(lower_blocks), _stream()), "lpgp_ozaki_split")

    def reconstruct(self, rows: slice, cols: slice) -> torch.Tensor:
        """FP64 value of the stored digits (tests / diagnostics; torch arithmetic, not a product path)."""
        kb0, kb1 = cols.start // self.kblock, (cols.stop + self.kblock - 1) // self.kblock
        out = torch.zeros((rows.stop - rows.start, cols.stop - cols.start), dtype=F64, device=self.buf.device)
        for s in range(self.nslices):
            d = self.buf[s, rows, cols]
            d = d.to(torch.int8).to(F64) if s == 0 else d.to(F64)
            out += d * 2.0 ** (-7 - 8 * s)
        e = self.exps[kb0:kb1, rows].to(F64).T.repeat_interleave(self.kblock, dim=1)
        e = e[:, cols.start - kb0 * self.kblock : cols.stop - kb0 * self.kblock]
        return out * torch.exp2(e)


def ozaki_gemm_nt(PA: OzakiPlanes, PB: OzakiPlanes, C: torch.Tensor, k: int, alpha: float = 1.0, beta: float = 0.0,
                  rowA0: int = 0, kA0: int = 0, rowB0: int = 0, kB0: int = 0) -> torch.Tensor:
    """C = beta C + alpha A B^T from digit planes (A = rows rowA0.. / columns kA0..kA0+k of PA, B likewise)."""
    m, n = C.shape
    rc = lib.lpgp_ozaki_gemm_nt(m, n, int(k), float(alpha), ctypes.byref(PA.struct), int(rowA0), int(kA0),
                                ctypes.byref(PB.struct), int(rowB0), int(kB0), float(beta), _ptr(C), _ld(C), _stream())
    check(rc, "lpgp_ozaki_gemm_nt")
    return C


# posterior-variance solver: S = 7 (DEFAULT): the O(M N^2) part of the solve is emulated on the INT8 tensor cores with 7
# digit planes = 55 bits below every row maximum, i.e. the FP64 operands exactly (measured 5e-14 from the DMMA result,
# 2x its speed); S = 6 keeps 47 bits (1e-11, 2.5x); {"ozaki_slices": 0} = native DMMA path.  Factors that are not
# eligible (fewer than 2 K-blocks, segments not multiples of 128 rows) always take the DMMA path.
VARIANCE_SOLVER = {"ozaki_slices": int(__import__("os").environ.get("LPGP_OZAKI_SLICES", "7")), "kblock": 1024}


if "LPGP_OZAKI_CLUSTER" in __import__("os").environ:  # A/B switch of the emulated GEMM's cluster size (1 or 2, default 2)
    check(lib.lpgp_set_option(5, int(__import__("os").environ["LPGP_OZAKI_CLUSTER"])), "lpgp_set_option(LPGP_OPT_OZAKI_CLUSTER)")
if "LPGP_OZAKI_CTA_PAIR" in __import__("os").environ:  # A/B switch: CTA-pair kernel (tcgen05 cta_group::2) on (1, default) or off (0)
    check(lib.lpgp_set_option(7, int(__import__("os").environ["LPGP_OZAKI_CTA_PAIR"])), "lpgp_set_option(LPGP_OPT_OZAKI_CTA_PAIR)")
if "LPGP_OZAKI_PAIR_LEVELS" in __import__("os").environ:  # A/B switch: two digit levels per pass (1, default) or one (0)
    check(lib.lpgp_set_option(6, int(__import__("os").environ["LPGP_OZAKI_PAIR_LEVELS"])), "lpgp_set_option(LPGP_OPT_OZAKI_PAIR_LEVELS)")


def set_variance_solver(ozaki_slices: int = 0, kblock: int = 1024) -> None:
    if not 0 <= int(ozaki_slices) <= _lib.OZAKI_MAX_SLICES:
        raise ValueError(f"ozaki_slices must be in 0..{_lib.OZAKI_MAX_SLICES}")
    VARIANCE_SOLVER["ozaki_slices"], VARIANCE_SOLVER["kblock"] = int(ozaki_slices), int(kblock)


# ------------------------------------------------------------------------------------------------------------
class ObsBlocks:
    """ctypes array of ``lpgp_obs_block`` keeping the referenced descriptors / tensors alive."""

    def __init__(self, descs, Xs, col_offs, extras=()):
        self.descs = list(descs)
        self.Xs = list(Xs)
        self.n = len(self.descs)
        # columns that are not pairwise kernel evaluations: integral observations
        # (column, (a, b), [(alpha, MaternIntegralDesc), ...]) -- one closed-form column each -- and dense column blocks
        # (column, width, fill) with ``fill(Xt) -> (M x width) device matrix`` (L2-projection observations)
        self.extras = list(extras)
        self.arr = (_lib.ObsBlock * self.n)()
        for i, (dsc, X, off) in enumerate(zip(self.descs, self.Xs, col_offs)):
            self.arr[i].desc = ctypes.pointer(dsc)
            self.arr[i].X = X.data_ptr()
            self.arr[i].n = X.shape[0]
            self.arr[i].col_off = int(off)

    @property
    def empty(self) -> bool:
        return self.n == 0 and not self.extras

    def add_integral_columns(self, Xt: torch.Tensor, K: torch.Tensor) -> None:
        """K[:, col] += sum alpha * int_a^b k(Xt, t) dt for every integral observation (after ``lpgp_crosscov``)."""
        for col, what, terms in self.extras:
            if callable(terms):  # dense block of `what` columns
                K[:, col : col + what].add_(terms(Xt))
                continue
            for alpha, dsc in terms:
                matern_integral(dsc, what[0], what[1], Xt, K[:, col:], out_stride=_ld(K), alpha=alpha, accumulate=True)


def post_mean(blocks: ObsBlocks, w: torch.Tensor, Xt: torch.Tensor, out: Optional[torch.Tensor] = None,
              accumulate: bool = False) -> torch.Tensor:
    m = Xt.shape[0]
    if out is None:
        out = torch.empty(m, dtype=F64, device=Xt.device)
    if blocks.n == 0:
        if not accumulate:
            out.zero_()
    else:
        rc = lib.lpgp_post_mean(blocks.arr, blocks.n, _ptr(w), _ptr(Xt), m, _ptr(out), int(accumulate), _stream())
        check(rc, "lpgp_post_mean")
    for col, what, terms in blocks.extras:
        if callable(terms):  # dense block of `what` columns: out += block @ w[col : col + what]
            gemv(terms(Xt), w[col : col + what].contiguous(), out, 1.0)
            continue
        for alpha, dsc in terms:  # integral observations: one closed-form column each, weight folded in
            matern_integral(dsc, what[0], what[1], Xt, out, alpha=alpha, w=w[col : col + 1], accumulate=True)
    return out


def crosscov(blocks: ObsBlocks, n: int, Xt: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    m = Xt.shape[0]
    if out is None:
        out = alloc_matrix(m, n)
    check(lib.lpgp_crosscov(blocks.arr, blocks.n, n, _ptr(Xt), m, _ptr(out), _ld(out), _stream()), "lpgp_crosscov")
    blocks.add_integral_columns(Xt, out)
    return out


def var_chunk_rows(n: int, m: int, max_bytes: int = 16 << 30) -> int:
    """Test points per chunk of the posterior-variance solve: the cross-covariance workspace (chunk x n doubles) takes
    up to a quarter of the free device memory, at most ``max_bytes``.  Larger chunks make the 128-wide leaf steps and
    the small GEMMs near the leaves of the blocked TRSM proportionally cheaper (m = 8192 -> 32768 rows at N = 64k)."""
    free, _ = torch.cuda.mem_get_info()
    budget = max(256 << 20, min(int(max_bytes), free // 4))
    rows = budget // (8 * round_up(max(int(n), 1), 16))
    rows = (rows // 128) * 128 if rows >= 128 else rows
    return int(max(1, min(int(m), max(rows, 256))))


def post_var(blocks: ObsBlocks, factor: DeviceFactor, Xt: torch.Tensor, prior_diag: float, chunk: int = 8192,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pointwise posterior variance, test points processed in chunks of ``chunk`` rows."""
    m = Xt.shape[0]
    if out is None:
        out = torch.empty(m, dtype=F64, device=Xt.device)
    chunk = max(1, min(chunk, m))
    K = alloc_matrix(chunk, factor.n)
    f = factor._struct()
    S, kblock = VARIANCE_SOLVER["ozaki_slices"], VARIANCE_SOLVER["kblock"]
    emulate = S > 0 and factor.ozaki_eligible(kblock)
    XP = OzakiPlanes(chunk, factor.n, S, kblock) if emulate else None
    for i0 in range(0, m, chunk):
        mc = min(chunk, m - i0)
        if blocks.extras or emulate:  # same three steps as lpgp_post_var, spelled out
            Kc = K[:mc]
            crosscov(blocks, factor.n, Xt[i0 : i0 + mc], out=Kc)
            if emulate:
                factor.trsm_rlt_ozaki(Kc, S, kblock, XP)
            else:
                factor.trsm_rlt(Kc)
            check(lib.lpgp_row_sumsq(_ptr(Kc), mc, factor.n, _ld(K), -1.0, float(prior_diag), _ptr(out[i0:]), _stream()),
                  "lpgp_row_sumsq")
            continue
        rc = lib.lpgp_post_var(blocks.arr, blocks.n, ctypes.byref(f), _ptr(Xt[i0:]), mc, float(prior_diag), _ptr(K),
                               _ld(K), _ptr(out[i0:]), _stream())
        check(rc, "lpgp_post_var")
    return out
