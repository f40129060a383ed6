"""Multi-GPU Cholesky of the Gram matrix: one process per GPU, NCCL over NVLink.

Layout: the lower triangle is cut into block rows of ``nb`` rows; block row ``i`` lives on rank ``i % P``
(1-D block-cyclic over rows), stored contiguously in that rank's ``A_loc`` (rows of local block ``l`` = global
block ``l * P + rank``).  Right-looking factorisation, per panel ``k`` (block column ``k``):

  1. the owner of block row ``k`` factors the nb x nb diagonal block in place (``lpgp_potrf`` on a view) and
     broadcasts ``L_kk`` together with its inverted 128 x 128 leaf blocks                      (NCCL broadcast, ~2.6 MB)
  2. every rank solves ITS rows of the panel, ``X <- X L_kk^{-T}`` (``lpgp_trsm_rlt``)        (no communication)
  3. the panel pieces are all-gathered and put into global row order                           (NCCL all-gather)
  4. every rank updates its block rows of the trailing matrix, ``A_loc -= X_loc P^T``, with ONE row-limited DMMA
     GEMM (tiles right of a row block's own diagonal block are skipped)                        (no communication)

Assembly needs no communication at all: every rank evaluates the Gram kernel for its own block rows.  After the
factorisation the factor is replicated (one broadcast per block row, straight into its final place) so that
posterior evaluation can shard the test points without further exchange.

The numerical building blocks are injected (``ops``): the product uses :class:`DeviceOps` (CUDA kernels through
``liblpgp.so``); the CPU test-suite substitutes a numpy double to exercise the ownership / packing / collective
logic under ``gloo`` with world_size 2.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

LEAF = 128


class BlockRowLayout:
    """Ownership bookkeeping of the 1-D block-cyclic row distribution."""

    def __init__(self, n: int, nb: int, world: int):
        if nb % LEAF:
            raise ValueError("nb must be a multiple of 128")
        if n % 2:
            raise ValueError("n must be even")
        self.n, self.nb, self.world = int(n), int(nb), int(world)
        self.nblk = (n + nb - 1) // nb

    def owner(self, i: int) -> int:
        return i % self.world

    def block_bounds(self, i: int) -> Tuple[int, int]:
        return i * self.nb, min(self.n, (i + 1) * self.nb)

    def block_size(self, i: int) -> int:
        lo, hi = self.block_bounds(i)
        return hi - lo

    def local_blocks(self, rank: int) -> List[int]:
        return list(range(rank, self.nblk, self.world))

    def n_local(self, rank: int) -> int:
        return sum(self.block_size(i) for i in self.local_blocks(rank))

    def local_row_offset(self, i: int) -> int:
        """first local row of global block ``i`` on its owner (only the last global block may be short)"""
        return (i // self.world) * self.nb

    def first_local_block_after(self, rank: int, k: int) -> int:
        """number of local blocks of ``rank`` with global index <= k"""
        return 0 if rank > k else (k - rank) // self.world + 1

    def rows_after(self, rank: int, k: int) -> int:
        return sum(self.block_size(i) for i in self.local_blocks(rank) if i > k)


class DeviceOps:
    """Local numerical kernels on the GPU (C ABI through ctypes) + the two CUDA streams of the panel pipeline."""

    def __init__(self):
        from . import _lib, backend

        self._lib, self.be = _lib, backend
        self.device = backend._require_cuda()  # pylint: disable=protected-access
        # panel chain (critical path) on a high-priority stream, bulk trailing updates on a normal one
        self.panel_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.update_stream = torch.cuda.Stream(device=self.device, priority=0)

    # -- streams / events (no-ops in the host test double) -----------------------------------------------------
    def on(self, which: str):
        return torch.cuda.stream(self.panel_stream if which == "panel" else self.update_stream)

    def fork(self) -> None:
        cur = torch.cuda.current_stream()
        self.panel_stream.wait_stream(cur)
        self.update_stream.wait_stream(cur)

    def join(self) -> None:
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.panel_stream)
        cur.wait_stream(self.update_stream)

    def record(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        return ev

    def wait(self, ev) -> None:
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    # -- buffers ---------------------------------------------------------------------------------------------------
    def empty(self, rows: int, cols: int) -> torch.Tensor:
        return self.be.alloc_matrix(rows, cols)

    def _factor_struct(self, L: torch.Tensor, dinv: torch.Tensor):
        f = self._lib.Factor()
        f.L = L.data_ptr()
        f.n = L.shape[0]
        f.ld = self.be._ld(L)  # pylint: disable=protected-access
        f.dinv = dinv.data_ptr()
        f.nseg = 1
        f.seg_off[0], f.seg_off[1] = 0, L.shape[0]
        return f

    def potrf_block(self, D: torch.Tensor, dinv: torch.Tensor, info_out: torch.Tensor) -> None:
        """Cholesky of the square view ``D`` in place, WITHOUT synchronising the host; ``dinv`` receives the
        inverted leaf blocks (+ 8 doubles of status area); ``info_out`` (1 double, device) receives the LAPACK
        info (0 = ok, > 0 = order of the first non-positive-definite leading minor of ``D``)."""
        f = self._factor_struct(D, dinv)
        rc = self._lib.lib.lpgp_potrf_async(ctypes.byref(f), self.be._stream())  # pylint: disable=protected-access
        self._lib.check(rc, "lpgp_potrf_async")
        info_out.copy_(dinv[-8:].view(torch.int32)[:1])

    def trsm_block(self, Lkk: torch.Tensor, dinv: torch.Tensor, X: torch.Tensor, refine: bool = False) -> None:
        """X <- X Lkk^{-T} in place; ``refine``: the residual-corrected panel solve of a factorisation
        (``lpgp_trsm_rlt_refined``, backward stable like LAPACK's dtrsm)."""
        if X.shape[0] == 0:
            return
        f = self._factor_struct(Lkk, dinv)
        fn = self._lib.lib.lpgp_trsm_rlt_refined if refine else self._lib.lib.lpgp_trsm_rlt
        rc = fn(ctypes.byref(f), Lkk.shape[0], ctypes.c_void_p(X.data_ptr()), X.shape[0],
                self.be._ld(X), self.be._stream())  # pylint: disable=protected-access
        self._lib.check(rc, "lpgp_trsm_rlt")

    def trsv_block(self, Lkk: torch.Tensor, dinv: torch.Tensor, b: torch.Tensor, trans: bool) -> None:
        """b <- Lkk^{-1} b (trans False) or Lkk^{-T} b (trans True) in place, one right-hand side."""
        f = self._factor_struct(Lkk, dinv)
        rc = self._lib.lib.lpgp_trsv(ctypes.byref(f), int(trans), ctypes.c_void_p(b.data_ptr()), self.be._stream())  # pylint: disable=protected-access
        self._lib.check(rc, "lpgp_trsv")

    def gemv(self, A: torch.Tensor, x: torch.Tensor, y: torch.Tensor, alpha: float, trans: bool) -> None:
        """y += alpha * A x  (A^T x if trans)."""
        if A.shape[0] and A.shape[1]:
            self.be.gemv(A, x, y, alpha=alpha, trans=trans)

    def gemm_update(self, C: torch.Tensor, A: torch.Tensor, B: torch.Tensor) -> None:
        """C -= A B^T (full tiles)."""
        if C.shape[0] and C.shape[1] and A.shape[1]:
            self.be.gemm_nt(A, B, C, alpha=-1.0, beta=1.0)

    def row_sumsq(self, A: torch.Tensor, scale: float, offset: float) -> torch.Tensor:
        return self.be.row_sumsq(A, scale, offset)

    # -- FP64 emulated on the INT8 tensor cores (csrc/ozaki.cu): the O(m n^2) part of the streamed solve -------------------
    def emu_slices(self) -> int:
        return int(self.be.VARIANCE_SOLVER["ozaki_slices"])

    def emu_planes(self, rows: int, cols: int, kblock: int):
        return self.be.OzakiPlanes(rows, cols, self.emu_slices(), kblock)

    def emu_split(self, P, A: torch.Tensor, col0: int) -> None:
        """digit planes of ``A`` (rows x ncols, ncols a multiple of the K-block) into columns col0.. of ``P``"""
        P.split(A, row_off=0, col0=col0)

    def emu_gemm_update(self, XP, RP, C: torch.Tensor, k: int) -> None:
        """C -= X[:, :k] R[:, :k]^T from the digit planes of X and R."""
        self.be.ozaki_gemm_nt(XP, RP, C, k, alpha=-1.0, beta=1.0)

    def update_limited(self, C: torch.Tensor, A: torch.Tensor, B: torch.Tensor, col_limit: torch.Tensor,
                       col_base: int) -> None:
        """C -= A B^T, each block of 128 rows restricted to the columns j with col_base + j < col_limit[block]."""
        if C.shape[0] == 0 or C.shape[1] == 0:
            return
        self.be.gemm_nt_limited(A, B, C, col_limit, alpha=-1.0, beta=1.0, col_base=col_base)


class DistributedCholesky:
    """Right-looking block-row cyclic Cholesky with one panel of lookahead.

    Two CUDA streams per rank: the *panel* stream (high priority) carries the critical path of panel ``k`` --
    diagonal-block factorisation on the owner, broadcast, local TRSM, all-gather, and the update of block column
    ``k+1`` --, the *update* stream carries the bulk of the trailing update (block columns ``>= k+2``), so that the
    panel chain of step ``k+1`` (latency-bound kernels and NCCL) hides behind the DMMA GEMMs of step ``k``.  The
    host never synchronises inside the loop (the LAPACK info travels with the broadcast and is read once at the
    end).  Every gathered panel IS a block column of the factor in global row order: when ``L_full`` is passed to
    :meth:`factor`, it is filled on the fly, i.e. replicating the factor costs no additional communication."""

    def __init__(self, n: int, nb: int = 1024, group=None, ops=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.layout = BlockRowLayout(n, nb, self.world)
        self.ops = DeviceOps() if ops is None else ops
        self.n, self.nb = n, nb
        self.A_loc = self.ops.empty(self.layout.n_local(self.rank), n)
        nleaves = (n + LEAF - 1) // LEAF
        # inverted diagonal leaf blocks of the WHOLE factor (replicated) + status area, as lpgp_factor expects
        self.dinv = torch.empty(nleaves * LEAF * LEAF + 8, dtype=torch.float64, device=self.A_loc.device)
        # global end column of the diagonal block of every local 128-row block (row-limited trailing updates)
        lim = []
        for i in self.layout.local_blocks(self.rank):
            lim += [self.layout.block_bounds(i)[1]] * ((self.layout.block_size(i) + LEAF - 1) // LEAF)
        self.col_limit = torch.tensor(lim or [0], dtype=torch.int32).to(self.A_loc.device)

    # -- views ---------------------------------------------------------------------------------------------
    def local_block_rows(self, i: int) -> torch.Tensor:
        off = self.layout.local_row_offset(i)
        return self.A_loc[off : off + self.layout.block_size(i)]

    def _bcast(self, t: torch.Tensor, src: int) -> None:
        if self.world > 1:
            dist.broadcast(t, src=dist.get_global_rank(self.group, src) if self.group is not None else src, group=self.group)

    # -- factorisation -------------------------------------------------------------------------------------
    def factor(self, L_full: Optional[torch.Tensor] = None, profile: Optional[dict] = None) -> None:
        """Factor in place (``A_loc`` holds this rank's block rows of L afterwards).  ``L_full`` (n x n row-major,
        optional): receives the lower triangle of the complete factor on every rank.  ``profile`` (a dict, optional):
        receives the CUDA-event time of every stage of the panel chain summed over the panels (``potrf``, ``bcast``,
        ``trsm``, ``gather``, ``rotate``, ``update_next`` on the panel stream; ``update_rest`` on the update stream) and
        the total -- the timeline that says what the critical path of the pipeline is made of."""
        lay, ops, P, rank, nb, n = self.layout, self.ops, self.world, self.rank, self.nb, self.n
        dev = self.A_loc.device
        nblk = lay.nblk
        jmax0 = (nblk - 1 + P - 1) // P if nblk > 1 else 0          # most blocks any rank owns below block 0
        wlen = (nb // LEAF) * LEAF * LEAF
        # preallocated, reused buffers (kept alive until the streams are joined)
        packs = [torch.zeros(nb * nb + wlen + 2, dtype=torch.float64, device=dev) for _ in range(2)]
        dtmp = torch.zeros(wlen + 8, dtype=torch.float64, device=dev)
        info_all = torch.zeros(1, dtype=torch.float64, device=dev)
        send = torch.zeros(max(jmax0, 1) * nb * nb, dtype=torch.float64, device=dev)
        recv = torch.zeros(max(jmax0, 1) * P * nb * nb, dtype=torch.float64, device=dev) if P > 1 else send
        panels = [torch.zeros(max(jmax0, 1) * P * nb * nb, dtype=torch.float64, device=dev) for _ in range(2)]
        ev_first = [None, None]   # update stream: block column k+2 of step k done
        ev_rest = [None, None]    # update stream: all of step k done (panel buffer k%2 free again)

        marks = []  # (stage, start event, end event)

        def mark(stage, ev0):
            """close `stage` (started at event ev0) at the current point of the current stream; returns the new event"""
            if profile is None:
                return None
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record(torch.cuda.current_stream())
            if ev0 is not None:
                marks.append((stage, ev0, ev1))
            return ev1

        ops.fork()
        t_begin = None
        if profile is not None:
            t_begin = torch.cuda.Event(enable_timing=True)
            t_begin.record(torch.cuda.current_stream())
        for k in range(nblk):
            k0, k1 = lay.block_bounds(k)
            bk = k1 - k0
            nleaf = (bk + LEAF - 1) // LEAF
            leaf0 = k0 // LEAF
            owner = lay.owner(k)
            pack = packs[k % 2]
            Lkk = pack[: bk * bk].view(bk, bk)
            Wk = pack[nb * nb : nb * nb + nleaf * LEAF * LEAF]
            dinv_k = self.dinv[leaf0 * LEAF * LEAF : (leaf0 + nleaf) * LEAF * LEAF]
            with ops.on("panel"):
                ops.wait(ev_rest[k % 2])  # step k-2 no longer reads pack / panel buffer k%2
                e = mark("", None)
                # (1) diagonal block on its owner; [L_kk | inverted leaves | info] travels in one broadcast
                if rank == owner:
                    D = self.local_block_rows(k)[:, k0:k1]
                    dt = dtmp[: nleaf * LEAF * LEAF + 8]
                    ops.potrf_block(D, dt, pack[-1:])
                    pack[-1:].add_(float(k0) * (pack[-1:] > 0))  # position inside the whole matrix
                    Lkk.copy_(D)
                    Wk.copy_(dt[: nleaf * LEAF * LEAF])
                e = mark("potrf", e)
                self._bcast(pack, owner)
                e = mark("bcast", e)
                info_all.copy_(torch.where(info_all > 0, info_all, pack[-1:]))
                dinv_k.copy_(Wk)
                if L_full is not None:
                    L_full[k0:k1, k0:k1].copy_(Lkk)
                if k == nblk - 1:
                    break
                # (2) my rows of the panel:  X <- X L_kk^{-T}
                first = lay.first_local_block_after(rank, k)
                r_lo = first * nb
                m_loc = lay.rows_after(rank, k)
                X = self.A_loc[r_lo : r_lo + m_loc, k0:k1]
                ops.trsm_block(Lkk, Wk, X, refine=True)
                e = mark("trsm", e)
                # (3) all-gather the panel pieces; slot (r, j) = j-th block below k of rank r.  Global block
                #     k+1+t sits in slot ((k+1+t) % P, t // P): rotating the rank axis and swapping it with the
                #     slot axis puts the panel into global row order.
                J = (nblk - 1 - k + P - 1) // P
                sview = send[: J * nb * bk].view(J * nb, bk)
                sview[:m_loc].copy_(X)
                pbuf = panels[k % 2]
                if P > 1:
                    rview = recv[: P * J * nb * bk]
                    dist.all_gather_into_tensor(rview, send[: J * nb * bk], group=self.group)
                    e = mark("gather", e)
                    R = rview.view(P, J, nb, bk)
                    pv = pbuf[: J * P * nb * bk].view(J, P, nb, bk)
                    s = (k + 1) % P
                    pv[:, : P - s].copy_(R[s:].permute(1, 0, 2, 3))
                    if s:
                        pv[:, P - s :].copy_(R[:s].permute(1, 0, 2, 3))
                else:
                    pbuf[: J * nb * bk].copy_(send[: J * nb * bk])
                panel = pbuf[: (n - k1) * bk].view(n - k1, bk)
                e = mark("rotate", e)
                ev_panel = ops.record()
                # (4a) block column k+1 (the next panel) right away, on the panel stream
                k2 = lay.block_bounds(k + 1)[1]
                lim = self.col_limit[r_lo // LEAF :]
                if m_loc > 0:
                    ops.wait(ev_first[(k + 1) % 2])  # step k-1 has finished with block column k+1
                    ops.update_limited(self.A_loc[r_lo : r_lo + m_loc, k1:k2], X, panel[: k2 - k1], lim, k1)
                e = mark("update_next", e)
            with ops.on("update"):
                ops.wait(ev_panel)
                eu = mark("", None)
                if L_full is not None:
                    L_full[k1:n, k0:k1].copy_(panel)
                # (4b) block column k+2, then everything to the right of it
                if m_loc > 0 and k2 < n:
                    k3 = lay.block_bounds(k + 2)[1]
                    ops.update_limited(self.A_loc[r_lo : r_lo + m_loc, k2:k3], X, panel[k2 - k1 : k3 - k1], lim, k2)
                    ev_first[k % 2] = ops.record()
                    if k3 < n:
                        ops.update_limited(self.A_loc[r_lo : r_lo + m_loc, k3:n], X, panel[k3 - k1 :], lim, k3)
                else:
                    ev_first[k % 2] = ops.record()
                ev_rest[k % 2] = ops.record()
                mark("update_rest", eu)
        ops.join()
        if profile is not None:
            t_end = torch.cuda.Event(enable_timing=True)
            t_end.record(torch.cuda.current_stream())
            torch.cuda.synchronize()
            profile.clear()
            for stage, e0, e1 in marks:
                profile[stage] = profile.get(stage, 0.0) + e0.elapsed_time(e1)
            profile["total"] = t_begin.elapsed_time(t_end)
            profile["panels"] = nblk
        info = int(info_all.item())
        if info > 0:  # every rank raises together (pn/linops/_linear_operator.py:823-839 semantics)
            import numpy as np

            raise np.linalg.LinAlgError(f"{info}-th leading minor of the array is not positive definite")

    # -- solves with the DISTRIBUTED factor (nothing replicated) ---------------------------------------------------
    def _dinv_block(self, k: int) -> torch.Tensor:
        k0, k1 = self.layout.block_bounds(k)
        nleaf = (k1 - k0 + LEAF - 1) // LEAF
        return self.dinv[(k0 // LEAF) * LEAF * LEAF : (k0 // LEAF + nleaf) * LEAF * LEAF]

    def solve(self, b: torch.Tensor) -> torch.Tensor:
        """x = G^{-1} b = L^{-T} L^{-1} b for ONE replicated right-hand side (the representer weights,
        _conditional.py:44), with L left in its block-row cyclic distribution.  Owner-computes: block k's owner
        does the GEMV with its rows and the substitution with its diagonal block, then broadcasts the nb new
        entries (forward); the transposed products of the backward sweep are accumulated per rank and summed with
        one small all-reduce per block.  O(n^2) flops, n/nb latency-bound steps."""
        lay, ops, rank = self.layout, self.ops, self.rank
        y = b.detach().clone().reshape(-1).contiguous()
        assert y.numel() == self.n
        for k in range(lay.nblk):  # forward: L y = b
            k0, k1 = lay.block_bounds(k)
            yk = y[k0:k1]
            if rank == lay.owner(k):
                rows = self.local_block_rows(k)
                ops.gemv(rows[:, :k0], y[:k0], yk, -1.0, False)
                ops.trsv_block(rows[:, k0:k1], self._dinv_block(k), yk, False)
            self._bcast(yk, lay.owner(k))
        acc = torch.zeros_like(y)
        for k in reversed(range(lay.nblk)):  # backward: L^T x = y
            k0, k1 = lay.block_bounds(k)
            yk = y[k0:k1]
            part = acc[k0:k1]
            if self.world > 1:
                dist.all_reduce(part, group=self.group)
            if rank == lay.owner(k):
                rows = self.local_block_rows(k)
                yk.sub_(part)
                ops.trsv_block(rows[:, k0:k1], self._dinv_block(k), yk, True)
            self._bcast(yk, lay.owner(k))
            if rank == lay.owner(k):
                ops.gemv(rows[:, :k0], yk, acc[:k0], 1.0, True)
        return y

    def solve_rows(self, K: torch.Tensor) -> torch.Tensor:
        """K <- K L^{-T} in place for the rows ``K`` (m_loc x n) held by THIS rank (cross-covariance rows of its
        test points, ``_conditional.py:245-251``), with the factor streamed instead of replicated: block row k of
        L is broadcast by its owner (double-buffered, on the panel stream) while every rank applies block row
        k-1 to its own rows -- left-looking,  K[:, k] <- (K[:, k] - K[:, :k0] L[k, :k0]^T) L_kk^{-T} -- so the
        N^2 M flops shard over the test points and the factor crosses NVLink exactly once per call.  Collective:
        every rank must call it (with its own, possibly empty, set of rows)."""
        lay, ops, P, rank, nb, n = self.layout, self.ops, self.world, self.rank, self.nb, self.n
        dev = self.A_loc.device
        bufs = [torch.empty(nb * n, dtype=torch.float64, device=dev) for _ in range(2)]
        ev_ready, ev_free = [None, None], [None, None]
        # INT8-emulated update: the streamed block row and the solved column blocks of K are split into
        # digit planes (K-block = nb columns) and the O(m n^2) products run on the tcgen05 tensor cores; the nb x nb
        # diagonal solves stay on DMMA.  Every rank splits the block rows it receives itself (15 B per entry, HBM-bound).
        emulate = (K.shape[0] > 0 and hasattr(ops, "emu_slices") and ops.emu_slices() > 0 and 128 <= nb <= 4096
                   and n >= 2 * nb and K.stride(0) % 2 == 0)
        XP = RP = None
        if emulate:
            XP = ops.emu_planes(K.shape[0], n, nb)
            RP = [ops.emu_planes(nb, n, nb) for _ in range(2)]
        ops.fork()
        for k in range(lay.nblk):
            k0, k1 = lay.block_bounds(k)
            bk = k1 - k0
            row = bufs[k % 2][: bk * k1].view(bk, k1)
            with ops.on("panel"):
                ops.wait(ev_free[k % 2])
                if rank == lay.owner(k):
                    row.copy_(self.local_block_rows(k)[:, :k1])
                self._bcast(row, lay.owner(k))
                ev_ready[k % 2] = ops.record()
            with ops.on("update"):
                ops.wait(ev_ready[k % 2])
                if K.shape[0] > 0:
                    Kk = K[:, k0:k1]
                    if emulate and k0 > 0:
                        ops.emu_split(RP[k % 2], row[:, :k0], 0)
                        ops.emu_gemm_update(XP, RP[k % 2], Kk, k0)
                    else:
                        ops.gemm_update(Kk, K[:, :k0], row[:, :k0])
                    ops.trsm_block(row[:, k0:k1], self._dinv_block(k), Kk)
                    if emulate and bk == nb and k + 1 < lay.nblk:
                        ops.emu_split(XP, Kk, k0)
                ev_free[k % 2] = ops.record()
        ops.join()
        return K

    def post_var(self, K: torch.Tensor, prior_diag: float) -> torch.Tensor:
        """var_i = prior_diag - || K_i L^{-T} ||^2 (``_conditional.py:223-231``); ``K`` is overwritten.  Collective."""
        return self.ops.row_sumsq(self.solve_rows(K), -1.0, prior_diag)

    # -- replication ---------------------------------------------------------------------------------------
    def replicate_into(self, L_full: torch.Tensor) -> None:
        """Every rank receives every block row of the factor, straight into ``L_full`` (n x n row-major).  Only
        needed when :meth:`factor` ran without ``L_full``."""
        lay = self.layout
        for i in range(lay.nblk):
            lo, hi = lay.block_bounds(i)
            slab = L_full[lo:hi]
            if lay.owner(i) == self.rank:
                slab[:, :hi].copy_(self.local_block_rows(i)[:, :hi])
            if self.world > 1:
                # the slab view covers whole (padded) rows of the buffer: contiguous memory
                flat = torch.as_strided(slab, (hi - lo, slab.stride(0)), (slab.stride(0), 1)) if hi - lo > 1 else slab
                self._bcast(flat, lay.owner(i))


class DistributedFactor:
    """The interface of ``backend.DeviceFactor`` that posterior evaluation needs, on top of a factor that STAYS
    distributed (block rows spread over the ranks; nothing replicated): N = 131,072 (137 GB Gram, BASELINE.json
    configs[4]) does not fit one GPU.  All methods are collective."""

    distributed = True

    def __init__(self, ch: DistributedCholesky):
        self.ch = ch
        self.n = ch.n

    def potrs(self, B: torch.Tensor) -> torch.Tensor:
        if B.dim() == 1:
            B = B.reshape(1, -1)
        for r in range(B.shape[0]):
            B[r].copy_(self.ch.solve(B[r]))
        return B

    def trsm_rlt(self, X: torch.Tensor, nlead: Optional[int] = None) -> torch.Tensor:
        if nlead is not None and nlead != self.n:
            raise NotImplementedError("partial solves with a distributed factor")
        return self.ch.solve_rows(X)

    def extended(self, new_size: int):
        raise NotImplementedError(
            "a posterior whose factor is distributed over several GPUs cannot be extended by bordering; "
            "pass all batches to from_observation_batches, or use replicate=True")

    def post_var(self, blocks, Xt: torch.Tensor, prior_diag: float, min_chunk_bytes: int = 4 << 30) -> torch.Tensor:
        """Pointwise posterior variance of THIS rank's test points ``Xt`` (m_loc x d, device): cross-covariance
        rows are assembled chunk by chunk (as many rows as fit half of the free device memory, so that the factor
        is streamed as few times as possible) and solved against the distributed factor.  Collective."""
        from . import backend

        n = self.n
        free = torch.cuda.mem_get_info()[0]
        # with the emulated update the digit planes of the chunk (7 B per entry) sit next to it (8 B per entry)
        emulated = hasattr(self.ch.ops, "emu_slices") and self.ch.ops.emu_slices() > 0
        budget = max(min_chunk_bytes, min(int((0.27 if emulated else 0.5) * free), 64 << 30))
        chunk = int(max(256, budget // (8 * backend.round_up(n, 16))))
        m = Xt.shape[0]
        out = torch.empty(m, dtype=torch.float64, device=Xt.device)
        K = backend.alloc_matrix(min(chunk, max(m, 1)), n)
        for p in range(self.passes(m, chunk)):
            lo, hi = min(m, p * chunk), min(m, (p + 1) * chunk)
            Kc = K[: hi - lo]
            if hi > lo:
                backend.crosscov(blocks, n, Xt[lo:hi], out=Kc)
            v = self.ch.post_var(Kc, prior_diag)
            if hi > lo:
                out[lo:hi].copy_(v)
        return out

    def passes(self, m_loc: int, chunk: int) -> int:
        """number of chunk passes every rank must make so that the collective solves line up"""
        t = torch.tensor([(m_loc + chunk - 1) // chunk], dtype=torch.int64, device=self.ch.A_loc.device)
        if self.ch.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.ch.group)
        return int(t.item())


# ================================================================================================================
# 2-D block-cyclic variant of the factorisation (the layout BASELINE.json's north star names)
# ================================================================================================================
class Grid2DLayout:
    """Block (i, j) of the lower triangle (``nb x nb``) lives on the rank at grid position (i % pr, j % pc) of a
    ``pr x pc`` process grid (rank = row * pc + column), stored at local block position (i // pr, j // pc)."""

    def __init__(self, n: int, nb: int, pr: int, pc: int):
        if nb % LEAF:
            raise ValueError("nb must be a multiple of 128")
        if n % nb:
            raise ValueError("the 2-D layout needs n to be a multiple of nb")
        self.n, self.nb, self.pr, self.pc = int(n), int(nb), int(pr), int(pc)
        self.nblk = n // nb

    def coords(self, rank: int) -> Tuple[int, int]:
        return rank // self.pc, rank % self.pc

    def rank_of(self, r: int, c: int) -> int:
        return r * self.pc + c

    def owner(self, i: int, j: int) -> int:
        return self.rank_of(i % self.pr, j % self.pc)

    def n_local_rows(self, r: int) -> int:
        return len(range(r, self.nblk, self.pr))

    def n_local_cols(self, c: int) -> int:
        return len(range(c, self.nblk, self.pc))

    @staticmethod
    def count_le(k: int, first: int, step: int) -> int:
        """number of indices first, first + step, ... that are <= k"""
        return 0 if k < first else (k - first) // step + 1


_GRID_GROUPS: dict = {}  # (pr, pc, world) -> (row groups, column groups) of the default process group


class BlockCyclic2DCholesky:
    """Right-looking Cholesky on a ``pr x pc`` process grid with one panel of lookahead (same two-stream pipeline as
    :class:`DistributedCholesky`).  Per panel ``k``:

      1. the owner of block (k, k) factors it and broadcasts ``[L_kk | inverted leaves | info]`` down its process COLUMN;
      2. the ``pr`` ranks of that column solve their blocks of the panel, ``X <- X L_kk^{-T}``;
      3. the panel is broadcast along the process ROWS (every rank receives the blocks ``L_ik`` of its own block rows),
         then the blocks whose index also belongs to a rank's block COLUMNS are all-gathered inside each process column
         (the "transposed" panel ``L_jk``, j = c mod pc);
      4. every rank updates its local blocks, ``A_ij -= L_ik L_jk^T`` for i >= j > k, with one row-limited DMMA GEMM
         (the local lower "staircase" is expressed through the per-row column limits of ``lpgp_gemm_nt_limited``).

    Compared with the 1-D block-row layout every rank receives ``(1/pr + 1/pc)`` of each panel instead of all of it, but
    only ``pr`` ranks share a panel's triangular solve.  Built to MEASURE the layout the north star names against the 1-D
    one on one NVSwitch domain (tools/dist_bench2d.py); posterior evaluation uses the 1-D class."""

    def __init__(self, n: int, nb: int, pr: int, pc: int, ops=None):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if pr * pc != self.world:
            raise ValueError(f"process grid {pr} x {pc} does not match the world size {self.world}")
        self.layout = lay = Grid2DLayout(n, nb, pr, pc)
        self.ops = DeviceOps() if ops is None else ops
        self.n, self.nb = n, nb
        self.r, self.c = lay.coords(self.rank)
        # sub-communicators: every rank creates every group, in the same order; cached per grid shape and warmed up
        # with one tiny collective each (NCCL builds a communicator lazily at its first use -- hundreds of milliseconds
        # that must not land inside a factorisation)
        self.row_group = self.col_group = None
        if self.world > 1:
            key = (pr, pc, self.world)
            if key not in _GRID_GROUPS:
                rows = [dist.new_group([lay.rank_of(r, c) for c in range(pc)]) for r in range(pr)]
                cols = [dist.new_group([lay.rank_of(r, c) for r in range(pr)]) for c in range(pc)]
                _GRID_GROUPS[key] = (rows, cols)
                warm = torch.zeros(1, dtype=torch.float64, device=self.ops.device)
                dist.all_reduce(warm, group=rows[self.r])
                dist.all_reduce(warm, group=cols[self.c])
            rows, cols = _GRID_GROUPS[key]
            self.row_group, self.col_group = rows[self.r], cols[self.c]
        self.nbr, self.nbc = lay.n_local_rows(self.r), lay.n_local_cols(self.c)
        self.A_loc = self.ops.empty(max(self.nbr, 1) * nb, max(self.nbc, 1) * nb)
        dev = self.A_loc.device
        # per local 128-row block: end (in local columns) of the blocks j <= i of its block row i
        lim = []
        for li in range(self.nbr):
            i = li * pr + self.r
            lim += [lay.count_le(i, self.c, pc) * nb] * (nb // LEAF)
        self.col_limit = torch.tensor(lim or [0], dtype=torch.int32).to(dev)

    def local_block(self, i: int, j: int) -> torch.Tensor:
        nb = self.nb
        li, lj = i // self.layout.pr, j // self.layout.pc
        return self.A_loc[li * nb : (li + 1) * nb, lj * nb : (lj + 1) * nb]

    def owns(self, i: int, j: int) -> bool:
        return self.layout.owner(i, j) == self.rank

    def factor(self, profile: Optional[dict] = None) -> None:
        lay, ops, nb, rank = self.layout, self.ops, self.nb, self.rank
        pr, pc, r, c, nblk = lay.pr, lay.pc, self.r, self.c, lay.nblk
        dev = self.A_loc.device
        wlen = (nb // LEAF) * LEAF * LEAF
        packs = [torch.zeros(nb * nb + wlen + 2, dtype=torch.float64, device=dev) for _ in range(2)]
        dtmp = torch.zeros(wlen + 8, dtype=torch.float64, device=dev)
        info_all = torch.zeros(1, dtype=torch.float64, device=dev)
        lrows = [torch.zeros(max(self.nbr, 1) * nb * nb, dtype=torch.float64, device=dev) for _ in range(2)]
        lcols = [torch.zeros(max(self.nbc, 1) * nb * nb, dtype=torch.float64, device=dev) for _ in range(2)]
        # blocks of my rows whose index is also one of my columns: i = r (mod pr) and i = c (mod pc)
        period = pr * pc // _gcd(pr, pc)
        cnt_max = (nblk + period - 1) // period + 1
        send = torch.zeros(cnt_max * nb * nb, dtype=torch.float64, device=dev)
        recv = torch.zeros(pr * cnt_max * nb * nb, dtype=torch.float64, device=dev)
        ev_first, ev_rest = [None, None], [None, None]
        t_begin = t_end = None
        if profile is not None:
            t_begin = torch.cuda.Event(enable_timing=True)
            t_begin.record(torch.cuda.current_stream())

        def both(ii: int, rr: int) -> bool:
            return ii % pr == rr and ii % pc == c

        ops.fork()
        for k in range(nblk):
            rk, ck = k % pr, k % pc
            in_col = c == ck
            lr0, lc0 = lay.count_le(k, r, pr), lay.count_le(k, c, pc)  # local rows / columns with index <= k
            m_loc, n_loc = (self.nbr - lr0) * nb, (self.nbc - lc0) * nb
            pack = packs[k % 2]
            Lkk = pack[: nb * nb].view(nb, nb)
            Wk = pack[nb * nb : nb * nb + wlen]
            with ops.on("panel"):
                ops.wait(ev_rest[k % 2])
                if in_col:
                    if r == rk:
                        D = self.local_block(k, k)
                        ops.potrf_block(D, dtmp, pack[-1:])
                        pack[-1:].add_(float(k * nb) * (pack[-1:] > 0))
                        Lkk.copy_(D)
                        Wk.copy_(dtmp[:wlen])
                    if self.col_group is not None and pr > 1:
                        dist.broadcast(pack, src=lay.rank_of(rk, ck), group=self.col_group)
                    info_all.copy_(torch.where(info_all > 0, info_all, pack[-1:]))
                if k == nblk - 1:
                    break
                Lrow = lrows[k % 2][: m_loc * nb].view(m_loc, nb)
                if in_col and m_loc > 0:
                    X = self.A_loc[lr0 * nb : lr0 * nb + m_loc, (k // pc) * nb : (k // pc + 1) * nb]
                    ops.trsm_block(Lkk, Wk, X, refine=True)
                    Lrow.copy_(X)
                # (3a) along my process row: the blocks L_ik of my block rows
                if self.row_group is not None and pc > 1 and m_loc > 0:
                    dist.broadcast(lrows[k % 2][: m_loc * nb], src=lay.rank_of(r, ck), group=self.row_group)
                # (3b) inside my process column: the blocks L_jk of my block columns (j = c mod pc), from whichever
                #      process row holds them (j mod pr)
                Lcol = lcols[k % 2][: n_loc * nb].view(n_loc, nb)
                if n_loc > 0:
                    # both index sets are arithmetic progressions with the period lcm(pr, pc): strided block copies
                    sv = send.view(cnt_max, nb, nb)
                    mine = [t for t in range(self.nbr - lr0) if both((lr0 + t) * pr + r, r)]
                    if mine:
                        sv[: len(mine)].copy_(Lrow.view(-1, nb, nb)[mine[0] :: period // pr][: len(mine)])
                    if self.col_group is not None and pr > 1:
                        dist.all_gather_into_tensor(recv, send, group=self.col_group)
                        rv = recv.view(pr, cnt_max, nb, nb)
                    else:
                        rv = sv.view(1, cnt_max, nb, nb)
                    Lc3 = Lcol.view(-1, nb, nb)
                    for rr in range(pr):
                        dst = [lj - lc0 for lj in range(lc0, self.nbc) if (lj * pc + c) % pr == rr]
                        if dst:
                            Lc3[dst[0] :: period // pc][: len(dst)].copy_(rv[rr, : len(dst)])
                ev_panel = ops.record()
                lim = self.col_limit[lr0 * (nb // LEAF) :]
                C = self.A_loc[lr0 * nb : lr0 * nb + m_loc, lc0 * nb : lc0 * nb + n_loc]
                first_cols = 0
                if m_loc > 0 and n_loc > 0 and (k + 1) % pc == c:  # (4a) block column k+1 = the next panel, right away
                    ops.wait(ev_first[(k + 1) % 2])
                    ops.update_limited(C[:, :nb], Lrow, Lcol[:nb], lim, lc0 * nb)
                    first_cols = 1
            with ops.on("update"):
                ops.wait(ev_panel)
                done = first_cols
                if m_loc > 0 and n_loc > done * nb and k + 2 < nblk and (k + 2) % pc == c:  # block column k+2 first
                    ops.update_limited(C[:, done * nb : (done + 1) * nb], Lrow, Lcol[done * nb : (done + 1) * nb], lim,
                                       (lc0 + done) * nb)
                    done += 1
                ev_first[k % 2] = ops.record()
                if m_loc > 0 and n_loc > done * nb:
                    ops.update_limited(C[:, done * nb :], Lrow, Lcol[done * nb :], lim, (lc0 + done) * nb)
                ev_rest[k % 2] = ops.record()
        ops.join()
        if profile is not None:
            t_end = torch.cuda.Event(enable_timing=True)
            t_end.record(torch.cuda.current_stream())
            torch.cuda.synchronize()
            profile["total"] = t_begin.elapsed_time(t_end)
        if self.world > 1:
            dist.all_reduce(info_all, op=dist.ReduceOp.MAX)
        info = int(info_all.item())
        if info > 0:
            import numpy as np

            raise np.linalg.LinAlgError(f"{info}-th leading minor of the array is not positive definite")

    def gather_full(self, L_full: torch.Tensor) -> None:
        """Every rank receives every block of the factor (validation at small sizes: one broadcast per block)."""
        lay, nb = self.layout, self.nb
        tmp = torch.zeros((nb, nb), dtype=torch.float64, device=self.A_loc.device)
        for i in range(lay.nblk):
            for j in range(i + 1):
                if self.owns(i, j):
                    tmp.copy_(self.local_block(i, j))
                if self.world > 1:
                    dist.broadcast(tmp, src=lay.owner(i, j))
                L_full[i * nb : (i + 1) * nb, j * nb : (j + 1) * nb].copy_(tmp)


def _gcd(a: int, b: int) -> int:
    while b:
        a, b = b, a % b
    return a
