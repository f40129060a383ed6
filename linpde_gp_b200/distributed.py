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
    """Local numerical kernels on the GPU (C ABI through ctypes)."""

    def __init__(self):
        from . import _lib, backend

        self._lib, self.be = _lib, backend
        self.device = backend._require_cuda()  # pylint: disable=protected-access

    def empty(self, rows: int, cols: int) -> torch.Tensor:
        return self.be.alloc_matrix(rows, cols)

    def zeros_i32(self, n: int) -> torch.Tensor:
        return torch.zeros(n, dtype=torch.int32, device=self.device)

    def _factor_struct(self, L: torch.Tensor, dinv: torch.Tensor):
        f = self._lib.Factor()
        f.L = L.data_ptr()
        f.n = L.shape[0]
        f.ld = self.be._ld(L)  # pylint: disable=protected-access
        f.dinv = dinv.data_ptr()
        f.nseg = 1
        f.seg_off[0], f.seg_off[1] = 0, L.shape[0]
        return f

    def potrf_block(self, D: torch.Tensor, dinv: torch.Tensor) -> int:
        """Cholesky of the square view ``D`` in place; ``dinv`` receives the inverted leaf blocks (+ status).
        Returns the LAPACK info (0 = ok, > 0 = order of the first non-positive-definite leading minor)."""
        f = self._factor_struct(D, dinv)
        rc = self._lib.lib.lpgp_potrf(ctypes.byref(f), self.be._stream())  # pylint: disable=protected-access
        if rc < 0:
            self._lib.check(rc, "lpgp_potrf")
        return int(rc)

    def trsm_block(self, Lkk: torch.Tensor, dinv: torch.Tensor, X: torch.Tensor) -> None:
        """X <- X Lkk^{-T} in place."""
        if X.shape[0] == 0:
            return
        f = self._factor_struct(Lkk, dinv)
        rc = self._lib.lib.lpgp_trsm_rlt(ctypes.byref(f), Lkk.shape[0], ctypes.c_void_p(X.data_ptr()), X.shape[0],
                                         self.be._ld(X), self.be._stream())  # pylint: disable=protected-access
        self._lib.check(rc, "lpgp_trsm_rlt")

    def update_limited(self, C: torch.Tensor, A: torch.Tensor, B: torch.Tensor, col_limit: torch.Tensor) -> None:
        """C -= A B^T, each block of 128 rows restricted to the columns < col_limit[block]."""
        if C.shape[0] == 0 or C.shape[1] == 0:
            return
        self.be.gemm_nt_limited(A, B, C, col_limit, alpha=-1.0, beta=1.0)


class DistributedCholesky:
    def __init__(self, n: int, nb: int = 512, group=None, ops=None):
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

    # -- views ---------------------------------------------------------------------------------------------
    def local_block_rows(self, i: int) -> torch.Tensor:
        off = self.layout.local_row_offset(i)
        return self.A_loc[off : off + self.layout.block_size(i)]

    def _bcast(self, t: torch.Tensor, src: int) -> None:
        if self.world > 1:
            dist.broadcast(t, src=dist.get_global_rank(self.group, src) if self.group is not None else src, group=self.group)

    # -- factorisation -------------------------------------------------------------------------------------
    def factor(self) -> None:
        lay, ops, P, rank, nb = self.layout, self.ops, self.world, self.rank, self.nb
        dev = self.A_loc.device
        pack = torch.zeros(nb * nb + nb * LEAF + 2, dtype=torch.float64, device=dev)  # [L_kk | W leaves | info]
        for k in range(lay.nblk):
            k0, k1 = lay.block_bounds(k)
            bk = k1 - k0
            nleaf = (bk + LEAF - 1) // LEAF
            leaf0 = k0 // LEAF
            owner = lay.owner(k)
            Lkk = pack[: bk * bk].view(bk, bk)
            Wk = pack[nb * nb : nb * nb + nleaf * LEAF * LEAF]
            dinv_k = self.dinv[leaf0 * LEAF * LEAF : (leaf0 + nleaf) * LEAF * LEAF + 8]
            if rank == owner:
                D = self.local_block_rows(k)[:, k0:k1]
                info = ops.potrf_block(D, dinv_k)
                pack[-1] = float(info + k0 if info > 0 else 0)
                Lkk.copy_(D)
                Wk.copy_(dinv_k[: nleaf * LEAF * LEAF])
            self._bcast(pack, owner)
            info = int(pack[-1].item())
            if info > 0:  # every rank raises together (pn/linops/_linear_operator.py:823-839 semantics)
                import numpy as np

                raise np.linalg.LinAlgError(f"{info}-th leading minor of the array is not positive definite")
            if rank != owner:
                dinv_k[: nleaf * LEAF * LEAF].copy_(Wk)
            if k == lay.nblk - 1:
                break
            # (2) my rows of the panel
            first = lay.first_local_block_after(rank, k)
            r_lo = first * nb
            m_loc = lay.rows_after(rank, k)
            X = self.A_loc[r_lo : r_lo + m_loc, k0:k1]
            ops.trsm_block(Lkk, pack[nb * nb : nb * nb + nleaf * LEAF * LEAF], X)
            # (3) all-gather the panel, global row order
            m_all = [lay.rows_after(r, k) for r in range(P)]
            m_max = max(m_all)
            send = torch.zeros((m_max, bk), dtype=torch.float64, device=dev)
            send[:m_loc].copy_(X)
            if P > 1:
                recv = torch.empty((P * m_max, bk), dtype=torch.float64, device=dev)
                dist.all_gather_into_tensor(recv, send, group=self.group)
                g = torch.arange(k1, self.n, device=dev)
                blk = torch.div(g, nb, rounding_mode="floor")
                rk = blk % P
                firsts = torch.tensor([lay.first_local_block_after(r, k) for r in range(P)], device=dev)
                idx = rk * m_max + (torch.div(blk, P, rounding_mode="floor") - firsts[rk]) * nb + g % nb
                panel = ops.empty(self.n - k1, bk)
                panel.copy_(recv.index_select(0, idx))
            else:
                panel = ops.empty(self.n - k1, bk)
                panel.copy_(send[:m_loc])
            # (4) trailing update of my block rows
            if m_loc > 0:
                C = self.A_loc[r_lo : r_lo + m_loc, k1 : self.n]
                lim = []
                for i in lay.local_blocks(rank):
                    if i > k:
                        lim += [lay.block_bounds(i)[1] - k1] * ((lay.block_size(i) + LEAF - 1) // LEAF)
                col_limit = torch.tensor(lim, dtype=torch.int32, device=dev)
                ops.update_limited(C, X, panel, col_limit)

    # -- replication ---------------------------------------------------------------------------------------
    def replicate_into(self, L_full: torch.Tensor) -> None:
        """Every rank receives every block row of the factor, straight into ``L_full`` (n x n row-major)."""
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
