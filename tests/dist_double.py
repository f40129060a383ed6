"""CPU test double of ``linpde_gp_b200.distributed.DeviceOps`` (plain torch on the host), used ONLY by the gloo
test to exercise the ownership / packing / collective logic of the distributed Cholesky without a GPU."""
import torch

LEAF = 128


class HostOps:
    device = torch.device("cpu")

    def empty(self, rows, cols):
        ld = max((cols + 15) // 16 * 16, 16)
        return torch.zeros((max(rows, 1), ld), dtype=torch.float64)[:rows, :cols]

    # streams / events do not exist on the host: the pipeline degenerates to program order
    def on(self, which):
        import contextlib

        return contextlib.nullcontext()

    def fork(self):
        pass

    def join(self):
        pass

    def record(self):
        return None

    def wait(self, ev):
        pass

    def potrf_block(self, D, dinv, info_out):
        n = D.shape[0]
        G = torch.tril(D) + torch.tril(D, -1).T
        info_out.zero_()
        try:
            L = torch.linalg.cholesky(G)
        except Exception:
            info_out.fill_(1.0)
            return
        D.copy_(torch.tril(L) + torch.triu(D, 1))
        nleaf = (n + LEAF - 1) // LEAF
        W = dinv[: nleaf * LEAF * LEAF].view(nleaf, LEAF, LEAF)
        W.zero_()
        for l in range(nleaf):
            lo, hi = l * LEAF, min(n, (l + 1) * LEAF)
            W[l] = torch.eye(LEAF, dtype=torch.float64)
            W[l, : hi - lo, : hi - lo] = torch.linalg.inv(L[lo:hi, lo:hi])

    def trsm_block(self, Lkk, dinv, X, refine=False):
        if X.shape[0] == 0:
            return
        X.copy_(torch.linalg.solve_triangular(torch.tril(Lkk), X.T.contiguous(), upper=False).T)

    def trsv_block(self, Lkk, dinv, b, trans):
        L = torch.tril(Lkk)
        b.copy_(torch.linalg.solve_triangular(L.T if trans else L, b.reshape(-1, 1), upper=bool(trans)).reshape(-1))

    def gemv(self, A, x, y, alpha, trans):
        if A.shape[0] and A.shape[1]:
            y.add_(alpha * ((A.T if trans else A) @ x))

    def gemm_update(self, C, A, B):
        if C.shape[0] and C.shape[1] and A.shape[1]:
            C.sub_(A @ B.T)

    def row_sumsq(self, A, scale, offset):
        return offset + scale * (A * A).sum(dim=1)

    def update_limited(self, C, A, B, col_limit, col_base):
        if C.shape[0] == 0 or C.shape[1] == 0:
            return
        full = A @ B.T
        for t in range((C.shape[0] + LEAF - 1) // LEAF):
            lim = max(int(col_limit[t]) - col_base, 0)
            rows = slice(t * LEAF, min(C.shape[0], (t + 1) * LEAF))
            # the device kernel works on whole 128-column tiles: columns up to the tile boundary may be touched
            lim_tile = min(C.shape[1], (lim + LEAF - 1) // LEAF * LEAF)
            C[rows, :lim_tile] -= full[rows, :lim_tile]
