"""CPU test double of the few ``linpde_gp_b200.backend`` calls behind the symbolic seam classes (cross-covariances,
functionals of cross-covariances, matrix-composed functionals) -- TEST INFRASTRUCTURE ONLY, used by
``tests/test_host_seams.py`` to run the HOST logic of those classes (atoms, shapes, layouts, `reverse`, accumulation,
alpha scaling) against the real-reference goldens without a GPU.

The double evaluates a kernel descriptor with ``tests/helpers.eval_desc_numpy`` (the numpy statement of
``include/lpgp.h::lpgp_kernel_desc``) and the GEMM with torch on the host.  It is installed by monkeypatching inside a
test and never imported by the product: the product itself has no CPU path (``backend._require_cuda``)."""
import numpy as np
import torch

from tests import helpers


def install(monkeypatch):
    """Patch ``backend`` for the duration of one test (pytest's ``monkeypatch`` undoes it)."""
    from linpde_gp_b200 import backend

    cpu = torch.device("cpu")

    def to_device(x, *, pinned=False):  # pylint: disable=unused-argument
        if isinstance(x, torch.Tensor):
            return x.to(dtype=torch.float64).contiguous()
        return torch.from_numpy(np.array(x, dtype=np.float64, order="C", copy=True))

    def points(X, d):
        return to_device(X).reshape(-1, d)

    def gram(desc, X0, X1=None, out=None, *, lower=False, accumulate=False, alpha=1.0):  # pylint: disable=unused-argument
        X1_ = X0 if X1 is None else X1
        K = torch.from_numpy(helpers.eval_desc_numpy(desc, X0.numpy(), X1_.numpy())) * alpha
        if out is None:
            out, accumulate = backend.alloc_matrix(X0.shape[0], X1_.shape[0]), False
        if accumulate:
            out.add_(K)
        else:
            out.copy_(K)
        return out

    def gemm_nt(A, B, C, alpha=1.0, beta=0.0, lower=False):  # pylint: disable=unused-argument
        P = alpha * (A @ B.T)
        C.copy_(P if beta == 0.0 else beta * C + P)  # beta == 0: C is not read (it may hold anything)
        return C

    monkeypatch.setattr(backend, "_require_cuda", lambda: cpu)
    monkeypatch.setattr(backend, "to_device", to_device)
    monkeypatch.setattr(backend, "points", points)
    monkeypatch.setattr(backend, "gram", gram)
    monkeypatch.setattr(backend, "gemm_nt", gemm_nt)
