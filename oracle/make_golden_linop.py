"""Golden vectors for matrix-composed functionals ``A @ linfunctl`` (tests/golden/seams_linop.npz), from the REAL reference.

TEST INFRASTRUCTURE ONLY (build container: needs /root/reference through ``oracle/refshim.py``).

    python -m oracle.make_golden_linop

Per seam kernel and operator of ``tests/golden/cases.py``: the reference's ``CompositeLinearFunctional(linop=A, ...)``
(src/linpde_gp/linfunctls/_arithmetic.py:92-174) applied to the kernel (``LinOpProcessVectorCrossCovariance``,
crosscov/_arithmetic.py:91-130, covfuncs/linfunctls/_registry.py:42-61), evaluated at test points.  Every vector is also
checked against the plain-matrix identity ``(unmixed object) @ A^T`` computed from the same reference objects.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from oracle.make_golden import GOLDEN, ref_kernel, ref_op  # noqa: E402
from tests.golden import cases as gcases  # noqa: E402


def main() -> float:
    pn, lg = refshim.load()  # pylint: disable=unused-variable
    from linpde_gp import linfunctls

    rng = np.random.default_rng(77)
    out = {"X": rng.uniform(0, 1, (37, 2)), "Xt": rng.uniform(0, 1, (5, 4, 2)), "X0": rng.uniform(0, 1, (3, 6, 2)),
           "A": rng.standard_normal((7, 37)), "B": rng.standard_normal((2, 7))}
    worst = 0.0
    for kname, kspec in gcases.SEAM_KERNELS.items():
        k = ref_kernel(kspec)
        for oname, ospec in gcases.SEAM_OPS.items():
            L = ref_op(ospec, (2,))
            fctl = (linfunctls._EvaluationFunctional((2,), (), out["X"]) if L is None  # pylint: disable=protected-access
                    else L.to_linfunctl(out["X"]))
            comp = out["B"] @ (out["A"] @ fctl)  # CompositeLinearFunctional(linop = B A, ...)
            assert type(comp).__name__ == "CompositeLinearFunctional" and comp.output_shape == (2,)
            pv = comp(k, argnum=1)
            assert type(pv).__name__ == "LinOpProcessVectorCrossCovariance" and not pv.reverse
            v = np.asarray(pv(out["Xt"]))
            assert v.shape == (5, 4, 2)
            out[f"pv__{kname}__{oname}"] = v
            plain = np.asarray(fctl(k, argnum=1)(out["Xt"]))
            worst = max(worst, np.max(np.abs(v - plain @ (out["B"] @ out["A"]).T)) / np.max(np.abs(v)))
            # (a second evaluation functional applied to it, f0(pv), cannot be generated: the reference's
            # LinOpProcessVectorCrossCovariance._evaluate_linop multiplies `linop @ covop` also for reverse=False, where
            # covop is (M, n): "Shape mismatch: Cannot multiply linear operators with shapes (2, 37) x (18, 37)",
            # crosscov/_arithmetic.py:127-130 -- the tests check that object against A @ (plain covariance) instead)
    np.savez(os.path.join(GOLDEN, "seams_linop.npz"), **out)
    print(f"seams_linop.npz: {len(out)} arrays, worst deviation from A @ (plain object) {worst:.2e}")
    return worst


if __name__ == "__main__":
    main()
