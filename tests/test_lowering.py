"""Host logic: the (kernel, L0, L1) -> lpgp_kernel_desc lowering, checked on CPU against the golden outputs of
the real reference by evaluating the descriptor's documented semantics in numpy."""
import json
import os

import numpy as np
import pytest

from tests import helpers
from tests.golden import cases as gcases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
_K = np.load(os.path.join(GOLDEN, "kernels.npz"))
SPECS = json.loads(bytes(_K["__specs__"]).decode())


# every reference case lowers to a device descriptor (product form, or the radial family for the isotropic
# multi-dimensional Matern kernels)
LOWERABLE = SPECS


@pytest.mark.parametrize("spec", LOWERABLE, ids=[s["name"] for s in LOWERABLE])
def test_descriptor_reproduces_reference(spec):
    desc = helpers.desc_from_spec(spec)
    shape = gcases.kernel_input_shape(spec["kernel"])
    X = gcases.sobol_points(shape)
    K_ref, d_ref = _K[spec["name"] + "__K"], _K[spec["name"] + "__diag"]
    K = helpers.eval_desc_numpy(desc, X[:32], X)
    scale = np.max(np.abs(K_ref))
    assert np.max(np.abs(K - K_ref)) <= 1e-13 * scale
    assert np.max(np.abs(desc.diag_value - d_ref)) <= 1e-13 * scale


def test_all_reference_cases_lower():
    assert len(LOWERABLE) == len(SPECS) == 117
    spec = next(s for s in SPECS if s["name"] == "matern2.5_3_dd_dd")
    desc = helpers.desc_from_spec(spec)
    assert desc.d == 3 and all(desc.dim_type[i] == 2 for i in range(3))  # LPGP_DIM_RADIAL


def test_isotropic_multid_matern_laplacian_is_rejected_like_the_reference():
    """diffops/_registry.py:270-280: no closed form (the reference falls to its jax path)."""
    from linpde_gp_b200.linfuncops import diffops
    from linpde_gp_b200.randprocs import covfuncs

    with pytest.raises(NotImplementedError):
        diffops.Laplacian((3,))(covfuncs.Matern((3,), nu=2.5), argnum=1)
    k = diffops.DirectionalDerivative(np.ones(3))(covfuncs.Matern((3,), nu=2.5), argnum=1)
    with pytest.raises(NotImplementedError):
        diffops.Laplacian((3,))(k, argnum=0)


def test_matern32_third_derivative_is_rejected_like_the_reference():
    from linpde_gp_b200._lowering import Factor1D, lower

    with pytest.raises(NotImplementedError):
        lower([Factor1D("matern", 1.0, nu=1.5)], {(1,): 1.0}, {(2,): 1.0})
