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


def _lowerable(spec):
    base = spec["kernel"]["base"]
    return not (base["kind"] == "matern" and int(np.prod(base.get("input_shape", ()) or (1,))) > 1)


LOWERABLE = [s for s in SPECS if _lowerable(s)]


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


def test_isotropic_multid_matern_is_rejected():
    spec = next(s for s in SPECS if s["name"] == "matern2.5_3_plain")
    with pytest.raises(NotImplementedError):
        helpers.desc_from_spec(spec)


def test_matern32_third_derivative_is_rejected_like_the_reference():
    from linpde_gp_b200._lowering import Factor1D, lower

    with pytest.raises(NotImplementedError):
        lower([Factor1D("matern", 1.0, nu=1.5)], {(1,): 1.0}, {(2,): 1.0})
