"""Host logic of the seam classes on a CPU test double of the device calls (``tests/host_double.py``): the same test bodies
the GPU suite runs against the real-reference goldens (``tests/test_gpu_seam_goldens.py``), here exercising the Python
side only -- atoms, shapes, layouts, ``reverse``, accumulation, the matrix of a matrix-composed functional -- so that the
CPU suite covers it.  (The numbers come from the numpy statement of the kernel descriptor, not from the CUDA kernels: this is
not a parity test of the device path.)"""
import pytest

from tests import host_double
from tests import test_gpu_seam_goldens as seam_tests
from tests.golden import cases as gcases


@pytest.mark.parametrize("kname", sorted(gcases.SEAM_KERNELS))
@pytest.mark.parametrize("oname", sorted(gcases.SEAM_OPS))
def test_crosscov_host_logic_matches_the_reference(monkeypatch, kname, oname):
    host_double.install(monkeypatch)
    seam_tests.test_crosscov_and_covariance_match_the_reference(kname, oname)


@pytest.mark.parametrize("kname", sorted(gcases.SEAM_KERNELS))
@pytest.mark.parametrize("oname", sorted(gcases.SEAM_OPS))
def test_matrix_composed_functional_host_logic_matches_the_reference(monkeypatch, kname, oname):
    host_double.install(monkeypatch)
    seam_tests.test_matrix_composed_functionals_match_the_reference(kname, oname)
