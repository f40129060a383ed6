"""Host side of the whole path on a CPU test double of the C ABI (``tests/host_double.py``): the SAME test bodies the GPU
suite runs (``tests/test_gpu_*.py``, imported below), here exercising the Python orchestration -- conditioning (block
assembly order, noise, bordered factor growth, weights, posterior mean / variance / covariance), ``DeviceFactor``
(in-place extension, copy-on-append), the seam classes and linear operators -- against the real-reference goldens, the
oracle and numpy on a CPU-only machine.  The numbers come from numpy statements of the C-ABI entry points, so this is not
a parity test of the CUDA kernels (that is ``pytest -m gpu``)."""
import pytest

from tests import host_double

# pylint: disable=unused-import
from tests.test_gpu_api import (  # noqa: F401
    test_append_beyond_capacity_moves_only_the_lower_triangle,
    test_append_in_place_shares_storage_and_keeps_old_posterior_valid,
    test_batch_shapes_are_flattened_c_order,
    test_conditioning_matches_reference_golden,
    test_experiment_0000_poisson_dirichlet_1d_flow,
    test_experiment_0001_poisson_dirichlet_2d_flow,
    test_gridded_conditioning_uses_kronecker_assembly_and_matches_pairwise,
    test_heat_ibvp_analytic_solution_within_two_sigma,
    test_heat_medium_against_oracle,
    test_integral_observation_of_scalar_process,
    test_iterative_equals_batch_conditioning,
    test_kronecker_linop_on_tensor_product_grids_matches_reference,
    test_L0kL1_matrix_and_call,
    test_linop_solve_and_cholesky,
    test_medium_poisson2d_against_oracle,
    test_multi_output_conditioning_matches_reference_golden,
    test_multi_output_prior_kernel_evaluation,
    test_not_positive_definite_raises_linalgerror,
    test_one_shot_batches_equal_sequential_conditioning,
    test_polynomial_prior_mean_is_pushed_through_operators_in_closed_form,
    test_posterior_objects_hold_no_reference_cycles,
    test_reference_heat_test_verbatim_through_problem_classes,
    test_sum_kernel_prior_posterior_covariance,
    test_symmetric_matrix_and_linop_matmul,
)
from tests.test_gpu_crosscov import (  # noqa: F401
    test_covariance_of_two_functionals,
    test_crosscov_with_integral_functionals,
    test_dirac_functional_conditioning_equals_evaluation_functional,
    test_pv_crosscov_arithmetic_and_operator_on_free_argument,
    test_pv_crosscov_of_point_evaluations,
)
from tests.test_gpu_linops import (  # noqa: F401
    test_block_diagonal_and_triangular_structures,
    test_block_matrix_matmul_todense_transpose,
    test_cholesky_and_inverse_3x3,
    test_cholesky_and_inverse_nested_5x5,
    test_concatenated_linear_operator,
    test_condition_normal_on_observations_matches_numpy,
    test_covariance_blocks_reuse_cached_factor,
    test_kronecker_structured_solve_cholesky_det,
    test_not_positive_definite_block_raises,
    test_spd_block_quantities,
)
from tests.test_gpu_projections import (  # noqa: F401
    test_conditioning_on_projection_then_point_observations_matches_reference,
    test_covariance_of_two_projections_matches_reference,
    test_projection_crosscov_matches_reference,
    test_projection_observation_of_an_expquad_process_interpolates,
    test_projection_of_functions,
)
from tests.test_gpu_seam_goldens import (  # noqa: F401
    test_block_matrix_2x2_matches_the_reference,
    test_crosscov_and_covariance_match_the_reference,
    test_matrix_composed_functionals_match_the_reference,
)


@pytest.fixture(autouse=True)
def _c_abi_double(monkeypatch):
    host_double.install_c_abi(monkeypatch)
