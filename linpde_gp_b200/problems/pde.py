"""Linear PDE boundary value problems (``linpde_gp.problems.pde``): the objects the reference's experiments and tests
use to build ``L``, ``X`` and ``Y`` for ``condition_on_observations`` (src/linpde_gp/problems/pde/_linear_pde.py:6-63,
_bvp.py:13-171, _poisson.py:14-134, _heat.py:16-144).  Host-side set-up code; nothing here touches the device."""
from __future__ import annotations

import dataclasses
from collections.abc import Sequence
from typing import Optional

import numpy as np

from .. import domains, functions, linfuncops
from ..linfuncops import diffops


class LinearPDE:
    """``diffop[u] = rhs`` on ``domain`` (``rhs`` defaults to zero)."""

    def __init__(self, domain, diffop, rhs=None):
        self._domain = domains.asdomain(domain)
        if tuple(diffop.input_domain_shape) != self._domain.shape:
            raise ValueError(
                "The shape of the domain of the differential operator's input function is not equal to the shape of "
                f"the given domain object ({diffop.input_domain_shape} != {self._domain.shape}).")
        self._diffop = diffop
        if rhs is None:
            rhs = functions.Zero(self._domain.shape, output_shape=diffop.output_codomain_shape)
        if rhs.input_shape != self._domain.shape:
            raise ValueError(f"right-hand side input shape {rhs.input_shape} != domain shape {self._domain.shape}")
        if rhs.output_shape != tuple(diffop.output_codomain_shape):
            raise ValueError(f"right-hand side output shape {rhs.output_shape} != {diffop.output_codomain_shape}")
        self._rhs = rhs

    domain = property(lambda self: self._domain)
    diffop = property(lambda self: self._diffop)
    rhs = property(lambda self: self._rhs)


class BoundaryCondition:
    """``operator[u] = values`` on ``boundary``."""

    def __init__(self, boundary, operator, values):
        self._boundary = domains.asdomain(boundary)
        if tuple(operator.input_domain_shape) != self._boundary.shape:
            raise ValueError(f"operator domain shape {operator.input_domain_shape} != boundary shape {self._boundary.shape}")
        self._operator = operator
        if not isinstance(values, functions.Function):
            values = functions.Constant(operator.output_domain_shape, values)
        if values.input_shape != tuple(operator.output_domain_shape) or values.output_shape != tuple(operator.output_codomain_shape):
            raise ValueError("shapes of the boundary values do not match the boundary operator")
        self._values = values

    boundary = property(lambda self: self._boundary)
    operator = property(lambda self: self._operator)
    values = property(lambda self: self._values)


class DirichletBoundaryCondition(BoundaryCondition):
    def __init__(self, boundary, values):
        boundary = domains.asdomain(boundary)
        codomain = values.output_shape if isinstance(values, functions.Function) else np.shape(values)
        super().__init__(boundary, linfuncops.Identity(boundary.shape, codomain), values)


def get_1d_dirichlet_boundary_observations(dirichlet_bcs):
    """``(X_bc, Y_bc)`` of the two end points of an interval problem (``_bvp.py:75-88``)."""
    if len(dirichlet_bcs) != 2 or not all(isinstance(bc.boundary, domains.Point) for bc in dirichlet_bcs):
        raise ValueError("expected the two point boundary conditions of an interval problem")
    X_bc = np.asarray([float(bc.boundary) for bc in dirichlet_bcs])
    Y_bc = np.asarray([bc.values(x) for bc, x in zip(dirichlet_bcs, X_bc)])
    return X_bc, Y_bc


@dataclasses.dataclass(frozen=True)
class BoundaryValueProblem:
    pde: LinearPDE
    boundary_conditions: Sequence
    solution: Optional[functions.Function] = None

    @property
    def domain(self):
        return self.pde.domain

    def __post_init__(self):
        for bc in self.boundary_conditions:
            if bc.boundary.shape != self.domain.shape:
                raise ValueError("The shape of the boundary must be equal to the shape of the domain")
        if self.solution is not None:
            if self.solution.input_shape != self.domain.shape:
                raise ValueError("The input shape of the solution function should be equal to the shape of the domain.")
            if self.solution.output_shape != tuple(self.pde.diffop.input_codomain_shape):
                raise ValueError("The output shape of the solution function should be equal to the output shape of the "
                                 "differential operator's input function.")


class InitialBoundaryValueProblem(BoundaryValueProblem):
    """Time-dependent problem on ``[t0, T] x spatial_domain`` with an initial condition at ``t0``."""

    def __init__(self, pde, initial_condition, boundary_conditions, solution=None):
        dom = pde.domain
        if not isinstance(dom, domains.CartesianProduct) or len(dom) != 2 or not isinstance(dom[0], domains.Interval):
            raise ValueError("the domain must be the product of a time interval and a spatial domain")
        if initial_condition.boundary != dom[1]:
            raise ValueError("the initial condition must live on the spatial domain")
        object.__setattr__(self, "_initial_condition", initial_condition)
        super().__init__(pde=pde, boundary_conditions=boundary_conditions, solution=solution)

    temporal_domain = property(lambda self: self.domain[0])
    t0 = property(lambda self: self.temporal_domain[0])
    T = property(lambda self: self.temporal_domain[1])
    spatial_domain = property(lambda self: self.domain[1])
    initial_condition = property(lambda self: self._initial_condition)

    @property
    def initial_domain(self):
        return domains.CartesianProduct(domains.Point(self.t0), self.spatial_domain)


# -- Poisson ------------------------------------------------------------------------------------------------------------
class PoissonEquation(LinearPDE):
    """``-alpha Laplace u = rhs`` (``_poisson.py:14-36``)."""

    def __init__(self, domain, rhs=None, alpha: float = 1.0):
        domain = domains.asdomain(domain)
        super().__init__(domain, -alpha * diffops.Laplacian(domain.shape), rhs)
        self._alpha = alpha

    alpha = property(lambda self: self._alpha)


class Solution_PoissonEquation_DirichletProblem_1D_RHSConstant(functions.Function):  # pylint: disable=invalid-name
    """Closed-form solution of ``-alpha u'' = rhs`` on ``[l, r]`` with ``u(l), u(r)`` given (``_poisson.py:98-134``)."""

    def __init__(self, domain, rhs, boundary_values, alpha=1.0):
        super().__init__((), ())
        domain = domains.asdomain(domain)
        if not isinstance(domain, domains.Interval):
            raise TypeError("We only support Interval domains.")
        self._l, self._r = (float(b) for b in domain)
        u_l, u_r = (float(v) for v in np.asarray(boundary_values))
        self._coeffs = (u_l, (u_r - u_l) / (self._r - self._l), 0.5 * float(rhs) / -float(alpha))

    def _evaluate(self, x):
        a = self._coeffs
        return (a[2] * (x - self._r) + a[1]) * (x - self._l) + a[0]


class PoissonEquationDirichletProblem(BoundaryValueProblem):
    def __init__(self, domain, *, rhs=None, alpha: float = 1.0, boundary_values=None, solution=None):
        pde = PoissonEquation(domain, rhs=rhs, alpha=alpha)
        if boundary_values is None:
            boundary_values = functions.Zero(pde.domain.shape, pde.diffop.input_codomain_shape)
        if pde.domain.shape == ():
            if not isinstance(pde.domain, domains.Interval):
                raise TypeError("In the scalar case, we only support Interval domains.")
            if isinstance(boundary_values, functions.Function):
                a, b = pde.domain
                boundary_values = (boundary_values(a), boundary_values(b))
            boundary_values = np.asarray(boundary_values, dtype=np.double)
            if solution is None and isinstance(pde.rhs, functions.Constant):
                solution = Solution_PoissonEquation_DirichletProblem_1D_RHSConstant(
                    pde.domain, rhs=pde.rhs.value, boundary_values=boundary_values, alpha=pde.alpha)
        if isinstance(boundary_values, functions.Function):
            bcs = tuple(DirichletBoundaryCondition(part, boundary_values) for part in pde.domain.boundary)
        else:
            bcs = tuple(DirichletBoundaryCondition(part, value)
                        for part, value in zip(pde.domain.boundary, np.asarray(boundary_values, dtype=np.double)))
        super().__init__(pde=pde, boundary_conditions=bcs, solution=solution)


# -- heat ---------------------------------------------------------------------------------------------------------------
class HeatEquation(LinearPDE):
    """``du/dt - alpha Laplace_x u = rhs`` on a space-time domain (``_heat.py:16-31``)."""

    def __init__(self, domain, rhs=None, alpha=1.0):
        domain = domains.asdomain(domain)
        self._alpha = float(alpha)
        super().__init__(domain, diffops.HeatOperator(domain.shape, alpha=self._alpha), rhs)

    alpha = property(lambda self: self._alpha)


class Solution_HeatEquation_DirichletProblem_1D_InitialTruncatedSineSeries_BoundaryZero(functions.Function):  # pylint: disable=invalid-name
    """``u(t, x) = sum_n c_n sin(w_n (x - l)) exp(-alpha w_n^2 (t - t0))`` (``_heat.py:96-131``)."""

    def __init__(self, t0, spatial_domain, initial_values, alpha):
        if not isinstance(spatial_domain, domains.Interval) or spatial_domain != initial_values.domain:
            raise ValueError("the initial values must be a sine series on the spatial interval")
        super().__init__((2,), ())
        self._t0, self._l, self._alpha = float(t0), float(spatial_domain[0]), float(alpha)
        self._initial_values = initial_values

    def _evaluate(self, txs):
        ts, xs = txs[..., :1], txs[..., 1:]
        w = self._initial_values.half_angular_frequencies
        return np.sum(self._initial_values.coefficients * np.sin(w * (xs - self._l))
                      * np.exp(self._alpha * w**2 * (self._t0 - ts)), axis=-1)


class HeatEquationDirichletProblem(InitialBoundaryValueProblem):
    """Heat equation with zero Dirichlet boundary values and initial values on the spatial domain (``_heat.py:34-93``)."""

    def __init__(self, t0, spatial_domain, T=float("inf"), rhs=None, alpha=1.0, initial_values=None, solution=None):
        spatial_domain = domains.asdomain(spatial_domain)
        domain = domains.CartesianProduct(domains.Interval(t0, T), spatial_domain)
        pde = HeatEquation(domain, rhs=rhs, alpha=alpha)
        if initial_values is None:
            initial_values = functions.Zero(spatial_domain.shape, ())
        if initial_values.input_shape != spatial_domain.shape or initial_values.output_shape != ():
            raise ValueError("the initial values must be a scalar function on the spatial domain")
        initial_condition = DirichletBoundaryCondition(domain[1], initial_values)
        bcs = tuple(DirichletBoundaryCondition(domains.CartesianProduct(domain[0], part), np.zeros(()))
                    for part in domain[1].boundary)
        if solution is None:
            if isinstance(initial_values, functions.Zero):
                solution = functions.Zero(domain.shape, ())
            elif (isinstance(domain[1], domains.Interval) and isinstance(initial_values, functions.TruncatedSineSeries)
                  and initial_values.domain == domain[1]):
                solution = Solution_HeatEquation_DirichletProblem_1D_InitialTruncatedSineSeries_BoundaryZero(
                    t0=t0, spatial_domain=domain[1], initial_values=initial_values, alpha=alpha)
        super().__init__(pde=pde, initial_condition=initial_condition, boundary_conditions=bcs, solution=solution)
