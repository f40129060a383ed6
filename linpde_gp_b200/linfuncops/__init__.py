"""``linpde_gp.linfuncops`` API (symbolic linear function operators) for the hot path."""
from . import diffops
from ._linfuncop import (
    Identity,
    LinearFunctionOperator,
    ScaledLinearFunctionOperator,
    SumLinearFunctionOperator,
)
from .diffops import LinearDifferentialOperator
