"""``linpde_gp.linfuncops`` API (symbolic linear function operators) for the hot path."""
from . import diffops
from ._linfuncop import (
    CompositeLinearFunctionOperator,
    Identity,
    LinearFunctionOperator,
    ScaledLinearFunctionOperator,
    SelectOutput,
    SumLinearFunctionOperator,
)
from .diffops import LinearDifferentialOperator
