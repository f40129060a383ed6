"""Projections onto finite-dimensional function spaces (``linpde_gp.linfunctls.projections``)."""
from . import l2  # noqa: F401
