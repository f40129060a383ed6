"""``linpde_gp.randprocs`` API of the hot path."""
from . import covfuncs, crosscov
from ._conditional import ConditionalGaussianProcess
from ._gaussian_process import GaussianProcess
