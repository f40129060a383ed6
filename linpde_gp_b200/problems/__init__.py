"""``linpde_gp.problems``: problem definitions used by the reference's experiments and tests (set-up code only)."""
from . import pde
