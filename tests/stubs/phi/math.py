"""phi.math re-exports phiml.math (phi/math/__init__.py:16)."""
from phiml.math import *  # noqa: F401,F403
from phiml.math import extrapolation, wrap, spatial, max  # noqa: F401
