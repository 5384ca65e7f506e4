"""`from phi.flow import *` of the test double: the names phi/flow.py:18-40 exports that the facade and its tests use."""
from phiml import math  # noqa: F401
from phiml.math import Solve, SolveTape, NotConverged, Diverged, extrapolation, spatial, batch, channel, dual, tensor, wrap  # noqa: F401
from phiml.math.extrapolation import ZERO, ONE, PERIODIC, ZERO_GRADIENT, BOUNDARY, combine_sides  # noqa: F401

from . import field  # noqa: F401
from .field import Field, CenteredGrid, StaggeredGrid, Box  # noqa: F401
from .physics import fluid, advect, diffuse  # noqa: F401
