"""phi.physics.fluid of the test double: the stock make_incompressible only records the call."""
from phiml.math import extrapolation

from .. import STOCK_CALLS


def _get_obstacles_for(obstacles, space):                    # phi/physics/fluid.py:85-91
    obstacles = [obstacles] if not isinstance(obstacles, (tuple, list)) else obstacles
    return list(obstacles)


def _pressure_extrapolation(vext):                            # phi/physics/fluid.py:264-274
    if vext == extrapolation.PERIODIC:
        return extrapolation.PERIODIC
    elif vext == extrapolation.BOUNDARY:
        return extrapolation.ZERO
    elif isinstance(vext, extrapolation.ConstantExtrapolation):
        return extrapolation.BOUNDARY
    return extrapolation.map(_pressure_extrapolation, vext)


def make_incompressible(velocity, obstacles=(), solve=None, active=None, order=2, correct_skew=False, wide_stencil=None):
    STOCK_CALLS.append(('fluid.make_incompressible', (velocity, obstacles, solve, active, order)))
    return 'stock make_incompressible'
