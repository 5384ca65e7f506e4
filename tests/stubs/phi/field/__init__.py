"""Field / CenteredGrid / StaggeredGrid of the test double (attribute names: phi/field/_field.py:51-474, phi/field/_grid.py:21-187)."""
from phiml import math
from phiml.math import Shape, spatial, extrapolation

from .. import STOCK_CALLS


class Box:
    """phi.geom.Box stand-in: lower / upper per named dimension, compared by value (phi/geom/_box.py)."""

    def __init__(self, **sizes):
        self.lower = {d: (float(v[0]) if isinstance(v, tuple) else 0.0) for d, v in sizes.items()}
        self.upper = {d: (float(v[1]) if isinstance(v, tuple) else float(v)) for d, v in sizes.items()}

    def __eq__(self, other):
        return isinstance(other, Box) and self.lower == other.lower and self.upper == other.upper

    def __hash__(self):
        return hash(tuple(sorted(self.upper.items())))


class _UniformGrid:
    is_uniform = True


class Field:
    def __init__(self, values, boundary, bounds: Box, resolution: Shape, staggered: bool):
        self._values, self._boundary, self._bounds, self._resolution, self._staggered = values, boundary, bounds, resolution, staggered

    values = property(lambda self: self._values)
    extrapolation = property(lambda self: self._boundary)
    boundary = property(lambda self: self._boundary)
    bounds = property(lambda self: self._bounds)
    resolution = property(lambda self: self._resolution)
    geometry = property(lambda self: _UniformGrid())
    is_grid = property(lambda self: True)
    is_staggered = property(lambda self: self._staggered)
    is_centered = property(lambda self: not self._staggered)
    shape = property(lambda self: self._values.shape)

    @property
    def dx(self):
        names = self._resolution.names
        return math.vec(**{d: (self._bounds.upper[d] - self._bounds.lower[d]) / self._resolution.get_size(d) for d in names})

    def with_values(self, values, **_):
        return Field(values, self._boundary, self._bounds, self._resolution, self._staggered)

    def with_boundary(self, boundary):
        return Field(self._values, boundary, self._bounds, self._resolution, self._staggered)

    with_extrapolation = with_boundary


def CenteredGrid(values=0., boundary=0., bounds=None, resolution=None, **resolution_):
    resolution = resolution if isinstance(resolution, Shape) else spatial(**resolution_)
    boundary = boundary if isinstance(boundary, extrapolation.Extrapolation) else extrapolation.ConstantExtrapolation(boundary)
    if not isinstance(values, math.Tensor):
        values = math.zeros(resolution) + values
    return Field(values, boundary, bounds, resolution, False)


def StaggeredGrid(values, boundary=0., bounds=None, resolution=None, **resolution_):
    resolution = resolution if isinstance(resolution, Shape) else spatial(**resolution_)
    boundary = boundary if isinstance(boundary, extrapolation.Extrapolation) else extrapolation.ConstantExtrapolation(boundary)
    return Field(values, boundary, bounds, resolution, True)


def laplace(u, *args, **kwargs):
    STOCK_CALLS.append(('field.laplace', (u,) + args))
    return 'stock laplace'


def divergence(field, *args, **kwargs):
    STOCK_CALLS.append(('field.divergence', (field,) + args))
    return 'stock divergence'
