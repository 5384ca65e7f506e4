"""
On-disk format of grid fields: the single-file `.npz` layout of `phi.field.write` / `phi.field.read`
(phi/field/_field_io.py:45-69 writer, :104-127 reader), restated on plain NumPy arrays so that trajectories written here
load in stock PhiFlow and vice versa.  Host-side only, no device code.

Keys of the archive (np.savez_compressed, _field_io.py:57-66):
    dim_names, dim_types, dim_item_names (object array), field_type ('CenteredGrid' | 'StaggeredGrid'),
    lower, upper, bounds_item_names, extrapolation (pickled dict, Extrapolation.to_dict), data
`data` of a staggered grid is `Field.staggered_tensor()` (phi/field/_field.py:586-604): every component padded with its own
extrapolation to (n+1) points per spatial axis and stacked along a trailing channel dim `vector`.

Pinned against the vendored PhiML (tests/golden/make_golden.py): the extrapolation dictionaries, the dim type strings
(`Shape.types` of PhiML 1.7.2; a newer PhiML may spell them differently - the reader accepts full names and first
letters) and the padding arithmetic of the staggered tensor (`math.pad`).  `phi` itself cannot be imported in the build
container, so a byte-level comparison with a file written by stock PhiFlow is not part of the tests.
"""
from typing import Dict, List, Sequence, Tuple

import numpy as np

PERIODIC, ZERO_GRADIENT = 'periodic', 'zg'
_TYPE_NAMES = {'b': 'batch', 's': 'spatial', 'c': 'channel', 'i': 'instance', 'd': 'dual'}


# ---- extrapolation <-> dict (PhiML/phiml/math/extrapolation.py:265-266, 437-438, 1260-1264) -------------------------------
def _side_to_dict(side) -> dict:
    if side == PERIODIC:
        return {'type': 'periodic'}
    if side == ZERO_GRADIENT:
        return {'type': 'zero-gradient'}
    if isinstance(side, (tuple, list)):
        return {'type': 'constant', 'value': np.asarray(side, np.float32)}
    v = float(side)
    return {'type': 'constant', 'value': np.asarray(int(v)) if v.is_integer() else np.asarray(v, np.float32)}


def _side_from_dict(d: dict):
    t = d['type']
    if t == 'periodic':
        return PERIODIC
    if t in ('zero-gradient', 'boundary'):
        return ZERO_GRADIENT
    if t == 'constant':
        v = np.asarray(d['value'])
        return float(v) if v.ndim == 0 else tuple(float(x) for x in v.ravel())
    raise NotImplementedError(f"extrapolation type '{t}' is outside the phiflow_b200 fast path")


def extrapolation_to_dict(default, sides: Dict[Tuple[str, bool], object], axes: Sequence[str]) -> dict:
    """`sides` maps (axis, is_upper) to 'periodic' | 'zg' | constant; `default` applies to sides that are not listed.
    Like extrapolation.combine_sides (extrapolation.py:1209-1240), equal sides collapse: a uniform boundary is written as
    that single extrapolation, an axis with equal lower/upper side under its plain name, otherwise as 'x-' / 'x+'."""
    per_axis = {a: (sides.get((a, False), default), sides.get((a, True), default)) for a in axes}
    flat = [s for lo_hi in per_axis.values() for s in lo_hi]
    if all(s == flat[0] for s in flat):
        return _side_to_dict(flat[0])
    dims = {}
    for a, (lo, hi) in per_axis.items():
        if lo == hi:
            dims[a] = _side_to_dict(lo)
        else:
            dims[a + '-'] = _side_to_dict(lo)
            dims[a + '+'] = _side_to_dict(hi)
    return {'type': 'mixed_v2', 'dims': dims}


def extrapolation_from_dict(d: dict):
    """Returns (default, sides) as taken by extrapolation_to_dict."""
    if d['type'] != 'mixed_v2':
        return _side_from_dict(d), {}
    sides = {}
    for key, sub in d['dims'].items():
        s = _side_from_dict(sub)
        if key[-1] in '+-':
            sides[(key[:-1], key[-1] == '+')] = s
        else:
            sides[(key, False)] = sides[(key, True)] = s
    return None, sides


# ---- staggered tensor (phi/field/_field.py:586-604, phi/field/_grid.py:179-187) ---------------------------------------------
def valid_outer_faces(lo_side, hi_side) -> Tuple[bool, bool]:
    """extrapolation.py:57-62 per side kind: which boundary faces a staggered component stores."""
    return (lo_side in (ZERO_GRADIENT, PERIODIC)), (hi_side == ZERO_GRADIENT)


def _pad_axis(a: np.ndarray, axis: int, lo: int, hi: int, lo_side, hi_side, component: int) -> np.ndarray:
    """math.pad of one axis by at most one cell per side with the side's extrapolation (extrapolation.py:291-325, 462-487,
    671-675): constant value, edge replicate, wrap."""
    def ghost(side, upper):
        n = a.shape[axis]
        if side == PERIODIC:
            src = 0 if upper else n - 1
        elif side == ZERO_GRADIENT:
            src = n - 1 if upper else 0
        else:
            v = side[component] if isinstance(side, (tuple, list)) else side
            shape = list(a.shape)
            shape[axis] = 1
            return np.full(shape, v, a.dtype)
        return np.take(a, [src], axis=axis)
    parts = ([ghost(lo_side, False)] if lo else []) + [a] + ([ghost(hi_side, True)] if hi else [])
    return np.concatenate(parts, axis=axis) if len(parts) > 1 else a


def staggered_tensor(comps: List[np.ndarray], sides_of, d: int) -> np.ndarray:
    """comps[c]: faces of component c as the reference stores them, spatial axes last in (x, y[, z]) order (any leading
    batch axes).  sides_of(axis_index) -> (lower side, upper side).  Returns (..., n_x+1, n_y+1[, n_z+1], d)."""
    padded = []
    for c, comp in enumerate(comps):
        a = np.asarray(comp)
        lead = a.ndim - d
        for ax in range(d):
            lo_side, hi_side = sides_of(ax)
            if ax == c:
                lo_valid, up_valid = valid_outer_faces(lo_side, hi_side)
                a = _pad_axis(a, lead + ax, int(not lo_valid), int(not up_valid), lo_side, hi_side, c)
            else:
                a = _pad_axis(a, lead + ax, 0, 1, lo_side, hi_side, c)
        padded.append(a)
    return np.stack(padded, axis=-1)


def unstack_staggered_tensor(data: np.ndarray, sides_of, d: int) -> List[np.ndarray]:
    """Inverse of staggered_tensor: slices the stored faces of every component out of the uniform array."""
    comps = []
    lead = data.ndim - 1 - d
    for c in range(d):
        sl = [slice(None)] * (data.ndim - 1)
        for ax in range(d):
            if ax == c:
                lo_valid, up_valid = valid_outer_faces(*sides_of(ax))
                sl[lead + ax] = slice(int(not lo_valid), -int(not up_valid) or None)
            else:
                sl[lead + ax] = slice(0, -1)
        comps.append(np.ascontiguousarray(data[..., c][tuple(sl)]))
    return comps


# ---- archive ------------------------------------------------------------------------------------------------------------------
def write_single_field(file: str, field_type: str, data: np.ndarray, dim_names: Sequence[str], dim_types: Sequence[str],
                       dim_item_names: Sequence, lower: Sequence[float], upper: Sequence[float],
                       bounds_item_names: Sequence[str], extrapolation: dict):
    """_field_io.py:45-69.  `file` gets the '.npz' suffix appended by NumPy when it has none, as in the reference."""
    assert field_type in ('CenteredGrid', 'StaggeredGrid'), field_type
    assert data.ndim == len(dim_names) == len(dim_types) == len(dim_item_names)
    items = np.empty(len(dim_item_names), dtype=object)
    for i, it in enumerate(dim_item_names):
        items[i] = None if it is None else tuple(it)
    np.savez_compressed(file,
                        dim_names=tuple(dim_names),
                        dim_types=tuple(dim_types),
                        dim_item_names=items,
                        field_type=field_type,
                        lower=np.asarray(lower, np.float32),
                        upper=np.asarray(upper, np.float32),
                        bounds_item_names=tuple(bounds_item_names),
                        extrapolation=extrapolation,
                        data=data)


def read_single_field(file: str) -> dict:
    """_field_io.py:104-127: returns the archive as a dict of plain Python / NumPy values."""
    with np.load(file if file.endswith('.npz') else file + '.npz', allow_pickle=True) as stored:
        ftype = str(stored['field_type'])
        if ftype not in ('CenteredGrid', 'StaggeredGrid'):
            raise NotImplementedError(f"{ftype} not implemented")
        data = stored['data']
        names = tuple(str(n) for n in stored['dim_names'])
        types = tuple(_TYPE_NAMES.get(str(t), str(t)) for t in stored['dim_types'])
        items = tuple(stored['dim_item_names']) if 'dim_item_names' in stored else (None,) * data.ndim
        items = tuple(None if it is None else tuple(str(s) for s in it) for it in items)
        spatial = tuple(n for n, t in zip(names, types) if t == 'spatial')
        bnames = stored['bounds_item_names'] if 'bounds_item_names' in stored else None
        bnames = spatial if bnames is None or bnames.shape == () else tuple(str(n) for n in bnames)
        lower, upper = np.asarray(stored['lower'], np.float64), np.asarray(stored['upper'], np.float64)
        if lower.ndim == 0:
            lower = np.full(len(bnames), float(lower))
        return {'field_type': ftype, 'data': data, 'dim_names': names, 'dim_types': types, 'dim_item_names': items,
                'lower': dict(zip(bnames, lower.tolist())), 'upper': dict(zip(bnames, upper.tolist())),
                'extrapolation': stored['extrapolation'][()]}
