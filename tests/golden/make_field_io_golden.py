"""
Generates tests/golden/phiml_field_io.npz  --  run ONLY in the build container, where the reference is mounted:

    python tests/golden/make_field_io_golden.py

Pins the pieces of the `.npz` field format (phi/field/_field_io.py:45-69) that live in the vendored PhiML 1.7.2:
`Extrapolation.to_dict()` of every boundary the fast path supports, the dim-type strings of a Shape, and the uniform
`staggered_tensor` of a staggered grid.  `Field.staggered_tensor` itself (phi/field/_field.py:586-604) cannot be imported
(phi 3.4.0 needs phiml >= 1.14), so its five lines are restated below ON phiml tensors: the widths come from phiml's
`valid_outer_faces`, the padding is executed by phiml's `math.pad`.
"""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings('ignore')
sys.path.insert(0, '/root/reference/PhiML')
from phiml import math  # noqa: E402
from phiml.math import extrapolation as E, spatial, channel, batch, tensor  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
NAMES = 'xyz'

SPECS = {
    'zero2': (((0.0, 0.0), (0.0, 0.0)), (5, 4)),
    'open2': ((('zg', 'zg'), ('zg', 'zg')), (5, 4)),
    'periodic2': ((('periodic', 'periodic'), ('periodic', 'periodic')), (6, 4)),
    'mixed2': ((('periodic', 'periodic'), (0.0, 'zg')), (6, 4)),
    'sides2': (((0.0, 'zg'), ('zg', 1.0)), (5, 3)),
    'zero3': (((0.0, 0.0),) * 3, (4, 3, 5)),
    'mixed3': ((('periodic', 'periodic'), (0.0, 0.0), (0.0, 'zg')), (4, 3, 5)),
}


def one(s):
    if s == 'periodic':
        return E.PERIODIC
    if s == 'zg':
        return E.ZERO_GRADIENT
    return E.ConstantExtrapolation(int(s) if float(s).is_integer() else s)      # ZERO / ONE are built from Python ints


def ext_from_spec(spec):
    # the way user code writes it: one extrapolation per axis, a (lower, upper) pair only where the sides differ
    return E.combine_sides(**{NAMES[ax]: (one(lo) if lo == hi else (one(lo), one(hi))) for ax, (lo, hi) in enumerate(spec)})


def main():
    out = {}
    rng = np.random.default_rng(5)
    for name, (spec, res) in SPECS.items():
        ext = ext_from_spec(spec)
        d = len(res)
        dims = NAMES[:d]
        out[f'ext/{name}'] = np.asarray(ext.to_dict(), dtype=object)
        out[f'spec/{name}'] = np.asarray(spec, dtype=object)
        out[f'res/{name}'] = np.asarray(res)
        padded = []
        for c, dim in enumerate(dims):
            lo_valid, up_valid = ext.valid_outer_faces(dim)
            shape = tuple(n + (int(lo_valid) + int(up_valid) - 1 if a == c else 0) for a, n in enumerate(res))
            comp = rng.standard_normal(shape).astype(np.float32)
            out[f'comp/{name}/{c}'] = comp
            widths = {k: (0, 1) for k in dims}
            widths[dim] = (int(not lo_valid), int(not up_valid))
            t = tensor(comp, spatial(*dims))
            padded.append(math.pad(t, widths, ext))                     # _field.py:599-602
        st = math.stack(padded, channel(vector=','.join(dims)))                   # _field.py:603
        out[f'staggered_tensor/{name}'] = st.numpy(tuple(dims) + ('vector',))
    s = batch(batch=2) & spatial(x=4, y=3) & channel(vector='x,y')
    out['shape/names'] = np.asarray(s.names)
    out['shape/types'] = np.asarray(s.types)
    out['ext/one'] = np.asarray(E.ONE.to_dict(), dtype=object)
    out['ext/half'] = np.asarray(E.ConstantExtrapolation(0.5).to_dict(), dtype=object)
    out['ext/vec'] = np.asarray(E.ConstantExtrapolation(math.wrap([1., 0.], channel(vector='x,y'))).to_dict(), dtype=object)
    np.savez_compressed(os.path.join(OUT, 'phiml_field_io.npz'), **out)
    print(f"wrote {len(out)} entries")


if __name__ == '__main__':
    main()
