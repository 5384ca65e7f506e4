#!/usr/bin/env python
"""Diagnostic: vectorised vs scalar advection of a centred field with a constant boundary, and which one agrees with the oracle."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_np as O  # noqa: E402
from phiflow_b200 import _ops as ops  # noqa: E402

vbc, res = (('periodic', 'periodic'),) * 3, (64, 12, 10)
rng = np.random.default_rng(41)
dx = tuple(100.0 / r for r in res)
for speed in (0.3, 6.0):
    for sname, side in (('zg', 'zg'), ('one', 1.0)):
        sbc = O.uniform_bc(3, side)
        dom = ops.Domain(res, dx, 1, vbc=vbc)
        v = [(speed * dx[c] / 0.5 * 0.5 * rng.standard_normal(s)).astype(np.float32) for c, s in enumerate(O.staggered_shapes(res, vbc))]
        s = rng.standard_normal(res).astype(np.float32)
        dv, ds = dom.faces_from_numpy(v, vbc), dom.centered_from_numpy(s)
        a = dom.centered_to_numpy(ops.advect_centered(dom, vbc, dv, sbc, ds, 0.5))
        os.environ['PHICUDA_SCALAR_KERNELS'] = '1'
        b = dom.centered_to_numpy(ops.advect_centered(dom, vbc, dv, sbc, ds, 0.5))
        os.environ.pop('PHICUDA_SCALAR_KERNELS')
        ref = O.semi_lagrangian_centered(s, sbc, v, vbc, (0.0,) * 3, (100.0,) * 3, 0.5)
        bad = np.argwhere(a != b)
        print(f"speed {speed} sbc {sname}: vec!=scalar at {len(bad)} cells; max|vec-ref|={np.abs(a - ref).max():.3e} max|scalar-ref|={np.abs(b - ref).max():.3e}")
        for idx in bad[:6]:
            i = tuple(idx)
            print("   cell", i, "vec", a[i], "scalar", b[i], "oracle", ref[i])
