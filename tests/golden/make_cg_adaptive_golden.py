"""
Generates tests/golden/phiml_cg_adaptive.npz  --  run ONLY in the build container (reference mounted):

    python tests/golden/make_cg_adaptive_golden.py

`Solve('CG-adaptive')` (PhiML/phiml/backend/_linalg.py:93-128) of the vendored PhiML on the pressure operator traced
from the restated `masked_laplace` (see make_golden.py): the solver the reference's `'auto'` policy and several notebooks
select instead of plain CG.  Fixtures for oracle.cg_adaptive.
Only systems with a Dirichlet side are recorded: on the rank-deficient periodic system the reference's CG-adaptive with its
random rank-1 `matrix_offset` needs 661 iterations at rtol 1e-3 and does not reach 1e-5 within 1000 (fp32, measured here),
so there is nothing stable to pin.
"""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings('ignore')
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as G  # noqa: E402  (imports phiml from /root/reference/PhiML)
from make_golden import math, Solve  # noqa: E402


def main():
    out = {}
    rng = np.random.default_rng(11)
    for name, spec, res, dx in [('open', G.BC_SETS_2D['open'], (16, 12), (0.5, 0.25)),
                                ('mixed', G.BC_SETS_2D['mixed'], (16, 12), (1.0, 1.0)),
                                ('mixed3', G.BC_SETS_3D['mixed'], (8, 6, 7), (1.0, 0.5, 1.0))]:
        d = len(res)
        vext0 = G.ext_from_spec(G.remove_const(spec))
        pext = G.ext_from_spec(G.pressure_ext(spec))
        rhs = rng.standard_normal(res).astype(np.float32)
        flexible = any(s == 'zg' for ax in spec for s in ax)
        if not flexible:
            rhs -= rhs.mean()
        lin = math.jit_compile_linear(G.masked_laplace, auxiliary_args='dx,pext,vext0')
        for rtol, tag in [(1e-3, 'r3'), (1e-5, 'r5')]:
            np.random.seed(7)
            solve = Solve('CG-adaptive', rtol, 1e-5, x0=G.to_tensor(np.zeros(res, np.float32)), max_iterations=1000,
                          rank_deficiency=None if flexible else 1)
            with math.SolveTape() as tape:
                x = math.solve_linear(lin, G.to_tensor(rhs), solve, dx=dx, pext=pext, vext0=vext0)
            info = tape[solve]
            out[f'{name}/{tag}/x'] = G.npy(x, d)
            out[f'{name}/{tag}/iterations'] = np.array(int(info.iterations))
            out[f'{name}/{tag}/function_evaluations'] = np.array(int(info.function_evaluations))
            out[f'{name}/{tag}/residual'] = G.npy(info.residual, d)
        out[f'{name}/bc'] = G.spec_to_arr(spec)
        out[f'{name}/dx'] = np.array(dx)
        out[f'{name}/rhs'] = rhs
    np.savez_compressed(os.path.join(G.OUT, 'phiml_cg_adaptive.npz'), **out)
    print(f"wrote {len(out)} arrays")


if __name__ == '__main__':
    main()
