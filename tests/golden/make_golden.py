"""
Generates tests/golden/*.npz  --  run ONLY in the build container, where the reference is mounted:

    PYTHONPATH=/root/reference/PhiML python tests/golden/make_golden.py

It imports the reference's own arithmetic layer (vendored `phiml` 1.7.2, NumPy backend) and records its outputs
on seeded inputs, so that tests on the GPU box (where /root/reference does not exist) can compare the oracle
(oracle/oracle_np.py) and the CUDA path against what the reference really computes.

The phi.field glue (stagger / divergence / masked_laplace) cannot be imported (phi 3.4.0 needs phiml>=1.14), so the
three small functions below restate it *on phiml Tensors with phiml ops*, following the cited lines; everything
numerical (pad, shift, differences, tracing to a sparse matrix, CG) is executed by the reference code itself.
"""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings('ignore')
sys.path.insert(0, '/root/reference/PhiML')
from phiml import math  # noqa: E402
from phiml.math import extrapolation as E, spatial, channel, instance, tensor, wrap, Solve  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
NAMES = 'xyz'


def ext_from_spec(spec):
    """spec: tuple over axes of (lo, hi) with 'periodic' | 'zg' | float  ->  phiml Extrapolation."""
    def one(s):
        if s == 'periodic':
            return E.PERIODIC
        if s == 'zg':
            return E.ZERO_GRADIENT
        return E.ConstantExtrapolation(s)
    sides = {}
    for ax, (lo, hi) in enumerate(spec):
        sides[NAMES[ax]] = (one(lo), one(hi))
    return E.combine_sides(**sides)


def to_tensor(a):
    return tensor(a, spatial(','.join(NAMES[:a.ndim])))


def order(a):
    return ','.join(NAMES[:a.ndim]) if hasattr(a, 'ndim') else ','.join(NAMES[:a.shape.spatial_rank])


def npy(t, d):
    return t.numpy(','.join(NAMES[:d]))


# ---- phi.field glue restated on phiml tensors -----------------------------------------------------

def stagger_grad(p, dx, pext, vext):
    """field.spatial_gradient(at='face') = stagger(), phi/field/_field_math.py:229-236, 535-581."""
    comps = []
    for i, dim in enumerate(p.shape.spatial.names):
        lo, up = vext.valid_outer_faces(dim)
        if lo and up:
            wl, wu = {dim: (1, 0)}, {dim: (0, 1)}
        elif lo and not up:
            wl, wu = {dim: (1, -1)}, {dim: (0, 0)}
        elif not lo and up:
            wl, wu = {dim: (0, 0)}, {dim: (-1, 1)}
        else:
            wl, wu = {dim: (0, -1)}, {dim: (-1, 0)}
        lower = math.pad(p, wl, pext)
        upper = math.pad(p, wu, pext)
        comps.append((upper - lower) / dx[i])
    return comps


def staggered_divergence(comps, dx, vext):
    """field.divergence, staggered order 2, phi/field/_field_math.py:617-626 (+ bake_extrapolation :20-39)."""
    total = None
    for i, comp in enumerate(comps):
        dim = NAMES[i]
        lo, up = vext.valid_outer_faces(dim)
        padded = math.pad(comp, {dim: (0 if lo else 1, 0 if up else 1)}, vext)
        term = math.spatial_gradient(padded, dx[i], 'forward', None, dims=dim, stack_dim=None)
        total = term if total is None else total + term
    return total


def pressure_ext(vext_spec):
    def conv(s):
        return 'periodic' if s == 'periodic' else (0.0 if s == 'zg' else 'zg')
    return tuple((conv(lo), conv(hi)) for lo, hi in vext_spec)


def masked_laplace(p, dx, pext, vext0):
    """fluid.masked_laplace without obstacles, phi/physics/fluid.py:197-202 (vext0 = v boundary with constant
    offsets removed, :200)."""
    return staggered_divergence(stagger_grad(p, dx, pext, vext0), dx, vext0)


def remove_const(spec):
    return tuple(tuple(0.0 if not isinstance(s, str) else s for s in ax) for ax in spec)


# ---- fixtures ----------------------------------------------------------------------------------------

BC_SETS_2D = {
    'zero': ((0.0, 0.0), (0.0, 0.0)),
    'open': (('zg', 'zg'), ('zg', 'zg')),
    'periodic': (('periodic', 'periodic'), ('periodic', 'periodic')),
    'mixed': (('zg', 'zg'), (0.0, 'zg')),              # tests/commit/physics/test_fluid.py:50-53
    'per_x_wall_y': (('periodic', 'periodic'), (0.0, 0.0)),
    'one': ((1.0, 1.0), (1.0, 1.0)),
}
BC_SETS_3D = {
    'zero': ((0.0, 0.0),) * 3,
    'open': (('zg', 'zg'),) * 3,
    'periodic': (('periodic', 'periodic'),) * 3,
    'mixed': (('periodic', 'periodic'), (0.0, 'zg'), ('zg', 0.0)),
}


def spec_to_arr(spec):
    """Encode a BC spec as a float array: nan = periodic, inf = zero-gradient, else constant."""
    return np.array([[np.nan if s == 'periodic' else (np.inf if s == 'zg' else s) for s in ax] for ax in spec], np.float64)


def main():
    rng = np.random.default_rng(1234)
    out = {}

    # 1. pad / laplace / forward gradient
    for name, spec in list(BC_SETS_2D.items()) + [(k + '3', v) for k, v in BC_SETS_3D.items()]:
        d = len(spec)
        shape = (6, 5) if d == 2 else (5, 4, 6)
        dx = (0.5, 0.25) if d == 2 else (0.5, 0.25, 2.0)
        a = rng.standard_normal(shape).astype(np.float32)
        ext = ext_from_spec(spec)
        t = to_tensor(a)
        out[f'pad/{name}/in'] = a
        out[f'pad/{name}/bc'] = spec_to_arr(spec)
        widths = {NAMES[i]: (2, 1) for i in range(d)}
        out[f'pad/{name}/out_2_1'] = npy(math.pad(t, widths, ext), d)
        widths = {NAMES[0]: (1, -1), NAMES[1]: (-1, 2)}
        out[f'pad/{name}/out_neg'] = npy(math.pad(t, widths, ext), d)
        out[f'laplace/{name}/dx'] = np.array(dx)
        out[f'laplace/{name}/out'] = npy(math.laplace(t, wrap(dx, channel(vector=','.join(NAMES[:d]))), padding=ext), d)

    # 2. staggered divergence, gradient at faces, pressure matrix
    for name, spec in list(BC_SETS_2D.items()) + [(k + '3', v) for k, v in BC_SETS_3D.items()]:
        if name.startswith('one'):
            continue
        d = len(spec)
        res = (6, 5) if d == 2 else (4, 3, 5)
        dx = (0.5, 0.25) if d == 2 else (0.5, 0.25, 2.0)
        vext = ext_from_spec(spec)
        pspec = pressure_ext(spec)
        pext = ext_from_spec(pspec)
        comps = []
        for c in range(d):
            lo, up = vext.valid_outer_faces(NAMES[c])
            s = list(res); s[c] = res[c] - 1 + int(lo) + int(up)
            comps.append(rng.standard_normal(s).astype(np.float32))
        div = staggered_divergence([to_tensor(c) for c in comps], dx, vext)
        p = rng.standard_normal(res).astype(np.float32)
        grad = stagger_grad(to_tensor(p), dx, pext, vext)
        out[f'fluid/{name}/bc'] = spec_to_arr(spec)
        out[f'fluid/{name}/dx'] = np.array(dx)
        for c in range(d):
            out[f'fluid/{name}/v{c}'] = comps[c]
            out[f'fluid/{name}/grad{c}'] = npy(grad[c], d)
        out[f'fluid/{name}/div'] = npy(div, d)
        out[f'fluid/{name}/p'] = p
        lin = math.jit_compile_linear(masked_laplace, auxiliary_args='dx,pext,vext0')
        vext0 = ext_from_spec(remove_const(spec))
        try:
            mat = lin.sparse_matrix(to_tensor(p), dx=dx, pext=pext, vext0=vext0)
            dense = math.dense(mat)
            n_tot = int(np.prod(res))
            order_ = ','.join(NAMES[:d]) + ',' + ','.join('~' + n for n in NAMES[:d])
            out[f'fluid/{name}/matrix'] = dense.numpy(order_).reshape(n_tot, n_tot)
        except NotImplementedError as err:   # phiml 1.7.2 cannot trace PERIODIC mixed with other sides
            print(f"matrix tracing not supported by the vendored phiml for '{name}': {err}")
        out[f'fluid/{name}/lap_p'] = npy(masked_laplace(to_tensor(p), dx, pext, vext0), d)

    # 3. grid_sample / closest_grid_values
    for name, spec in list(BC_SETS_2D.items()) + [(k + '3', v) for k, v in BC_SETS_3D.items()]:
        d = len(spec)
        shape = (6, 5) if d == 2 else (5, 4, 6)
        a = rng.standard_normal(shape).astype(np.float32)
        n_pts = 400
        coords = (rng.uniform(-3.0, np.array(shape) + 2.0, size=(n_pts, d))).astype(np.float32)
        coords[:20] = np.round(coords[:20])                    # exact integers (floor edge cases)
        coords[20:30] = np.round(coords[20:30]) + 0.5
        ext = ext_from_spec(spec)
        ct = tensor(coords, instance('points'), channel(vector=','.join(NAMES[:d])))
        res = math.grid_sample(to_tensor(a), ct, ext)
        out[f'sample/{name}/bc'] = spec_to_arr(spec)
        out[f'sample/{name}/grid'] = a
        out[f'sample/{name}/coords'] = coords
        out[f'sample/{name}/out'] = res.numpy('points')
        closest = math.closest_grid_values(to_tensor(a), ct, ext)
        out[f'sample/{name}/closest'] = closest.numpy(['points'] + [f'closest_{NAMES[i]}' for i in range(d)])

    # 4. sample_subgrid (half-cell shifts)
    a = rng.standard_normal((7, 6)).astype(np.float32)
    out['subgrid/in'] = a
    out['subgrid/x_half'] = math.sample_subgrid(to_tensor(a), wrap((0.5, 0), channel('vector')), spatial(x=6, y=6)).numpy('x,y')
    out['subgrid/xy_half'] = math.sample_subgrid(to_tensor(a), wrap((0.5, 0.5), channel('vector')), spatial(x=6, y=5)).numpy('x,y')
    out['subgrid/y_half_off1'] = math.sample_subgrid(to_tensor(a), wrap((1, 0.5), channel('vector')), spatial(x=6, y=5)).numpy('x,y')

    # 5. CG solve through the reference's solve_linear (sparse matrix path, NumPy backend)
    for name, spec, res, dx in [('periodic', BC_SETS_2D['periodic'], (16, 12), (1.0, 1.0)),
                                ('zero', BC_SETS_2D['zero'], (16, 12), (0.5, 0.25)),
                                ('open', BC_SETS_2D['open'], (16, 12), (0.5, 0.25)),
                                ('mixed', BC_SETS_2D['mixed'], (16, 12), (1.0, 1.0)),
                                ('periodic3', BC_SETS_3D['periodic'], (12, 10, 8), (1.0, 1.0, 1.0)),
                                ('mixed3', BC_SETS_3D['mixed'], (8, 6, 7), (1.0, 0.5, 1.0))]:
        d = len(res)
        vext0 = ext_from_spec(remove_const(spec))
        pext = ext_from_spec(pressure_ext(spec))
        rhs = rng.standard_normal(res).astype(np.float32)
        flexible = any(s == 'zg' for ax in spec for s in ax)
        if not flexible:
            rhs -= rhs.mean()
        lin = math.jit_compile_linear(masked_laplace, auxiliary_args='dx,pext,vext0')
        for rtol, tag in [(1e-3, 'r3'), (1e-5, 'r5')]:
            np.random.seed(7)
            solve = Solve('CG', rtol, 1e-5, x0=to_tensor(np.zeros(res, np.float32)), max_iterations=1000,
                          rank_deficiency=None if flexible else 1)
            with math.SolveTape() as tape:
                x = math.solve_linear(lin, to_tensor(rhs), solve, dx=dx, pext=pext, vext0=vext0)
            info = tape[solve]
            out[f'cg/{name}/{tag}/x'] = npy(x, d)
            out[f'cg/{name}/{tag}/iterations'] = np.array(int(info.iterations))
            out[f'cg/{name}/{tag}/residual'] = npy(info.residual, d)
        out[f'cg/{name}/bc'] = spec_to_arr(spec)
        out[f'cg/{name}/dx'] = np.array(dx)
        out[f'cg/{name}/rhs'] = rhs

    # 6. static obstacles: hard_bcs = stagger(accessible, minimum), masked_laplace with where(active, div, p)
    def accessible_ext(spec):
        conv = lambda s_: 'periodic' if s_ == 'periodic' else (1.0 if s_ == 'zg' else 0.0)
        return tuple((conv(lo), conv(hi)) for lo, hi in spec)

    def hard_bcs(acc, aext, vext):
        comps = []
        for dim in acc.shape.spatial.names:
            lo, up = vext.valid_outer_faces(dim)
            if lo and up:
                wl, wu = {dim: (1, 0)}, {dim: (0, 1)}
            elif lo and not up:
                wl, wu = {dim: (1, -1)}, {dim: (0, 0)}
            elif not lo and up:
                wl, wu = {dim: (0, 0)}, {dim: (-1, 1)}
            else:
                wl, wu = {dim: (0, -1)}, {dim: (-1, 0)}
            comps.append(math.minimum(math.pad(acc, wl, aext), math.pad(acc, wu, aext)))
        return comps

    def masked_laplace_obs(p, dx, pext, vext0, hard, active):
        grad = [g_ * h_ for g_, h_ in zip(stagger_grad(p, dx, pext, vext0), hard)]
        return math.where(active, staggered_divergence(grad, dx, vext0), p)

    for name, spec in [('zero', BC_SETS_2D['zero']), ('open', BC_SETS_2D['open']), ('periodic', BC_SETS_2D['periodic']),
                       ('mixed', BC_SETS_2D['mixed']), ('zero3', BC_SETS_3D['zero']), ('periodic3', BC_SETS_3D['periodic'])]:
        d = len(spec)
        res = (7, 6) if d == 2 else (5, 4, 4)
        dx = (0.5, 0.25) if d == 2 else (0.5, 0.25, 2.0)
        acc = np.ones(res, np.float32)
        if d == 2:
            acc[2:4, 1:4] = 0
        else:
            acc[1:3, 1:3, 1:3] = 0
        vext = ext_from_spec(spec)
        vext0 = ext_from_spec(remove_const(spec))
        pext = ext_from_spec(pressure_ext(spec))
        aext = ext_from_spec(accessible_ext(spec))
        hard = hard_bcs(to_tensor(acc), aext, vext)
        pr = rng.standard_normal(res).astype(np.float32)
        out[f'obst/{name}/bc'] = spec_to_arr(spec)
        out[f'obst/{name}/dx'] = np.array(dx)
        out[f'obst/{name}/accessible'] = acc
        for c in range(d):
            out[f'obst/{name}/hard{c}'] = npy(hard[c], d)
        out[f'obst/{name}/p'] = pr
        active = to_tensor(acc)
        out[f'obst/{name}/lap_p'] = npy(masked_laplace_obs(to_tensor(pr), dx, pext, vext0, hard, active), d)
        try:
            lin = math.jit_compile_linear(masked_laplace_obs, auxiliary_args='dx,pext,vext0,hard,active')
            mat = lin.sparse_matrix(to_tensor(pr), dx=dx, pext=pext, vext0=vext0, hard=hard, active=active)
            n_tot = int(np.prod(res))
            order_ = ','.join(NAMES[:d]) + ',' + ','.join('~' + n_ for n_ in NAMES[:d])
            out[f'obst/{name}/matrix'] = math.dense(mat).numpy(order_).reshape(n_tot, n_tot)
        except Exception as err:
            print(f"masked matrix tracing failed for '{name}': {type(err).__name__}: {err}")

    np.savez_compressed(os.path.join(OUT, 'phiml_golden.npz'), **out)
    print(f"wrote {len(out)} arrays to {os.path.join(OUT, 'phiml_golden.npz')}")


if __name__ == '__main__':
    main()
