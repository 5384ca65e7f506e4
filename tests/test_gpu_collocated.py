"""
CenteredGrid (collocated) velocities on the GPU: wide-stencil operator and make_incompressible with CG-adaptive
(phi/physics/fluid.py:154-155, 197-202) against the oracle restatement, which is pinned against the
vendored PhiML in tests/golden/phiml_collocated.npz (gradient, divergence, traced matrix).
Reference tests mirrored: tests/commit/physics/test_fluid.py:17-28, 34-36 (CenteredGrid ZERO / BOUNDARY, divergence < 5e-5).
"""
import numpy as np
import pytest
import torch

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from phiflow_b200 import _ops as ops

BCS = {
    'zero': ((0.0, 0.0), (0.0, 0.0)),
    'open': (('zg', 'zg'), ('zg', 'zg')),
    'periodic': (('periodic', 'periodic'), ('periodic', 'periodic')),
    'mixed': (('zg', 'zg'), (0.0, 'zg')),
    'zero3': ((0.0, 0.0),) * 3,
    'mixed3': (('periodic', 'periodic'), (0.0, 'zg'), ('zg', 0.0)),
}


def _centered_list(dom, arrays):
    return [dom.centered_from_numpy(a) for a in arrays]


@pytest.mark.parametrize('name', sorted(BCS))
def test_wide_laplace_matches_oracle(name):
    vbc = BCS[name]
    d = len(vbc)
    res = (17, 12) if d == 2 else (11, 9, 7)
    dx = (0.5, 0.25) if d == 2 else (0.5, 0.25, 2.0)
    rng = np.random.default_rng(51)
    batch = 2
    dom = ops.Domain(res, dx, batch)
    p = rng.standard_normal((batch,) + res).astype(np.float32)
    out = dom.centered_to_numpy(ops.wide_laplace(dom, vbc, dom.centered_from_numpy(p)), squeeze=False)
    ref = np.stack([O.wide_laplace(p[b], dx, O.pressure_bc(vbc), vbc) for b in range(batch)])
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-5 * np.abs(ref).max())


@pytest.mark.parametrize('name', sorted(BCS))
def test_make_incompressible_centered_matches_oracle(name):
    vbc = BCS[name]
    d = len(vbc)
    res = (16, 20) if d == 2 else (12, 10, 8)
    dx = tuple(100.0 / r for r in res)
    rng = np.random.default_rng(52)
    batch = 2
    # smooth input, as in the reference test (buoyancy of a smoke blob): white noise has components outside the range of the
    # singular, non-symmetric wide-stencil operator and makes CG-adaptive diverge - in the oracle as well (measured), so its
    # iterates are not a meaningful reference there
    pts = O.points_of((0.0,) * d, (100.0,) * d, res)
    v = []
    for c in range(d):
        comp = np.zeros((batch,) + res, np.float32)
        for b in range(batch):
            centre = rng.uniform(25.0, 75.0, d).astype(np.float32)
            comp[b] = (0.1 if c == d - 1 else 0.03) * np.exp(-np.sum((pts - centre) ** 2, -1) / np.float32(2 * 12.0 ** 2))
        v.append(comp)
    dom = ops.Domain(res, dx, batch)
    dv = _centered_list(dom, v)
    rank_def = not O.is_flexible(vbc)
    # the same offset on both sides (the reference draws it from an unseeded random probe; its value only matters at the 1e-7 level)
    A = O.wide_poisson_matrix(res, dx, vbc)
    c = O.estimate_matrix_offset(A, int(np.prod(res)), np.random.default_rng(0)) if rank_def else None
    dv, dp = ops.make_incompressible_centered(dom, vbc, dv, None, rtol=1e-5, atol=1e-5, max_iter=1000, matrix_offset=c)
    info = ops.read_results(dom)
    got_v = [dom.centered_to_numpy(t, squeeze=False) for t in dv]
    for b in range(batch):
        vb = [a[b] for a in v]
        div = O.divergence_centered(vb, dx, O.component_bcs(vbc, d))
        if rank_def:
            div = div - np.mean(div, dtype=np.float32)
        ref = O.cg_adaptive(A, div, np.zeros(res, np.float32), 1e-5, 1e-5, 1000, c)
        assert not ref['diverged'] and info['diverged'][b] == 0
        got_b = [got_v[comp][b] for comp in range(d)]
        if ref['converged'] and ref['iterations'] <= 200:
            # short solves: the fp32 iterates of both sides stay together
            assert info['converged'][b] == 1
            assert abs(int(info['iterations'][b]) - ref['iterations']) <= max(3, ref['iterations'] // 6), (info['iterations'][b], ref['iterations'])
            grad = O.gradient_centered(ref['x'].reshape(res), dx, O.pressure_bc(vbc))
            for comp in range(d):
                np.testing.assert_allclose(got_b[comp], vb[comp] - grad[comp], rtol=0, atol=2e-4 * max(np.abs(vb[comp]).max(), 1e-3))
        else:
            # closed boxes: the wide operator has checkerboard null modes besides the constant one, the rank-1
            # offset removes only the latter, and CG-adaptive runs into max_iterations in the oracle as well (measured: 1000 iterations,
            # divergence reduced by 50x-600x depending on the input).  Iterates decorrelate in fp32 over hundreds of iterations on a
            # non-symmetric operator, so only what the projection is for is checked: finite, not flagged, divergence clearly reduced.
            # (The reference's own closed-box scenario is test_reference_test_fluid_centered below: < 5e-5.)
            after = O.divergence_centered(got_b, dx, O.component_bcs(vbc, d))
            after = after - np.mean(after, dtype=np.float32)
            assert np.isfinite(after).all() and np.abs(after).max() < 0.3 * np.abs(div).max(), (np.abs(after).max(), np.abs(div).max())


@pytest.mark.parametrize('name', ['zero', 'open'])
def test_reference_test_fluid_centered(name):
    """tests/commit/physics/test_fluid.py:17-28, 34-36: two rounds of buoyancy + make_incompressible on a 16 x 20 CenteredGrid;
    the centred divergence ends below 5e-5."""
    vbc = BCS[name]
    res, dx = (16, 20), (100 / 16, 100 / 20)
    pts = O.points_of((0.0, 0.0), (100.0, 100.0), res)
    smoke = (np.sum((pts - np.array([40.0, 10.0], np.float32)) ** 2, -1) <= 25.0).astype(np.float32)
    dom = ops.Domain(res, dx, 1)
    dv = [dom.alloc_centered(), dom.alloc_centered()]
    dsmoke = dom.centered_from_numpy(smoke)
    for _ in range(2):
        dv[1] += dsmoke * 0.1
        dv, dp = ops.make_incompressible_centered(dom, vbc, dv, None)
        info = ops.read_results(dom)
        assert info['diverged'][0] == 0 and (info['converged'][0] == 1 or not O.is_flexible(vbc))
    v = [dom.centered_to_numpy(t) for t in dv]
    div = O.divergence_centered(v, dx, O.component_bcs(vbc, 2))
    assert np.abs(div).max() < 5e-5, np.abs(div).max()
