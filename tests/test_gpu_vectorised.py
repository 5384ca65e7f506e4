"""
The vectorised / fused kernels of the step (csrc/fused_kernels.cu: k_div_vec, k_gradsub_vec, k_advect_centered_vec,
k_advect_staggered_vec with the buoyancy / inflow epilogues) against the one-thread-per-sample kernels they replace
(PHICUDA_SCALAR_KERNELS=1), which tests/test_gpu_kernels.py pins against the oracle for every boundary type.  Both
families execute the same fp32 operations in the same order, so at full-size lines (nx = 512, several cells of displacement
per step, all boundary kinds, slab-like extents) the results must agree to the last bit - up to fused-multiply-add
contraction of the two epilogues, hence a tolerance of 2 ulp of the largest value.
The oracle comparisons of the same entry points (small grids, every boundary type) run in tests/test_gpu_kernels.py.
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from phiflow_b200 import _ops as ops

EPS = float(np.finfo(np.float32).eps)
CASES = {
    'periodic3': ((('periodic', 'periodic'),) * 3, (512, 24, 20)),
    'wall_open3': ((('periodic', 'periodic'), (0.0, 0.0), (0.0, 'zg')), (512, 17, 12)),
    'mixed3': ((('periodic', 'periodic'), (0.0, 'zg'), ('zg', 0.0)), (260, 19, 11)),
    'open3': ((('zg', 'zg'),) * 3, (133, 10, 9)),
    'inflow3': (((1.5, 'zg'), (0.0, 0.0), (-0.5, 0.25)), (96, 12, 10)),            # non-zero boundary constants
    'periodic2': ((('periodic', 'periodic'),) * 2, (512, 64)),
    'mixed2': ((('zg', 'zg'), (0.0, 'zg')), (150, 37)),
    'zero2': (((0.0, 0.0), (0.0, 0.0)), (256, 48)),
}


class scalar_kernels:
    def __enter__(self):
        os.environ['PHICUDA_SCALAR_KERNELS'] = '1'

    def __exit__(self, *exc):
        os.environ.pop('PHICUDA_SCALAR_KERNELS', None)


def _state(name, batch=2, speed=6.0):
    vbc, res = CASES[name]
    d = len(res)
    rng = np.random.default_rng(41)
    dx = tuple(100.0 / r for r in res)
    dom = ops.Domain(res, dx, batch, vbc=vbc)
    # displacements of up to ~`speed` cells per step (dt = 0.5): back-traces leave the tile, cross boundaries and wrap
    v = [(speed * dx[c] / 0.5 * 0.5 * rng.standard_normal((batch,) + s)).astype(np.float32)
         for c, s in enumerate(O.staggered_shapes(res, vbc))]
    s = rng.standard_normal((batch,) + res).astype(np.float32)
    return vbc, res, d, dx, dom, v, s


def _close(a, b, scale):
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=0, atol=2 * EPS * scale)


@pytest.mark.parametrize('name', sorted(CASES))
def test_divergence_and_grad_sub_match_scalar_kernels(name):
    vbc, res, d, dx, dom, v, s = _state(name)
    dv = dom.faces_from_numpy(v, vbc)
    dp = dom.centered_from_numpy(s)
    div = ops.divergence(dom, vbc, dv)
    g1 = [t.clone() for t in dv]
    ops.grad_sub(dom, vbc, g1, dp)
    with scalar_kernels():
        div0 = ops.divergence(dom, vbc, dv)
        g0 = [t.clone() for t in dv]
        ops.grad_sub(dom, vbc, g0, dp)
    assert torch.equal(div, div0)
    for c in range(d):
        assert torch.equal(g1[c], g0[c])


@pytest.mark.parametrize('name', sorted(CASES))
@pytest.mark.parametrize('sname', ['zg', 'zero', 'one'])
def test_advect_centered_matches_scalar_kernel(name, sname):
    vbc, res, d, dx, dom, v, s = _state(name)
    sbc = O.uniform_bc(d, {'zg': 'zg', 'zero': 0.0, 'one': 1.0}[sname])
    dv, ds = dom.faces_from_numpy(v, vbc), dom.centered_from_numpy(s)
    for dt in (0.5, -0.5):
        a = ops.advect_centered(dom, vbc, dv, sbc, ds, dt)
        with scalar_kernels():
            b = ops.advect_centered(dom, vbc, dv, sbc, ds, dt)
        assert torch.equal(a, b)
    m1 = ops.mac_cormack_centered(dom, vbc, dv, sbc, ds, 0.5)
    with scalar_kernels():
        m0 = ops.mac_cormack_centered(dom, vbc, dv, sbc, ds, 0.5)
    assert torch.equal(m1, m0)


@pytest.mark.parametrize('name', sorted(CASES))
def test_advect_staggered_matches_scalar_kernels(name):
    vbc, res, d, dx, dom, v, s = _state(name)
    dv = dom.faces_from_numpy(v, vbc)
    a = ops.advect_staggered(dom, vbc, dv, vbc, dv, 0.5)
    with scalar_kernels():
        b = ops.advect_staggered(dom, vbc, dv, vbc, dv, 0.5)
    for c in range(d):
        assert torch.equal(a[c], b[c]), f"component {c}: max diff {float((a[c] - b[c]).abs().max())}"


@pytest.mark.parametrize('name', sorted(CASES))
@pytest.mark.parametrize('mac', [False, True])
def test_fused_step_matches_round1_sequence(name, mac):
    """phicuda_plume_step_f32 (5 launches: inflow / buoyancy as advection epilogues, projected velocity written in place of
    the input) against the unfused sequence of the scalar kernels; 2 steps so the second one starts from fused output."""
    vbc, res, d, dx, dom, v, s = _state(name, batch=1, speed=3.0)
    sbc = O.uniform_bc(d, 'zg')
    buoy = (0.0, 0.1) if d == 2 else (0.02, 0.0, 0.1)
    inflow = np.abs(s) * np.float32(0.3)
    prm = ops.cg_params(vbc, rtol=1e-4, atol=1e-6, max_iter=40)       # a fixed, short solve: identical inputs -> identical iterates
    outs = []
    for scalar in (False, True):
        dv, ds = dom.faces_from_numpy(v, vbc), dom.centered_from_numpy(s)
        dpp, dinf = dom.alloc_centered(), dom.centered_from_numpy(inflow)
        for _ in range(2):
            if scalar:
                with scalar_kernels():
                    ops.plume_step(dom, vbc, sbc, dv, ds, dpp, dinf, 0.5, 0.2, buoy, prm, mac_cormack=mac)
            else:
                ops.plume_step(dom, vbc, sbc, dv, ds, dpp, dinf, 0.5, 0.2, buoy, prm, mac_cormack=mac)
        outs.append((dv, ds, dpp, ops.read_results(dom)['iterations'][0]))
    (v1, s1, p1, it1), (v0, s0, p0, it0) = outs
    assert it1 == it0
    _close(s1, s0, float(s0.abs().max()))
    for c in range(d):
        # the velocities went through 2 x 40 CG iterations whose input differs by the epilogue's fma rounding
        np.testing.assert_allclose(v1[c].cpu().numpy(), v0[c].cpu().numpy(), rtol=0, atol=1e-4 * float(v0[c].abs().max()))
    np.testing.assert_allclose(p1.cpu().numpy(), p0.cpu().numpy(), rtol=0, atol=1e-3 * float(p0.abs().max()) + 1e-6)


def test_plume_step_records_cg_events():
    vbc, res, d, dx, dom, v, s = _state('periodic3', batch=1, speed=1.0)
    sbc = O.uniform_bc(3, 'zg')
    dv, ds, dpp = dom.faces_from_numpy(v, vbc), dom.centered_from_numpy(s), dom.alloc_centered()
    prm = ops.cg_params(vbc, rtol=1e-3, atol=1e-6)
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    ops.plume_step(dom, vbc, sbc, dv, ds, dpp, None, 0.5, 0.0, (0.0, 0.0, 0.1), prm, cg_events=ev)
    t1.record()
    torch.cuda.synchronize()
    cg_ms, all_ms = ev[0].elapsed_time(ev[1]), t0.elapsed_time(t1)
    assert 0.0 < cg_ms <= all_ms
