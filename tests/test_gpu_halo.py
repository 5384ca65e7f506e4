"""
PHI_BC_HALO (z-slab sides that border another rank) on ONE GPU: a periodic grid is run once with PERIODIC z sides and once as
a "slab" whose halo planes hold the wrapped neighbour planes, exactly what the halo exchange of phiflow_b200.dist delivers.
Every kernel that reads across a slab face must give bit-identical owned planes in both runs (the multi-GPU parity proper is
tests/tools/dist_check.py / tests/test_gpu_dist.py, which need >= 2 GPUs).
"""
import numpy as np
import pytest
import torch

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from phiflow_b200 import _ops as ops

H = 4
XY = {'periodic': ('periodic', 'periodic'), 'wall': (0.0, 0.0), 'open': ('zg', 'zg')}


def _setup(xy, res=(96, 20, 12), speed=2.5, batch=1):
    vbc_p = (XY[xy], XY[xy], ('periodic', 'periodic'))
    vbc_h = (XY[xy], XY[xy], ('halo', 'halo'))
    dx = tuple(100.0 / r for r in res)
    dom_p = ops.Domain(res, dx, batch, vbc=vbc_p)
    dom_h = ops.Domain(res, dx, batch, vbc=vbc_h, halo=H)
    assert dom_h.fext[:2] == dom_p.fext[:2] and dom_h.fext[2] == dom_p.fext[2] + 2 * H
    g = torch.Generator(device='cpu').manual_seed(7)
    return vbc_p, vbc_h, dx, dom_p, dom_h, g


def _slab(t_p):
    """periodic array (b, nz, y, x) -> slab array with H wrapped halo planes below and above"""
    return torch.cat([t_p[:, -H:], t_p, t_p[:, :H]], dim=1).contiguous()


def _owned(t_h):
    return t_h[:, H:t_h.shape[1] - H]


def _rand_faces(dom, g, scale):
    return [(torch.randn(dom._shape(dom.fext), generator=g) * scale).cuda() for _ in range(3)]


@pytest.mark.parametrize('xy', sorted(XY))
def test_stencils_read_slab_halo_planes(xy):
    vbc_p, vbc_h, dx, dom_p, dom_h, g = _setup(xy)
    v_p = _rand_faces(dom_p, g, 1.0)
    v_h = [_slab(t) for t in v_p]
    p_p = torch.randn(dom_p._shape(dom_p.cext), generator=g).cuda()
    p_h = _slab(p_p)
    assert torch.equal(_owned(ops.divergence(dom_h, vbc_h, v_h)), ops.divergence(dom_p, vbc_p, v_p))
    g_p, g_h = [t.clone() for t in v_p], [t.clone() for t in v_h]
    ops.grad_sub(dom_p, vbc_p, g_p, p_p)
    ops.grad_sub(dom_h, vbc_h, g_h, p_h)
    for c in range(3):
        assert torch.equal(_owned(g_h[c]), g_p[c]), f"grad_sub component {c}"
    sbc_p, sbc_h = (('zg', 'zg'),) * 2 + (('periodic', 'periodic'),), (('zg', 'zg'),) * 2 + (('halo', 'halo'),)
    b_p, b_h = [t.clone() for t in v_p], [t.clone() for t in v_h]
    ops.add_buoyancy(dom_p, vbc_p, sbc_p, p_p, (0.05, 0.0, 0.1), 0.5, b_p)
    ops.add_buoyancy(dom_h, vbc_h, sbc_h, p_h, (0.05, 0.0, 0.1), 0.5, b_h)
    for c in range(3):
        assert torch.equal(_owned(b_h[c]), b_p[c]), f"buoyancy component {c}"
    lbc_p, lbc_h = (XY[xy] if xy != 'wall' else ('zg', 'zg'),) * 2 + (('periodic', 'periodic'),), (XY[xy] if xy != 'wall' else ('zg', 'zg'),) * 2 + (('halo', 'halo'),)
    assert torch.equal(_owned(ops.laplace(dom_h, lbc_h, p_h)), ops.laplace(dom_p, lbc_p, p_p))
    assert torch.equal(ops.max_abs_velocity(dom_h, vbc_h, v_h), ops.max_abs_velocity(dom_p, vbc_p, v_p))


@pytest.mark.parametrize('xy', sorted(XY))
@pytest.mark.parametrize('speed', [0.3, 2.7])
def test_advection_reads_slab_halo_planes(xy, speed):
    """Back-traces of up to H - 1 cells leave the owned planes through the slab faces (h = ceil(disp) + 1 <= H planes)."""
    vbc_p, vbc_h, dx, dom_p, dom_h, g = _setup(xy)
    v_p = [t.clamp_(-1.0, 1.0) * (speed * dx[c] / 0.5) for c, t in enumerate(_rand_faces(dom_p, g, 0.5))]
    v_h = [_slab(t) for t in v_p]
    s_p = torch.randn(dom_p._shape(dom_p.cext), generator=g).cuda()
    s_h = _slab(s_p)
    sbc_p, sbc_h = (('zg', 'zg'),) * 2 + (('periodic', 'periodic'),), (('zg', 'zg'),) * 2 + (('halo', 'halo'),)
    a_p = ops.advect_centered(dom_p, vbc_p, v_p, sbc_p, s_p, 0.5)
    a_h = ops.advect_centered(dom_h, vbc_h, v_h, sbc_h, s_h, 0.5)
    assert torch.equal(_owned(a_h), a_p)
    w_p = ops.advect_staggered(dom_p, vbc_p, v_p, vbc_p, v_p, 0.5)
    w_h = ops.advect_staggered(dom_h, vbc_h, v_h, vbc_h, v_h, 0.5)
    for c in range(3):
        assert torch.equal(_owned(w_h[c]), w_p[c]), f"component {c}: {float((_owned(w_h[c]) - w_p[c]).abs().max())}"
