#!/usr/bin/env python
"""Launches the hot-path kernels a few times for ncu:  tools/prof_kernels.py N [noring] [scalar] [stencils]
   default: laplace + a short CG solve;  stencils: the vectorised divergence / grad_sub / advection kernels on a smooth velocity
   field that moves samples a few cells per step (the regime of the developed plume)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phiflow_b200 import _ops as ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
if 'noring' in sys.argv:
    os.environ['PHICUDA_NO_RING'] = '1'
if 'scalar' in sys.argv:
    os.environ['PHICUDA_SCALAR_KERNELS'] = '1'
vbc = (('periodic', 'periodic'),) * 3
sbc = (('zg', 'zg'),) * 3
dom = ops.Domain((n, n, n), (100.0 / n,) * 3, 1, vbc=vbc)
x = torch.randn(dom._shape(dom.cext), device='cuda')
y = torch.empty_like(x)
if 'stencils' in sys.argv:
    ax = (torch.arange(n, device='cuda', dtype=torch.float32) + 0.5) / n * 6.2831853
    z, yy, xx = torch.meshgrid(ax, ax, ax, indexing='ij')
    w0 = 3.0 * (100.0 / n) / 0.5                                  # ~3 cells per step
    v = [(w0 * 0.3 * torch.sin(yy) * torch.cos(z) + 0.01 * torch.randn_like(xx)).unsqueeze(0).contiguous(),
         (w0 * 0.2 * torch.cos(xx) * torch.sin(2 * z) + 0.01 * torch.randn_like(xx)).unsqueeze(0).contiguous(),
         (w0 * (0.8 + 0.15 * torch.sin(xx) * torch.cos(yy)) + 0.01 * torch.randn_like(xx)).unsqueeze(0).contiguous()]
    del xx, yy, z
    v2 = dom.alloc_faces()
    for _ in range(3):
        ops.divergence(dom, vbc, v, out=y)
        ops.grad_sub(dom, vbc, v2, x)
        ops.advect_centered(dom, vbc, v, sbc, x, 0.5, out=y)
        ops.advect_staggered(dom, vbc, v, vbc, v, 0.5, out=v2)
    torch.cuda.synchronize()
    sys.exit(0)
for _ in range(3):
    ops.laplace(dom, vbc, x, out=y)
rhs = torch.randn(dom._shape(dom.cext), device='cuda')
p = dom.alloc_centered()
prm = ops.cg_params(vbc, rtol=1e-30, atol=0.0, max_iter=int(os.environ.get('CG_ITERS', '6')))
for _ in range(2):
    p.zero_()
    ops.cg_poisson(dom, vbc, rhs, p, prm)
torch.cuda.synchronize()
print(ops.read_results(dom))
