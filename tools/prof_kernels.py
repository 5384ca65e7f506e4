#!/usr/bin/env python
"""Launches the flagship kernels a few times for ncu (tools/prof_kernels.py N [noring])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phiflow_b200 import _ops as ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
if 'noring' in sys.argv:
    os.environ['PHICUDA_NO_RING'] = '1'
vbc = (('periodic', 'periodic'),) * 3
dom = ops.Domain((n, n, n), (1.0, 1.0, 1.0), 1, vbc=vbc)
x = torch.randn(dom._shape(dom.cext), device='cuda')
y = torch.empty_like(x)
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
