#!/usr/bin/env python
"""Kernel micro-benchmarks on one GPU: laplace and CG iteration bandwidth, TMA ring vs register-marching kernels.
Usage: python tools/microbench.py [sizes...]      (default 256 512)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phiflow_b200 import _ops as ops  # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    sizes = [a for a in sys.argv[1:] if a[0].isdigit()] or ['256', '512']
    dims2 = '--2d' in sys.argv
    for n in sizes:
        for ring in ((True,) if '--ring-only' in sys.argv else (True, False)):
            os.environ['PHICUDA_NO_RING'] = '0' if ring else '1'
            shape = tuple(int(v) for v in n.split('x'))
            if dims2:
                vbc = (('periodic', 'periodic'),) * 2
                batch = 64
                shape = shape * 2 if len(shape) == 1 else shape
                dom = ops.Domain(shape, (1.0, 1.0), batch, vbc=vbc)
                cells = batch * shape[0] * shape[1]
            else:
                vbc = (('periodic', 'periodic'),) * 3
                shape = shape * 3 if len(shape) == 1 else shape
                dom = ops.Domain(shape, (1.0, 1.0, 1.0), 1, vbc=vbc)
                cells = shape[0] * shape[1] * shape[2]
            x = torch.randn(dom._shape(dom.cext), device='cuda')
            y = torch.empty_like(x)
            ms = timed(lambda: ops.laplace(dom, vbc, x, out=y), 20)
            lap = 8.0 * cells / ms / 1e6
            rhs = torch.randn(dom._shape(dom.cext), device='cuda')
            p = dom.alloc_centered()
            iters = 40
            prm = ops.cg_params(vbc, rtol=1e-30, atol=0.0, max_iter=iters)

            def solve():
                p.zero_()
                ops.cg_poisson(dom, vbc, rhs, p, prm)
            ms_cg = timed(solve, 3)
            info = ops.read_results(dom)
            it = float(np.mean(info['iterations']))
            cg = cells * (30.0 * it + 32.0) / ms_cg / 1e6
            adv = ''
            if ring:
                v = [0.3 * torch.randn(dom._shape(dom.fext), device='cuda') for _ in range(dom.dim)]
                v2 = dom.alloc_faces()
                sbc = (('zg', 'zg'),) * dom.dim
                ms_a = timed(lambda: ops.advect_staggered(dom, vbc, v, vbc, v, 0.5, out=v2), 5)
                ms_c = timed(lambda: ops.advect_centered(dom, vbc, v, sbc, rhs, 0.5, out=y), 5)
                adv = f" | advect staggered {ms_a:.3f} ms = {20.0 * dom.dim * cells / ms_a / 1e6:.0f} GB/s, centred {ms_c:.3f} ms = {20.0 * cells / ms_c / 1e6:.0f} GB/s"
                del v, v2
            print(f"n={n} {'2d x64' if dims2 else '3d'} ring={int(ring)}: laplace {ms:.4f} ms = {lap:.0f} GB/s | "
                  f"CG {it:.0f} it in {ms_cg:.2f} ms = {ms_cg / max(it, 1) * 1e3:.1f} us/it = {cg:.0f} GB/s (30 B/cell/it){adv}", flush=True)
            del x, y, rhs, p, dom
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
