#!/usr/bin/env python
"""Per-iteration time of the distributed CG solve on z-slabs:  torchrun --nproc-per-node N tools/dist_cg_bench.py [planes_per_rank]
   (512 x 512 x planes_per_rank per rank, periodic; 40 iterations, 3 repetitions; honours the PHICUDA_RING_* knobs)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phiflow_b200 import _ops as ops  # noqa: E402
from phiflow_b200.dist import Slab  # noqa: E402


def main():
    dist.init_process_group('nccl')
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    world, rank = dist.get_world_size(), dist.get_rank()
    planes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = 512
    vbc = (('periodic', 'periodic'),) * 3
    slab = Slab((n, n, planes * world), (1.0, 1.0, 1.0), vbc, halo=2, device=dev)
    d = slab.dom
    g = torch.Generator(device='cpu').manual_seed(rank)
    rhs = d.alloc_centered()
    rhs[:, slab.halo:slab.halo + planes] = torch.randn((1, planes, n, n), generator=g).to(dev)
    p = d.alloc_centered()
    iters = 40
    prm = ops.cg_params(vbc, rtol=1e-30, atol=0.0, max_iter=iters)
    times = []
    for rep in range(4):
        p.zero_()
        slab.exchange([p], 1)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        slab.cg_poisson(rhs, p, prm)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep > 0:
            times.append(float(t.item()))
    info = ops.last_launch_info()
    if rank == 0:
        ms = min(times)
        cells = float(n) * n * planes
        print(f"dist CG world={world} slab 512x512x{planes}: {ms / iters * 1e3:.1f} us/iteration per rank "
              f"({cells * (30.0 * iters + 32.0) / ms / 1e6:.0f} GB/s per rank) split={info['split']} ZC={info['ZC']} nzc={info['nzc']} "
              f"units={info['total_units']} grid={info['grid_ctas']} SPLIT_ENV={os.environ.get('PHICUDA_RING_SPLIT', 'auto')}", flush=True)
    slab.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
