"""bench.py arm for N > 1 GPUs: the 512^3 plume on z-slabs (strong scaling), one process per GPU under torchrun."""
import json
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _ops as ops
from .dist import Slab, SlabPlume
from ._clocks import ClockSampler

DT, INFLOW_RATE, BUOYANCY = 0.5, 0.2, (0.0, 0.0, 0.1)
RTOL, ATOL, MAX_ITER = 1e-3, 1e-5, 1000


def run_e2e(sim, slab, args, dev, snap):
    """Per rank: the slab state (v, s, p) lives in pinned HOST arrays in the reference's (x, y, z) order; every step uploads
    it, transposes to the device layout, steps (halo exchange + kernels), transposes back and downloads it."""
    H, nz = slab.halo, slab.nz
    names = ('vx', 'vy', 'vz', 's', 'p')
    cur = dict(zip(names, snap))             # owned planes at the start of the timed steps
    host = {k: torch.zeros(tuple(reversed(cur[k].shape)), dtype=torch.float32).pin_memory() for k in names}
    for k in names:
        host[k].copy_(cur[k].permute(2, 1, 0))
    torch.cuda.synchronize()
    nbytes = sum(h.numel() * 4 for h in host.values()) * slab.world
    steps = args.steps                       # the same steps as the device-resident leg, replayed from its start state

    def one():
        H = sim.slab.halo
        cur = {'vx': sim.v[0], 'vy': sim.v[1], 'vz': sim.v[2], 's': sim.s, 'p': sim.p}
        for k in names:
            cur[k][0, H:H + nz].copy_(host[k].to(dev, non_blocking=True).permute(2, 1, 0))
        sim.step()
        cur = {'vx': sim.v[0], 'vy': sim.v[1], 'vz': sim.v[2], 's': sim.s, 'p': sim.p}
        for k in names:
            host[k].copy_(cur[k][0, H:H + nz].permute(2, 1, 0), non_blocking=True)
    one()                                    # untimed warm-up of the copy path, then back to the start state
    torch.cuda.synchronize()
    for k in names:
        host[k].copy_(cur[k].permute(2, 1, 0))
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    its = []
    t0.record()
    for _ in range(steps):
        one()
        its.append(sim.slab.result_tensor()[:1].clone())
    t1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([t0.elapsed_time(t1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return {"value": 1e3 / float(ms.item()), "unit": "steps/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": steps,
            "cg_iterations_per_step": float(torch.cat(its).float().mean().item()),
            "note": "per rank: slab state (v, s, p) in pinned host arrays in reference (x,y,z) order; upload + transpose + step + transpose + download"}


def run(args, metric):
    """Never exits without a JSON line: any exception on any rank is reported as {"error": ..., "traceback": ...}."""
    rank = int(os.environ.get('RANK', '0'))
    try:
        _run(args, metric)
    except BaseException as err:  # noqa: BLE001  (the driver only sees stdout + the exit code)
        import sys
        import traceback
        tb = traceback.format_exc()
        sys.stderr.write(f"[rank {rank}] {tb}\n")
        print(json.dumps({"metric": metric, "n_gpus": int(os.environ.get('WORLD_SIZE', '1')), "rank": rank,
                          "error": f"{type(err).__name__}: {err}", "traceback": tb[-2000:]}), flush=True)
        raise SystemExit(1)


def _run(args, metric):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    n = args.size
    dx = tuple(100.0 / n for _ in range(3))
    vbc = (('periodic', 'periodic'),) * 3
    sbc = (('zg', 'zg'),) * 3
    slab = Slab((n, n, n), dx, vbc, halo=args.halo, device=dev)
    prm = ops.cg_params(vbc, rtol=RTOL, atol=ATOL, max_iter=MAX_ITER)
    sim = SlabPlume(slab, sbc, DT, INFLOW_RATE, BUOYANCY, prm)
    H, nz, z0 = slab.halo, slab.nz, slab.z0
    # identical initial condition to the single-GPU run: the same seeded noise, sliced to the slab
    g = torch.Generator().manual_seed(0)
    for c in range(3):
        host = torch.randn((1, n, n, n), generator=g, dtype=torch.float32)[:, z0:z0 + nz].mul_(0.01).contiguous().pin_memory()
        sim.v[c][:, H:H + nz].copy_(host, non_blocking=True)
        del host
    ax = (torch.arange(n, device=dev, dtype=torch.float32) + 0.5) * dx[0]
    zz, yy, xx = torch.meshgrid(ax[z0:z0 + nz], ax, ax, indexing='ij')
    distc = torch.sqrt(torch.clamp((xx - 50.0) ** 2 + (yy - 50.0) ** 2 + (zz - 9.5) ** 2, min=1e-6))
    cell_r = float(np.sqrt(3 * (dx[0] * 0.5) ** 2))
    sim.inflow[:, H:H + nz] = torch.clamp(0.5 - (distc - 5.0) / cell_r, 0, 1)
    del xx, yy, zz, distc
    sim.project()
    sim.p.zero_()

    res_dev = None
    iters_host = torch.zeros((args.warmup + args.steps, 6), dtype=torch.int32).pin_memory()
    for i in range(args.warmup):
        sim.step()
        iters_host[i].copy_(sim.slab.result_tensor()[:6], non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    cg_events = []
    # the e2e leg replays the first timed steps from this state, so both legs do the same CG iterations
    snap = [t[0, sim.slab.halo:sim.slab.halo + nz].clone() for t in (sim.v[0], sim.v[1], sim.v[2], sim.s, sim.p)]
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(args.steps):
        sim.step(cg_events)
        iters_host[args.warmup + i].copy_(sim.slab.result_tensor()[:6], non_blocking=True)
    end.record()
    torch.cuda.synchronize()
    clocks = sampler.summary() if sampler else None
    dist.barrier()
    torch.cuda.synchronize()
    ms_local = torch.tensor([start.elapsed_time(end)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms_local, op=dist.ReduceOp.MAX)
    ms = float(ms_local.item()) / args.steps
    cg_ms = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in cg_events]))], device=dev, dtype=torch.float64)
    dist.all_reduce(cg_ms, op=dist.ReduceOp.MAX)
    iters = iters_host[args.warmup:, 0].numpy().astype(np.int64)
    slab = sim.slab                      # a regrow (halo wider than allocated) replaces the slab object
    e2e = run_e2e(sim, slab, args, dev, snap)
    del snap
    if rank == 0:
        peak = 6572.9
        try:
            peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'])
        except Exception:
            pass
        cells = float(n) ** 3
        cg_gbs = float(np.sum(cells * (30.0 * iters + 32.0)) / (float(cg_ms.item()) * len(iters) * 1e-3) / 1e9)
        line = {"metric": metric, "value": 1e3 / ms, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"3-D smoke plume {n}^3 fp32 periodic, CG rtol=1e-3 warm start, z-slabs of {slab.nz} planes on {world} GPUs "
                                       f"(BASELINE configs[3])",
                           "cg_iterations_per_step": float(np.mean(iters)), "cg_ms_per_step": float(cg_ms.item()),
                           "halo_planes_allocated": slab.halo, "halo_planes_used": sim.halo_used,
                           "max_displacement_cells": sim.max_displacement, "halo_regrown": sim.regrown,
                           "halo_rule": "h = ceil(max|v_z| dt/dz) + 1 from an all-reduced device max before every step",
                           "comm": "CG: in-kernel NVLink peer stores (halo planes + mailbox all-reduce); other halos: NCCL send/recv",
                           "l2": "arrays exceed L2, no flush"},
                "clocks": clocks, "gpu_launches": sim.launches_per_step * args.steps * world,
                "roofline": {"bound": "hbm", "kernel": "k_cg_ring<3> (all ranks)", "achieved": cg_gbs, "peak": peak * world, "unit": "GB/s",
                             "frac": cg_gbs / (peak * world), "traffic": None,
                             "algorithmic_bytes": "cells*(30*iterations+32) per solve, aggregate over ranks"},
                "e2e": e2e}
        print(json.dumps(line))
    sim.slab.close()
    dist.destroy_process_group()
