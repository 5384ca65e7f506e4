#!/usr/bin/env python
"""
bench.py  --  512^3 smoke-plume steps/s on N B200s (BASELINE.json metric), laplace / CG HBM roofline, CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # our arm   (N>1: launched under torch.distributed.run)
    python bench.py --impl reference --steps K --warmup W    # reference arm: the oracle port on the host cores

One "step" = incompressible_step on the whole grid:
    s' = semi_lagrangian(s, v, dt) + inflow ; v* = semi_lagrangian(v, v, dt) + dt*buoyancy(s') ;
    v', p' = make_incompressible(v*, Solve('CG', 1e-3, x0=p))
Workload = BASELINE.json configs[3] at N GPUs ("3-D smoke plume 512^3 fp32, periodic, z-slab decomposed"): strong scaling.
Inputs are larger than L2 (every array is 512 MiB), so no L2 flush is needed between timed iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "smoke_plume_512^3_steps_per_sec"
DT, INFLOW_RATE, BUOYANCY = 0.5, 0.2, (0.0, 0.0, 0.1)
RTOL, ATOL, MAX_ITER = 1e-3, 1e-5, 1000


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the CG kernel in this bench (ncu --set full capture of the
    same command, committed as profiles/r1_cg_traffic.json); None when no capture is committed."""
    path = os.path.join(ROOT, 'profiles', 'r1_cg_traffic.json')
    try:
        return float(json.load(open(path))['dram_bytes_per_launch'])
    except Exception:
        return None


from phiflow_b200._clocks import ClockSampler  # noqa: E402


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port (NumPy + SciPy CSR CG, i.e. what the reference's NumPy backend runs)
# ------------------------------------------------------------------------------------------------------------------
def cpu_sample(n: int, steps: int, warmup: int, full: int):
    from oracle import oracle_np as O
    res = (n, n, n)
    lower, upper = (0.0,) * 3, (100.0,) * 3
    dx = tuple(100.0 / n for _ in range(3))
    vbc, sbc = O.uniform_bc(3, O.PERIODIC), O.uniform_bc(3, O.ZG)
    inflow = O.sphere_soft_mask((50.0, 50.0, 9.5), 5.0, lower, upper, res)
    rng = np.random.default_rng(0)
    v = [(0.01 * rng.standard_normal(s)).astype(np.float32) for s in O.staggered_shapes(res, vbc)]
    A = O.poisson_matrix(res, dx, O.pressure_bc(vbc))        # the reference builds it once per shape (tracing, untimed)
    v, p, _ = O.make_incompressible(v, vbc, res, dx, RTOL, ATOL, MAX_ITER, matrix=A)
    s = np.zeros(res, np.float32)
    iters = []
    t0 = None
    for i in range(warmup + steps):
        if i == warmup:
            t0 = time.perf_counter()
        v, s, p, info = O.plume_step(v, s, p, DT, vbc, sbc, lower, upper, res, inflow, INFLOW_RATE, BUOYANCY,
                                     rtol=RTOL, atol=ATOL, max_iter=MAX_ITER, matrix=A)
        if i >= warmup:
            iters.append(info['iterations'])
    el = time.perf_counter() - t0
    sample_sps = steps / el
    return {"value": sample_sps * (n ** 3) / float(full ** 3), "unit": "steps/s", "cores": 1, "kind": "port",
            "sample": f"{steps} steps of the same plume at {n}^3 (1/{(full / n) ** 3:.1f} of the cells), NumPy/SciPy oracle port, "
                      f"{sample_sps:.4f} steps/s measured, scaled by cells; CG iterations/step {np.mean(iters):.1f}; "
                      f"the reference's NumPy path is single-threaded ({os.cpu_count()} cores present)"}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    base = cpu_sample(args.cpu_size, args.steps, min(args.warmup, 1), args.size)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 / base["value"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"3-D smoke plume {args.size}^3 fp32 periodic (BASELINE configs[3])", "sample_grid": args.cpu_size},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
class PlumeSim:
    """Device-resident plume state + the step sequenced call by call through the C ABI (so the CG launch can be timed)."""

    def __init__(self, n, device):
        import torch
        from phiflow_b200 import _ops as ops
        self.torch, self.ops = torch, ops
        self.n = n
        self.vbc = (('periodic', 'periodic'),) * 3
        self.sbc = (('zg', 'zg'),) * 3
        dx = tuple(100.0 / n for _ in range(3))
        self.dom = ops.Domain((n, n, n), dx, 1, vbc=self.vbc, device=device)
        g = torch.Generator().manual_seed(0)
        self.v = []
        for c in range(3):
            host = torch.randn((1, n, n, n), generator=g, dtype=torch.float32).mul_(0.01).pin_memory()
            self.v.append(host.to(device, non_blocking=True))
        self.v2 = self.dom.alloc_faces()
        self.s, self.s2 = self.dom.alloc_centered(), self.dom.alloc_centered()
        self.p, self.div = self.dom.alloc_centered(), self.dom.alloc_centered()
        # inflow mask (setup, not per-step compute): soft sphere, phi/geom/_geom.py:278-308
        ax = (torch.arange(n, device=device, dtype=torch.float32) + 0.5) * dx[0]
        z, y, x = torch.meshgrid(ax, ax, ax, indexing='ij')
        dist = torch.sqrt(torch.clamp((x - 50.0) ** 2 + (y - 50.0) ** 2 + (z - 9.5) ** 2, min=1e-6))
        cell_r = float(np.sqrt(3 * (dx[0] * 0.5) ** 2))
        self.inflow = torch.clamp(0.5 - (dist - 5.0) / cell_r, 0, 1).reshape(1, n, n, n).contiguous()
        del x, y, z, dist
        self.prm = ops.cg_params(self.vbc, rtol=RTOL, atol=ATOL, max_iter=MAX_ITER)
        ops.make_incompressible(self.dom, self.vbc, self.v, self.p, self.prm)
        self.p.zero_()
        self.launches_per_step = 9
        self.cg_events = []

    def step(self, time_cg=False):
        ops, dom = self.ops, self.dom
        ops.advect_centered(dom, self.vbc, self.v, self.sbc, self.s, DT, out=self.s2)
        ops.axpy_centered(dom, INFLOW_RATE, self.inflow, self.s2)
        ops.advect_staggered(dom, self.vbc, self.v, self.vbc, self.v, DT, out=self.v2)
        ops.add_buoyancy(dom, self.vbc, self.sbc, self.s2, BUOYANCY, DT, self.v2)
        self.s, self.s2 = self.s2, self.s
        self.v, self.v2 = self.v2, self.v
        ops.divergence(dom, self.vbc, self.v, out=self.div)
        if time_cg:
            e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            e0.record()
        ops.cg_poisson(dom, self.vbc, self.div, self.p, self.prm)
        if time_cg:
            e1.record()
            self.cg_events.append((e0, e1))
        ops.grad_sub(dom, self.vbc, self.v, self.p)


def run_ours(args):
    import torch
    from phiflow_b200 import _ops as ops
    if args.gpus > 1 or int(os.environ.get('WORLD_SIZE', '1')) > 1:
        from phiflow_b200 import dist_bench
        return dist_bench.run(args, METRIC)
    torch.cuda.set_device(0)
    dev = torch.device('cuda:0')
    n = args.size
    peak, peak_kind = load_peaks()
    sim = PlumeSim(n, dev)
    cells = float(n) ** 3
    _, res_dev = sim.dom.workspace()
    res_host = torch.zeros((args.warmup + args.steps, 6), dtype=torch.int32).pin_memory()

    for i in range(args.warmup):
        sim.step()
        res_host[i].copy_(res_dev[:6], non_blocking=True)
    torch.cuda.synchronize()
    # the e2e leg replays the first timed steps from this state, so both legs do the same CG iterations
    snap = [t.clone() for t in (sim.v[0], sim.v[1], sim.v[2], sim.s, sim.p)]
    sampler = ClockSampler(0)
    sampler.start()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for i in range(args.steps):
        sim.step(time_cg=True)
        res_host[args.warmup + i].copy_(res_dev[:6], non_blocking=True)
    end.record()
    torch.cuda.synchronize()
    clocks = sampler.summary()
    ms = start.elapsed_time(end) / args.steps
    iters = res_host[args.warmup:, 0].numpy().astype(np.int64)
    cg_ms = np.array([a.elapsed_time(b) for a, b in sim.cg_events])
    # algorithmic bytes of one solve: 30 B/cell/iteration (pass A 12 + pass B 12 / 24 alternating: the x update is applied
    # every second iteration) + 32 B/cell of setup/teardown passes
    cg_bytes = cells * (30.0 * iters + 32.0)
    cg_gbs = float(np.sum(cg_bytes) / np.sum(cg_ms * 1e-3) / 1e9)

    # laplace micro-benchmark (the metric's second half): 8 B/cell
    x = torch.randn((1, n, n, n), device=dev, dtype=torch.float32)
    y = torch.empty_like(x)
    lbc = (('periodic', 'periodic'),) * 3
    for _ in range(3):
        ops.laplace(sim.dom, lbc, x, out=y)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    l0.record()
    for _ in range(20):
        ops.laplace(sim.dom, lbc, x, out=y)
    l1.record()
    torch.cuda.synchronize()
    lap_ms = l0.elapsed_time(l1) / 20
    lap_gbs = 8.0 * cells / (lap_ms * 1e-3) / 1e9
    del x, y

    # end to end: state held in HOST (pinned) buffers in the reference's (x, y, z) array order; every step uploads it,
    # transposes to the device layout, steps, transposes back and downloads it
    e2e = run_e2e(sim, args, torch, snap)
    del snap

    base = cpu_sample(args.cpu_size, 3, 1, n) if not args.no_cpu else None
    line = {"metric": METRIC, "value": 1e3 / ms, "unit": "steps/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"3-D smoke plume {n}^3 fp32 periodic, CG rtol=1e-3 warm start (BASELINE configs[3] at 1 GPU)",
                       "cg_iterations_per_step": float(np.mean(iters)), "cg_ms_per_step": float(np.mean(cg_ms)),
                       "l2": "inputs (512 MiB per array) exceed L2, no flush"},
            "clocks": clocks, "gpu_launches": sim.launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "kernel": "k_cg_ring<3,false> (persistent CG solve)", "achieved": cg_gbs, "peak": peak, "unit": "GB/s",
                         "frac": cg_gbs / peak, "traffic": load_traffic(), "peak_kind": peak_kind,
                         "algorithmic_bytes": "cells*(30*iterations+32) per solve"},
            "laplace": {"achieved": lap_gbs, "peak": peak, "frac": lap_gbs / peak, "unit": "GB/s", "ms": lap_ms,
                        "algorithmic_bytes": "8 B/cell"},
            "e2e": e2e}
    if base:
        line["cpu_baseline"] = base
    print(json.dumps(line))


def run_e2e(sim, args, torch, snap):
    n = sim.n
    dev = sim.s.device
    host = {k: torch.zeros((n, n, n), dtype=torch.float32).pin_memory() for k in ('vx', 'vy', 'vz', 's', 'p')}
    # seed the host state from the device state at the start of the timed steps (reference layout x, y, z)
    cur = dict(zip(('vx', 'vy', 'vz', 's', 'p'), snap))
    for k, t in cur.items():
        host[k].copy_(t[0].permute(2, 1, 0))
    torch.cuda.synchronize()
    nbytes = sum(h.numel() * 4 for h in host.values())
    steps = max(2, min(args.steps, 5))

    def one():
        dv = [host[k].to(dev, non_blocking=True).permute(2, 1, 0).contiguous().unsqueeze(0) for k in ('vx', 'vy', 'vz')]
        sim.v = dv
        sim.s = host['s'].to(dev, non_blocking=True).permute(2, 1, 0).contiguous().unsqueeze(0)
        sim.p = host['p'].to(dev, non_blocking=True).permute(2, 1, 0).contiguous().unsqueeze(0)
        sim.step()
        for k, t in (('vx', sim.v[0]), ('vy', sim.v[1]), ('vz', sim.v[2]), ('s', sim.s), ('p', sim.p)):
            host[k].copy_(t[0].permute(2, 1, 0), non_blocking=True)
    one()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _, res_dev = sim.dom.workspace()
    its = []
    t0.record()
    for _ in range(steps):
        one()
        its.append(res_dev[:1].clone())
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    return {"value": 1e3 / ms, "unit": "steps/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": steps,
            "cg_iterations_per_step": float(torch.cat(its).float().mean().item()),
            "note": "full state (v, s, p) in pinned host arrays in reference (x,y,z) order; upload + transpose + step + transpose + download"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--cpu-size', type=int, default=96, dest='cpu_size')
    ap.add_argument('--no-cpu', action='store_true', dest='no_cpu')
    ap.add_argument('--halo', type=int, default=16, help='z-slab halo planes allocated per side (N>1); grows on demand')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != 'reference' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
