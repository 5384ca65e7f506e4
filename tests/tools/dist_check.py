#!/usr/bin/env python
"""
Multi-GPU parity check:  torchrun --nproc-per-node N tests/tools/dist_check.py
Every rank runs its z-slab of a small plume; rank 0 additionally runs the whole grid on its own GPU with the
single-GPU kernels, and the gathered slab results must agree with it (the distributed solve only changes the order of
the dot-product reductions).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from phiflow_b200 import _ops as ops  # noqa: E402
from phiflow_b200.dist import Slab, SlabPlume  # noqa: E402


def gather_centered(slab, t):
    H, nz = slab.halo, slab.nz
    own = t[:, H:H + nz].contiguous()
    parts = [torch.empty_like(own) for _ in range(slab.world)]
    dist.all_gather(parts, own)
    return torch.cat(parts, dim=1)


def run_case(name, vbc, sbc, res, steps, rtol):
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device())
    dx = tuple(100.0 / r for r in res)
    rng = np.random.default_rng(5)
    nx, ny, nz = res
    # global initial state in device layout (z, y, x); staggered arrays are filled on every index (unused slots are ignored)
    v0 = [(0.05 * rng.standard_normal((1, nz + 1, ny + 1, nx + 4))).astype(np.float32) for _ in range(3)]
    zc, yc, xc = np.meshgrid(np.arange(nz) + .5, np.arange(ny) + .5, np.arange(nx) + .5, indexing='ij')
    inflow = np.clip(1.5 - np.sqrt((xc * dx[0] - 50) ** 2 + (yc * dx[1] - 50) ** 2 + (zc * dx[2] - 30) ** 2) / 10.0, 0, 1).astype(np.float32)[None]
    prm = ops.cg_params(vbc, rtol=rtol, atol=1e-7, max_iter=2000)

    slab = Slab(res, dx, vbc, halo=3, device=dev)
    sim = SlabPlume(slab, sbc, 0.5, 0.2, (0.0, 0.0, 0.1), prm)
    H, nzl, z0 = slab.halo, slab.nz, slab.z0
    d = slab.dom
    for c in range(3):
        ez, ey, ex = d.fext[2] - 2 * H, d.fext[1], d.fext[0]
        sim.v[c][:, H:H + ez] = torch.from_numpy(v0[c][:, z0:z0 + ez, :ey, :ex]).to(dev)
    sim.inflow[:, H:H + nzl, :, :nx] = torch.from_numpy(inflow[:, z0:z0 + nzl]).to(dev)
    sim.project()
    iters = []
    for _ in range(steps):
        sim.step()
        iters.append(int(sim.slab.results()['iterations'][0]))
        if rank == 0:
            print(f"[{name}] dist solve result: {sim.slab.results()}", flush=True)
    slab = sim.slab
    d, H = slab.dom, slab.halo
    s_all = gather_centered(slab, sim.s)
    p_all = gather_centered(slab, sim.p)
    div = ops.divergence(d, slab.vbc, sim.v)          # needs v halo of the upper neighbour
    slab.exchange(sim.v, 1)
    div = ops.divergence(d, slab.vbc, sim.v)
    div_all = gather_centered(slab, div)
    v_all = []
    for c in range(3):
        own = sim.v[c][:, H:H + nzl].contiguous()
        parts = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(parts, own)
        v_all.append(torch.cat(parts, dim=1))
    ok = True
    if rank == 0:
        dom = ops.Domain(res, dx, 1, vbc=vbc, device=dev)
        v = []
        for c in range(3):
            t = dom.alloc_faces()[0]
            ez, ey, ex = dom.fext[2], dom.fext[1], dom.fext[0]
            t[:] = torch.from_numpy(v0[c][:, :ez, :ey, :ex]).to(dev)
            v.append(t)
        s, p = dom.alloc_centered(), dom.alloc_centered()
        infl = dom.alloc_centered()
        infl[:, :, :, :nx] = torch.from_numpy(inflow).to(dev)
        ops.make_incompressible(dom, vbc, v, p, prm)
        ref_iters = []
        for _ in range(steps):
            ops.plume_step(dom, vbc, sbc, v, s, p, infl, 0.5, 0.2, (0.0, 0.0, 0.1), prm)
            ref_iters.append(int(ops.read_results(dom)['iterations'][0]))
            print(f"[{name}] single solve result: {ops.read_results(dom)}", flush=True)
        ds = float((s_all - s).abs().max()); sm = float(s.abs().max())
        dp = float((p_all - p).abs().max()); pm = float(p.abs().max())
        dmax = float(div_all.abs().max())
        ref_div = ops.divergence(dom, vbc, v)
        loc = np.unravel_index(int(div_all.abs().argmax()), div_all.shape)
        dv = [float((v_all[c][:, :nz, :dom.fext[1], :dom.fext[0]] - v[c][:, :nz]).abs().max()) for c in range(3)]
        print(f"[{name}] ref max|div|={float(ref_div.abs().max()):.3e} dist-div argmax at (b,z,y,x)={loc} max|v diff| per comp={dv}", flush=True)
        vmax = max(float(c.abs().max()) for c in v)
        print(f"[{name}] world={world} iters dist={iters} single={ref_iters} | max|s diff|={ds:.3e} (max {sm:.3e}) "
              f"max|p diff|={dp:.3e} (max {pm:.3e}) max|div|={dmax:.3e} vmax={vmax:.3e}", flush=True)
        ok = ds <= 1e-3 * max(sm, 1e-3) and dp <= 50 * rtol * max(pm, 1e-6) and all(abs(a - b) <= max(3, b // 10) for a, b in zip(iters, ref_iters))
        ok = ok and dmax <= 50 * rtol * vmax * sum(2.0 / h for h in dx)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    slab.close()
    return bool(flag.item())


def run_oracle_case(name, disp_cells, steps, rtol=1e-5, start_halo=3):
    """CFL-derived advection halo against the ORACLE (not only against the single-GPU kernels): a periodic plume whose
    z velocity moves samples `disp_cells` cells per step - more than the 4 planes round 1 hard-coded, and more than the
    `start_halo` planes allocated at the start, so the state is re-allocated with a wider halo on the way
    (reference: the unbounded back-trace of phi/physics/advect.py:20-24, 156-179)."""
    from oracle import oracle_np as O
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device())
    res = (64, 48, 16 * world)
    nx, ny, nz = res
    dx = tuple(100.0 / r for r in res)
    vbc, sbc = (('periodic', 'periodic'),) * 3, (('zg', 'zg'),) * 3
    dt = 0.5
    xs, ys, zs = np.meshgrid((np.arange(nx) + .5) / nx, (np.arange(ny) + .5) / ny, (np.arange(nz) + .5) / nz, indexing='ij')
    w0 = disp_cells * dx[2] / dt
    rng = np.random.default_rng(11)
    v = [(0.3 * w0 * np.sin(2 * np.pi * ys) * np.cos(2 * np.pi * zs) + 0.02 * w0 * rng.standard_normal(res)).astype(np.float32),
         (0.2 * w0 * np.cos(2 * np.pi * xs) * np.sin(4 * np.pi * zs) + 0.02 * w0 * rng.standard_normal(res)).astype(np.float32),
         (w0 * (0.8 + 0.15 * np.sin(2 * np.pi * xs) * np.cos(2 * np.pi * ys)) + 0.02 * w0 * rng.standard_normal(res)).astype(np.float32)]
    inflow = O.sphere_soft_mask((50.0, 50.0, 30.0), 15.0, (0.0,) * 3, (100.0,) * 3, res)
    s0 = (0.5 + 0.5 * np.sin(2 * np.pi * xs) * np.sin(2 * np.pi * ys) * np.cos(2 * np.pi * zs)).astype(np.float32)
    prm = ops.cg_params(vbc, rtol=rtol, atol=1e-7, max_iter=3000)
    slab = Slab(res, dx, vbc, halo=start_halo, device=dev)
    sim = SlabPlume(slab, sbc, dt, 0.2, (0.0, 0.0, 0.1), prm)
    H, nzl, z0 = slab.halo, slab.nz, slab.z0
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a[:, :, z0:z0 + nzl].transpose(2, 1, 0))).to(dev)
    for c in range(3):
        sim.v[c][0, H:H + nzl, :ny, :nx] = to_dev(v[c])
    sim.s[0, H:H + nzl, :ny, :nx] = to_dev(s0)
    sim.inflow[0, H:H + nzl, :ny, :nx] = to_dev(inflow)
    sim.project()
    for _ in range(steps):
        sim.step()
    slab = sim.slab
    own = lambda t: t[:, slab.halo:slab.halo + nzl, :ny, :nx].contiguous()
    def gather(t):
        parts = [torch.empty_like(own(t)) for _ in range(world)]
        dist.all_gather(parts, own(t))
        return torch.cat(parts, dim=1)[0].permute(2, 1, 0).cpu().numpy()
    s_all = gather(sim.s)
    v_all = [gather(sim.v[c]) for c in range(3)]
    ok = True
    if rank == 0:
        A = O.poisson_matrix(res, dx, O.pressure_bc(vbc))
        vr, pr, _ = O.make_incompressible(v, vbc, res, dx, rtol, 1e-7, 3000, matrix=A, use_matrix_offset=False)
        sr = s0
        for _ in range(steps):
            vr, sr, pr, info = O.plume_step(vr, sr, pr, dt, vbc, sbc, (0.0,) * 3, (100.0,) * 3, res, inflow, 0.2, (0.0, 0.0, 0.1),
                                            rtol=rtol, atol=1e-7, max_iter=3000, use_matrix_offset=False, matrix=A)
        vmax = max(float(np.abs(c).max()) for c in vr)
        ds = float(np.abs(s_all - sr).max())
        dv = [float(np.abs(v_all[c] - vr[c]).max()) for c in range(3)]
        need = int(np.ceil(sim.max_displacement)) + 1
        print(f"[{name}] world={world} vs ORACLE: max|s diff|={ds:.3e} (max {float(np.abs(sr).max()):.3e}) max|v diff|={dv} (vmax {vmax:.3e}) "
              f"displacement={sim.max_displacement:.2f} cells, halo used={sim.halo_used}, allocated={slab.halo}, regrown={sim.regrown}", flush=True)
        ok = ds <= 5e-4 * max(float(np.abs(sr).max()), 1e-3) and all(d <= 2e-3 * vmax for d in dv)
        ok = ok and sim.max_displacement > 4.0 and sim.halo_used >= need and sim.regrown >= 1
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    slab.close()
    return bool(flag.item())


def run_obstacle_case(name, rtol=1e-5):
    """N4 on z-slabs: make_incompressible with a static obstacle that straddles a slab face, against the single-GPU masked
    projection (the masked TMA-ring CG on both sides; same iterates up to the order of the dot-product reductions)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device())
    res = (128, 24, 8 * world)
    nx, ny, nz = res
    dx = tuple(50.0 / r for r in res)
    vbc = ((0.0, 0.0), ('periodic', 'periodic'), ('periodic', 'periodic'))
    rng = np.random.default_rng(17)
    acc = np.ones((nz, ny, nx), np.float32)
    acc[nz // 2 - 3:nz // 2 + 2, 6:14, 40:70] = 0           # crosses the interface between two slabs
    acc[:2, 2:5, 5:20] = 0                                   # touches the periodic wrap in z
    v0 = [(0.1 * rng.standard_normal((1, nz, ny, nx + 4))).astype(np.float32) for _ in range(3)]
    prm = ops.cg_params(vbc, rtol=rtol, atol=1e-7, max_iter=3000)
    slab = Slab(res, dx, vbc, halo=2, device=dev)
    d, H, nzl, z0 = slab.dom, slab.halo, slab.nz, slab.z0
    v = d.alloc_faces()
    for c in range(3):
        v[c][:, H:H + nzl, :, :d.fext[0]] = torch.from_numpy(v0[c][:, z0:z0 + nzl, :, :d.fext[0]]).to(dev)
    a = d.alloc_centered()
    a[0, H:H + nzl, :, :nx] = torch.from_numpy(acc[z0:z0 + nzl]).to(dev)
    slab.exchange([a], H)
    p, div = d.alloc_centered(), d.alloc_centered()
    slab.make_incompressible(v, p, div, prm, accessible=a)
    it_dist = int(slab.results()['iterations'][0])
    info = ops.last_launch_info()
    own = lambda t: t[:, H:H + nzl].contiguous()
    def gather(t):
        parts = [torch.empty_like(own(t)) for _ in range(world)]
        dist.all_gather(parts, own(t))
        return torch.cat(parts, dim=1)
    p_all = gather(p)
    v_all = [gather(t) for t in v]
    ok = True
    if rank == 0:
        dom = ops.Domain(res, dx, 1, vbc=vbc, device=dev)
        vs = dom.alloc_faces()
        for c in range(3):
            vs[c][:, :, :, :dom.fext[0]] = torch.from_numpy(v0[c][:, :, :, :dom.fext[0]]).to(dev)
        accs = dom.alloc_centered()
        accs[0, :, :, :nx] = torch.from_numpy(acc).to(dev)
        vs, ps = ops.make_incompressible(dom, vbc, vs, None, prm, accessible=accs)
        r = ops.read_results(dom)
        sinfo = ops.last_launch_info()
        dp = float((p_all - ps).abs().max()); pm = float(ps.abs().max())
        dv = [float((v_all[c] - vs[c]).abs().max()) for c in range(3)]
        vmax = max(float(t.abs().max()) for t in vs)
        print(f"[{name}] world={world} masked ring: dist kernel masked={info['masked']} dist={info['dist']}, single masked={sinfo['masked']} kernel={sinfo['kernel']} | "
              f"iters dist={it_dist} single={int(r['iterations'][0])} max|p diff|={dp:.3e} (max {pm:.3e}) max|v diff|={dv} (vmax {vmax:.3e})", flush=True)
        ok = info['masked'] == 1 and sinfo['masked'] == 1 and abs(it_dist - int(r['iterations'][0])) <= 3 and dp <= 50 * rtol * max(pm, 1e-6) \
            and all(x <= 50 * rtol * vmax for x in dv) and int(r['converged'][0]) == 1
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    slab.close()
    return bool(flag.item())


def main():
    dist.init_process_group('nccl')
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    world = dist.get_world_size()
    per = (('periodic', 'periodic'),) * 3
    mixed = (('periodic', 'periodic'), (0.0, 0.0), (0.0, 'zg'))
    zg = (('zg', 'zg'),) * 3
    ok = True
    ok &= run_case('periodic', per, zg, (64, 48, 16 * world), 3, 1e-4)
    ok &= run_case('mixed-z-wall-open', mixed, zg, (128, 24, 8 * world), 2, 1e-4)
    ok &= run_oracle_case('cfl-halo-vs-oracle', 6.4, 3)
    ok &= run_obstacle_case('obstacle-across-slabs')
    if dist.get_rank() == 0:
        print('DIST_CHECK', 'PASS' if ok else 'FAIL', flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
