"""
Multi-GPU execution of the hot path: one process per GPU, z-slab decomposition.

Rank r owns planes [r*nz/P, (r+1)*nz/P) of every array plus `halo` planes on each side.  Interior slab faces get the
boundary kind 'halo' (PHI_BC_HALO): kernels read neighbour values from the halo planes, which this module keeps current
with grouped send/recv over torch.distributed (NCCL on GPUs, gloo in the CPU tests of the host logic).
The pressure solve itself needs no host-side communication: phicuda_cg_poisson_dist_f32 runs the whole CG in one
persistent kernel per rank that exchanges halo planes and dot products through NVLink peer memory (csrc/comm.cu).
Batched 2-D runs shard the batch axis instead - entries are independent systems, no communication at all.
"""
import ctypes as C
import math
from typing import List, Sequence

import torch
import torch.distributed as dist

from . import _lib, _ops as ops

HALO, PERIODIC = 'halo', 'periodic'


def local_bc(spec, rank: int, world: int):
    """Boundary spec of a slab: z sides that border another rank become 'halo'."""
    if world == 1:
        return spec
    if isinstance(spec, list):
        return [local_bc(s, rank, world) for s in spec]
    lo, hi = spec[2]
    periodic = lo == PERIODIC
    zlo = HALO if (periodic or rank > 0) else lo
    zhi = HALO if (periodic or rank < world - 1) else hi
    return (spec[0], spec[1], (zlo, zhi))


def batch_shard(total_batch: int, rank: int = None, world: int = None):
    """Batched 2-D runs (BASELINE configs[4]) shard the batch axis: returns (first entry, count) of this rank.
    Batch entries are independent systems (PhiML/phiml/backend/_linalg.py:72-87), so the sharded run needs NO collective;
    ranks only meet at the final timing barrier."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    base, rem = divmod(int(total_batch), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def advection_halo(max_abs_vz: float, dt: float, dz: float, bc_const_max: float = 0.0) -> int:
    """Halo planes the semi-Lagrangian passes of one step need: h = ceil(max|v_z| * |dt| / dz) + 1.
    The reference's back-trace `points - dt * v(points)` is unbounded (phi/physics/advect.py:20-24, 156-179); the sample at
    z - disp reads the planes floor(z - disp) and floor(z - disp) + 1, and the velocity at a face centre is an average of
    stored values and boundary constants, hence bounded by max(max|v_z|, |constants|).  Non-finite input -> ValueError."""
    vmax = max(float(max_abs_vz), float(bc_const_max))
    if not math.isfinite(vmax):
        raise ValueError("velocity is not finite: the advection halo cannot be bounded")
    disp = vmax * abs(float(dt)) / float(dz)
    return int(math.ceil(disp * (1.0 + 1e-6))) + 1


class HaloTooWide(_lib.Unsupported):
    """The back-trace reaches further than one neighbouring slab (h > planes per rank): outside the z-slab fast path."""

    def __init__(self, need, have):
        RuntimeError.__init__(self, f"semi-Lagrangian back-trace needs {need} halo planes but a slab has only {have} planes; "
                                    f"use fewer ranks or a smaller dt (PHI_ERR_UNSUPPORTED: raised BEFORE the step computes)")
        self.code = _lib.ERR_UNSUPPORTED


class Slab:
    """Geometry + communication of one rank's z-slab."""

    def __init__(self, global_res: Sequence[int], dx: Sequence[float], vbc, halo: int = 2, device='cuda', group=None):
        assert len(global_res) == 3, "z-slab decomposition is for 3-D grids"
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        nx, ny, nz = global_res
        assert nz % self.world == 0, f"nz={nz} must be divisible by the number of ranks {self.world}"
        self.global_res = tuple(global_res)
        self.dx = tuple(float(h) for h in dx)
        self.nz = nz // self.world
        self.z0 = self.rank * self.nz
        self.halo = min(halo, self.nz) if self.world > 1 else 0      # allocated planes per side = widest possible exchange
        self.vbc_global = vbc
        self.vbc = local_bc(vbc, self.rank, self.world)
        spec = vbc[0] if isinstance(vbc, list) else vbc
        periodic = spec[2][0] == PERIODIC
        self.lower = None if self.world == 1 else ((self.rank - 1) % self.world if (periodic or self.rank > 0) else None)
        self.upper = None if self.world == 1 else ((self.rank + 1) % self.world if (periodic or self.rank < self.world - 1) else None)
        self.dom = ops.Domain((nx, ny, self.nz), dx, 1, vbc=self.vbc, device=device, halo=self.halo)
        self._comm = None
        # largest |constant| of the z component's boundary (enters face-centre averages next to walls)
        zspec = (vbc[2] if isinstance(vbc, list) else vbc)
        self.bc_const_max = max([abs(float(side)) for ax in zspec for side in ax if not isinstance(side, str)] + [0.0])

    def bc(self, spec):
        return local_bc(spec, self.rank, self.world)

    # ---- halo exchange --------------------------------------------------------------------------------------------
    def exchange(self, tensors: List[torch.Tensor], width: int):
        """Fills `width` halo planes on both sides of every tensor from the neighbouring slabs.  z planes are contiguous
        (DESIGN.md section 2), so with batch 1 the owned edge planes are sent and the halo planes received in place - no
        staging copies; batched tensors go through contiguous staging buffers."""
        if self.world == 1:
            return
        H, nz = self.halo, self.nz
        if not 1 <= width <= H:
            raise ValueError(f"halo exchange of {width} planes but {H} are allocated")
        ops_ = []
        copy_back = []

        def send(view, peer):
            ops_.append(dist.P2POp(dist.isend, view if view.is_contiguous() else view.contiguous(), peer, self.group))

        def recv(view, peer):
            if view.is_contiguous():
                ops_.append(dist.P2POp(dist.irecv, view, peer, self.group))
            else:
                buf = torch.empty_like(view, memory_format=torch.contiguous_format)
                ops_.append(dist.P2POp(dist.irecv, buf, peer, self.group))
                copy_back.append((view, buf))
        for t in tensors:
            # order matters when lower == upper (two ranks, periodic): sends and receives pair up in issue order
            if self.upper is not None:
                send(t[:, H + nz - width:H + nz], self.upper)
            if self.lower is not None:
                send(t[:, H:H + width], self.lower)
        for t in tensors:
            if self.lower is not None:
                recv(t[:, H - width:H], self.lower)
            if self.upper is not None:
                recv(t[:, H + nz:H + nz + width], self.upper)
        for req in dist.batch_isend_irecv(ops_):
            req.wait()
        for view, buf in copy_back:
            view.copy_(buf)

    # ---- distributed pressure solve -----------------------------------------------------------------------------------
    def _communicator(self):
        if self._comm is None:
            lib = _lib.load()
            handle = (C.c_ubyte * 64)()
            comm = C.c_void_p()
            _lib.check(lib.phicuda_comm_create(self.rank, self.world, C.byref(self.dom.grid), C.byref(comm), handle))
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.dom.device)
            gathered = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine, group=self.group)
            allh = (C.c_ubyte * (64 * self.world))(*torch.cat(gathered).cpu().tolist())
            _lib.check(lib.phicuda_comm_connect(comm, allh))
            dist.barrier(group=self.group)
            self._comm = comm
            self._result = torch.zeros(6, dtype=torch.int32, device=self.dom.device)
        return self._comm

    def cg_poisson(self, rhs: torch.Tensor, x: torch.Tensor, prm, accessible: torch.Tensor = None):
        """CG over all slabs; x (with valid halo planes) is updated in place.  Single rank: the ordinary solve.
        accessible: static obstacle mask with valid halo planes (N4)."""
        if self.world == 1:
            assert accessible is None, "single-rank masked solves go through ops.make_incompressible(accessible=...)"
            return ops.cg_poisson(self.dom, self.vbc, rhs, x, prm)
        comm = self._communicator()
        dom = self.dom
        if accessible is not None:
            _lib.check(_lib.load().phicuda_cg_poisson_dist_masked_f32(C.byref(dom.grid), C.byref(ops.make_vbc(self.vbc, 3)), ops._ptr(rhs, dom.coff),
                                                                      ops._ptr(x, dom.coff), ops._ptr(accessible, dom.coff), C.byref(prm),
                                                                      ops._ptr(self._result), comm, ops._stream()))
            return x
        _lib.check(_lib.load().phicuda_cg_poisson_dist_f32(C.byref(dom.grid), C.byref(ops.make_vbc(self.vbc, 3)), ops._ptr(rhs, dom.coff),
                                                           ops._ptr(x, dom.coff), C.byref(prm), ops._ptr(self._result), comm, ops._stream()))
        return x

    def make_incompressible(self, v: List[torch.Tensor], p: torch.Tensor, div: torch.Tensor, prm, accessible: torch.Tensor = None,
                            vmask: List[torch.Tensor] = None):
        """fluid.make_incompressible on z-slabs (phi/physics/fluid.py:94-162), optionally with static obstacles (N4): the caller
        passes the accessible mask and the face factors of apply_boundary_conditions with valid halo planes."""
        if accessible is not None and vmask is not None:
            ops.mul_faces(self.dom, self.vbc, v, vmask)
        self.exchange(v, 1)
        ops.divergence(self.dom, self.vbc, v, out=div, accessible=accessible)
        self.exchange([p], 1)
        self.cg_poisson(div, p, prm, accessible=accessible)
        self.exchange([p], 1)
        ops.grad_sub(self.dom, self.vbc, v, p, accessible=accessible)
        return v, p

    def results(self):
        if self.world == 1:
            return ops.read_results(self.dom)
        return self._result.cpu().numpy().view(ops._RESULT_DTYPE)

    def result_tensor(self):
        return self.dom.workspace()[1] if self.world == 1 else self._result

    def close(self):
        if self._comm is not None:
            torch.cuda.synchronize(self.dom.device)
            _lib.load().phicuda_comm_destroy(self._comm)
            self._comm = None


class SlabPlume:
    """incompressible_step on z-slabs.  State lives on the device of each rank.

    Advection halo: before every step the ranks agree on max|v_z| (device reduction + all_reduce(MAX)) and exchange
    h = ceil(max|v_z| dt / dz) + 1 planes of v and s (`advection_halo`).  When h exceeds the allocated planes the state is
    re-allocated with a wider halo (collective, all ranks take the same decision); when it exceeds the slab thickness the
    step raises HaloTooWide BEFORE any kernel runs.  `halo_used` / `max_displacement` record the largest values seen."""

    def __init__(self, slab: Slab, sbc, dt, inflow_rate, buoyancy, prm, adv_halo=None):
        self.slab, self.dom = slab, slab.dom
        self.sbc_global = sbc
        self.vbc, self.sbc = slab.vbc, slab.bc(sbc)
        self.dt, self.inflow_rate, self.buoyancy, self.prm = dt, inflow_rate, buoyancy, prm
        self.fixed_adv_halo = adv_halo            # diagnostics only: force a width instead of the CFL-derived one
        d = self.dom
        self.v, self.v2 = d.alloc_faces(), d.alloc_faces()
        self.s, self.s2 = d.alloc_centered(), d.alloc_centered()
        self.p, self.div = d.alloc_centered(), d.alloc_centered()
        self.inflow = d.alloc_centered()
        self._vmax = torch.zeros(3, dtype=torch.float32, device=d.device)
        self.launches_per_step = 10
        self.halo_used = 0
        self.max_displacement = 0.0
        self.regrown = 0

    # ---- halo sizing ------------------------------------------------------------------------------------------------
    def required_halo(self) -> int:
        """Collective: planes the advection of the coming step reads beyond the slab (same value on every rank)."""
        s, d = self.slab, self.dom
        if d.device.type == 'cuda':
            ops.max_abs_velocity(d, self.vbc, self.v, out=self._vmax)
        else:                                       # host-logic tests (gloo): same reduction with torch
            H, nz = s.halo, s.nz
            self._vmax.copy_(torch.stack([c[:, H:H + nz + 1].abs().max() for c in self.v]))
        if s.world > 1:
            dist.all_reduce(self._vmax, op=dist.ReduceOp.MAX, group=s.group)
        vz = float(self._vmax[2].item())            # the one host round trip of the step: the width decides what is exchanged
        h = advection_halo(vz, self.dt, s.dx[2], s.bc_const_max)
        self.max_displacement = max(self.max_displacement, max(vz, s.bc_const_max) * abs(self.dt) / s.dx[2])
        return h

    def ensure_halo(self, h: int):
        s = self.slab
        if s.world == 1 or h <= s.halo:
            return
        if h > s.nz:
            raise HaloTooWide(h, s.nz)
        self._regrow(min(s.nz, max(h, 2 * s.halo)))

    def _regrow(self, new_halo: int):
        """Re-allocates the state with `new_halo` planes per side, keeping the owned planes (collective)."""
        old, oH, nz = self.slab, self.slab.halo, self.slab.nz
        old.close()
        slab = Slab(old.global_res, old.dx, old.vbc_global, halo=new_halo, device=old.dom.device, group=old.group)
        nH, d = slab.halo, slab.dom

        def move(t, alloc):
            n = alloc()
            keep = t.shape[1] - 2 * oH               # owned planes (+ the stored upper boundary face plane, if any)
            n[:, nH:nH + keep] = t[:, oH:oH + keep]
            return n
        self.v = [move(t, lambda: d.alloc_faces()[0]) for t in self.v]
        self.v2 = d.alloc_faces()
        self.s, self.s2 = move(self.s, d.alloc_centered), d.alloc_centered()
        self.p, self.div = move(self.p, d.alloc_centered), d.alloc_centered()
        self.inflow = move(self.inflow, d.alloc_centered)
        self.slab, self.dom = slab, d
        self.vbc, self.sbc = slab.vbc, slab.bc(self.sbc_global)
        self.regrown += 1

    # ---- the step -----------------------------------------------------------------------------------------------------
    def project(self):
        s = self.slab
        s.exchange(self.v, 1)
        ops.divergence(self.dom, self.vbc, self.v, out=self.div)
        s.exchange([self.p], 1)
        s.cg_poisson(self.div, self.p, self.prm)
        s.exchange([self.p], 1)
        ops.grad_sub(self.dom, self.vbc, self.v, self.p)

    def step(self, cg_events=None):
        if self.slab.world > 1:
            h = self.required_halo() if self.fixed_adv_halo is None else self.fixed_adv_halo
            self.ensure_halo(h)                    # may re-allocate; raises HaloTooWide before anything is computed
            self.halo_used = max(self.halo_used, h)
            self.slab.exchange(self.v + [self.s], h)
        s, d = self.slab, self.dom
        ops.advect_centered(d, self.vbc, self.v, self.sbc, self.s, self.dt, out=self.s2)
        ops.axpy_centered(d, self.inflow_rate, self.inflow, self.s2)
        ops.advect_staggered(d, self.vbc, self.v, self.vbc, self.v, self.dt, out=self.v2)
        s.exchange([self.s2], 1)
        ops.add_buoyancy(d, self.vbc, self.sbc, self.s2, self.buoyancy, self.dt, self.v2)
        self.s, self.s2 = self.s2, self.s
        self.v, self.v2 = self.v2, self.v
        s.exchange(self.v, 1)
        ops.divergence(d, self.vbc, self.v, out=self.div)
        s.exchange([self.p], 1)
        if cg_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        s.cg_poisson(self.div, self.p, self.prm)
        if cg_events is not None:
            e1.record()
            cg_events.append((e0, e1))
        s.exchange([self.p], 1)
        ops.grad_sub(d, self.vbc, self.v, self.p)
