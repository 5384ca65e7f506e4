"""
Oracle-backed stand-in for `phiflow_b200._ops` (the array-level engine over libphicuda.so) with the SAME interface: device-layout
tensors in, device-layout tensors out or updated in place - on the CPU, computed by oracle/oracle_np.py.  Test infrastructure only:
it lets the `-m "not gpu"` tests execute the HOST logic that sits above the C ABI (phi_cuda adapter / backend / facade, the
phi.flow-like mirror `phiflow_b200/flow.py`, the example scripts) in a container without a GPU.  It says nothing about the kernels;
those are compared with the oracle through the C ABI in the `-m gpu` tests.  The product never imports this file.
"""
import numpy as np
import torch

from oracle import oracle_np as O
from phiflow_b200 import _ops


class OracleEngine:
    Domain = _ops.Domain
    cg_params = staticmethod(_ops.cg_params)
    last = None

    @staticmethod
    def _geom(dom):
        lower = (0.0,) * dom.dim
        upper = tuple(dom.res[a] * dom.dx[a] for a in range(dom.dim))
        return lower, upper

    @classmethod
    def make_incompressible(cls, dom, vspec, v, p=None, prm=None, accessible=None):
        p = dom.alloc_centered() if p is None else p
        prm = prm or _ops.cg_params(vspec)
        if accessible is not None:
            cls.make_incompressible_masked(dom, vspec, v, p, prm, accessible)
            return v, p
        comps = dom.faces_to_numpy(v, vspec, squeeze=False)
        p0 = dom.centered_to_numpy(p, squeeze=False)
        outs, ps, infos = [], [], []
        solver = O.cg_adaptive if prm.method == 1 else O.cg
        for b in range(dom.batch):
            vb = [c[b] for c in comps]
            # same sequence as oracle.make_incompressible, with the solver the engine was asked for
            div = O.divergence_staggered(vb, dom.dx, O.component_bcs(vspec, dom.dim))
            if not O.is_flexible(vspec):
                div = div - np.mean(div, dtype=np.float32)
            Amat = O.poisson_matrix(dom.res, dom.dx, O.pressure_bc(vspec))
            info = solver(Amat, div, p0[b], prm.rtol, prm.atol, prm.max_iter, None)
            pb = info['x'].reshape(dom.res)
            grad = O.gradient_faces(pb, dom.dx, O.pressure_bc(vspec), vspec)
            outs.append([a - g for a, g in zip(vb, grad)]); ps.append(pb); infos.append(info)
        new = dom.faces_from_numpy([np.stack([o[c] for o in outs]) for c in range(dom.dim)], vspec)
        for c in range(dom.dim):
            v[c].copy_(new[c])
        p.copy_(dom.centered_from_numpy(np.stack(ps)))
        rec = np.zeros(dom.batch, dtype=_ops._RESULT_DTYPE)
        for b, info in enumerate(infos):
            rec[b] = (info['iterations'], int(info['converged']), int(info['diverged']), info['residual_sq'], info['tol_sq'], 0.0)
        cls.last = rec
        return v, p

    @classmethod
    def make_incompressible_centered(cls, dom, vspec, v, p, rtol=1e-5, atol=1e-5, max_iter=1000, matrix_offset=None):
        comps = [dom.centered_to_numpy(t, squeeze=False) for t in v]
        outs, ps, rec = [], [], np.zeros(dom.batch, dtype=_ops._RESULT_DTYPE)
        for b in range(dom.batch):
            vb, pb, info = O.make_incompressible_centered([c[b] for c in comps], vspec, dom.res, dom.dx, rtol, atol, max_iter)
            outs.append(vb); ps.append(pb)
            rec[b] = (info['iterations'], int(info['converged']), int(info['diverged']), info['residual_sq'], info['tol_sq'], 0.0)
        cls.last = rec
        new = [dom.centered_from_numpy(np.stack([o[c] for o in outs])) for c in range(dom.dim)]
        return new, dom.centered_from_numpy(np.stack(ps))

    @classmethod
    def read_results(cls, dom):
        return cls.last

    @classmethod
    def advect_staggered(cls, dom, vspec, v, fspec, f, dt):
        lower, upper = cls._geom(dom)
        vc, fc = dom.faces_to_numpy(v, vspec, squeeze=False), dom.faces_to_numpy(f, fspec, squeeze=False)
        out = [O.semi_lagrangian_staggered([c[b] for c in fc], fspec, [c[b] for c in vc], vspec, dom.res, lower, upper, dt) for b in range(dom.batch)]
        return dom.faces_from_numpy([np.stack([o[c] for o in out]) for c in range(dom.dim)], fspec)

    @classmethod
    def advect_centered(cls, dom, vspec, v, sspec, s, dt):
        lower, upper = cls._geom(dom)
        vc, sc = dom.faces_to_numpy(v, vspec, squeeze=False), dom.centered_to_numpy(s, squeeze=False)
        return dom.centered_from_numpy(np.stack([O.semi_lagrangian_centered(sc[b], sspec, [c[b] for c in vc], vspec, lower, upper, dt)
                                                 for b in range(dom.batch)]))

    @classmethod
    def mac_cormack_centered(cls, dom, vspec, v, sspec, s, dt, correction_strength=1.0):
        lower, upper = cls._geom(dom)
        vc, sc = dom.faces_to_numpy(v, vspec, squeeze=False), dom.centered_to_numpy(s, squeeze=False)
        return dom.centered_from_numpy(np.stack([O.mac_cormack_centered(sc[b], sspec, [c[b] for c in vc], vspec, lower, upper, dt, correction_strength)
                                                 for b in range(dom.batch)]))

    @classmethod
    def laplace(cls, dom, spec, x):
        a = dom.centered_to_numpy(x, squeeze=False)
        return dom.centered_from_numpy(np.stack([O.laplace(a[b], dom.dx, spec) for b in range(dom.batch)]))

    @classmethod
    def grid_sample(cls, dom, bc, grid, coords):
        g = dom.centered_to_numpy(grid, squeeze=False)                     # (batch, x, y[, z])
        c = coords.numpy()
        return torch.from_numpy(np.stack([O.grid_sample(g[b], c[b], bc) for b in range(dom.batch)]).astype(np.float32))

    @classmethod
    def cg_poisson(cls, dom, vspec, rhs, x, prm):
        y, x0 = dom.centered_to_numpy(rhs, squeeze=False), dom.centered_to_numpy(x, squeeze=False)
        Amat = O.poisson_matrix(dom.res, dom.dx, O.pressure_bc(vspec))
        solver = O.cg_adaptive if prm.method == 1 else O.cg
        rec, xs = np.zeros(dom.batch, dtype=_ops._RESULT_DTYPE), []
        for b in range(dom.batch):
            info = solver(Amat, y[b], x0[b], prm.rtol, prm.atol, prm.max_iter, None)
            xs.append(info['x'].reshape(dom.res))
            rec[b] = (info['iterations'], int(info['converged']), int(info['diverged']), info['residual_sq'], info['tol_sq'], 0.0)
        cls.last = rec
        x.copy_(dom.centered_from_numpy(np.stack(xs)))
        return x

    @classmethod
    def divergence(cls, dom, vspec, v):
        vc = dom.faces_to_numpy(v, vspec, squeeze=False)
        return dom.centered_from_numpy(np.stack([O.divergence_staggered([c[b] for c in vc], dom.dx, O.component_bcs(vspec, dom.dim))
                                                 for b in range(dom.batch)]))

    # ---- the rest of the engine surface used by phiflow_b200/flow.py ---------------------------------------------------------------
    stored_faces = staticmethod(_ops.stored_faces)
    PhiCgParams = _ops.PhiCgParams

    @classmethod
    def _record(cls, dom, infos):
        rec = np.zeros(dom.batch, dtype=_ops._RESULT_DTYPE)
        for b, info in enumerate(infos):
            rec[b] = (info['iterations'], int(info['converged']), int(info['diverged']), info['residual_sq'], info['tol_sq'], 0.0)
        cls.last = rec

    @classmethod
    def make_incompressible_masked(cls, dom, vspec, v, p, prm, accessible):
        comps = dom.faces_to_numpy(v, vspec, squeeze=False)
        p0, acc = dom.centered_to_numpy(p, squeeze=False), dom.centered_to_numpy(accessible, squeeze=False)
        outs, ps, infos = [], [], []
        for b in range(dom.batch):
            vb, pb, info = O.make_incompressible_obstacles([c[b] for c in comps], vspec, dom.res, dom.dx, acc[b], None, prm.rtol, prm.atol,
                                                           prm.max_iter, x0=p0[b])
            outs.append(vb); ps.append(pb); infos.append(info)
        new = dom.faces_from_numpy([np.stack([o[c] for o in outs]) for c in range(dom.dim)], vspec)
        for c in range(dom.dim):
            v[c].copy_(new[c])
        p.copy_(dom.centered_from_numpy(np.stack(ps)))
        cls._record(dom, infos)

    @classmethod
    def laplace_axpy(cls, dom, bc, x, coeff, out=None):
        a = dom.centered_to_numpy(x, squeeze=False)
        res = dom.centered_from_numpy(np.stack([(a[b] + np.float32(coeff) * O.laplace(a[b], dom.dx, bc)).astype(np.float32) for b in range(dom.batch)]))
        if out is not None:
            out.copy_(res)
            return out
        return res

    @classmethod
    def laplace_axpy_faces(cls, dom, vbc, v, coeff, substeps=1):
        """The REAL host composition of phiflow_b200._ops (slicing of the stored faces into per-component domains) over this
        engine's laplace_axpy."""
        real = _ops.laplace_axpy_faces
        saved = _ops.require_cuda, _ops.laplace_axpy
        _ops.require_cuda, _ops.laplace_axpy = (lambda: None), cls.laplace_axpy
        try:
            return real(dom, vbc, v, coeff, substeps)
        finally:
            _ops.require_cuda, _ops.laplace_axpy = saved

    @classmethod
    def grad_sub(cls, dom, vspec, v, p, accessible=None):
        assert accessible is None
        comps, pp = dom.faces_to_numpy(v, vspec, squeeze=False), dom.centered_to_numpy(p, squeeze=False)
        new = []
        for c in range(dom.dim):
            new.append(np.stack([comps[c][b] - O.gradient_faces(pp[b], dom.dx, O.pressure_bc(vspec), vspec)[c] for b in range(dom.batch)]))
        for c, t in enumerate(dom.faces_from_numpy(new, vspec)):
            v[c].copy_(t)
        return v

    @classmethod
    def mul_faces(cls, dom, vspec, v, mask):
        for c in range(dom.dim):
            v[c].mul_(mask[c])
        return v

    @classmethod
    def add_buoyancy(cls, dom, vspec, sbc, s, factor, dt, v):
        comps, ss = dom.faces_to_numpy(v, vspec, squeeze=False), dom.centered_to_numpy(s, squeeze=False)
        new = []
        for c in range(dom.dim):
            f = np.float32(factor[c] if c < len(factor) else 0.0)
            new.append(np.stack([comps[c][b] + np.float32(dt) * O.centered_to_faces(ss[b] * f, sbc, vspec)[c] for b in range(dom.batch)]))
        for c, t in enumerate(dom.faces_from_numpy(new, vspec)):
            v[c].copy_(t)
        return v

    @classmethod
    def plume_step(cls, dom, vspec, sbc, v, s, p, inflow, dt, inflow_rate, buoyancy, prm, mac_cormack=False, cg_events=None, static_scalar=False):
        assert not static_scalar, "forced step: not in the stand-in"
        lower, upper = cls._geom(dom)
        comps = dom.faces_to_numpy(v, vspec, squeeze=False)
        ss, pp = dom.centered_to_numpy(s, squeeze=False), dom.centered_to_numpy(p, squeeze=False)
        infl = dom.centered_to_numpy(inflow, squeeze=False) if inflow is not None else np.zeros_like(ss)
        Amat = O.poisson_matrix(dom.res, dom.dx, O.pressure_bc(vspec))
        outs, s_new, p_new, infos = [], [], [], []
        for b in range(dom.batch):
            vb, sb, pb, info = O.plume_step([c[b] for c in comps], ss[b], pp[b], dt, vspec, sbc, lower, upper, dom.res, infl[b], inflow_rate,
                                            tuple(buoyancy), rtol=prm.rtol, atol=prm.atol, max_iter=prm.max_iter, use_matrix_offset=False,
                                            matrix=Amat, smoke_advection='mac_cormack' if mac_cormack else 'semi_lagrangian')
            outs.append(vb); s_new.append(sb); p_new.append(pb); infos.append(info)
        for c, t in enumerate(dom.faces_from_numpy([np.stack([o[c] for o in outs]) for c in range(dom.dim)], vspec)):
            v[c].copy_(t)
        s.copy_(dom.centered_from_numpy(np.stack(s_new)))
        p.copy_(dom.centered_from_numpy(np.stack(p_new)))
        cls._record(dom, infos)
        return v, s, p
