"""
`PhiCudaBackend` - the reference's plugin interface for compute libraries: a `phiml.backend.Backend` registered with
`BACKENDS.append(...)` / `set_global_default_backend(...)` (PhiML/phiml/backend/_backend.py:1700, 1874-1902).

Storage and every generic tensor operation are PyTorch's (CUDA tensors; the class derives from phiml's TorchBackend), the methods
the hot path dispatches to are libphicuda's:

    grid_sample(grid, coordinates, mode)             -> phicuda_grid_sample_f32        (_backend.py:1578-1593)
    linear_solve('CG' | 'CG-adaptive' | 'auto', lin, ...) with `lin` a `PoissonOperator` tag
                                                     -> phicuda_cg_poisson_f32         (_backend.py:1408-1464, _linalg.py:52-128)
    jit_compile(f)                                   -> f   (ctypes launches must not be traced; CUDA graphs replace JIT)

Everything else - and every call these overrides do not recognise - behaves exactly like the stock torch backend: the overrides
return `NotImplemented` / defer to super(), which is the reference's own fall-through convention (PhiML/phiml/math/_ops.py:983-1015).
"""
import numpy as np
import torch

from phiml.backend import Backend, BACKENDS  # noqa: F401
from phiml.backend._backend import SolveResult
from phiml.backend.torch._torch_backend import TorchBackend

from . import _adapter


class PoissonOperator:
    """Tag object passed as `lin` to Backend.linear_solve: "the pressure Poisson operator of fluid.masked_laplace for this grid"
    (phi/physics/fluid.py:165-202).  The reference passes the traced CSR matrix here; the fast path needs the grid, not a matrix."""

    def __init__(self, resolution, dx, vspec):
        self.resolution, self.dx, self.vspec = tuple(resolution), tuple(float(h) for h in dx), vspec

    def __call__(self, x):  # pragma: no cover - never evaluated as a function
        raise NotImplementedError("PoissonOperator is a tag for PhiCudaBackend.linear_solve")


class PhiCudaBackend(TorchBackend):

    def __init__(self):
        TorchBackend.__init__(self)
        self._name = 'phicuda'
        gpus = [d for d in self._devices if d.device_type == 'GPU'] if hasattr(self, '_devices') else []
        if gpus:
            self._default_device = gpus[0]

    @property
    def name(self):
        return 'phicuda'

    # --- JIT: never trace ctypes launches ---
    def jit_compile(self, f):
        return f

    def grid_sample(self, grid, coordinates, extrapolation: str):
        grid, coordinates = self.as_tensor(grid), self.as_tensor(coordinates)
        if isinstance(grid, torch.Tensor) and grid.device.type == torch.device(_adapter.DEVICE).type:
            result = _adapter.grid_sample_native(grid, coordinates.to(grid.device), extrapolation)
            if result is not NotImplemented:
                return result
        return TorchBackend.grid_sample(self, grid, coordinates, extrapolation)

    def linear_solve(self, method, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset) -> SolveResult:
        if not isinstance(lin, PoissonOperator) or pre is not None or method not in _adapter.FAST_SOLVERS or len(max_iter) > 1:
            if isinstance(lin, PoissonOperator):
                raise NotImplementedError(f"PoissonOperator with method={method}, preconditioner={pre}: use the stock path")
            return TorchBackend.linear_solve(self, method, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset)
        ops = _adapter.ENGINE
        res = lin.resolution
        d = len(res)
        y, x0 = self.as_tensor(y), self.as_tensor(x0)
        batch = y.shape[0]
        dom = ops.Domain(res, lin.dx, batch, vbc=lin.vspec, device=_adapter.DEVICE)
        idx = (slice(None),) + tuple(slice(0, res[a]) for a in range(d - 1, -1, -1))
        # natives are flattened in the reference's (x, y, z) order: x outermost
        to_dev = lambda t: t.reshape(batch, *res).permute(0, *range(d, 0, -1)).to(device=_adapter.DEVICE, dtype=torch.float32)
        rhs, x = dom.alloc_centered(), dom.alloc_centered()
        rhs[idx] = to_dev(y)
        x[idx] = to_dev(x0)
        rt, at = float(np.max(np.asarray(rtol))), float(np.max(np.asarray(atol)))
        adaptive = method in ('auto', 'CG-adaptive')
        prm = ops.cg_params(lin.vspec, rtol=rt, atol=at, max_iter=int(np.max(max_iter)), method='CG-adaptive' if adaptive else 'CG', balance=False)
        ops.cg_poisson(dom, lin.vspec, rhs, x, prm)
        info = ops.read_results(dom)
        xs = x[idx].permute(0, *range(d, 0, -1)).reshape(batch, -1)
        residual = torch.zeros_like(xs)            # the kernel keeps |r|^2 only; SolveInfo.residual consumers get the norm via `message`
        name = f"phicuda {'CG-adaptive' if adaptive else 'CG'} (persistent TMA-ring kernel)"
        return SolveResult(name, xs, residual, torch.as_tensor(info['iterations'].copy()), torch.as_tensor(info['iterations'].copy() + 1),
                           torch.as_tensor(info['converged'].astype(bool)), torch.as_tensor(info['diverged'].astype(bool)),
                           [f"|r|^2={float(r):.3e}" for r in info['residual_sq']])


PHICUDA = None


def get_backend() -> 'PhiCudaBackend':
    """The singleton, registered in phiml's BACKENDS list on first use."""
    global PHICUDA
    if PHICUDA is None:
        PHICUDA = PhiCudaBackend()
        if PHICUDA not in BACKENDS:
            BACKENDS.append(PHICUDA)
    return PHICUDA
