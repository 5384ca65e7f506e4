"""
phi.cuda for PhiFlow: the reference-side plugin over libphicuda.so (boundary 1).

    _adapter.py   phiml Extrapolation -> PhiBC specs, named-dim Tensors <-> device layout, eligibility, the hot-path functions on
                  phiml Tensors (make_incompressible, semi_lagrangian, laplace, divergence, grid_sample)
    _backend.py   `PhiCudaBackend`: a phiml Backend (torch tensors as storage) whose grid_sample / linear_solve run libphicuda kernels
    flow.py       the façade: `from phiflow_b200.phi_cuda.flow import *` (or, installed into the phi tree, `from phi.cuda.flow import *`)
                  = `from phi.flow import *` with fluid / advect / field swapped for fast-path wrappers that fall through when a
                  case is not eligible (pattern of phi/torch/flow.py:14-35)

Importing this package needs PhiML (and flow.py needs PhiFlow) to be installed; phiflow_b200 itself does not.
"""
from ._adapter import NotEligible, eligible, to_spec, to_vspec  # noqa: F401
