"""
tests/test_gpu_flow_api.py (the reference's own hot-path tests re-run against the phi.flow-like mirror) once more WITHOUT a GPU: the
same test functions, with the engine under `phiflow_b200/flow.py` replaced by the oracle-backed stand-in (tests/oracle_engine.py).
Covers the host layer of the mirror in the `-m "not gpu"` run; the `-m gpu` run of the original module covers the kernels.
"""
import pytest

import phiflow_b200.flow as flow
import test_gpu_flow_api as G
from oracle_engine import OracleEngine

# the original module star-imports the mirror only where CUDA is available; give it the same names here
for _name in flow.__all__:
    G.__dict__.setdefault(_name, getattr(flow, _name))
for _name in ('fluid', 'advect', 'diffuse', 'field', 'math', 'extrapolation'):
    G.__dict__.setdefault(_name, getattr(flow, _name))


@pytest.fixture(autouse=True, scope='module')
def oracle_engine_under_the_mirror():
    saved = flow.ops, flow._DEVICE
    flow.ops = OracleEngine
    flow.set_device('cpu')
    try:
        yield
    finally:
        flow.ops = saved[0]
        flow.set_device(saved[1])


from test_gpu_flow_api import (  # noqa: E402,F401
    test_make_incompressible_staggered, test_make_incompressible_batched, test_advection_identities, test_self_advect_staggered,
    test_staggered_grid_sizes_by_extrapolation, test_staggered_grid_with_extrapolation, test_explicit_diffusion_known_answer,
    test_divergence_and_laplace_fields, test_not_converged_and_suppress_and_tape, test_incompressible_step_matches_sequenced_calls,
    test_make_incompressible_with_obstacles, test_scene_trajectory_roundtrip)
