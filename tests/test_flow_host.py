"""
Host logic of the phi.flow-like mirror (`phiflow_b200/flow.py`) and the example scripts, executed WITHOUT a GPU: the engine module
`phiflow_b200._ops` is replaced by the oracle-backed stand-in of tests/oracle_engine.py (same interface, CPU tensors in the device
layout).  What this covers is everything above the C ABI - boundary dictionaries incl. vector constants, stored-face bookkeeping, the
per-component slicing of `_ops.laplace_axpy_faces` (the real function runs, over the stand-in's laplace), field arithmetic, Solve /
warm starts, obstacles, Scene output, and that the example scripts run end to end.  The kernels are the `-m gpu` tests' business
(tests/test_gpu_vector_boundaries.py runs the same notebook steps through libphicuda.so).
"""
import importlib.util
import os

import numpy as np
import pytest

from oracle import oracle_np as O
from oracle_engine import OracleEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def F():
    import phiflow_b200.flow as flow
    saved = flow.ops, flow._DEVICE
    flow.ops = OracleEngine
    flow.set_device('cpu')
    try:
        yield flow
    finally:
        flow.ops = saved[0]
        flow.set_device(saved[1])


def example(name):
    spec = importlib.util.spec_from_file_location(f'example_{name}', os.path.join(ROOT, 'examples', name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_lid_driven_cavity_through_the_mirror(F):
    """Lid_Driven_Cavity.ipynb: vector-valued lid constant, staggered diffusion, default Solve() = CG-adaptive with warm start."""
    v = F.StaggeredGrid(0, {'x': 0, 'y-': 0, 'y+': F.vec(x=1, y=0)}, x=24, y=16)
    vspec = [((0.0, 0.0), (0.0, 1.0)), ((0.0, 0.0), (0.0, 0.0))]
    assert v.vspec == vspec and [a.shape for a in v.numpy()] == [(23, 16), (24, 15)]
    res, dx, lower, upper = (24, 16), (1.0, 1.0), (0.0, 0.0), (24.0, 16.0)
    ref = [np.zeros(s, np.float32) for s in O.staggered_shapes(res, vspec)]
    Amat = O.poisson_matrix(res, dx, O.pressure_bc(vspec))
    p, p_ref = None, np.zeros(res, np.float32)
    for _ in range(4):
        v = F.advect.semi_lagrangian(v, v, 1.0)
        v = F.diffuse.explicit(v, 0.1, 1.0)
        v, p = F.fluid.make_incompressible(v, solve=F.Solve(x0=p))
        ref = O.semi_lagrangian_staggered(ref, vspec, ref, vspec, res, lower, upper, 1.0)
        ref = O.diffuse_explicit(ref, vspec, dx, 0.1, 1.0)
        div = O.divergence_staggered(ref, dx, O.component_bcs(vspec, 2))
        div = div - np.mean(div, dtype=np.float32)
        info = O.cg_adaptive(Amat, div, p_ref, 1e-5, 1e-5, 1000, None)          # Solve() = 'auto' = CG-adaptive (_backend.py:1446-1447)
        p_ref = info['x'].reshape(res)
        ref = [a - g for a, g in zip(ref, O.gradient_faces(p_ref, dx, O.pressure_bc(vspec), vspec))]
    for got, want in zip(v.numpy(), ref):
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)
    assert float(np.abs(ref[0]).max()) > 0.1 and p.boundary.spec(('x', 'y')) == (('zg', 'zg'), ('zg', 'zg'))


@pytest.mark.parametrize('name', ['cavity3', 'inflow3', 'open3', 'periodic3', 'mixed3'])
def test_staggered_diffusion_slicing_3d(F, name):
    """_ops.laplace_axpy_faces: which stored faces go into the per-component domain and come back (walls: faces 1..n-1, periodic: 0..n-1,
    open: 0..n), batch of 2, two substeps."""
    vbc = {
        'cavity3': [((0.0, 0.0), (0.0, 1.0), (0.0, 0.0)), ((0.0, 0.0),) * 3, ((0.0, 0.0), (0.0, 0.25), (0.0, 0.0))],
        'inflow3': [((0.5, 'zg'), (0.0, 0.0), ('periodic', 'periodic')), ((-0.25, 'zg'), (0.0, 0.0), ('periodic', 'periodic')),
                    ((0.125, 'zg'), (0.0, 0.0), ('periodic', 'periodic'))],
        'open3': O.uniform_bc(3, 'zg'), 'periodic3': O.uniform_bc(3, 'periodic'),
        'mixed3': (('periodic', 'periodic'), (0.0, 'zg'), ('zg', 0.0)),
    }[name]
    res, dx = (7, 6, 5), (0.5, 1.0, 2.0)
    dom = OracleEngine.Domain(res, dx, 2, vbc=vbc, device='cpu')
    rng = np.random.default_rng(0)
    v = [rng.standard_normal((2,) + s).astype(np.float32) for s in O.staggered_shapes(res, vbc)]
    dv = dom.faces_from_numpy(v, vbc)
    out = dom.faces_to_numpy(OracleEngine.laplace_axpy_faces(dom, vbc, dv, 0.01, substeps=2), vbc, squeeze=False)
    for c in range(3):
        for b in range(2):
            want = O.diffuse_explicit(v[c][b], O.component_bcs(vbc, 3)[c], dx, 0.02, 1.0, substeps=2)
            np.testing.assert_allclose(out[c][b], want, rtol=0, atol=1e-6)
    # padding of the device arrays (beyond the stored faces) stays zero
    for c in range(3):
        total = float(np.abs(dv[c].numpy()).sum())
        assert abs(total - float(sum(np.abs(v[c][b]).sum() for b in range(2)))) < 1e-2


def test_smoke_plume_example_fused_equals_unfused(F, tmp_path):
    ex = example('smoke_plume')
    v1, s1, p1 = ex.main(res=32, steps=4, scene_dir=str(tmp_path), fused=True)
    v2, s2, p2 = ex.main(res=32, steps=4, scene_dir=None, fused=False)
    assert float(s1.numpy().sum()) > 0.5
    np.testing.assert_allclose(s1.numpy(), s2.numpy(), rtol=0, atol=1e-5)
    for a, b in zip(v1.numpy(), v2.numpy()):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-4)
    # the trajectory is a reference-format scene: sim_000000/smoke_000000.npz, velocity_000000.npz
    scene = F.Scene.list(str(tmp_path))[0]
    back = scene.read('smoke', frame=0)
    assert back.res == (32, 32) and os.path.isfile(os.path.join(scene.path, 'velocity_000000.npz'))


def test_lid_driven_cavity_example(F):
    v, p = example('lid_driven_cavity').main(steps=5, x=20, y=12)
    vx, vy = v.numpy()
    assert vx.shape == (19, 12) and vy.shape == (20, 11)
    assert float(vx[:, -1].mean()) > 0.05 and float(np.abs(F.field.divergence(v).numpy()).max()) < 1e-4


def test_batched_smoke_obstacle_example(F):
    v, s, p = example('batched_smoke_obstacle').main(res=24, steps=3)
    smoke = s.numpy()
    assert smoke.shape == (3, 24, 24)
    totals = smoke.reshape(3, -1).sum(1)
    assert totals[0] < totals[1] < totals[2]                      # inflow rates .1 < .2 < .3
    # no flux into the obstacle: faces inside it carry no velocity after the masked projection
    vx = v.numpy()[0]
    inside = F.Box(x=(35, 65), y=(50, 70)).lies_inside(v.face_points(0))
    assert float(np.abs(vx[:, inside]).max()) < 1e-5


def test_iterate_and_jit_compile_like_the_notebooks(F):
    """Lid_Driven_Cavity.ipynb as written: `@jit_compile def step(v, p, dt=1., viscosity=.1)` and
    `v_trj, p_trj = iterate(step, batch(time=4), v0, None)` (math.iterate, PhiML/phiml/math/_functional.py:1241-1300)."""
    @F.jit_compile
    def step(v, p, dt=1., viscosity=.1):
        v = F.advect.semi_lagrangian(v, v, dt)
        v = F.diffuse.explicit(v, viscosity, dt)
        v, p = F.fluid.make_incompressible(v, solve=F.Solve(x0=p))
        return v, p

    v0 = F.StaggeredGrid(0, {'x': 0, 'y-': 0, 'y+': F.vec(x=1, y=0)}, x=16, y=12)
    v_trj, p_trj = F.iterate(step, F.batch(time=4), v0, None)
    assert len(v_trj) == 5 and v_trj[0] is v0 and p_trj[0] is None
    v_end, p_end = F.iterate(step, 4, v0, None)
    for a, b in zip(v_trj[-1].numpy(), v_end.numpy()):
        np.testing.assert_array_equal(a, b)
    # substeps: two calls per recorded entry; keyword arguments reach the step
    v_sub, _ = F.iterate(step, F.batch(time=2), v0, None, substeps=2, dt=1.)
    for a, b in zip(v_sub[-1].numpy(), v_end.numpy()):
        np.testing.assert_array_equal(a, b)
    calls = []
    assert F.iterate(lambda x: calls.append(x) or x + 1, 3, 0, range=lambda n: range(n)) == 3 and calls == [0, 1, 2]
