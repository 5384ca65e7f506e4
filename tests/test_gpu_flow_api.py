"""
The reference's own hot-path tests re-run against `phiflow_b200.flow` - the user-facing mirror of
`phi.flow` for this path.  Each test names the reference test it follows.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from phiflow_b200.flow import *  # noqa: F401,F403
    from phiflow_b200.flow import fluid, advect, diffuse, field, math, extrapolation


def _boundaries():
    return {'ZERO': ZERO, 'BOUNDARY': BOUNDARY, 'PERIODIC': PERIODIC,
            'mixed': combine_sides(x=BOUNDARY, y=(ZERO, BOUNDARY))}


@pytest.mark.parametrize('name', ['ZERO', 'BOUNDARY', 'PERIODIC', 'mixed'])
def test_make_incompressible_staggered(name):
    """tests/commit/physics/test_fluid.py:19-53 (_test_make_incompressible, StaggeredGrid cases)."""
    ext = _boundaries()[name]
    bounds = Box(x=(0, 100), y=(0, 100))
    smoke = CenteredGrid(Sphere(x=50, y=10, radius=5), ext, bounds, x=16, y=20)
    velocity = StaggeredGrid(0, ext, bounds, x=16, y=20)
    for _ in range(2):
        velocity += resample(smoke * (0, 0.1), to=velocity)
        velocity, pressure = fluid.make_incompressible(velocity)
    div = field.divergence(velocity).numpy()
    assert np.abs(div).max() < 5e-5
    assert np.abs(np.concatenate([c.ravel() for c in velocity.numpy()])).max() > 1e-4      # the plume actually moves


def test_make_incompressible_batched():
    """tests/commit/physics/test_fluid.py:38-40 (batch3=3, batch2=2 -> 6 independent systems)."""
    bounds = Box(x=(0, 100), y=(0, 100))
    rng = np.random.default_rng(0)
    masks = np.stack([Sphere(x=float(cx), y=10, radius=5).lies_inside(CenteredGrid(0, ZERO, bounds, x=16, y=20).points()).astype(np.float32)
                      for cx in rng.uniform(0, 100, 6)])
    smoke = CenteredGrid(masks, ZERO, bounds, batch=6, x=16, y=20)
    velocity = StaggeredGrid(0, ZERO, bounds, batch=6, x=16, y=20)
    for _ in range(2):
        velocity += resample(smoke * (0, 0.1), to=velocity)
        velocity, _ = fluid.make_incompressible(velocity)
    assert np.abs(field.divergence(velocity).numpy()).max() < 5e-5
    # batch entries are independent: entry 0 equals the un-batched run
    smoke0 = CenteredGrid(masks[0], ZERO, bounds, x=16, y=20)
    v0 = StaggeredGrid(0, ZERO, bounds, x=16, y=20)
    for _ in range(2):
        v0 += resample(smoke0 * (0, 0.1), to=v0)
        v0, _ = fluid.make_incompressible(v0)
    for a, b in zip(velocity.numpy(), v0.numpy()):
        np.testing.assert_allclose(a[0], b, atol=1e-5)


def test_advection_identities():
    """tests/commit/physics/test_advect.py:12-30: adv(s, v, 0) == adv(s, 0*v, 1) == s for centred and staggered fields."""
    rng = np.random.default_rng(1)
    v = StaggeredGrid([rng.standard_normal((3, 3)).astype(np.float32), rng.standard_normal((4, 2)).astype(np.float32)], ZERO, x=4, y=3)
    s = CenteredGrid(rng.standard_normal((4, 3)).astype(np.float32), ZERO_GRADIENT, x=4, y=3)
    for adv in (advect.advect, advect.semi_lagrangian, advect.mac_cormack):
        np.testing.assert_allclose(adv(s, v, 0).numpy(), s.numpy(), atol=1e-5)
        np.testing.assert_allclose(adv(s, v * 0, 1).numpy(), s.numpy(), atol=1e-5)
    for adv in (advect.advect, advect.semi_lagrangian):
        for a, b in zip(adv(v, v, 0).numpy(), v.numpy()):
            np.testing.assert_allclose(a, b, atol=1e-5)
        for a, b in zip(adv(v, v * 0, 1).numpy(), v.numpy()):
            np.testing.assert_allclose(a, b, atol=1e-5)


def test_self_advect_staggered():
    """tests/commit/physics/test_advect.py:41-45 (known answer)."""
    v0 = StaggeredGrid(Box(x=(.9, 2.6), y=(.9, 2)), 0, x=4, y=3) * (0, 1)
    v = advect.semi_lagrangian(v0, v0, 1)
    np.testing.assert_allclose(v['x'].numpy(), 0, atol=1e-6)
    np.testing.assert_allclose(v['y'].numpy().T, [[0, 0, 0, 0], [0, 1, 1, 0]], atol=1e-6)


def test_staggered_grid_sizes_by_extrapolation():
    """tests/commit/field/test__grid.py:25-37"""
    s = StaggeredGrid(0, ZERO, x=20, y=10)
    assert [c.shape for c in s.numpy()] == [(19, 10), (20, 9)]
    s = StaggeredGrid(0, PERIODIC, x=20, y=10)
    assert [c.shape for c in s.numpy()] == [(20, 10), (20, 10)]
    s = StaggeredGrid(0, BOUNDARY, x=20, y=10)
    assert [c.shape for c in s.numpy()] == [(21, 10), (20, 11)]


def test_staggered_grid_with_extrapolation():
    """tests/commit/field/test__grid.py:85-94: BOUNDARY -> ZERO -> BOUNDARY re-pads the boundary faces with 0."""
    rng = np.random.default_rng(2)
    grid = StaggeredGrid([rng.standard_normal((21, 10)).astype(np.float32), rng.standard_normal((20, 11)).astype(np.float32)], BOUNDARY, x=20, y=10)
    grid_0 = grid.with_extrapolation(ZERO)
    assert [c.shape for c in grid_0.numpy()] == [(19, 10), (20, 9)]
    grid_ = grid_0.with_extrapolation(BOUNDARY)
    vx, vy = grid_.numpy()
    assert np.all(vx[0] == 0) and np.all(vx[-1] == 0) and np.all(vy[:, 0] == 0) and np.all(vy[:, -1] == 0)
    np.testing.assert_array_equal(vx[1:-1], grid.numpy()[0][1:-1])


def test_explicit_diffusion_known_answer():
    """tests/commit/physics/test_diffuse.py:68-72: stencil [[0,1,0],[1,-3,1],[0,1,0]]."""
    grid = CenteredGrid(np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], np.float32), 0, x=3, y=3)
    result = diffuse.explicit(grid, 1, 1).numpy()
    np.testing.assert_allclose(result, [[0, 1, 0], [1, -3, 1], [0, 1, 0]], atol=1e-6)


def test_divergence_and_laplace_fields():
    """field.laplace ghost cells / boundaries (PhiML/tests/commit/math/test__nd.py:26-67 shape & interior checks)."""
    a = np.arange(12, dtype=np.float32).reshape(4, 3) ** 2
    lap = field.laplace(CenteredGrid(a, ZERO_GRADIENT, x=4, y=3)).numpy()
    assert lap.shape == (4, 3)
    np.testing.assert_allclose(lap[1:-1, 1], (a[:-2, 1] + a[2:, 1] - 2 * a[1:-1, 1]) + (a[1:-1, 0] + a[1:-1, 2] - 2 * a[1:-1, 1]), atol=1e-4)


def test_not_converged_and_suppress_and_tape():
    """PhiML/tests/commit/math/test__optimize.py:95-143: SolveTape contract, NotConverged raised unless suppressed."""
    rng = np.random.default_rng(3)
    v = StaggeredGrid([rng.standard_normal((33, 32)).astype(np.float32), rng.standard_normal((32, 33)).astype(np.float32)], BOUNDARY, x=32, y=32)
    with pytest.raises(NotConverged):
        fluid.make_incompressible(v, solve=Solve('CG', 1e-6, 0, max_iterations=3))
    solve = Solve('CG', 1e-6, 0, max_iterations=3, suppress=[NotConverged])
    with math.SolveTape() as solves:
        v2, p = fluid.make_incompressible(v, solve=solve)
    assert int(solves[solve].iterations[0]) == 3 and not solves[solve].converged[0]
    with pytest.raises(NotImplementedError):
        fluid.make_incompressible(v, solve=Solve('biCG-stab(2)'))


def test_incompressible_step_matches_sequenced_calls():
    """incompressible_step == the notebook step written with the individual functions (Smoke_Plume.ipynb:58-68)."""
    domain = Box(x=100, y=100)
    inflow = resample(Sphere(x=50, y=9.5, radius=5), to=CenteredGrid(0, ZERO_GRADIENT, domain, x=32, y=32), soft=True)
    v = StaggeredGrid(0, 0, domain, x=32, y=32)
    s = CenteredGrid(0, ZERO_GRADIENT, domain, x=32, y=32)
    p = None
    v1, s1, p1 = v, s, None
    for _ in range(3):
        v, s, p = fluid.incompressible_step(v, s, p, 0.5, inflow, 0.2, (0, 0.1), Solve('CG', 1e-3, x0=None), smoke_advection='mac_cormack')
        s1 = advect.mac_cormack(s1, v1, 0.5) + inflow * 0.2
        v1 = advect.semi_lagrangian(v1, v1, 0.5) + resample(s1 * (0, 0.1), to=v1) * 0.5
        v1, p1 = fluid.make_incompressible(v1, (), Solve('CG', 1e-3, x0=p1))
    np.testing.assert_allclose(s.numpy(), s1.numpy(), atol=1e-5)
    for a, b in zip(v.numpy(), v1.numpy()):
        np.testing.assert_allclose(a, b, atol=1e-5)
    assert s.numpy().max() > 0.1


def test_make_incompressible_with_obstacles():
    """Obstacles through the user-facing API (examples/grids/Batched_Smoke: closed box with a Box obstacle)."""
    bounds = Box(x=(0, 100), y=(0, 100))
    obstacle = Box(x=(40, 60), y=(40, 50))
    smoke = CenteredGrid(Sphere(x=50, y=20, radius=8), ZERO_GRADIENT, bounds, x=32, y=32)
    velocity = StaggeredGrid(0, ZERO, bounds, x=32, y=32)
    for _ in range(3):
        velocity = advect.semi_lagrangian(velocity, velocity, 1.0) + resample(smoke * (0, 0.5), to=velocity)
        velocity, pressure = fluid.make_incompressible(velocity, [obstacle], Solve('CG', 1e-5, 1e-6))
    vx, vy = velocity.numpy()
    # faces well inside the obstacle carry no flow
    fy = velocity.face_points(1)
    inside = (fy[..., 0] > 43) & (fy[..., 0] < 57) & (fy[..., 1] > 43) & (fy[..., 1] < 47)
    assert inside.any() and np.abs(vy[inside]).max() < 1e-6
    # the fluid cells are divergence-free and the plume flows around the obstacle
    centres = CenteredGrid(0, ZERO, bounds, x=32, y=32).points()
    fluid_cells = ~obstacle.lies_inside(centres)
    div = field.divergence(velocity).numpy()
    assert np.abs(div[fluid_cells]).max() < 1e-4
    assert np.abs(vy).max() > 1e-2 and np.abs(vx).max() > 1e-3
    assert np.abs(pressure.numpy()[~fluid_cells]).max() == 0.0


def test_scene_trajectory_roundtrip(tmp_path):
    """Scene.create / write / read (phi/field/_scene.py:107-153, 354-426): a short smoke trajectory on disk, read back bit-exact."""
    bounds = Box(x=(0, 100), y=(0, 100))
    smoke = CenteredGrid(Sphere(x=50, y=20, radius=8), ZERO_GRADIENT, bounds, x=24, y=20)
    velocity = StaggeredGrid(0, ZERO, bounds, x=24, y=20)
    scene = Scene.create(str(tmp_path), copy_calling_script=False)
    scene.put_properties(dt=1.0, solver='CG')
    frames = {}
    for frame in range(3):
        velocity = advect.semi_lagrangian(velocity, velocity, 1.0) + resample(smoke * (0, 0.5), to=velocity)
        velocity, _ = fluid.make_incompressible(velocity, (), Solve('CG', 1e-4, 1e-6))
        smoke = advect.mac_cormack(smoke, velocity, 1.0)
        scene.write({'smoke': smoke, 'velocity': velocity}, frame=frame)
        frames[frame] = (smoke.numpy().copy(), [c.copy() for c in velocity.numpy()])
    assert scene.fieldnames == ('smoke', 'velocity') and scene.frames == (0, 1, 2) and scene.complete_frames == (0, 1, 2)
    again = Scene.at(scene.path)
    assert again.properties == {'dt': 1.0, 'solver': 'CG'}
    for frame, (s_ref, v_ref) in frames.items():
        s, v = again.read('smoke', 'velocity', frame=frame)
        assert isinstance(s, CenteredGrid) and isinstance(v, StaggeredGrid)
        np.testing.assert_array_equal(s.numpy(), s_ref)
        for a, b in zip(v.numpy(), v_ref):
            np.testing.assert_array_equal(a, b)
        assert v.boundary == velocity.boundary and s.boundary == smoke.boundary
