"""
The `.npz` field format (phi/field/_field_io.py:45-127) - host-side, runs on CPU.
Pinned against fixtures recorded from the vendored PhiML (tests/golden/make_field_io_golden.py): extrapolation dictionaries,
dim-type strings, `valid_outer_faces` and the padding arithmetic of `Field.staggered_tensor`.
"""
import os

import numpy as np
import pytest

from phiflow_b200 import field_io
from oracle import oracle_np as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'phiml_field_io.npz'), allow_pickle=True)
NAMES = sorted(k.split('/', 1)[1] for k in GOLD.files if k.startswith('spec/'))


def _spec(name):
    return tuple((lo, hi) for lo, hi in GOLD[f'spec/{name}'].tolist())


def _sides(spec, axes):
    return {(a, up): (s if isinstance(s, str) else float(s)) for a, lohi in zip(axes, spec) for up, s in zip((False, True), lohi)}


def _same_dict(a, b):
    if isinstance(a, dict):
        return isinstance(b, dict) and a.keys() == b.keys() and all(_same_dict(a[k], b[k]) for k in a)
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        a, b = np.asarray(a), np.asarray(b)
        return a.shape == b.shape and a.dtype.kind == b.dtype.kind and np.array_equal(a, b)
    return a == b


@pytest.mark.parametrize('name', NAMES)
def test_extrapolation_dict_matches_reference(name):
    spec = _spec(name)
    axes = 'xyz'[:len(spec)]
    ref = GOLD[f'ext/{name}'][()]
    got = field_io.extrapolation_to_dict(None, _sides(spec, axes), axes)
    assert _same_dict(got, ref), (got, ref)
    default, sides = field_io.extrapolation_from_dict(ref)
    back = {(a, up): sides.get((a, up), default) for a in axes for up in (False, True)}
    assert back == _sides(spec, axes)


def test_constant_extrapolation_dicts():
    assert _same_dict(field_io.extrapolation_to_dict(1.0, {}, 'xy'), GOLD['ext/one'][()])
    assert _same_dict(field_io.extrapolation_to_dict(0.5, {}, 'xy'), GOLD['ext/half'][()])
    assert _same_dict(field_io.extrapolation_to_dict((1.0, 0.0), {}, 'xy'), GOLD['ext/vec'][()])
    assert field_io.extrapolation_from_dict(GOLD['ext/vec'][()]) == ((1.0, 0.0), {})


def test_dim_type_strings():
    assert tuple(GOLD['shape/names']) == ('batch', 'x', 'y', 'vector')
    assert tuple(GOLD['shape/types']) == ('batch', 'spatial', 'spatial', 'channel')      # what write() stores in dim_types


@pytest.mark.parametrize('name', NAMES)
def test_staggered_tensor_matches_reference_and_oracle(name):
    spec = _spec(name)
    d = len(spec)
    comps = [GOLD[f'comp/{name}/{c}'] for c in range(d)]
    ref = GOLD[f'staggered_tensor/{name}']
    res = tuple(int(n) for n in GOLD[f'res/{name}'])
    assert ref.shape == tuple(n + 1 for n in res) + (d,)
    got = field_io.staggered_tensor(comps, lambda ax: spec[ax], d)
    np.testing.assert_array_equal(got, ref)
    # stored face counts agree with the oracle's restatement of valid_outer_faces (extrapolation.py:57-62)
    shapes = O.staggered_shapes(res, spec)
    assert [c.shape for c in comps] == [tuple(s) for s in shapes]
    back = field_io.unstack_staggered_tensor(ref, lambda ax: spec[ax], d)
    for a, b in zip(back, comps):
        np.testing.assert_array_equal(a, b)


def test_batched_staggered_tensor_round_trip():
    spec = _spec('mixed2')
    rng = np.random.default_rng(1)
    comps = [rng.standard_normal((3,) + GOLD[f'comp/mixed2/{c}'].shape).astype(np.float32) for c in range(2)]
    st = field_io.staggered_tensor(comps, lambda ax: spec[ax], 2)
    assert st.shape == (3, 7, 5, 2)
    for b in range(3):
        np.testing.assert_array_equal(st[b], field_io.staggered_tensor([c[b] for c in comps], lambda ax: spec[ax], 2))
    for a, b in zip(field_io.unstack_staggered_tensor(st, lambda ax: spec[ax], 2), comps):
        np.testing.assert_array_equal(a, b)


def test_archive_keys_and_round_trip(tmp_path):
    spec = _spec('mixed3')
    comps = [GOLD[f'comp/mixed3/{c}'] for c in range(3)]
    data = field_io.staggered_tensor(comps, lambda ax: spec[ax], 3)
    ext = field_io.extrapolation_to_dict(None, _sides(spec, 'xyz'), 'xyz')
    f = str(tmp_path / 'velocity')
    field_io.write_single_field(f, 'StaggeredGrid', data, ('x', 'y', 'z', 'vector'), ('spatial',) * 3 + ('channel',),
                                (None, None, None, ('x', 'y', 'z')), (0, 0, 0), (4, 3, 5), 'xyz', ext)
    with np.load(f + '.npz', allow_pickle=True) as st:
        # exactly the keys of phi/field/_field_io.py:57-66
        assert sorted(st.files) == sorted(['dim_names', 'dim_types', 'dim_item_names', 'field_type', 'lower', 'upper',
                                           'bounds_item_names', 'extrapolation', 'data'])
        assert str(st['field_type']) == 'StaggeredGrid' and st['dim_item_names'][3] == ('x', 'y', 'z')
        assert _same_dict(st['extrapolation'][()], GOLD['ext/mixed3'][()])
    back = field_io.read_single_field(f)
    assert back['field_type'] == 'StaggeredGrid' and back['dim_names'] == ('x', 'y', 'z', 'vector')
    assert back['lower'] == {'x': 0.0, 'y': 0.0, 'z': 0.0} and back['upper'] == {'x': 4.0, 'y': 3.0, 'z': 5.0}
    np.testing.assert_array_equal(back['data'], data)
    with pytest.raises(NotImplementedError):
        field_io.extrapolation_from_dict({'type': 'symmetric'})


def test_flow_write_read_on_cpu_containers(tmp_path):
    """flow.write / flow.read (field.write / field.read of the reference) with fields held on the CPU: containers only."""
    from phiflow_b200 import flow as F
    F.set_device('cpu')
    try:
        rng = np.random.default_rng(3)
        b = F.combine_sides(x=F.PERIODIC, y=(F.ZERO, F.ZERO_GRADIENT))
        v = F.StaggeredGrid([rng.standard_normal(s).astype(np.float32) for s in [(6, 4), (6, 4)]], b, F.Box(x=(0, 3), y=(1, 2)), x=6, y=4)
        F.write(v, str(tmp_path / 'v'))
        v2 = F.read(str(tmp_path / 'v'))
        assert isinstance(v2, F.StaggeredGrid) and v2.boundary == v.boundary and v2.res == v.res
        assert v2.lower == v.lower and v2.upper == v.upper
        for a, c in zip(v.numpy(), v2.numpy()):
            np.testing.assert_array_equal(a, c)
        s = F.CenteredGrid(rng.standard_normal((2, 6, 4)).astype(np.float32), F.ZERO_GRADIENT, F.Box(x=3, y=2), batch=2, x=6, y=4)
        F.write(s, str(tmp_path / 's.npz'))
        with np.load(str(tmp_path / 's.npz'), allow_pickle=True) as st:
            assert tuple(st['dim_names']) == ('batch', 'x', 'y') and tuple(st['dim_types']) == ('batch', 'spatial', 'spatial')
        s2 = F.read(str(tmp_path / 's.npz'))
        assert isinstance(s2, F.CenteredGrid) and s2.batch == 2 and s2.boundary == s.boundary
        np.testing.assert_array_equal(s.numpy(), s2.numpy())
        import torch
        if not torch.cuda.is_available():
            with pytest.raises(RuntimeError):                 # kernels still need the GPU: no CPU fallback
                F.field.laplace(s2)
    finally:
        F.set_device('cuda')
