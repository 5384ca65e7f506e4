"""Scene directory layout (phi/field/_scene.py:23-24, 34-49, 107-153, 303-386) - host-side, with a stub field writer."""
import json
import os

import numpy as np
import pytest

from phiflow_b200.scene import Scene, field_filename, slugify


def _writer(fld, file):
    np.savez_compressed(file, data=np.asarray(fld))


def _reader(file):
    return np.load(file)['data']


def test_scene_ids_filenames_and_properties(tmp_path):
    parent = str(tmp_path / 'runs')
    s0 = Scene.create(parent, writer=_writer, reader=_reader)
    s1 = Scene.create(parent, writer=_writer, reader=_reader)
    assert os.path.basename(s0.path) == 'sim_000000' and os.path.basename(s1.path) == 'sim_000001'
    os.makedirs(os.path.join(parent, 'sim_000007'))
    assert os.path.basename(Scene.create(parent).path) == 'sim_000008'          # next id after the largest existing one
    assert [os.path.basename(s.path) for s in Scene.list(parent)] == ['sim_000000', 'sim_000001', 'sim_000007', 'sim_000008']
    assert Scene.at(parent, 1).path == s1.path
    with pytest.raises(IOError):
        Scene.at(parent, 99)
    # field files: <slug(name)>_<frame:06d>.npz
    assert os.path.basename(field_filename(s0.path, 'Velocity', 12)) == 'velocity_000012.npz'
    assert slugify('Smoke Density!') == 'smoke-density' and slugify('Φ field') == 'phi-field'
    s0.write({'velocity': np.arange(6.0).reshape(2, 3), 'smoke': np.ones(4)}, frame=0)
    s0.write(velocity=np.zeros((2, 3)), frame=3)
    assert s0.fieldnames == ('smoke', 'velocity')
    assert s0.frames == (0, 3) and s0.complete_frames == (0,)
    np.testing.assert_array_equal(s0.read('velocity', frame=0), np.arange(6.0).reshape(2, 3))
    v, sm = s0.read('velocity', 'smoke', frame=0)
    assert v.shape == (2, 3) and sm.shape == (4,)
    # description.json
    s0.put_properties(dt=0.5, resolution=[64, 64])
    s0.put_property('solver', 'CG')
    assert json.load(open(os.path.join(s0.path, 'description.json'))) == {'dt': 0.5, 'resolution': [64, 64], 'solver': 'CG'}
    assert Scene.at(s0.path).properties['dt'] == 0.5
    s1.remove()
    assert not s1.exists()
    with pytest.raises(RuntimeError):
        Scene.at(s0.path).write_field(np.zeros(2), 'x', 0)
