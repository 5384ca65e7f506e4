"""
`Scene`: the reference's on-disk layout for simulation trajectories (phi/field/_scene.py:52-426; last part of row N4).

    <parent>/sim_000123/                      Scene.create(parent) picks the next free id       (_scene.py:107-153)
        description.json                      scene.put_properties(dt=0.5, ...)                 (_scene.py:303-352)
        <name>_<frame:06d>.npz                scene.write({'velocity': v, 'smoke': s}, frame=7) (_scene.py:354-386, 23-24)

One .npz per field and frame in the reference's field format (phiflow_b200/field_io.py, phi/field/_field_io.py:45-127), so a
trajectory written here opens with stock PhiFlow's `Scene.at(...).read(...)` and vice versa.  Single scenes only (the
reference's batched scenes - a Tensor of paths - are not needed for writing trajectories of the fast path).
Host-side and storage only: the field writer / reader is passed in (phiflow_b200.flow uses its own `write` / `read`).
"""
import json
import os
import re
import shutil
import sys
from typing import Callable, Dict, Optional, Tuple

_GREEK = {'α': 'alpha', 'β': 'beta', 'γ': 'gamma', 'δ': 'delta', 'ε': 'epsilon', 'λ': 'lambda', 'μ': 'mu', 'ν': 'nu', 'π': 'pi', 'ρ': 'rho',
          'σ': 'sigma', 'τ': 'tau', 'φ': 'phi', 'ω': 'omega', 'Φ': 'Phi', 'Δ': 'Delta'}


def slugify(value: str) -> str:
    """_scene.py:524-533: lower-case, alphanumerics, hyphens for spaces; greek letters spelled out."""
    for letter, name in _GREEK.items():
        value = value.replace(letter, name)
    value = re.sub(r'[^\w\s-]', '', value).strip().lower()
    return re.sub(r'[-\s]+', '-', value)


def _slugify_filename(name: str) -> str:
    name = name.replace('._', '.').replace('.', '_')                     # _scene.py:517-521
    return name[1:] if name.startswith('_') else name


def field_filename(scene_path: str, name: str, frame: int) -> str:
    return os.path.join(scene_path, f"{slugify(_slugify_filename(name))}_{int(frame):06d}.npz")      # _scene.py:23-24


class Scene:
    def __init__(self, path: str, writer: Optional[Callable] = None, reader: Optional[Callable] = None):
        self.path = path
        self._writer, self._reader = writer, reader
        self._properties: Optional[dict] = None

    # ---- creation / lookup -------------------------------------------------------------------------------------------------
    @staticmethod
    def create(parent_directory: str, name: str = 'sim', copy_calling_script: bool = False, writer=None, reader=None) -> 'Scene':
        parent = os.path.expanduser(parent_directory)
        if not os.path.isdir(parent):
            os.makedirs(parent)
            next_id = 0
        else:
            ids = [int(f[len(name) + 1:]) for f in os.listdir(parent) if f.startswith(f"{name}_") and f[len(name) + 1:].isdigit()]
            next_id = max([-1] + ids) + 1
        scene = Scene(os.path.join(parent_directory, f"{name}_{next_id:06d}"), writer, reader)
        os.makedirs(scene.path)
        if copy_calling_script:
            scene.copy_calling_script()
        return scene

    @staticmethod
    def at(directory: str, id: Optional[int] = None, writer=None, reader=None) -> 'Scene':
        path = os.path.join(directory, f"sim_{int(id):06d}") if id is not None else directory
        if not os.path.isdir(os.path.expanduser(path)):
            raise IOError(f"There is no scene at '{path}'")
        return Scene(path, writer, reader)

    @staticmethod
    def list(parent_directory: str, name: str = 'sim', writer=None, reader=None) -> Tuple['Scene', ...]:
        parent = os.path.expanduser(parent_directory)
        if not os.path.isdir(parent):
            return ()
        names = sorted(f for f in os.listdir(parent) if f.startswith(f"{name}_") and os.path.isdir(os.path.join(parent, f)))
        return tuple(Scene(os.path.join(parent_directory, f), writer, reader) for f in names)

    def exists(self) -> bool:
        return os.path.isdir(self.path)

    def remove(self):
        if self.exists():
            shutil.rmtree(self.path)

    def subpath(self, name: str, create: bool = False) -> str:
        path = os.path.join(self.path, name)
        if create and not os.path.isdir(path):
            os.makedirs(path)
        return path

    def copy_calling_script(self):
        """_scene.py:428-447: the script that started the run goes to <scene>/src/."""
        script = os.path.abspath(sys.argv[0]) if sys.argv and sys.argv[0] else None
        if script and os.path.isfile(script):
            shutil.copy(script, os.path.join(self.subpath('src', create=True), os.path.basename(script)))

    # ---- properties: description.json ------------------------------------------------------------------------------------------
    @property
    def properties(self) -> dict:
        if self._properties is None:
            file = os.path.join(self.path, 'description.json')
            self._properties = json.load(open(file)) if os.path.isfile(file) else {}
        return self._properties

    def put_properties(self, update: Optional[dict] = None, **kw):
        props = self.properties
        props.update(update or {})
        props.update(kw)
        with open(os.path.join(self.path, 'description.json'), 'w') as out:
            json.dump(props, out, indent=2)

    def put_property(self, key, value):
        self.put_properties({key: value})

    # ---- fields --------------------------------------------------------------------------------------------------------------
    def write(self, data: Optional[Dict[str, object]] = None, frame: int = 0, **kw_data):
        data = dict(data or {})
        data.update(kw_data)
        for name, fld in data.items():
            self.write_field(fld, name, frame)

    def write_field(self, fld, name: str, frame: int):
        if self._writer is None:
            raise RuntimeError("Scene has no field writer (create it through phiflow_b200.flow.Scene)")
        self._writer(fld, field_filename(self.path, name, frame))

    def read_field(self, name: str, frame: int):
        if self._reader is None:
            raise RuntimeError("Scene has no field reader (create it through phiflow_b200.flow.Scene)")
        return self._reader(field_filename(self.path, name, frame))

    def read(self, *names, frame: int = 0):
        if len(names) == 1 and isinstance(names[0], (tuple, list)):
            names = tuple(names[0])
        result = [self.read_field(n, frame) for n in names]
        return result[0] if len(names) == 1 else result

    @property
    def fieldnames(self) -> tuple:
        return tuple(sorted({f[:-11] for f in os.listdir(self.path) if f.endswith('.npz')}))          # _scene.py:34-36

    def _frames_of(self, fieldname: str) -> set:
        return {int(f[-10:-4]) for f in os.listdir(self.path) if f.startswith(fieldname + '_') and f.endswith('.npz')}

    @property
    def frames(self) -> tuple:
        sets = [self._frames_of(n) for n in self.fieldnames]
        return tuple(sorted(set().union(*sets))) if sets else ()

    @property
    def complete_frames(self) -> tuple:
        sets = [self._frames_of(n) for n in self.fieldnames]
        return tuple(sorted(set.intersection(*sets))) if sets else ()

    def __repr__(self):
        return self.path
