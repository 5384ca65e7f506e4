"""
CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol include/phicuda.h declares,
the ctypes mirrors have the C struct sizes, and argument validation reports errors (no compute calls without a GPU).
"""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from phiflow_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'phicuda.h')


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(phicuda_\w+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in phicuda.h but not exported"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == names
    assert lib.phicuda_abi_version() == 2


def test_struct_sizes_match_c():
    src = r'''
    #include <stdio.h>
    #include "phicuda.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(PhiGrid), sizeof(PhiBC), sizeof(PhiVBC), sizeof(PhiCgParams),
               sizeof(PhiCgResult), sizeof(PhiPlumeParams), sizeof(PhiLaunchInfo));
        return 0;
    }'''
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, 's.c')
        exe = os.path.join(tmp, 's')
        with open(c, 'w') as f:
            f.write(src)
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    mirrors = [_lib.PhiGrid, _lib.PhiBC, _lib.PhiVBC, _lib.PhiCgParams, _lib.PhiCgResult, _lib.PhiPlumeParams, _lib.PhiLaunchInfo]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_argument_validation_reports_errors_without_gpu():
    lib = _lib.load()
    g = _lib.PhiGrid()
    g.dim = 4
    assert lib.phicuda_cg_workspace_bytes(C.byref(g)) == 0
    assert 'dim' in _lib.last_error()
    g.dim, g.batch = 2, 1
    g.n[0], g.n[1], g.cext[0], g.cext[1], g.fext[0], g.fext[1] = 10, 8, 10, 8, 12, 8      # cext[0] not a multiple of 4
    g.dx[0] = g.dx[1] = 1.0
    assert lib.phicuda_cg_workspace_bytes(C.byref(g)) == 0
    assert 'multiples of 4' in _lib.last_error()
    g.cext[0] = 12
    assert lib.phicuda_cg_workspace_bytes(C.byref(g)) > 3 * 12 * 8 * 4
    bc = _lib.PhiBC()
    bc.lo[0], bc.hi[0] = _lib.BC_PERIODIC, _lib.BC_CONST
    code = lib.phicuda_laplace_f32(C.byref(g), C.byref(bc), None, None, None)
    assert code == _lib.ERR_INVALID and 'PERIODIC' in _lib.last_error()
    with pytest.raises(_lib.PhiCudaError):
        _lib.check(code)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from phiflow_b200 import _ops
    dom = _ops.Domain((8, 8), (1.0, 1.0), device='cpu')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _ops.laplace(dom, ((0.0, 0.0), (0.0, 0.0)), dom.alloc_centered())
