"""
Multi-GPU parity under pytest: runs tests/tools/dist_check.py on 2 GPUs when the box has them (one process per GPU, NCCL).
Covers: distributed CG == single-GPU CG, z-slab plume steps == single-GPU steps, and the CFL-derived advection halo
(displacement > 4 cells, halo re-allocation on the way) against the ORACLE.
"""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
def test_dist_check_two_gpus():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'tools', 'dist_check.py')]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and 'DIST_CHECK PASS' in out.stdout, out.stdout[-4000:] + out.stderr[-4000:]
