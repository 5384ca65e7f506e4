"""
Host-side logic of the multi-GPU path on CPU: slab geometry, boundary rewriting and the halo exchange, run with
world_size 2 and 3 over gloo.  The kernels themselves are covered by tools/dist_check.py on GPUs.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, periodic, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from phiflow_b200.dist import Slab
        res = (8, 6, 4 * world)
        zside = ('periodic', 'periodic') if periodic else (0.0, 'zg')
        vbc = (('periodic', 'periodic'), (0.0, 0.0), zside)
        slab = Slab(res, (1.0, 1.0, 1.0), vbc, halo=2, device='cpu')
        # boundary rewriting
        zlo, zhi = slab.vbc[2]
        assert zlo == ('halo' if (periodic or rank > 0) else 0.0)
        assert zhi == ('halo' if (periodic or rank < world - 1) else 'zg')
        assert slab.dom.cext[2] == slab.nz + 4 and slab.dom.grid.halo == 2
        # a global field whose value encodes the global plane index
        H, nz = slab.halo, slab.nz
        t = slab.dom.alloc_centered()
        f = slab.dom.alloc_faces()
        for zl in range(nz):
            t[:, H + zl] = float(slab.z0 + zl)
            for c in range(3):
                f[c][:, H + zl] = float(100 * c + slab.z0 + zl)
        slab.exchange([t] + f, 2)
        gz = res[2]
        for k in (1, 2):
            lo_expect = slab.z0 - k
            hi_expect = slab.z0 + nz - 1 + k
            if periodic:
                lo_expect %= gz; hi_expect %= gz
            if slab.lower is not None:
                assert float(t[0, H - k, 0, 0]) == lo_expect, (rank, k, float(t[0, H - k, 0, 0]), lo_expect)
                assert float(f[2][0, H - k, 1, 1]) == 200 + lo_expect
            else:
                assert float(t[0, H - k, 0, 0]) == 0.0            # physical boundary: halo untouched
            if slab.upper is not None:
                assert float(t[0, H + nz - 1 + k, 0, 0]) == hi_expect
                assert float(f[1][0, H + nz - 1 + k, 2, 3]) == 100 + hi_expect
            else:
                assert float(t[0, H + nz - 1 + k, 0, 0]) == 0.0
        # width-1 exchange leaves the outer halo plane alone
        t[:, H:H + nz] += 1000.0
        slab.exchange([t], 1)
        if slab.lower is not None:
            assert float(t[0, H - 1, 0, 0]) >= 1000.0 and float(t[0, H - 2, 0, 0]) < 1000.0
        q.put((rank, 'ok'))
    except Exception as err:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,periodic', [(2, True), (2, False), (3, True)])
def test_slab_halo_exchange_gloo(world, periodic):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, periodic, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == 'ok', f"rank {rank}: {msg}"


def test_batch_shard_covers_every_entry_once():
    """BASELINE configs[4]: batch=512 over 8 ranks -> 64 each; ragged batches are spread over the first ranks."""
    from phiflow_b200.dist import batch_shard
    assert [batch_shard(512, r, 8) for r in range(8)] == [(64 * r, 64) for r in range(8)]
    parts = [batch_shard(13, r, 4) for r in range(4)]
    assert parts == [(0, 4), (4, 3), (7, 3), (10, 3)]
    covered = [i for first, n in parts for i in range(first, first + n)]
    assert covered == list(range(13))
    assert batch_shard(5, 0, 1) == (0, 5)


def test_local_bc_single_rank_is_identity():
    from phiflow_b200.dist import local_bc
    spec = (('periodic', 'periodic'), (0.0, 'zg'), ('zg', 0.0))
    assert local_bc(spec, 0, 1) == spec
    assert local_bc(spec, 1, 4)[2] == ('halo', 'halo')
    assert local_bc(spec, 0, 4)[2] == ('zg', 'halo')
    assert local_bc(spec, 3, 4)[2] == ('halo', 0.0)
    assert local_bc([spec, spec, spec], 3, 4)[1][2] == ('halo', 0.0)
