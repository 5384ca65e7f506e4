"""
Host-side logic of the multi-GPU path on CPU: slab geometry, boundary rewriting and the halo exchange, run with
world_size 2 and 3 over gloo.  The kernels themselves are covered by tests/tools/dist_check.py on GPUs.
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


def test_advection_halo_rule():
    """h = ceil(max|v_z| dt / dz) + 1; the reference's back-trace is unbounded (advect.py:20-24)."""
    from phiflow_b200.dist import advection_halo
    assert advection_halo(0.0, 0.5, 1.0) == 1
    assert advection_halo(0.1, 0.5, 100.0 / 512) == 2            # 0.256 cells
    assert advection_halo(1.95, 0.5, 100.0 / 512) == 6           # 4.99 cells -> 5 + 1
    assert advection_halo(2.0, 0.5, 100.0 / 512) == 7            # 5.12 cells -> 6 + 1
    assert advection_halo(2.0, -0.5, 100.0 / 512) == 7           # MacCormack backward pass: |dt|
    assert advection_halo(0.0, 1.0, 1.0, bc_const_max=2.9) == 4   # a constant inflow boundary moves samples too
    assert advection_halo(3.0, 1.0, 1.0) == 5                    # exact integers round up (fp32 rounding of dt*v/dx in the kernel)
    with pytest.raises(ValueError):
        advection_halo(float('inf'), 0.5, 1.0)
    with pytest.raises(ValueError):
        advection_halo(float('nan'), 0.5, 1.0)


def _halo_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from phiflow_b200.dist import Slab, SlabPlume, HaloTooWide
        from phiflow_b200._lib import Unsupported
        res = (8, 6, 8 * world)
        vbc = (('periodic', 'periodic'),) * 3
        sbc = (('zg', 'zg'),) * 3
        slab = Slab(res, (1.0, 1.0, 2.0), vbc, halo=2, device='cpu')
        sim = SlabPlume(slab, sbc, 0.5, 0.2, (0.0, 0.0, 0.1), None)
        H, nz = slab.halo, slab.nz
        for zl in range(nz):
            sim.s[:, H + zl] = float(slab.z0 + zl)
            sim.v[2][:, H + zl] = 0.25
        # only rank 1 holds the fast cell: every rank must still agree on the width (all_reduce MAX)
        if rank == 1:
            sim.v[2][0, H + 3, 2, 2] = -13.0                    # 13 * 0.5 / 2 = 3.25 cells -> h = 5
        h = sim.required_halo()
        assert h == 5, h
        assert abs(sim.max_displacement - 3.25) < 1e-6
        sim.ensure_halo(h)                                       # 5 > 2 allocated -> regrow to max(5, 4) = 5
        assert sim.slab.halo == 5 and sim.regrown == 1 and sim.dom.cext[2] == nz + 10
        H2 = sim.slab.halo
        for zl in range(nz):                                     # owned planes survive the re-allocation
            assert float(sim.s[0, H2 + zl, 0, 0]) == float(slab.z0 + zl)
        assert float(sim.v[2][0, H2 + 3, 2, 2]) == (-13.0 if rank == 1 else 0.25)
        sim.slab.exchange(sim.v + [sim.s], h)
        gz = res[2]
        for k in range(1, h + 1):
            assert float(sim.s[0, H2 - k, 0, 0]) == (slab.z0 - k) % gz
            assert float(sim.s[0, H2 + nz - 1 + k, 0, 0]) == (slab.z0 + nz - 1 + k) % gz
        # wider than a slab: refused BEFORE anything is computed, as PHI_ERR_UNSUPPORTED
        if rank == 0:
            sim.v[2][0, H2, 0, 0] = 39.0                         # 9.75 cells -> h = 11 > 8 planes per rank
        h = sim.required_halo()
        assert h == 11
        try:
            sim.ensure_halo(h)
            raise AssertionError('HaloTooWide not raised')
        except HaloTooWide as err:
            assert isinstance(err, Unsupported) and err.code == -2 and '11' in str(err)
        with pytest.raises(ValueError):
            sim.slab.exchange([sim.s], sim.slab.halo + 1)
        q.put((rank, 'ok'))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_cfl_halo_regrow_and_refusal_gloo():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == 'ok', f"rank {rank}: {msg}"
