"""CPU suite — host logic of depth-slab sharded whole-volume inference: slab partition, level
geometry, halo exchange and the small reductions, with world_size 2 over gloo and with the in-process
ThreadComm; no CUDA kernels involved (the slab arithmetic itself is covered by tests/test_gpu_slab.py)."""
import importlib
import os
import sys
import threading

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "3d-brain-tumor-segmentation_b200"


def test_slab_bounds(b3d):
    sb = b3d.slab_bounds
    assert [b - a for a, b in sb(160, 8)] == [24, 24, 24, 24, 16, 16, 16, 16]
    assert sb(160, 4) == [(0, 40), (40, 80), (80, 120), (120, 160)]
    assert sb(160, 1) == [(0, 160)]
    assert [b - a for a, b in sb(256, 8)] == [32] * 8                               # skull-strip volume
    for world in (1, 2, 3, 5, 8):
        b = sb(160, world)
        assert b[0][0] == 0 and b[-1][1] == 160 and all(p[1] == q[0] for p, q in zip(b, b[1:]))
        assert all(a % 8 == 0 for a, _ in b)
    with pytest.raises(ValueError):
        sb(150, 2)
    with pytest.raises(ValueError):
        sb(16, 4)


def _halo_reference(x, d0, d1, before, after):
    """slices [d0-before, d1+after) of the whole volume, zeros outside"""
    D = x.shape[1]
    out = torch.zeros((1, before + d1 - d0 + after) + tuple(x.shape[2:]), dtype=x.dtype)
    lo, hi = max(d0 - before, 0), min(d1 + after, D)
    out[:, lo - (d0 - before):hi - (d0 - before)] = x[:, lo:hi]
    return out


def _check_rank(b3d, comm, x, lvl_shapes):
    """Every rank: geometry per level, halos for the three conv kinds, reductions."""
    ctx = b3d.SlabContext(comm, x.shape[1])
    ok = True
    for lvl in range(3):
        xl = x[:, ::2 ** lvl].contiguous()                    # stand-in for the level-`lvl` activation
        loc = xl[:, ctx.d0 >> lvl:ctx.d1 >> lvl].contiguous()
        g0, dg = ctx.geometry(loc)
        ok &= (g0, dg) == (ctx.d0 >> lvl, x.shape[1] >> lvl)
        ok &= ctx.global_voxels(loc) == dg * x.shape[2] * x.shape[3]
        for before, after in ((1, 1), (0, 1), (1, 0)):
            got = ctx.with_halo(loc, before, after)
            ok &= torch.equal(got, _halo_reference(xl, ctx.d0 >> lvl, ctx.d1 >> lvl, before, after))
    t = torch.full((4,), float(comm.rank + 1), dtype=torch.float64)
    comm.all_reduce_sum(t)
    ok &= bool(torch.all(t == sum(range(1, comm.world + 1))))
    loc = x[:, ctx.d0:ctx.d1].contiguous()
    ok &= torch.equal(comm.all_gather_cat(loc, [b - a for a, b in ctx.bounds]), x)
    ok &= ctx.stats["halo_exchanges"] == 9
    return bool(ok)


def test_thread_comm_halos_and_reductions(b3d):
    x = torch.randn(1, 48, 3, 5, 4)
    for world in (1, 2, 3):
        comms = b3d.ThreadComm.make(world)
        res = [None] * world

        def work(r):
            res[r] = _check_rank(b3d, comms[r], x, None)

        ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert res == [True] * world


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    b3d = importlib.import_module(PKG)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 32, 3, 5, 4, generator=g)
    q.put((rank, _check_rank(b3d, b3d.DistComm(), x, None)))
    dist.destroy_process_group()


def test_dist_comm_world2_gloo(b3d):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_slab_context_rejects_foreign_shapes(b3d):
    ctx = b3d.SlabContext(b3d.ThreadComm.make(1)[0], 32)
    with pytest.raises(ValueError):
        ctx.geometry(torch.zeros(1, 12, 2, 2, 4))
    with pytest.raises(ValueError):
        ctx.geometry(torch.zeros(2, 32, 2, 2, 4))
