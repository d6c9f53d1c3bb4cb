"""GPU, >= 2 devices: the sharded path over NCCL (one process per GPU).  Chain ids are
global, so the 2-GPU nested-sampling run must be IDENTICAL to the 1-GPU run; skipped on a
single-GPU box (the CPU tier covers the same logic with gloo, tests/test_dist_gloo.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, outdir, fused=False, sample='rwalk'):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from dynesty_b200 import _lib, likelihoods as DL, nested
    from dynesty_b200.dist import Comm
    torch.cuda.set_device(rank)
    comm = None
    if world > 1:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
        comm = Comm()
    ctx = _lib.Context(rank)
    if comm is not None and fused:
        comm.attach_peer(ctx, 400, 10)       # gather fused into the chain kernels (NVLink peer windows)
    m = DL.gauss_corr(10, 0.4, 5.0)
    s = nested.NestedSampler(m, nlive=400, bound='multi', sample=sample, queue_size=400, seed=11, ctx=ctx, comm=comm)
    res = s.run_nested(dlogz=0.5)
    np.savez(os.path.join(outdir, 'r%d_w%d_%s%d.npz' % (rank, world, sample, fused)), logz=res.logz, logl=res.logl,
             ncall=res.ncall)
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize('fused,sample', [(False, 'rwalk'), (True, 'rwalk'), (True, 'rslice'), (True, 'unif')])
def test_two_gpu_run_identical_to_one_gpu(tmp_path, fused, sample):
    """fused=False: one NCCL all-gather per output array; fused=True: the chain kernels store
    finished chains into every rank's exchange window and synchronise in-kernel."""
    import torch.multiprocessing as mp
    mp.spawn(_run, args=(1, 0, str(tmp_path), fused, sample), nprocs=1, join=True)
    mp.spawn(_run, args=(2, _free_port(), str(tmp_path), fused, sample), nprocs=2, join=True)
    a, b0, b1 = (np.load(tmp_path / (f % (sample, fused))) for f in ('r0_w1_%s%d.npz', 'r0_w2_%s%d.npz', 'r1_w2_%s%d.npz'))
    for k in ('logz', 'logl', 'ncall'):
        assert np.array_equal(b0[k], b1[k])
        assert np.array_equal(a[k], b0[k])
    assert abs(a['logz'][-1] - (-10 * np.log(10.))) < 1.0


def _run_sharded_bound(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from dynesty_b200 import _lib, bounding as B, ops
    from dynesty_b200.dist import Comm
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    comm = Comm()
    ctx = _lib.Context(rank)
    rng = np.random.default_rng(21)
    pts = 0.5 + 0.04 * rng.standard_normal((8000, 200))                 # C4's shape: 8000 x 200, sliced eigensolver
    lo, hi = (len(pts) * rank) // world, (len(pts) * (rank + 1)) // world
    b = B.B200Ellipsoid(200, ctx=ctx)
    b.update_sharded(pts[lo:hi], comm)
    out = dict(ctr=b.ctr, cov=b.cov, am=b.am, logvol=b.logvol)
    if rank == 0:
        o = ops.bounding_ellipsoid(pts, ctx=ctx)                        # all rows on one GPU
        out.update(ctr1=o['ctr'], cov1=o['cov'], am1=o['am'], logvol1=o['logvol'])
    np.savez(os.path.join(outdir, 'sb%d.npz' % rank), **out)
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_sharded_single_ellipsoid(tmp_path):
    """SURVEY 8e, bound='single' (C4): rows sharded over the GPUs, NCCL all-reduce of the moments and of fmax -> the
    ellipsoid of ALL rows, identical on both ranks."""
    import torch.multiprocessing as mp
    mp.spawn(_run_sharded_bound, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / 'sb0.npz'), np.load(tmp_path / 'sb1.npz')
    for k in ('ctr', 'cov', 'am', 'logvol'):
        assert np.array_equal(a[k], b[k])
    np.testing.assert_allclose(a['ctr'], a['ctr1'], rtol=1e-12)
    np.testing.assert_allclose(a['cov'], a['cov1'], rtol=1e-9, atol=1e-12 * np.abs(a['cov1']).max())
    np.testing.assert_allclose(a['am'], a['am1'], rtol=1e-6, atol=1e-9 * np.abs(a['am1']).max())
    assert abs(float(a['logvol']) - float(a['logvol1'])) < 1e-7


def _run_replicas(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import json
    import torch
    import torch.distributed as dist
    from dynesty_b200 import likelihoods as DL, replicas
    from dynesty_b200.dist import Comm
    torch.cuda.set_device(rank)
    comm = None
    if world > 1:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
        comm = Comm()
    m = DL.gauss_corr(8, 0.4, 5.0)
    outs, wall = replicas.run_replicas(m, range(30, 36), nlive=300, bound='multi', sample='rwalk', sampler_kwargs=dict(walks=20),
                                       device=rank, max_in_flight=3, comm=comm, dlogz=0.5, batch=10)
    if rank == 0:
        with open(os.path.join(outdir, 'rep_w%d.json' % world), 'w') as f:
            json.dump([[o['seed'], o['logz'], o['ncall'], o['niter']] for o in outs], f)
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_replicas_equal_one_gpu(tmp_path):
    """Replicas shard over ranks with no data-path collective: the ensemble on 2 GPUs is the ensemble on 1 GPU."""
    import json
    import torch.multiprocessing as mp
    mp.spawn(_run_replicas, args=(1, 0, str(tmp_path)), nprocs=1, join=True)
    mp.spawn(_run_replicas, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a = json.load(open(tmp_path / 'rep_w1.json'))
    b = json.load(open(tmp_path / 'rep_w2.json'))
    assert a == b and len(a) == 6
