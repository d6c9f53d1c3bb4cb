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
