"""CPU tier: the N>1 path (chains of a queue fill sharded over ranks + all-gather) with
world_size=2 on the gloo backend.  Kernels are replaced by the oracle-backed stand-in
(tests/fake_backend.py); what is pinned is that the sharded run is IDENTICAL to the
single-rank run (global chain ids -> same Philox streams) and that every rank ends with the
same state."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, outdir, sample='rwalk'):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import fake_backend
    from dynesty_b200 import ops, likelihoods as DL, nested
    from dynesty_b200.dist import Comm

    class MP:                       # minimal monkeypatch object
        def setattr(self, obj, name, val):
            setattr(obj, name, val)
    fake_backend.install(MP())
    comm = None
    if world > 1:
        dist.init_process_group('gloo', rank=rank, world_size=world)
        comm = Comm()
    m = DL.gauss_test3d()
    s = nested.NestedSampler(m, nlive=60, bound='multi', sample=sample, walks=8, slices=3, queue_size=20, seed=5,
                             comm=comm)
    res = s.run_nested(dlogz=None, maxiter=250)
    np.savez(os.path.join(outdir, '%s_r%d_w%d.npz' % (sample, rank, world)), logz=res.logz, logl=res.logl,
             samples=res.samples, ncall=res.ncall)
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize('sample', ['rwalk', 'rslice', 'unif'])
def test_sharded_run_matches_single_rank(tmp_path, sample):
    """rslice / unif fills return an unsigned `flags` array: the all-gather must carry it."""
    import torch.multiprocessing as mp
    _run(0, 1, 0, str(tmp_path), sample)
    port = _free_port()
    mp.spawn(_run, args=(2, port, str(tmp_path), sample), nprocs=2, join=True)
    a = np.load(tmp_path / ('%s_r0_w1.npz' % sample))
    b0 = np.load(tmp_path / ('%s_r0_w2.npz' % sample))
    b1 = np.load(tmp_path / ('%s_r1_w2.npz' % sample))
    for k in ('logz', 'logl', 'samples', 'ncall'):
        assert np.array_equal(b0[k], b1[k])            # replicated host state
        assert np.array_equal(a[k], b0[k])             # sharding does not change the run


def test_comm_shard_ranges():
    from dynesty_b200.dist import Comm
    c = Comm.__new__(Comm)
    c.world, c.rank = 4, 2
    assert c.shard(20) == (10, 15)
    with pytest.raises(AssertionError):
        c.shard(18)
