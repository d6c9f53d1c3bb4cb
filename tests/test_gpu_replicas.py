"""GPU: concurrent replicas (dynesty_b200/replicas.py) -- every replica owns a context (stream, scratch, resident
bound, device run state); running them concurrently must not change any of them."""
import numpy as np
import pytest

from dynesty_b200 import likelihoods as DL, nested, replicas, _lib

pytestmark = pytest.mark.gpu


def test_concurrent_replicas_equal_solo_runs():
    m = DL.gauss_corr(8, 0.4, 5.0)
    kw = dict(nlive=300, bound='multi', sample='rwalk', sampler_kwargs=dict(walks=20), dlogz=0.5, batch=10)
    seeds = list(range(40, 48))
    outs, wall = replicas.run_replicas(m, seeds, max_in_flight=8, **kw)
    solo, _ = replicas.run_replicas(m, seeds[2:4], max_in_flight=1, **kw)
    for a, b in zip(outs[2:4], solo):
        assert (a['logz'], a['ncall'], a['niter'], a['nbound']) == (b['logz'], b['ncall'], b['niter'], b['nbound'])
    ctx = _lib.Context(0)          # and the same run through the plain sampler API on a fresh context
    s = nested.NestedSampler(m, nlive=300, bound='multi', sample='rwalk', walks=20, seed=seeds[5], ctx=ctx)
    r = s.run_nested(loop='device', dlogz=0.5, batch=10)
    assert float(r.logz[-1]) == outs[5]['logz'] and r.ncall == outs[5]['ncall']
    ctx.close()
    lz = np.array([o['logz'] for o in outs])
    truth = m.logz_truth
    assert abs(lz.mean() - truth) < 4 * lz.std(ddof=1) / np.sqrt(len(lz)) + 0.3
    assert len({o['logz'] for o in outs}) == len(outs)        # different seeds, different runs


@pytest.mark.parametrize('sample,kw', [('rslice', dict(slices=5)), ('unif', {})])
def test_replicas_other_samplers(sample, kw):
    m = DL.gauss_test3d()
    outs, _ = replicas.run_replicas(m, [1, 2, 3, 4], nlive=200, bound='multi', sample=sample, sampler_kwargs=kw,
                                    max_in_flight=4, dlogz=0.5)
    lz = np.array([o['logz'] for o in outs])
    err = np.mean([o['logzerr'] for o in outs])
    assert abs(lz.mean() - m.logz_truth) < 3 * err / 2 + 0.1
