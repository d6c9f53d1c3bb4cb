"""CPU tier: oracle/friends.py (RadFriends / SupFriends restated) against tests/golden/friends.npz, which the
UNMODIFIED reference generated (oracle/make_golden.py gen_friends; its draws replayed on the Philox stream).
Groundwork for SURVEY.md 8(f) row 3 -- there is no CUDA counterpart yet."""
import os

import numpy as np
import pytest

from oracle import friends as F, philox
from conftest import GOLDEN

SEED = 56432


@pytest.fixture(scope='module')
def g():
    return np.load(os.path.join(GOLDEN, 'friends.npz'))


def close(a, b, rtol=1e-9):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize('cname', ['blob', 'two'])
@pytest.mark.parametrize('kind', ['balls', 'cubes'])
def test_friends_update_and_queries(g, cname, kind):
    p = 'fr_%s_%s_' % (cname, kind)
    pts = g[p + 'points']
    n = pts.shape[1]
    b = F.Friends(n, kind)
    for rep in (1, 2):
        sub = pts if rep == 1 else pts[::-1][:len(pts) - 10]
        b.update(sub)                                             # leave-one-out radius
        q = p + 'u%d_' % rep
        close(b.cov, g[q + 'cov'])
        close(b.am, g[q + 'am'], rtol=1e-7)
        close(b.axes, g[q + 'axes'], rtol=1e-8)
        close(b.axes_inv, g[q + 'axes_inv'], rtol=1e-7)
        assert abs(b.logvol - float(g[q + 'logvol'])) < 1e-8
        b.scale_to_logvol(b.logvol + np.log(1.25))                # the Sampler's enlarge step (sampler.py:506-508)
    xs = g[p + 'query']
    assert np.array_equal(np.array([b.overlap(x) for x in xs]), g[p + 'overlap'])
    assert np.array_equal(np.array([b.contains(x) for x in xs]), g[p + 'contains'])
    pt = np.dot(b.ctrs, b.axes_inv)
    close(F.loo_radius(pt, kind), g[p + 'loo'], rtol=1e-7)
    boot = [F.bootstrap_radius(pt, kind, philox.ChainStream(SEED, 400 + r).integers(len(pt), len(pt))) for r in range(3)]
    close(np.array(boot), g[p + 'boot'], rtol=1e-7)
    draws, qs = [], []
    for c in range(30):
        draws.append(b.sample(philox.ChainStream(SEED, 500 + c)))
        x, qq = b.sample(philox.ChainStream(SEED, 600 + c), return_q=True)
        draws.append(x)
        qs.append(qq)
    close(np.array(draws), g[p + 'draws'], rtol=1e-8)
    assert np.array_equal(np.array(qs), g[p + 'draw_q'])
    assert all(b.contains(x) for x in draws)
    b.scale_to_logvol(b.logvol + 0.3)
    close(b.am, g[p + 'scaled_am'], rtol=1e-7)
    close(b.axes, g[p + 'scaled_axes'], rtol=1e-8)


def test_clusters_are_connected_components():
    rng = np.random.default_rng(3)
    pts = np.concatenate([0.2 + 0.01 * rng.standard_normal((40, 2)), 0.8 + 0.01 * rng.standard_normal((35, 2))])
    am = np.linalg.inv(np.cov(pts[:40], rowvar=False)) / 9.0       # distance 1 = 3 sigma of one blob
    lab = F.components_within(pts, am)
    assert lab.max() == 1 and len(set(lab[:40])) == 1 and len(set(lab[40:])) == 1 and lab[0] != lab[-1]
    try:
        from scipy import cluster, spatial
    except ImportError:
        return
    ref = cluster.hierarchy.fcluster(cluster.hierarchy.single(spatial.distance.pdist(pts, 'mahalanobis', VI=am)), 1.0,
                                     criterion='distance')
    assert len(set(zip(lab, ref))) == 2                            # same partition, labels aside
