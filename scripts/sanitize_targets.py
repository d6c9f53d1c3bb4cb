"""Small instances of EVERY kernel family, for compute-sanitizer (memcheck / racecheck / synccheck) -- VERDICT r1 item 8:
lock-step DMMA rwalk (16 <= n <= 64), streamed DMMA rwalk (n > 64), warp rwalk (ncdim < n, periodic / reflective dims),
slice / rslice incl. doubling, unif (two ellipsoids), unit-cube, membership, bound construction (Cholesky candidates,
Jacobi ladder, k-means on clusters, bootstrap), the sliced eigensolver + its repair ladder (n = 130), improve_covar,
scale_to_logvol, device rounds (fused step kernel, all samplers, graph replay, device bound update), the peer exchange
with two contexts acting as two ranks on one device, RadFriends / SupFriends.
usage (GPU box): compute-sanitizer --tool memcheck python scripts/sanitize_targets.py"""
import math
import sys

sys.path.insert(0, '.')
import numpy as np

from dynesty_b200 import _lib, ops, likelihoods as DL, bounding as B, nested

rng = np.random.default_rng(5)


def cloud(N, n, s=0.05):
    return 0.5 + s * rng.standard_normal((N, n))


def chains(model, n, N, Q, sampler, steps, **kw):
    u = cloud(N, n)
    b = B.B200MultiEllipsoid(kw.pop('ncdim', n))
    b.update(u[:, :b.ndim], rstate=rng)
    b.scale_to_logvol(b.logvol + math.log(1.25))
    b.make_resident()
    _, l = model.evaluate(u)
    thr = float(np.quantile(l, 0.3))
    st = np.resize(u[l > thr], (Q, n))
    ell = b.random_ells(rng, Q)
    mid = model.model_id()
    if sampler == 'rwalk':
        return ops.rwalk_batch(mid, st, thr, 0.5, steps, 3, ell=ell, ncdim=b.ndim, **kw)
    fn = ops.rslice_batch if sampler == 'rslice' else ops.slice_batch
    return fn(mid, st, thr, 0.7, steps, 3, ell=ell, **kw)


print('rwalk'); sys.stdout.flush()
chains(DL.gauss_corr(20, 0.4, 5.0), 20, 200, 40, 'rwalk', 6)                              # lock-step DMMA
chains(DL.iid_normal_ppf(72), 72, 400, 24, 'rwalk', 4)                                      # streamed DMMA
chains(DL.gauss_corr(6, 0.4, 5.0), 6, 120, 20, 'rwalk', 6, ncdim=4, dimflags=np.array([1, 2, 0, 0, 0, 0], dtype=np.uint8))
print('slice'); sys.stdout.flush()
chains(DL.eggbox(5), 5, 200, 24, 'rslice', 3)
chains(DL.shells(4), 4, 200, 24, 'rslice', 3, doubling=True)
chains(DL.gauss_corr(4, 0.6, 5.0), 4, 120, 16, 'slice', 2)
print('unif / unitcube / membership'); sys.stdout.flush()
m2 = DL.shells(2)
pp = rng.random((4000, 2))
_, l2 = m2.evaluate(pp)
pp = pp[np.argsort(l2)[-300:]]
b2 = B.B200MultiEllipsoid(2)
b2.update(pp, rstate=rng, bootstrap=2)
b2.make_resident()
ops.unif_batch(m2.model_id(), 64, 2, float(np.sort(l2)[-300]), 1)
b2.samples(50, rstate=rng)
b2.monte_carlo_logvol(200, rstate=rng)
ops.unitcube_batch(DL.gauss_test3d().model_id(), 32, 3, -20.0, 1)
ops.membership(pp, b2.ctrs, b2.ams, want_d2=True)
print('bounds'); sys.stdout.flush()
ops.bounding_ellipsoid(cloud(500, 130, 0.02))                                               # sliced eigensolver
S = np.cov(cloud(300, 130), rowvar=False)
lam, V = np.linalg.eigh(S)
lam[:20] = 0.0
ops.improve_covar((V * lam) @ V.T)                                                          # its repair ladder
ops.improve_covar(np.diag([1., -1., 2.]))
two = np.concatenate([0.3 + 0.02 * rng.standard_normal((300, 8)), 0.7 + 0.02 * rng.standard_normal((300, 8))])
o = ops.multi_decompose(two)                                                                 # k-means, Cholesky candidates, ladder
e = B.B200Ellipsoid(8)
e.update(two, rstate=rng, bootstrap=2)
ops.moments(two)
print('device rounds'); sys.stdout.flush()
for sample, kw in (('rwalk', dict(walks=8)), ('rslice', dict(slices=3)), ('slice', dict(slices=1)), ('unif', {})):
    s = nested.NestedSampler(DL.gauss_corr(6, 0.4, 5.0), nlive=160, bound='multi', sample=sample, seed=3, **kw)
    s.run_nested(loop='device', batch=8, dlogz=None, maxiter=900)                            # > 16 rounds per call: graph replay
s = nested.NestedSampler(DL.gauss_corr(20, 0.4, 5.0), nlive=200, bound='single', sample='rwalk', walks=6, seed=3)
s.run_nested(loop='device', batch=10, dlogz=None, maxiter=1200)
print('peer'); sys.stdout.flush()
import torch
n, Q = 6, 24
half = Q // 2
ctxs = [_lib.Context(0), _lib.Context(0)]
wins = []
for c in ctxs:
    c.peer_export(c.peer_window_bytes(Q, n))
    wins.append(c.peer_result()[0])
m6 = DL.gauss_corr(n, 0.4, 5.0)
u = cloud(200, n)
bb = B.B200MultiEllipsoid(n)
bb.update(u, rstate=rng)
_, l = m6.evaluate(u)
thr = float(np.quantile(l, 0.3))
st = np.resize(u[l > thr], (Q, n))
d_u0 = torch.from_numpy(st).cuda()
mids = []
for r, c in enumerate(ctxs):
    c.peer_import_raw(r, 2, wins)
    bb.make_resident(c)
    mids.append(m6.model_id(c))
torch.cuda.synchronize()
for rep in range(2):
    for r, c in enumerate(ctxs):
        c.set_pointer_mode(_lib.PTR_DEVICE)
        ops.rwalk_batch(mids[r], d_u0[r * half:(r + 1) * half], thr, 0.5, 6, 9, chain0=r * half, ctx=c, out=ops.NO_OUT,
                        peer=(r * half, Q))
        c.set_pointer_mode(_lib.PTR_HOST)
    g = []
    for c in ctxs:
        c.peer_check()
        g.append(c.peer_gathered(Q, n, ['n_accept', 'n_reject', 'ncall']))
    assert np.array_equal(g[0]['u'], g[1]['u'])
for c in ctxs:
    c.close()
print('friends'); sys.stdout.flush()
for kind in ('balls', 'cubes'):
    f = (B.B200RadFriends if kind == 'balls' else B.B200SupFriends)(3)
    pts = np.concatenate([0.25 + 0.02 * rng.standard_normal((60, 3)), 0.75 + 0.02 * rng.standard_normal((60, 3))])
    f.update(pts)
    f.scale_to_logvol(f.logvol + math.log(1.25))
    f.update(np.ascontiguousarray(pts[::-1]), bootstrap=2, rstate=rng)
    f.overlap_many(pts[:10])
    f.samples(20, rstate=rng)
    s = nested.NestedSampler(DL.gauss_test3d(), nlive=60, bound=kind, sample='unif', queue_size=8, seed=2)
    s.run_nested(dlogz=None, maxiter=150)
print('sanitize targets done')
