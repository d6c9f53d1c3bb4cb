"""TEST-ONLY stand-in for ``dynesty_b200.ops`` built from the oracle.

The product has no CPU path.  To exercise the HOST logic (``dynesty_b200.nested``, the
Bound / InternalSampler mirrors, the dynesty drop-in seams, the multi-rank sharding) in
the CPU-only test tier, the fixture ``fake_ops`` monkeypatches the array-level functions
of ``dynesty_b200.ops`` with oracle-backed equivalents that honour the same contract
(same arguments, same outputs, same Philox streams).  Nothing outside tests/ imports this.
"""
import numpy as np

from oracle import bounding as OB, samplers as OS, philox, likelihoods as OL

_state = {}
_models = []


def _to_oracle_model(dm, prior_kind):
    p = {}
    if prior_kind == OL.PRIOR_UNIFORM:
        p.update(lo=dm.prior_p0, width=dm.prior_p1)
    elif prior_kind == OL.PRIOR_NORMAL_PPF:
        p.update(mu=dm.prior_p0, sigma=dm.prior_p1)
    k = dm.like_kind
    if k == OL.LIKE_GAUSS_PREC:
        p.update(mean=dm.like_vec0, prec=dm.like_mat, lnorm=dm.s[0])
    elif k == OL.LIKE_GAUSS_DIAG:
        p.update(mean=dm.like_vec0, ivar=dm.like_vec1, lnorm=dm.s[0])
    elif k == OL.LIKE_EGGBOX:
        p.update(tmax=dm.s[0], power=dm.s[1])
    elif k == OL.LIKE_REGION2D:
        p.update(shape=int(dm.s[0]))
    else:
        p.update(c1=dm.like_vec0, c2=dm.like_vec1, r=dm.s[0], w=dm.s[1])
    return OL.Model(dm.ndim, prior_kind, k, **p)


def model_ids(self, ctx=None):
    if not hasattr(self, '_fake_ids'):
        _models.append(_to_oracle_model(self, self.prior_kind))
        _models.append(_to_oracle_model(self, OL.PRIOR_IDENTITY))
        self._fake_ids = (len(_models) - 2, len(_models) - 1)
    return self._fake_ids


def model_eval(model, u, want_v=True, ctx=None):
    m = _models[model]
    u = np.atleast_2d(np.asarray(u, dtype=float))
    v = m.prior_transform(u)
    return (v if want_v else None), np.asarray(m.loglike(v), dtype=float).reshape(len(u))


def membership(x, ctrs, ams, strict=True, want_d2=False, ctx=None):
    x = np.atleast_2d(np.asarray(x, dtype=float))
    ctrs = np.atleast_2d(ctrs)
    ams = np.asarray(ams).reshape(len(ctrs), ctrs.shape[1], ctrs.shape[1])
    d = x[:, None, :] - ctrs[None]
    d2 = np.einsum('mki,kij,mkj->mk', d, ams, d)
    mask = d2 < 1 if strict else d2 <= 1
    out = (mask, mask.sum(1).astype(np.int32))
    return out + (d2,) if want_d2 else out


def _ell_out(e):
    lam, vec = np.linalg.eigh(e.cov)
    return dict(ctr=e.ctr, cov=e.cov, am=e.am, axes=vec * np.sqrt(lam), axlens=np.sqrt(lam),
                logvol=float(e.logvol), warn=0)


def bounding_ellipsoid(points, ctx=None):
    return _ell_out(OB.bounding_ellipsoid(points))


def multi_decompose(points, max_ells=None, ctx=None):
    me, members = OB.multi_update(points)
    outs = [_ell_out(e) for e in me.ells]
    labels = np.empty(len(points), dtype=np.int32)
    for k, m in enumerate(members):
        labels[m] = k
    return dict(nells=me.nells, labels=labels, warn=0,
                ctrs=np.array([o['ctr'] for o in outs]), covs=np.array([o['cov'] for o in outs]),
                ams=np.array([o['am'] for o in outs]), axes=np.array([o['axes'] for o in outs]),
                axlens=np.array([o['axlens'] for o in outs]), logvols=np.array([o['logvol'] for o in outs]))


def moments(points, ctx=None):
    points = np.asarray(points, dtype=float)
    cov = np.cov(points, rowvar=False) if len(points) > 1 else np.zeros((points.shape[1],) * 2)
    return points.mean(axis=0), np.atleast_2d(cov)


def improve_covar(covar, ctx=None):
    good, cov, am, axes, st = OB.improve_covar_mat(np.asarray(covar, dtype=float))
    return good, cov, am, axes, (1 if st else 0)


def scale_to_logvol(covs, ams, axes, axlens, logvols, targets, ctx=None):
    for k in range(len(logvols)):
        e = OB.Ell.__new__(OB.Ell)
        e.ndim = covs.shape[1]
        e.ctr, e.cov, e.am, e.axes, e.axlens, e.logvol = None, covs[k], ams[k], axes[k], axlens[k], logvols[k]
        e.scale_to_logvol(float(np.asarray(targets)[k]))
        covs[k], ams[k], axes[k], axlens[k], logvols[k] = e.cov, e.am, e.axes, e.axlens, e.logvol


def bootstrap_expand(points, multi, nboot, seed, chain0, ctx=None):
    out = np.empty(nboot)
    for r in range(nboot):
        s = philox.ChainStream(seed, chain0 + r)
        sel = OB.bootstrap_split(len(points), s.integers(len(points), len(points)))
        out[r] = OB.bootstrap_expand(np.asarray(points), sel, bool(multi))
    return out


class FakeContext:
    """What the host code reads off a ``_lib.Context`` (no device behind it)."""
    serial = 0
    device = 0
    resident_key = None
    friends_key = None

    def __init__(self, device=0):
        self.device = device

    def close(self):
        pass


_fake_ctx = FakeContext()


def default_context(device=None):
    return _fake_ctx


def ensure_resident(bound, ctx=None):
    c = ctx if ctx is not None else _fake_ctx
    if c.resident_key is None or c.resident_key != bound.version:
        bound.make_resident()


def bound_set(axes, ctrs=None, ams=None, logvols=None, ctx=None, key=None):
    (ctx if ctx is not None else _fake_ctx).resident_key = key
    axes = np.asarray(axes, dtype=float)
    if axes.ndim == 2:
        axes = axes[None]
    _state['axes'] = axes.copy()
    _state['ctrs'] = None if ctrs is None else np.array(ctrs, dtype=float).reshape(len(axes), -1)
    _state['ams'] = None if ams is None else np.array(ams, dtype=float).reshape(axes.shape)
    _state['logvols'] = None if logvols is None else np.array(logvols, dtype=float).reshape(-1)


# ---- RadFriends / SupFriends backed by oracle.friends ------------------------------------------------
def friends_update(points, kind, am_prev=None, use_clustering=True, nboot=0, seed=0, chain0=0, ctx=None):
    from oracle import friends as OF
    points = np.asarray(points, dtype=float)
    f = OF.Friends(points.shape[1], kind)
    if am_prev is not None:
        f.am = np.asarray(am_prev, dtype=float)
    idxs = None
    if nboot:
        idxs = [philox.ChainStream(seed, chain0 + r).integers(len(points), len(points)) for r in range(nboot)]
    ncl = int(OF.components_within(points, f.am).max()) + 1 if (use_clustering and am_prev is not None) else 1
    r = f.update(points, bootstrap_idxs=idxs, use_clustering=bool(use_clustering and am_prev is not None))
    return dict(cov=f.cov, am=f.am, axes=f.axes, axes_inv=f.axes_inv, logvol=float(f.logvol), radius=r, nclusters=ncl)


def _friends_obj():
    from oracle import friends as OF
    st = _state['friends']
    f = OF.Friends(st['ctrs'].shape[1], st['kind'])
    f.ctrs, f.axes, f.axes_inv = st['ctrs'], st['axes'], st['axes_inv']
    return f


def friends_set(kind, ctrs, axes, axes_inv, ctx=None, key=None):
    _state['friends'] = dict(kind=kind, ctrs=np.array(ctrs, dtype=float), axes=np.array(axes, dtype=float),
                             axes_inv=np.array(axes_inv, dtype=float))
    (ctx if ctx is not None else _fake_ctx).friends_key = key


def friends_overlap(x, ctx=None):
    f = _friends_obj()
    return np.array([f.overlap(r) for r in np.atleast_2d(x)], dtype=np.int32)


def friends_unif_batch(model, nchain, ndim, loglstar, seed, chain0=0, dimflags=None, ctx=None, draw_only=False,
                       mixture=False):
    f = _friends_obj()
    Q, n = int(nchain), int(ndim)
    o = dict(u=np.empty((Q, n)), v=np.empty((Q, n)), logl=np.zeros(Q), ncall=np.zeros(Q, dtype=np.int32),
             nprop=np.zeros(Q, dtype=np.int32), flags=np.zeros(Q, dtype=np.uint32))
    m = None if draw_only else _models[model]
    for i in range(Q):
        s = philox.ChainStream(seed, chain0 + i)
        if draw_only:
            r = f.sample(s, return_q=bool(mixture))
            x, q = r if mixture else (r, 1)
            o['u'][i] = o['v'][i] = x
            o['ncall'][i], o['nprop'][i] = q, 1
            continue
        nc = npr = 0
        while True:
            x = f.sample(s)
            npr += 1
            if not OS.unitcheck(x, None):
                continue
            v = m.prior_transform(x)
            l = float(m.loglike(v))
            nc += 1
            if l > loglstar:
                break
        o['u'][i], o['v'][i], o['logl'][i], o['ncall'][i], o['nprop'][i] = x, v, l, nc, npr
    return o


def dimflags_from(ndim, periodic=None, reflective=None):
    if periodic is None and reflective is None:
        return None
    f = np.zeros(ndim, dtype=np.uint8)
    if periodic is not None:
        f[np.asarray(periodic, dtype=int)] |= 1
    if reflective is not None:
        f[np.asarray(reflective, dtype=int)] |= 2
    return f


def rwalk_batch(model, u0, loglstar, scale, walks, seed, chain0=0, ncdim=None, ell=None, dimflags=None,
                ctx=None, peer=None, out=None, start_rows=None):
    m = _models[model]
    u0 = np.atleast_2d(u0)
    if start_rows is not None:          # b2n_set_start_rows: u0 is the live set, chain q starts from row start_rows[q]
        idx = np.asarray(start_rows)
        if idx.min() < 0 or idx.max() >= len(u0):
            raise ValueError("start row index out of range")
        u0 = u0[idx]
    Q, n = u0.shape
    per = ref = nb = None
    if dimflags is not None:
        f = np.asarray(dimflags)
        per = np.nonzero(f & 1)[0] if (f & 1).any() else None
        ref = np.nonzero(f & 2)[0] if (f & 2).any() else None
        nb = f == 0
    o = dict(u=np.empty((Q, n)), v=np.empty((Q, n)), logl=np.empty(Q), n_accept=np.empty(Q, dtype=np.int32),
             n_reject=np.empty(Q, dtype=np.int32), ncall=np.empty(Q, dtype=np.int32))
    for i in range(Q):
        ax = _state['axes'][0 if ell is None else ell[i]]
        r = OS.rwalk_chain(u0[i], loglstar, ax, scale, m, philox.ChainStream(seed, chain0 + i), walks,
                           periodic=per, reflective=ref, nonbounded=nb)
        o['u'][i], o['v'][i], o['logl'][i] = r['u'], r['v'], r['logl']
        o['n_accept'][i], o['n_reject'][i], o['ncall'][i] = r['n_accept'], r['n_reject'], r['ncall']
    return o


def _slice(fn, model, u0, loglstar, scale, slices, seed, chain0, doubling, ell):
    m = _models[model]
    u0 = np.atleast_2d(u0)
    Q, n = u0.shape
    o = dict(u=np.empty((Q, n)), v=np.empty((Q, n)), logl=np.empty(Q), n_expand=np.empty(Q, dtype=np.int32),
             n_contract=np.empty(Q, dtype=np.int32), ncall=np.empty(Q, dtype=np.int32),
             flags=np.zeros(Q, dtype=np.uint32))
    for i in range(Q):
        ax = _state['axes'][0 if ell is None else ell[i]]
        r = fn(u0[i], loglstar, ax, scale, m, philox.ChainStream(seed, chain0 + i), slices, doubling=doubling)
        o['u'][i], o['v'][i], o['logl'][i] = r['u'], r['v'], r['logl']
        o['n_expand'][i], o['n_contract'][i], o['ncall'][i] = r['n_expand'], r['n_contract'], r['ncall']
        o['flags'][i] = 2 if r['expansion_warning_set'] else 0
    return o


def rslice_batch(model, u0, loglstar, scale, slices, seed, chain0=0, doubling=False, ell=None, ctx=None,
                 peer=None):
    return _slice(OS.rslice_chain, model, u0, loglstar, scale, slices, seed, chain0, doubling, ell)


def slice_batch(model, u0, loglstar, scale, slices, seed, chain0=0, doubling=False, ell=None, ctx=None,
                peer=None):
    return _slice(OS.slice_chain, model, u0, loglstar, scale, slices, seed, chain0, doubling, ell)


def unif_batch(model, nchain, ndim, loglstar, seed, chain0=0, ncdim=None, dimflags=None, ctx=None,
               draw_only=False, mixture=False, peer=None):
    from scipy.special import logsumexp
    K = len(_state['axes'])
    me = OB.MultiEll.__new__(OB.MultiEll)
    me.ells = []
    for k in range(K):
        e = OB.Ell.__new__(OB.Ell)
        e.ctr, e.am, e.axes = _state['ctrs'][k], _state['ams'][k], _state['axes'][k]
        e.ndim = len(e.ctr)
        me.ells.append(e)
    me.nells, me.ctrs, me.ams, me.logvol_ells = K, _state['ctrs'], _state['ams'], _state['logvols']
    me.logvol = logsumexp(me.logvol_ells)
    Q, n = int(nchain), int(ndim)
    o = dict(u=np.empty((Q, n)), v=np.empty((Q, n)), logl=np.empty(Q), ncall=np.empty(Q, dtype=np.int32),
             nprop=np.empty(Q, dtype=np.int32), flags=np.zeros(Q, dtype=np.uint32))
    if draw_only:
        class _Flat:
            ndim = n

            @staticmethod
            def prior_transform(u):
                return np.full_like(u, 0.5)

            @staticmethod
            def loglike(v):
                return 0.0
        # draw-only: accept the first bound draw irrespective of the cube
        for i in range(Q):
            s = philox.ChainStream(seed, chain0 + i)
            if K == 1:
                x = me.ells[0].ctr + me.ells[0].axes @ OS.randsphere(n, s)
            else:
                cum = np.cumsum(np.exp(me.logvol_ells - me.logvol))
                while True:
                    idx = min(int(np.searchsorted(cum, s.uniform())), K - 1)
                    x = me.ells[idx].ctr + me.ells[idx].axes @ OS.randsphere(n, s)
                    q = int((me.mahal2(x)[0] < 1).sum())
                    if q <= 1 or s.uniform() < 1. / q:
                        break
            o['u'][i] = o['v'][i] = x
            o['logl'][i], o['ncall'][i], o['nprop'][i] = 0.0, 0, 1
        return o
    m = _models[model]
    nb = None if dimflags is None else (np.asarray(dimflags) == 0)
    for i in range(Q):
        r = OS.unif_chain(loglstar, me, m, philox.ChainStream(seed, chain0 + i), n, nonbounded=nb)
        o['u'][i], o['v'][i], o['logl'][i], o['ncall'][i] = r['u'], r['v'], r['logl'], r['ncall']
        o['nprop'][i] = r['ncall']
    return o


# ---- device-resident rounds (ops.ns_*) backed by oracle.nsloop.BatchNS ---------------------------
def _ns_bound():
    return dict(ctrs=_state['ctrs'], ams=_state['ams'], axes=_state['axes'], logvols=_state['logvols'],
                strict=_state['ns_cfg']['strict'])


def ns_create(model, nlive, ndim, batch, sampler, steps, seed, chain0=0, ncdim=None, strict_contains=True,
              facc=0.5, dlogz=0.01, maxiter=None, maxcall=None, update_interval=1 << 62, dimflags=None,
              dead_capacity=None, ctx=None, unit_cube_phase=False, first_min_ncall=0, first_min_eff=100., it0=0,
              logl_max=None):
    assert dimflags is None, "fake backend: periodic/reflective not wired for ns rounds"
    _state['ns_cfg'] = dict(model=_models[model], batch=batch, sampler=('rwalk', 'rslice', 'slice', 'unif')[sampler],
                            steps=steps, seed=seed, chain0=chain0, facc=facc, dlogz=dlogz,
                            maxiter=maxiter if maxiter is not None else 1 << 62,
                            maxcall=maxcall if maxcall is not None else 1 << 62,
                            update_interval=update_interval, strict=bool(strict_contains),
                            unit_cube_phase=bool(unit_cube_phase), first_min_ncall=first_min_ncall,
                            first_min_eff=first_min_eff, it0=it0,
                            logl_max=np.inf if logl_max is None else logl_max, ncdim=int(ncdim or ndim))


def ns_set_state(live_u, live_v, live_logl, logvol, logz, loglstar, ncall, scale, ctx=None):
    from oracle import nsloop
    c = _state['ns_cfg']
    _state['ns'] = nsloop.BatchNS(c['model'], live_u, live_v, live_logl, c['batch'], c['sampler'], c['steps'],
                                  c['seed'], chain0=c['chain0'], facc=c['facc'], scale=scale, logvol=logvol,
                                  logz=logz, loglstar=loglstar, ncall=ncall, update_interval=c['update_interval'],
                                  dlogz=c['dlogz'], maxiter=c['maxiter'], maxcall=c['maxcall'], bound=None,
                                  unit_cube_phase=c['unit_cube_phase'], first_min_ncall=c['first_min_ncall'],
                                  first_min_eff=c['first_min_eff'], it0=c['it0'], logl_max=c['logl_max'])


def ns_status(ctx=None):
    b = _state['ns']
    return dict(it=b.it, ncall=b.ncall, rounds=b.round, logz=b.logz, logvol=b.logvol, loglstar=b.loglstar,
                lmax=float(b.live_logl.max()), delta_logz=b.delta_logz, scale=b.scale, done=b.done,
                need_bound=b.need_bound, doubling=int(b.doubling), error=b.error,
                ncall_last_update=b.ncall_last_update)


def ns_run(max_rounds, check_every=0, ctx=None):
    b = _state['ns']
    if b.phase == 1:
        b.bound = _ns_bound()
    for _ in range(max_rounds):
        if not b.step():
            break
    return ns_status()


def ns_set_counters(rounds, ncall_last_update, doubling, ctx=None):
    b = _state['ns']
    b.round, b.ncall_last_update, b.doubling = int(rounds), int(ncall_last_update), bool(doubling)


def ns_bound_updated(ctx=None):
    _state['ns'].bound_updated(_ns_bound())


def ns_reserve_dead(capacity, ctx=None):
    pass


def ns_update_bound(multi, enlarge=1.0, ctx=None):
    """b2n_ns_update_bound: fit to the run's live points, enlarge (scalar branch: every ellipsoid shifted by the
    same ln enlarge, bounding.py:487-489), make resident."""
    import math
    b = _state['ns']
    pts = b.live_u[:, :_state['ns_cfg']['ncdim']]
    if multi:
        o = multi_decompose(pts)
    else:
        e = bounding_ellipsoid(pts)
        o = dict(nells=1, ctrs=e['ctr'][None], covs=e['cov'][None], ams=e['am'][None], axes=e['axes'][None],
                 axlens=e['axlens'][None], logvols=np.array([e['logvol']]))
    for k in ('ctrs', 'covs', 'ams', 'axes', 'axlens', 'logvols'):
        o[k] = np.array(o[k], dtype=float)
    if enlarge != 1.0:
        scale_to_logvol(o['covs'], o['ams'], o['axes'], o['axlens'], o['logvols'], o['logvols'] + math.log(enlarge))
    bound_set(o['axes'], o['ctrs'], o['ams'], o['logvols'], ctx=ctx, key=('ns', len(_state.setdefault('ns_hist', []))))
    _state['ns_hist'].append(1)
    _state['ns_bound'] = o
    from scipy.special import logsumexp
    return o['nells'], float(logsumexp(o['logvols'])), 0


def ns_get_bound(nells, ncdim, ctx=None):
    o = _state['ns_bound']
    return {k: o[k].copy() for k in ('ctrs', 'covs', 'ams', 'axes', 'axlens', 'logvols')}


def unitcube_batch(model, nchain, ndim, loglstar, seed, chain0=0, ctx=None, peer=None):
    m = _models[model]
    Q, n = int(nchain), int(ndim)
    o = dict(u=np.empty((Q, n)), v=np.empty((Q, n)), logl=np.empty(Q), ncall=np.empty(Q, dtype=np.int32))
    for i in range(Q):
        r = OS.unitcube_chain(loglstar, m, philox.ChainStream(seed, chain0 + i), n)
        o['u'][i], o['v'][i], o['logl'][i], o['ncall'][i] = r['u'], r['v'], r['logl'], r['ncall']
    return o


def ns_get_live(nlive, ndim, ctx=None, only_u=False):
    b = _state['ns']
    if only_u:
        return b.live_u.copy()
    return b.live_u.copy(), b.live_v.copy(), b.live_logl.copy()


def ns_get_dead(first, count, ndim, ctx=None, positions=True):
    u, v, l, lv, nc = _state['ns'].dead_arrays()
    sl = slice(first, first + count)
    if not positions:
        return np.empty((0, ndim)), np.empty((0, ndim)), l[sl], lv[sl], nc[sl].astype(np.int32)
    return u[sl], v[sl], l[sl], lv[sl], nc[sl].astype(np.int32)


def ns_destroy(ctx=None):
    _state.pop('ns', None)


FUNCS = ['ns_set_counters', 'ns_create', 'ns_set_state', 'ns_status', 'ns_run', 'ns_bound_updated', 'ns_reserve_dead',
         'ns_get_live', 'ns_get_dead', 'ns_destroy', 'ns_update_bound', 'ns_get_bound', 'unitcube_batch', 'friends_update', 'friends_set', 'friends_overlap',
         'friends_unif_batch', 'moments', 'improve_covar', 'model_eval', 'membership', 'bounding_ellipsoid', 'multi_decompose', 'scale_to_logvol',
         'bootstrap_expand', 'bound_set', 'ensure_resident', 'dimflags_from', 'rwalk_batch', 'rslice_batch', 'slice_batch',
         'unif_batch']


def install(monkeypatch):
    from dynesty_b200 import ops, likelihoods, _lib
    _state.clear()          # no resident bound until the code under test uploads one
    _fake_ctx.resident_key = None
    monkeypatch.setattr(_lib, 'default_context', default_context)
    monkeypatch.setattr(_lib, 'Context', lambda device=0: _fake_ctx)
    g = globals()
    for name in FUNCS:
        monkeypatch.setattr(ops, name, g[name])
    monkeypatch.setattr(likelihoods.DeviceModel, 'ids', model_ids)
