"""B200 bounds: drop-in mirrors of dynesty's ``Ellipsoid`` / ``MultiEllipsoid``.

Interface = the reference's ``Bound`` duck-type (bounding.py:76-122) plus the public
attributes its other code reads (``ctr/cov/am/axes/axlens`` and ``ells/ctrs/covs/ams/
logvol_ells/nells``, bounding.py:201-240, 440-476; plotting.py:1710, 2035).  Every
numeric method is one call into libb200nest.so; the objects hold plain numpy arrays,
so ``copy.deepcopy`` (sampler.py:510) and ``pickle`` (utils.py:2343) just work.
"""
import itertools
import math
import warnings

import numpy as np

from . import ops, _lib
from ._compat import BoundBase

__all__ = ['B200Ellipsoid', 'B200MultiEllipsoid', 'B200RadFriends', 'B200SupFriends', 'TaggedAxes']


class TaggedAxes(np.ndarray):
    """``get_random_axes`` result: the (ncdim, ncdim) axes matrix that also remembers
    which ellipsoid of which bound it belongs to, so that the batched samplers can
    address the device-resident copy instead of re-uploading Q matrices per queue fill
    (the reference pickles the matrix into every task, sampler.py:708-717)."""

    def __new__(cls, arr, bound, ell):
        obj = np.asarray(arr).view(cls)
        obj.bound = bound
        obj.ell = int(ell)
        return obj

    def __array_finalize__(self, obj):
        self.bound = getattr(obj, 'bound', None)
        self.ell = getattr(obj, 'ell', 0)

    def __reduce__(self):            # pickles as a plain array
        return (np.asarray, (np.array(self),))

    def __deepcopy__(self, memo):
        return np.array(self)


# Version tokens: process-wide unique and increasing, a fresh one after EVERY change of a bound's arrays and
# after every unpickle / deepcopy.  `Context.resident_key == bound.version` therefore means "these very
# ellipsoids are the resident bound of that ctx" no matter which object, sampler or driver uploaded them
# (one resident bound per ctx; round-1 ADVICE: per-object caches keyed on id() went stale).
_version = itertools.count(1)


def _resolve_ctx(ctx, device):
    """The Context a bound works on: the one it was given, else the default context of its device."""
    if ctx is not None:
        return ctx
    return _lib.default_context(device)


def _seed_from(rstate):
    """64-bit Philox seed drawn from the caller's numpy Generator."""
    if rstate is None:
        rstate = np.random.default_rng()
    return int(rstate.integers(0, 2**63 - 1))


class _EllView:
    """Read-only view of one ellipsoid of a B200MultiEllipsoid (plotting compatibility)."""

    def __init__(self, parent, k):
        self.ndim = parent.ndim
        self.ctr, self.cov, self.am = parent.ctrs[k], parent.covs[k], parent.ams[k]
        self.axes, self.axlens = parent.axes_all[k], parent.axlens_all[k]
        self.logvol = float(parent.logvol_ells[k])
        self.funit = 1


class B200MultiEllipsoid(BoundBase):
    """``MultiEllipsoid`` (bounding.py:420-731) built and queried on the GPU."""

    def __init__(self, ndim, ctx=None):
        super().__init__(ndim)
        self._ctx = ctx
        self._device = getattr(ctx, 'device', None)     # survives pickling (the handle does not)
        n = ndim
        # Ellipsoid(ndim) default: centre 0 (sic), cov = I n/4 (bounding.py:203-205)
        cov = np.identity(n) * n / 4
        self.nells = 1
        self.ctrs = np.zeros((1, n))
        self.covs = cov[None].copy()
        self.ams = np.linalg.inv(cov)[None]
        self.axes_all = (np.identity(n) * math.sqrt(n / 4.))[None]
        self.axlens_all = np.full((1, n), math.sqrt(n / 4.))
        from scipy.special import gammaln
        pref = n * math.log(2.) + n * gammaln(1.5) - gammaln(n / 2. + 1)
        self.logvol_ells = np.array([pref + 0.5 * n * math.log(n / 4.)])
        self.logvol = float(self.logvol_ells[0])
        self.funit = 1
        self.labels = None
        self.version = next(_version)      # new token whenever the arrays change (device cache key)

    # -- pickling: drop the (per-process) context handle, keep the device index ------------
    def __getstate__(self):
        d = self.__dict__.copy()
        d['_ctx'] = None
        d.pop('_contains_cache', None)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.__dict__.setdefault('_device', None)
        self.version = next(_version)

    # -- deepcopy (sampler.py:510, 589 copy the bound on every update): same process, so the copy stays
    #    on the SAME context (a copy that silently moved to the default context would run on another
    #    stream / GPU than the sampler it belongs to)
    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k != '_contains_cache':
                new.__dict__[k] = v if k == '_ctx' else copy.deepcopy(v, memo)
        new.version = next(_version)
        return new

    @property
    def ctx(self):
        return _resolve_ctx(self._ctx, self._device)

    @property
    def ells(self):
        return [_EllView(self, k) for k in range(self.nells)]

    def _refresh_logvol(self):
        from scipy.special import logsumexp
        self.logvol = float(logsumexp(self.logvol_ells))     # ignores overlap (:466)
        self.version = next(_version)

    # -- Bound interface -----------------------------------------------------------------
    # The reference's Sampler.propose_live asks ``bound.contains(u)`` once per queue slot (sampler.py:485), i.e.
    # `queue_size` single-point queries per fill.  One launch + round trip per point would dominate a fill, so
    # membership is computed IN BATCHES on the GPU and memoised per bound version: `update` prefetches the points
    # it was fitted to (the live points -- the only start points there are), the samplers prefetch the end points
    # of every queue they evolve (the live points of the future); ``contains`` is then a dictionary lookup and
    # falls back to a one-point launch only for a point nobody announced.
    def _cache(self, strict):
        c = self.__dict__.get('_contains_cache')
        if c is None or c['version'] != self.version or c['strict'] != strict:
            pts = None if c is None else c.get('points')
            c = dict(version=self.version, strict=strict, hit={}, points=None)
            self.__dict__['_contains_cache'] = c
            if pts is not None and pts.shape[1] == self.ndim:
                self._prefetch(pts, strict)               # same points, new ellipsoids (e.g. after the enlarge)
        return c

    def _prefetch(self, x, strict=True):
        x = np.ascontiguousarray(x, dtype=float)
        if x.ndim != 2 or x.shape[1] != self.ndim or len(x) == 0:
            return
        c = self._cache(strict)
        q = ops.membership(x, self.ctrs, self.ams, strict=strict, ctx=self.ctx)[1]
        inside = (q > 0).tolist()
        hit = c['hit']
        if len(hit) > 8 * max(len(x), 4096):              # bounded memory: drop what was announced long ago
            hit.clear()
        for row, ok in zip(x, inside):
            hit[row.tobytes()] = ok
        c['points'] = x if c['points'] is None or len(x) >= len(c['points']) else c['points']

    def prefetch_contains(self, x):
        """Announce points whose membership will be asked for (batched on the GPU, memoised)."""
        self._prefetch(np.asarray(x, dtype=float)[:, :self.ndim], self.__dict__.get('_strict', True))

    def _contains_one(self, x, strict):
        x = np.ascontiguousarray(x, dtype=float)
        c = self._cache(strict)
        r = c['hit'].get(x.tobytes())
        if r is None:
            r = bool(ops.membership(x[None], self.ctrs, self.ams, strict=strict, ctx=self.ctx)[1][0] > 0)
            c['hit'][x.tobytes()] = r
        return r

    def contains(self, x):
        """bounding.py:520-523 (strict <)."""
        return self._contains_one(x, True)

    def contains_many(self, x):
        """Batched ``contains`` (one launch for a whole queue of start points)."""
        return ops.membership(x, self.ctrs, self.ams, strict=True, ctx=self.ctx)[1] > 0

    def within(self, x, j=None):
        mask, _ = ops.membership(np.asarray(x, dtype=float)[None], self.ctrs, self.ams, ctx=self.ctx)
        m = mask[0].copy()
        if j is not None:
            m[j] = False
        return np.nonzero(m)[0]

    def overlap(self, x, j=None):
        return len(self.within(x, j=j))

    def make_resident(self, ctx=None):
        """Upload the ellipsoids as the resident bound of `ctx` (default: the bound's own context)."""
        ops.bound_set(self.axes_all, self.ctrs, self.ams, self.logvol_ells, ctx=ctx if ctx is not None else self.ctx,
                      key=self.version)

    def samples(self, nsamples, rstate=None):
        """bounding.py:592-606: uniform draws from the union (no cube test)."""
        ops.ensure_resident(self, self.ctx)
        o = ops.unif_batch(-1, nsamples, self.ndim, -np.inf, _seed_from(rstate), draw_only=True,
                           ncdim=self.ndim, ctx=self.ctx)
        return o['u']

    def sample(self, rstate=None, return_q=False):
        x = self.samples(1, rstate=rstate)[0]
        idx = int(np.argmin(np.einsum('ki,kij,kj->k', x - self.ctrs, self.ams, x - self.ctrs)))
        if return_q:
            return x, idx, self.overlap(x)
        return x, idx

    def monte_carlo_logvol(self, ndraws=10000, rstate=None, return_overlap=True):
        """bounding.py:608-630: MC estimate of the log-volume of the UNION (and of its
        fractional overlap with the unit cube) from volume-weighted draws and their q."""
        ops.ensure_resident(self, self.ctx)
        o = ops.unif_batch(-1, ndraws, self.ndim, -np.inf, _seed_from(rstate), draw_only=True, mixture=True,
                           ncdim=self.ndim, ctx=self.ctx)
        w = 1. / o['ncall']                                   # 1 / q
        logvol = math.log(w.sum() / ndraws) + self.logvol
        if not return_overlap:
            return logvol
        x = o['u']
        inside = np.all((x > 0) & (x < 1), axis=1)
        return logvol, float((w * inside).sum() / w.sum())

    def unitcube_overlap(self, ndraws=10000, rstate=None):
        """bounding.py:336-343 (single ellipsoid) / the overlap half of monte_carlo_logvol."""
        x = self.samples(ndraws, rstate=rstate)
        return float(np.all((x > 0) & (x < 1), axis=1).mean())

    def get_random_axes(self, rstate):
        """bounding.py:726-731: axes of an ellipsoid picked with probability ~ volume."""
        if self.nells == 1:
            k = 0
        else:
            probs = np.exp(self.logvol_ells - self.logvol)
            k = min(int(np.searchsorted(np.cumsum(probs), rstate.random())), self.nells - 1)
        return TaggedAxes(self.axes_all[k], self, k)

    def random_ells(self, rstate, size):
        """Vectorised ``get_random_axes``: `size` volume-weighted ellipsoid indices."""
        if self.nells == 1:
            return np.zeros(size, dtype=np.int32)
        probs = np.exp(self.logvol_ells - self.logvol)
        k = np.searchsorted(np.cumsum(probs), rstate.random(size))
        return np.minimum(k, self.nells - 1).astype(np.int32)

    def scale_to_logvol(self, logvol):
        """bounding.py:478-495: scalar = new total, iterable = per-ellipsoid targets."""
        if np.ndim(logvol) > 0:
            target = np.asarray(logvol, dtype=float)
        else:
            target = self.logvol_ells + (float(logvol) - self.logvol)
        ops.scale_to_logvol(self.covs, self.ams, self.axes_all, self.axlens_all, self.logvol_ells,
                            target, ctx=self.ctx)
        self._refresh_logvol()

    def update(self, points, rstate=None, bootstrap=0, pool=None, mc_integrate=False):
        """bounding.py:632-724.  `pool` is ignored: the bootstrap replicas run on the GPU."""
        points = np.ascontiguousarray(points, dtype=float)
        npoints, ndim = points.shape
        if npoints == 1:
            raise RuntimeError('Cannot compute the bounding ellipsoid of a single point.')
        o = ops.multi_decompose(points, ctx=self.ctx)
        if o['warn'] & _lib.WARN_IDENTITY_FALLBACK:
            warnings.warn("Failed to guarantee the ellipsoid axes will be non-singular. "
                          "Defaulting to a sphere.")
        self.nells = o['nells']
        self.ctrs, self.covs, self.ams = o['ctrs'], o['covs'], o['ams']
        self.axes_all, self.axlens_all, self.logvol_ells = o['axes'], o['axlens'], o['logvols']
        self.labels = o['labels']
        self._refresh_logvol()
        self.__dict__['_contains_cache'] = dict(version=-1, strict=True, hit={}, points=points)   # re-evaluated lazily
        if bootstrap > 0:
            expands = ops.bootstrap_expand(points, True, int(bootstrap), _seed_from(rstate), 0, ctx=self.ctx)
            expand = float(expands.max())
            if math.log10(expand) * ndim > 2:                              # :705-714
                warnings.warn('The enlargement factor for the ellipsoidal bounds determined '
                              'from bootstrapping is very large.')
            if expand > 1.:
                self.scale_to_logvol(self.logvol_ells + ndim * math.log(expand))
        if mc_integrate:                                                       # :720-724
            self.logvol, self.funit = self.monte_carlo_logvol(rstate=rstate, return_overlap=True)


class B200Ellipsoid(BoundBase):
    """``Ellipsoid`` (bounding.py:182-417) built and queried on the GPU."""

    def __init__(self, ndim, ctx=None):
        super().__init__(ndim)
        self._m = B200MultiEllipsoid(ndim, ctx=ctx)
        self.funit = 1

    def __getattr__(self, name):
        m = self.__dict__.get('_m')
        if m is None:
            raise AttributeError(name)
        one = {'ctr': 'ctrs', 'cov': 'covs', 'am': 'ams', 'axes': 'axes_all', 'axlens': 'axlens_all'}
        if name in one:
            return getattr(m, one[name])[0]
        if name in ('version', 'nells', 'ctrs', 'ams', 'covs', 'axes_all', 'axlens_all', 'logvol_ells'):
            return getattr(m, name)
        raise AttributeError(name)

    @property
    def logvol(self):
        return float(self._m.logvol_ells[0])

    @logvol.setter
    def logvol(self, v):       # BoundBase.__init__ assigns 0 before _m exists
        pass

    def contains(self, x):
        """bounding.py:302-305 (non-strict: distance <= 1); memoised like MultiEllipsoid.contains."""
        return self._m._contains_one(x, False)

    def prefetch_contains(self, x):
        self._m._prefetch(np.asarray(x, dtype=float)[:, :self.ndim], False)

    def contains_many(self, x):
        return ops.membership(x, self._m.ctrs, self._m.ams, strict=False, ctx=self._m.ctx)[1] > 0

    def distance_many(self, x):
        return np.sqrt(ops.membership(x, self._m.ctrs, self._m.ams, want_d2=True, ctx=self._m.ctx)[2][:, 0])

    def make_resident(self, ctx=None):
        self._m.make_resident(ctx)

    @property
    def ctx(self):
        return self._m.ctx

    def samples(self, nsamples, rstate=None):
        return self._m.samples(nsamples, rstate=rstate)

    def sample(self, rstate=None):
        return self._m.samples(1, rstate=rstate)[0]

    def get_random_axes(self, rstate):
        return TaggedAxes(self._m.axes_all[0], self, 0)

    def random_ells(self, rstate, size):
        return np.zeros(size, dtype=np.int32)

    def scale_to_logvol(self, logvol):
        self._m.scale_to_logvol(np.array([float(logvol)]))

    def update(self, points, rstate=None, bootstrap=0, pool=None, mc_integrate=False):
        """bounding.py:345-414."""
        points = np.ascontiguousarray(points, dtype=float)
        o = ops.bounding_ellipsoid(points, ctx=self._m.ctx)
        if o['warn'] & _lib.WARN_IDENTITY_FALLBACK:
            warnings.warn("Failed to guarantee the ellipsoid axes will be non-singular. "
                          "Defaulting to a sphere.")
        m = self._m
        m.nells = 1
        m.ctrs, m.covs, m.ams = o['ctr'][None], o['cov'][None], o['am'][None]
        m.axes_all, m.axlens_all = o['axes'][None], o['axlens'][None]
        m.logvol_ells = np.array([o['logvol']])
        m._refresh_logvol()
        m.__dict__['_contains_cache'] = dict(version=-1, strict=False, hit={}, points=points)
        if bootstrap > 0:
            expands = ops.bootstrap_expand(points, False, int(bootstrap), _seed_from(rstate), 0, ctx=m.ctx)
            expand = float(expands.max())
            if expand > 1.:
                self.scale_to_logvol(self.logvol + self.ndim * math.log(expand))
        if mc_integrate:                                                       # :411-414
            self.funit = self.unitcube_overlap(rstate=rstate)

    def unitcube_overlap(self, ndraws=10000, rstate=None):
        """bounding.py:336-343."""
        return self._m.unitcube_overlap(ndraws, rstate=rstate)

    def update_sharded(self, points_shard, comm):
        """``bounding_ellipsoid`` (bounding.py:1387-1461) of a live set whose ROWS are sharded over the ranks of
        `comm`: every rank reduces its own rows on its GPU, two small all-reduces carry
        (count, sum x, scatter) and max_i delta^T am delta, and every rank finishes the identical ellipsoid (the
        repair ladder and the eigen-decomposition are deterministic, so no broadcast is needed).
        Communication per update: n^2 + n + 1 doubles (sum) and 1 double (max) per pass, instead of N n doubles."""
        pts = np.ascontiguousarray(points_shard, dtype=float)
        n = self.ndim
        m = self._m
        mean_r, cov_r = ops.moments(pts, ctx=m.ctx)                                   # this rank's rows, on its GPU
        nr = float(len(pts))
        tot = comm.allreduce_sum(np.concatenate([[nr], nr * mean_r]))                  # (count, sum x)
        N = tot[0]
        if N < 2:
            raise ValueError('Cannot compute a bounding ellipsoid of a single point.')
        mean = tot[1:] / N
        d = mean_r - mean
        S = comm.allreduce_sum((nr - 1.0) * cov_r + nr * np.outer(d, d))               # scatter about the global mean
        covar = S / (N - 1.0)
        for i in range(2):                                                              # bounding.py:1424-1457
            good, covar, am, axes, warn = ops.improve_covar(covar, ctx=m.ctx)
            if warn & _lib.WARN_IDENTITY_FALLBACK:
                warnings.warn("Failed to guarantee the ellipsoid axes will be non-singular. Defaulting to a sphere.")
            d2 = ops.membership(pts, mean[None], am[None], want_d2=True, ctx=m.ctx)[2][:, 0]
            fmax = comm.max(float(d2.max()))
            one_minus_a_bit = 1. - 1e-3
            if i == 0 and fmax > one_minus_a_bit:
                mult = fmax / one_minus_a_bit
                covar = covar * mult
                am = am / mult
                axes = axes * math.sqrt(mult)
            if i == 1 and fmax >= 1:
                raise RuntimeError("Failed to initialize the ellipsoid to contain all the points")
            if good:
                break
        axlens = np.linalg.norm(axes, axis=0)                                           # axes = V sqrt(lambda) (:227-230)
        from scipy.special import gammaln
        pref = n * math.log(2.) + n * gammaln(1.5) - gammaln(n / 2. + 1)
        m.nells = 1
        m.ctrs, m.covs, m.ams = mean[None], covar[None], am[None]
        m.axes_all, m.axlens_all = axes[None], axlens[None]
        m.logvol_ells = np.array([pref + float(np.log(axlens).sum())])
        m._refresh_logvol()


class _B200Friends(BoundBase):
    """``RadFriends`` / ``SupFriends`` (bounding.py:734-996 / 999-1263): one ball / cube of common shape around every
    live point, built and queried on the GPU (csrc/b2n_friends.cu).  Attributes as in the reference: ``ctrs, cov, am,
    axes, axes_inv, logvol, funit``; ``need_centers`` makes the Sampler hand over the live points before it samples
    (sampler.py:479-482)."""
    kind = None

    def __init__(self, ndim, cov=None, ctx=None):
        super().__init__(ndim)
        self._ctx = ctx
        self._device = getattr(ctx, 'device', None)
        self.need_centers = True
        self.ctrs = np.empty((0, ndim))
        from scipy.special import gammaln
        self._pref = (ndim * math.log(2.) + ndim * gammaln(1.5) - gammaln(ndim / 2. + 1)) if self.kind == 'balls' \
            else ndim * math.log(2.)                                      # :761 / :1027
        cov = np.identity(ndim) if cov is None else np.array(cov, dtype=float)
        lam, vec = np.linalg.eigh(cov)        # (constructor only: the initial metric of a fresh object, :749-762)
        self.cov = cov
        self.am = (vec / lam) @ vec.T
        self.axes = (vec * np.sqrt(lam)) @ vec.T
        self.axes_inv = (vec / np.sqrt(lam)) @ vec.T
        self.logvol = float(self._pref + 0.5 * np.log(lam).sum())
        self.funit = 1
        self.radius = 1.0
        self.version = next(_version)

    def __getstate__(self):
        d = self.__dict__.copy()
        d['_ctx'] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.version = next(_version)

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = v if k == '_ctx' else copy.deepcopy(v, memo)
        new.version = next(_version)
        return new

    @property
    def ctx(self):
        return _resolve_ctx(self._ctx, self._device)

    # -- residency: (ctrs, axes, axes_inv) on the device.  `ctrs` is a plain attribute the reference's Sampler
    #    re-assigns to its live-point array before every proposal (sampler.py:481) -- the SAME array object, mutated
    #    in place as live points are replaced -- so it is uploaded afresh for every batched operation (N x n doubles).
    def _resident(self):
        c = self.ctx
        ops.friends_set(self.kind, np.asarray(self.ctrs, dtype=float)[:, :self.ndim], self.axes, self.axes_inv, ctx=c,
                        key=self.version)
        return c

    def make_resident(self, ctx=None):
        """for the chain samplers: the common axes as a one-ellipsoid resident bound (get_random_axes)"""
        ops.bound_set(self.axes[None], ctx=ctx if ctx is not None else self.ctx, key=self.version)

    def scale_to_logvol(self, logvol):
        """bounding.py:765-774 / 1031-1040."""
        f = math.exp((float(logvol) - self.logvol) / self.ndim)
        self.cov = self.cov * f**2
        self.am = self.am / f**2
        self.axes = self.axes * f
        self.axes_inv = self.axes_inv / f
        self.logvol = float(logvol)
        self.version = next(_version)

    def overlap(self, x):
        return int(ops.friends_overlap(np.asarray(x, dtype=float)[None], ctx=self._resident())[0])

    def overlap_many(self, x):
        return ops.friends_overlap(x, ctx=self._resident())

    def contains(self, x):
        """bounding.py:792-795 / 1059-1062.  The Sampler asks this for its START points (sampler.py:485), which are
        rows of `ctrs` themselves: a centre lies in its own ball / cube (its distance to itself is exactly 0), so a
        query that IS a row of the centre array is answered without a launch; anything else goes to the GPU."""
        x = np.asarray(x)
        c = self.ctrs
        if isinstance(c, np.ndarray) and len(c) and np.may_share_memory(x, c) and x.shape == (self.ndim,):
            return True
        return self.overlap(x) > 0

    def contains_many(self, x):
        return self.overlap_many(x) > 0

    def samples(self, nsamples, rstate=None):
        o = ops.friends_unif_batch(-1, nsamples, self.ndim, -np.inf, _seed_from(rstate), draw_only=True, ctx=self._resident())
        return o['u']

    def sample(self, rstate=None, return_q=False):
        o = ops.friends_unif_batch(-1, 1, self.ndim, -np.inf, _seed_from(rstate), draw_only=True, mixture=return_q,
                                   ctx=self._resident())
        return (o['u'][0], int(o['ncall'][0])) if return_q else o['u'][0]

    def monte_carlo_logvol(self, ndraws=10000, rstate=None, return_overlap=True):
        """bounding.py:842-869 / 1110-1137."""
        o = ops.friends_unif_batch(-1, ndraws, self.ndim, -np.inf, _seed_from(rstate), draw_only=True, mixture=True,
                                   ctx=self._resident())
        w = 1. / o['ncall']
        logvol = math.log(w.sum() / ndraws * len(self.ctrs)) + self.logvol
        if not return_overlap:
            return logvol
        x = o['u']
        inside = np.all((x > 0) & (x < 1), axis=1)
        return logvol, float((w * inside).sum() / w.sum())

    def get_random_axes(self, rstate):
        return TaggedAxes(self.axes, self, 0)

    def random_ells(self, rstate, size):
        return np.zeros(size, dtype=np.int32)

    def update(self, points, rstate=None, bootstrap=0, pool=None, mc_integrate=False, use_clustering=True):
        """bounding.py:874-958 / 1142-1226.  `pool` is ignored: the bootstrap realisations run on the GPU."""
        points = np.ascontiguousarray(points, dtype=float)
        o = ops.friends_update(points, self.kind, am_prev=self.am, use_clustering=use_clustering, nboot=int(bootstrap),
                               seed=_seed_from(rstate) if bootstrap else 0, ctx=self.ctx)
        self.cov, self.am, self.axes, self.axes_inv = o['cov'], o['am'], o['axes'], o['axes_inv']
        self.logvol, self.radius, self.nclusters = o['logvol'], o['radius'], o['nclusters']
        self.ctrs = points
        self.version = next(_version)
        if mc_integrate:
            self.funit = self.monte_carlo_logvol(rstate=rstate, return_overlap=True)[1]


class B200RadFriends(_B200Friends):
    """``RadFriends`` (bounding.py:734-996): N-balls, Euclidean norm."""
    kind = 'balls'


class B200SupFriends(_B200Friends):
    """``SupFriends`` (bounding.py:999-1263): N-cubes, Chebyshev norm."""
    kind = 'cubes'
