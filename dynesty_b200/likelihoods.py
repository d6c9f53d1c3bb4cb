"""Device models: the prior-transform / log-likelihood pairs the proposal kernels
can evaluate in-kernel ("device-side likelihood callback").

The reference accepts arbitrary Python callables (dynesty.py:584-614) and calls
them once per proposal on the host.  Here a model is a descriptor from a closed
registry (include/b200nest.h, ``b2n_model_desc``) whose parameters live in HBM.
A ``DeviceModel`` is ALSO a pair of host callables (``prior_transform`` /
``loglikelihood``), evaluated on the GPU through ``b2n_model_eval`` -- so the
same object can be handed to dynesty's own ``NestedSampler`` as
``loglikelihood=model.loglikelihood, prior_transform=model.prior_transform``
(unit-cube warm-up phase, initial live points) while the B200 samplers pick up
the descriptor for the in-kernel evaluation.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import ModelDesc, ptr, f64


class DeviceModel:
    def __init__(self, ndim, prior_kind, like_kind, prior_p0=None, prior_p1=None, like_vec0=None,
                 like_vec1=None, like_mat=None, s0=0.0, s1=0.0, s2=0.0, name='model'):
        self.ndim = int(ndim)
        self.name = name
        self.prior_kind, self.like_kind = int(prior_kind), int(like_kind)
        opt = lambda a, shape: None if a is None else f64(np.broadcast_to(a, shape))
        n = self.ndim
        self.prior_p0, self.prior_p1 = opt(prior_p0, (n,)), opt(prior_p1, (n,))
        self.like_vec0, self.like_vec1 = opt(like_vec0, (n,)), opt(like_vec1, (n,))
        self.like_mat = opt(like_mat, (n, n))
        self.s = (float(s0), float(s1), float(s2))
        self._ids = {}          # ctx -> (full id, likelihood-only id)

    # -- pickling: device handles are per-process, re-created lazily ------------
    def __getstate__(self):
        d = self.__dict__.copy()
        d['_ids'] = {}
        return d

    def _desc(self, prior_kind):
        d = ModelDesc()
        d.ndim, d.prior_kind, d.like_kind = self.ndim, prior_kind, self.like_kind
        d.prior_p0, d.prior_p1 = ptr(self.prior_p0), ptr(self.prior_p1)
        d.like_vec0, d.like_vec1, d.like_mat = ptr(self.like_vec0), ptr(self.like_vec1), ptr(self.like_mat)
        d.like_s0, d.like_s1, d.like_s2 = self.s
        return d

    def ids(self, ctx=None):
        ctx = ctx if ctx is not None else _lib.default_context()
        key = ctx.serial           # (not id(ctx): an address can be reused after a Context is freed)
        if key not in self._ids:
            out = []
            for pk in (self.prior_kind, _lib.PRIOR_IDENTITY):
                mid = C.c_int32(-1)
                ctx.check(ctx.lib.b2n_model_create(ctx.h, C.byref(self._desc(pk)), C.byref(mid)))
                out.append(mid.value)
            self._ids[key] = tuple(out)
        return self._ids[key]

    def model_id(self, ctx=None):
        return self.ids(ctx)[0]

    # -- host callables (GPU-evaluated) ------------------------------------------
    def evaluate(self, u, ctx=None):
        """(v, logl) of unit-cube points u (M, ndim) in one launch."""
        from . import ops
        return ops.model_eval(self.ids(ctx)[0], u, ctx=ctx)

    def prior_transform(self, u):
        from . import ops
        u = np.asarray(u, dtype=float)
        v, _ = ops.model_eval(self.ids()[0], u.reshape(-1, self.ndim))
        return v.reshape(u.shape)

    def loglikelihood(self, v):
        from . import ops
        v = np.asarray(v, dtype=float)
        _, l = ops.model_eval(self.ids()[1], v.reshape(-1, self.ndim), want_v=False)
        return float(l[0]) if v.ndim == 1 else l


# ---- the BASELINE.json problem families -------------------------------------------
def gauss_corr(ndim, rho=0.4, halfwidth=5.0):
    """C2: correlated normal, prior U(-h, h)^n (demos/Examples -- 25-D Correlated Normal.ipynb)."""
    Cm = np.full((ndim, ndim), float(rho))
    np.fill_diagonal(Cm, 1.0)
    prec = np.linalg.inv(Cm)
    lnorm = -0.5 * (math.log(2 * math.pi) * ndim + np.linalg.slogdet(Cm)[1])
    m = DeviceModel(ndim, _lib.PRIOR_UNIFORM, _lib.LIKE_GAUSS_PREC, prior_p0=-halfwidth,
                    prior_p1=2 * halfwidth, like_vec0=0.0, like_mat=prec, s0=lnorm,
                    name='gauss_corr%d' % ndim)
    m.logz_truth = -ndim * math.log(2 * halfwidth)
    return m


def gauss_test3d():
    """C1: tests/test_gau.py:67-102."""
    n = 3
    Cm = np.full((n, n), 0.95)
    np.fill_diagonal(Cm, 1.0)
    lnorm = -0.5 * (math.log(2 * math.pi) * n + np.linalg.slogdet(Cm)[1])
    m = DeviceModel(n, _lib.PRIOR_UNIFORM, _lib.LIKE_GAUSS_PREC, prior_p0=-10., prior_p1=20.,
                    like_vec0=np.linspace(-1, 1, n), like_mat=np.linalg.inv(Cm), s0=lnorm,
                    name='gauss_test3d')
    m.logz_truth = n * (-math.log(20.))
    return m


def iid_normal_ppf(ndim):
    """C4: iid N(0,1) likelihood with a standard-normal ppf prior
    (demos/Examples -- 200-D Multivariate Normal.ipynb)."""
    lnorm = -0.5 * math.log(2 * math.pi) * ndim
    m = DeviceModel(ndim, _lib.PRIOR_NORMAL_PPF, _lib.LIKE_GAUSS_DIAG, prior_p0=0., prior_p1=1.,
                    like_vec0=0., like_vec1=1., s0=lnorm, name='iid_normal%d' % ndim)
    m.logz_truth = lnorm - 0.5 * ndim * math.log(2)
    return m


def eggbox(ndim, tmax=5.0 * math.pi, power=5.0):
    """C3: demos/Examples -- Eggbox.ipynb generalised to ndim."""
    m = DeviceModel(ndim, _lib.PRIOR_IDENTITY, _lib.LIKE_EGGBOX, s0=tmax, s1=power,
                    name='eggbox%d' % ndim)
    m.logz_truth = 235.856 if ndim == 2 else None      # tests/test_egg.py:29-46
    return m


def shells(ndim, r=2.0, w=0.1, c=3.5, halfwidth=6.0):
    """C5: demos/Examples -- Gaussian Shells.ipynb."""
    c1 = np.zeros(ndim)
    c1[0] = -c
    c2 = np.zeros(ndim)
    c2[0] = c
    m = DeviceModel(ndim, _lib.PRIOR_UNIFORM, _lib.LIKE_SHELLS, prior_p0=-halfwidth,
                    prior_p1=2 * halfwidth, like_vec0=c1, like_vec1=c2, s0=r, s1=w,
                    name='shells%d' % ndim)
    m.logz_truth = {2: -1.75, 5: -5.67, 10: -14.59}.get(ndim)
    return m


def region2d(shape='diamond', ndim=2):
    """The hard-edged regions of the reference's sampler-uniformity harness (tests/test_sampling.py:8-23):
    ``diamond_logl`` / ``checker_logl`` on the first two coordinates, identity prior, the remaining
    dimensions free.  Used with loglstar = 0: the samplers must leave the uniform distribution on
    {logl > 0} invariant."""
    m = DeviceModel(ndim, _lib.PRIOR_IDENTITY, _lib.LIKE_REGION2D, s0={'diamond': 0.0, 'checkerboard': 1.0}[shape],
                    name='region2d_%s%d' % (shape, ndim))
    m.logz_truth = None
    return m
