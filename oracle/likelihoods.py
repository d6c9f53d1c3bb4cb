"""Host (numpy) restatement of the device likelihood / prior-transform registry.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference takes arbitrary Python callables (dynesty.py:578-614); the B200
path evaluates the likelihood *inside* the proposal kernels, so it supports a
closed registry of device models (include/b200nest.h, ``b2n_model``).  The
formulas below follow the reference's own demo / test problem definitions:

  GAUSS_PREC + UNIFORM   demos/Examples -- 25-D Correlated Normal.ipynb cell 1,
                         tests/test_gau.py:67-102
  GAUSS_DIAG + NORMAL_PPF demos/Examples -- 200-D Multivariate Normal.ipynb cell 1
  EGGBOX + IDENTITY      demos/Examples -- Eggbox.ipynb cell 1, tests/test_egg.py
  SHELLS + UNIFORM       demos/Examples -- Gaussian Shells.ipynb cell 1
"""
import math
import numpy as np
from scipy.special import ndtri

PRIOR_IDENTITY, PRIOR_UNIFORM, PRIOR_NORMAL_PPF = 0, 1, 2
LIKE_GAUSS_PREC, LIKE_GAUSS_DIAG, LIKE_EGGBOX, LIKE_SHELLS, LIKE_REGION2D = 0, 1, 2, 3, 4


class Model:
    """prior_kind/prior params + like_kind/like params, host evaluation."""

    def __init__(self, ndim, prior_kind, like_kind, **p):
        self.ndim = ndim
        self.prior_kind = prior_kind
        self.like_kind = like_kind
        self.p = p

    # ---- prior transform u -> v ------------------------------------------
    def prior_transform(self, u):
        u = np.asarray(u, dtype=float)
        if self.prior_kind == PRIOR_IDENTITY:
            return u.copy()
        if self.prior_kind == PRIOR_UNIFORM:
            return self.p['lo'] + self.p['width'] * u
        if self.prior_kind == PRIOR_NORMAL_PPF:
            return self.p['mu'] + self.p['sigma'] * ndtri(u)
        raise ValueError(self.prior_kind)

    # ---- log-likelihood v -> logl  (v: (..., n)) -----------------------------
    def loglike(self, v):
        v = np.asarray(v, dtype=float)
        k = self.like_kind
        if k == LIKE_GAUSS_PREC:
            d = v - self.p['mean']
            return -0.5 * np.einsum('...i,ij,...j->...', d, self.p['prec'],
                                    d) + self.p['lnorm']
        if k == LIKE_GAUSS_DIAG:
            d = v - self.p['mean']
            return -0.5 * np.sum(self.p['ivar'] * d * d, axis=-1) + self.p['lnorm']
        if k == LIKE_EGGBOX:
            t = 2.0 * self.p['tmax'] * v - self.p['tmax']
            return (2.0 + np.prod(np.cos(t / 2.0), axis=-1))**self.p['power']
        if k == LIKE_SHELLS:
            r, w = self.p['r'], self.p['w']
            const = math.log(1. / math.sqrt(2. * math.pi * w**2))
            d1 = np.sqrt(np.sum((v - self.p['c1'])**2, axis=-1))
            d2 = np.sqrt(np.sum((v - self.p['c2'])**2, axis=-1))
            return np.logaddexp(const - (d1 - r)**2 / (2. * w**2),
                                const - (d2 - r)**2 / (2. * w**2))
        if k == LIKE_REGION2D:
            # tests/test_sampling.py:8-23 (diamond_logl / checker_logl) on the first two coordinates
            x, y = v[..., 0], v[..., 1]
            if self.p['shape'] == 0:
                x1, y1 = np.abs(x - 0.5), np.abs(y - 0.5)
                D2 = (x1 - 0.5)**2 + (y1 - 0.5)**2
                out = (np.minimum(x, y) < 0) | (np.maximum(x, y) > 1)
                return np.where((D2 > 0.25) & ~out, D2 - 0.25, -np.inf)
            mult = 16 * 2 * np.pi
            return np.where((x >= 0) & (x <= 1) & (y >= 0) & (y < 1), np.sin(x * mult) * np.sin(y * mult), -np.inf)
        raise ValueError(k)


def region2d(shape='diamond', ndim=2):
    """tests/test_sampling.py:8-23."""
    return Model(ndim, PRIOR_IDENTITY, LIKE_REGION2D, shape={'diamond': 0, 'checkerboard': 1}[shape])


def gauss_corr(ndim, rho, halfwidth):
    """C2 family: mean 0, unit variances, correlation rho; prior U(-h, h)^n."""
    C = np.full((ndim, ndim), rho)
    np.fill_diagonal(C, 1.0)
    prec = np.linalg.inv(C)
    lnorm = -0.5 * (math.log(2 * math.pi) * ndim + np.linalg.slogdet(C)[1])
    return Model(ndim, PRIOR_UNIFORM, LIKE_GAUSS_PREC,
                 lo=np.full(ndim, -halfwidth), width=np.full(ndim, 2. * halfwidth),
                 mean=np.zeros(ndim), prec=prec, lnorm=lnorm)


def gauss_test3d():
    """C1: tests/test_gau.py:67-102 (mean linspace(-1,1,3), 0.95 off-diagonal)."""
    ndim = 3
    mean = np.linspace(-1, 1, ndim)
    C = np.full((ndim, ndim), 0.95)
    np.fill_diagonal(C, 1.0)
    prec = np.linalg.inv(C)
    lnorm = -0.5 * (math.log(2 * math.pi) * ndim + np.linalg.slogdet(C)[1])
    return Model(ndim, PRIOR_UNIFORM, LIKE_GAUSS_PREC, lo=np.full(ndim, -10.),
                 width=np.full(ndim, 20.), mean=mean, prec=prec, lnorm=lnorm)


def iid_normal_ppf(ndim):
    """C4: iid N(0,1) likelihood, standard-normal prior via ppf."""
    lnorm = -0.5 * math.log(2 * math.pi) * ndim
    return Model(ndim, PRIOR_NORMAL_PPF, LIKE_GAUSS_DIAG, mu=np.zeros(ndim),
                 sigma=np.ones(ndim), mean=np.zeros(ndim), ivar=np.ones(ndim),
                 lnorm=lnorm)


def eggbox(ndim, tmax=5.0 * math.pi, power=5.0):
    """C3."""
    return Model(ndim, PRIOR_IDENTITY, LIKE_EGGBOX, tmax=tmax, power=power)


def shells(ndim, r=2.0, w=0.1, c=3.5, halfwidth=6.0):
    """C5."""
    c1 = np.zeros(ndim)
    c1[0] = -c
    c2 = np.zeros(ndim)
    c2[0] = c
    return Model(ndim, PRIOR_UNIFORM, LIKE_SHELLS, lo=np.full(ndim, -halfwidth),
                 width=np.full(ndim, 2 * halfwidth), c1=c1, c2=c2, r=r, w=w)
