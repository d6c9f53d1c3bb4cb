"""Shared test helpers (oracle models -> device models, fixtures)."""
import numpy as np

from oracle import likelihoods as OL

SEED = 56432


def device_model(m):
    """DeviceModel with the same parameters as an oracle Model."""
    from dynesty_b200.likelihoods import DeviceModel
    p = m.p
    if m.like_kind == OL.LIKE_GAUSS_PREC:
        kw = dict(like_vec0=p['mean'], like_mat=p['prec'], s0=p['lnorm'])
    elif m.like_kind == OL.LIKE_GAUSS_DIAG:
        kw = dict(like_vec0=p['mean'], like_vec1=p['ivar'], s0=p['lnorm'])
    elif m.like_kind == OL.LIKE_EGGBOX:
        kw = dict(s0=p['tmax'], s1=p['power'])
    else:
        kw = dict(like_vec0=p['c1'], like_vec1=p['c2'], s0=p['r'], s1=p['w'])
    if m.prior_kind == OL.PRIOR_UNIFORM:
        kw.update(prior_p0=p['lo'], prior_p1=p['width'])
    elif m.prior_kind == OL.PRIOR_NORMAL_PPF:
        kw.update(prior_p0=p['mu'], prior_p1=p['sigma'])
    return DeviceModel(m.ndim, m.prior_kind, m.like_kind, **kw)


M6 = OL.gauss_corr(6, 0.4, 5.)
MODELS = {
    'g6': M6, 'g6nc': M6,
    'wall': OL.Model(6, OL.PRIOR_UNIFORM, OL.LIKE_GAUSS_PREC, lo=np.full(6, -5.),
                     width=np.full(6, 10.), mean=np.r_[-4.2, -4.2, -4.2, 0, 0, 0.],
                     prec=M6.p['prec'], lnorm=M6.p['lnorm']),
    'g50': OL.gauss_corr(50, 0.4, 5.), 'n200': OL.iid_normal_ppf(200),
    'egg': OL.eggbox(5), 'shell': OL.shells(4), 'g4': OL.gauss_corr(4, 0.6, 5.),
    'g3': OL.gauss_test3d(), 'g3nc': OL.gauss_test3d(), 'shell2': OL.shells(2),
}


def close(a, b, rtol=1e-9, atol=0):
    scale = max(np.abs(b).max(), 1e-300)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol + rtol * scale)
