"""GPU: device model registry (prior transform + log-likelihood) vs the oracle."""
import numpy as np
import pytest

from helpers import MODELS, device_model, close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['g6', 'wall', 'g50', 'n200', 'egg', 'shell', 'g3', 'shell2'])
def test_model_eval(name):
    m = MODELS[name]
    dm = device_model(m)
    rng = np.random.default_rng(3)
    u = rng.random((513, m.ndim))
    if name in ('g6', 'g50', 'g3'):
        u = 0.5 + 0.1 * (u - 0.5)
    v, logl = dm.evaluate(u)
    v0 = m.prior_transform(u)
    close(v, v0, rtol=1e-13)
    l0 = m.loglike(v0)
    np.testing.assert_allclose(logl, l0, rtol=1e-11, atol=1e-11)
    # host-callable views used by the dynesty drop-in (single point)
    assert abs(dm.loglikelihood(v0[7]) - l0[7]) <= 1e-11 * max(1, abs(l0[7]))
    close(dm.prior_transform(u[7]), v0[7], rtol=1e-13)
