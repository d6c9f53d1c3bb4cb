"""GPU: whole nested-sampling runs at the BASELINE sizes (BASELINE.json configs C2, C3, C4, C5) with the rounds
on the device, against (i) runs of the UNMODIFIED reference recorded in profiles/ref_*.jsonl (scripts/ref_*_run.py,
build container, CPU) and (ii) the analytic evidences.

Stated tolerances.  logZ of a nested-sampling run is a random variable with run-to-run scatter sigma ~ sqrt(H/nlive)
(~0.1 at C2, ~0.25 at C4, both measured on the reference); a mean over R runs is compared with the reference mean
within 3 (sigma_ref^2/R_ref + sigma^2/R)^(1/2) + 0.1, the +0.1 being north_star's own tolerance.  Both the
reference and this code are +0.4..0.5 above the analytic value at C2 (a 50-D random walk shaped by
a covariance estimated from 2000 points -- the reference's bias, reproduced), so the analytic value is only
bracketed."""
import json
import os

import numpy as np
import pytest

from dynesty_b200 import likelihoods as DL, replicas

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref(name):
    path = os.path.join(ROOT, 'profiles', name)
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return [json.loads(x) for x in f if x.strip()]


def test_c2_logz_matches_reference_runs():
    """C2: 50-D correlated Gaussian, multi / rwalk (walks 70), nlive 2000, batch nlive/40 -- 8 concurrent replicas."""
    m = DL.gauss_corr(50, 0.4, 5.0)
    outs, wall = replicas.run_replicas(m, range(56432, 56440), nlive=2000, bound='multi', sample='rwalk',
                                       sampler_kwargs=dict(walks=70), max_in_flight=8)
    lz = np.array([o['logz'] for o in outs])
    ref = _ref('ref_c2_rwalk_nlive2000.jsonl')
    assert ref is not None
    rz = np.array([r['logz'] for r in ref])
    tol = 3 * np.sqrt(rz.var(ddof=1) / len(rz) + lz.var(ddof=1) / len(lz)) + 0.1
    assert abs(lz.mean() - rz.mean()) < tol, (lz.mean(), rz.mean(), tol)
    assert lz.std(ddof=1) < 0.45                                  # per-run scatter: logzerr ~ 0.6 is the quoted error
    assert -115.129 - 0.3 < lz.mean() < -115.129 + 0.9            # bracket of the analytic value (see docstring)
    assert all(1.1e5 < o['niter'] < 1.5e5 for o in outs)          # the reference: 1.25e5 iterations


def test_c4_logz_single_ellipsoid_200d():
    """C4: 200-D iid normal, normal-ppf prior, single / rwalk (walks 220), nlive 8000.  Analytic value -253.102; the
    UNMODIFIED reference gives -250.18 / -249.75 (profiles/ref_c4_rwalk_nlive8000.jsonl): +3 from the truth, chains of
    220 steps do not decorrelate in 200-D -- the reference's bias, not the kernels' (with walks = 440 the device rounds
    give -252.1, profiles/r2c_c4_sweep.jsonl).  The device rounds reproduce the REFERENCE's value at any round size
    (batch 20 .. 400: -249.5 .. -250.1, same file): that is the parity statement, tolerance 1.0 (the runs' quoted
    logzerr is 1.4, the scatter of the sweep 0.3)."""
    m = DL.iid_normal_ppf(200)
    outs, wall = replicas.run_replicas(m, [1, 2, 3, 4], nlive=8000, bound='single', sample='rwalk',
                                       sampler_kwargs=dict(walks=220), max_in_flight=4)
    lz = np.array([o['logz'] for o in outs])
    ref = _ref('ref_c4_rwalk_nlive8000.jsonl')
    assert ref
    rz = np.array([r['logz'] for r in ref])
    assert abs(lz.mean() - rz.mean()) < 1.0, (lz, rz)
    assert lz.std(ddof=1) < 1.0
    assert all(1.2e5 < o['niter'] < 1.9e5 for o in outs)               # the reference: 1.48e5 iterations


def test_c5_static_shells_10d():
    """C5's likelihood as a static run: 10-D Gaussian shells, multi / rslice, nlive 500 (analytic -14.59,
    demos/Examples -- Gaussian Shells.ipynb:837)."""
    m = DL.shells(10)
    outs, _ = replicas.run_replicas(m, range(8), nlive=500, bound='multi', sample='rslice', max_in_flight=8)
    lz = np.array([o['logz'] for o in outs])
    err = np.mean([o['logzerr'] for o in outs])
    assert abs(lz.mean() - m.logz_truth) < 3 * err / np.sqrt(len(lz)) + 0.15, (lz.mean(), err)


def test_c3_eggbox_logz_trajectory_vs_reference():
    """C3: 25-D eggbox, multi / rslice (slices 28), nlive 4000 -- no analytic truth and a nearly flat likelihood, so
    BASELINE.md section 3 compares logZ at a FIXED iteration count with the reference run to the same maxiter
    (profiles/ref_c3_rslice_nlive4000.jsonl, scripts/ref_c3_run.py)."""
    ref = _ref('ref_c3_rslice_nlive4000.jsonl')
    if not ref:
        pytest.skip('profiles/ref_c3_rslice_nlive4000.jsonl not recorded')
    maxiter = int(ref[0]['maxiter'])
    m = DL.eggbox(25)
    outs, _ = replicas.run_replicas(m, range(4), nlive=4000, bound='multi', sample='rslice',
                                    sampler_kwargs=dict(slices=28), max_in_flight=4, dlogz=None, maxiter=maxiter,
                                    add_live=False)
    lz = np.array([o['logz'] for o in outs])
    rz = np.array([r['logz_dead'] for r in ref])
    # logZ of the dead points only after `maxiter` iterations = ln sum L_i w_i ~ 32 + ln(1 - X_end): the reference's
    # three seeds agree to 4e-5.  A round removes its K points from a shrinking live set (ln X falls by
    # ln((N-K+1)/(N+1)) per round, not K ln(N/(N+1))), so X_end = e^-10.5 here against e^-10.0 for the reference:
    # a difference of 2e-5 in logZ.  Stated tolerance 1e-3.
    assert abs(lz.mean() - rz.mean()) < 1e-3, (lz, rz)
    assert all(o['niter'] == maxiter for o in outs)


def test_c5_dynamic_shells_device_batches():
    """C5: 10-D Gaussian shells, DynamicNestedSampler(bound='multi', sample='rslice'), nlive_init = nlive_batch = 500
    (dynesty.py:701, dynamicsampler.py:1796) -- the baseline AND every batch as device rounds (dynesty_b200/dynamic.py).
    Analytic evidence -14.59; the batches must add samples (n_effective grows), live counts add up in the overlap."""
    from dynesty_b200 import dynamic as D
    m = DL.shells(10)
    lz, neff = [], []
    for seed in (1, 2, 3):
        d = D.DynamicNestedSampler(m, nlive=500, bound='multi', sample='rslice', seed=seed)
        r0 = d.sample_initial()
        n0 = D.n_effective(r0.logwt)
        res = d.run_nested(maxbatch=4, n_effective=1e9)
        assert d.batch == 4 and D.n_effective_of(res) > 1.5 * n0
        assert res.samples_n.max() >= 1000 and np.all(np.diff(res.logl) >= 0)
        lz.append(float(res.logz[-1]))
        neff.append(D.n_effective_of(res))
        err = float(res.logzerr[-1])
    lz = np.array(lz)
    assert abs(lz.mean() - m.logz_truth) < 3 * err / np.sqrt(len(lz)) + 0.15, (lz, err)
