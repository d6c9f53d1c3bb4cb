"""Batched-replacement nested-sampling rounds, restated on the CPU (numpy + the oracle chains).

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker of csrc/b2n_ns.cu, never on the
product path.

What is restated.  One *round* of ``b2n_ns_run`` (include/b200nest.h): remove the K lowest
live points at once, evolve K chains from uniformly chosen survivors at the threshold of the
K-th lowest, put every chain end point into a freed slot.  The pieces and the reference lines
they follow (py/dynesty/):

  worst points / threshold / termination   sampler.py:1095-1120, 1131-1140
  start row + ellipsoid of a chain          sampler.py:469-491 (propose_live),
                                            bounding.py:726-731 (get_random_axes)
  bound.contains(start) else forced update  sampler.py:485-489
  evidence increment                        utils.py:1470-1492 (progress_integration); for the
                                            j-th removal of a round the live count is N-j, the
                                            reference's rule for a shrinking live set
                                            (sampler.py:780-914: ln X -= ln((m+1)/m))
  tune with update=True once per round      internal_samplers.py:460-493, 1209-1239
  bound update due                          sampler.py:648-651

With K = 1 a round IS one iteration of the reference's loop with queue_size = 1.  Parity pins:
the chains are the oracle chains (pinned to the reference through tests/golden/chains.npz); the
quadrature is checked against the reference's ``utils.compute_integrals`` in
tests/test_oracle_nsloop.py.  The round structure itself (K > 1) has no counterpart in the
reference -- for that part "parity" is statistical (logZ against analytic truths and against
the reference's own runs), stated here and in DESIGN.md.

Random streams ("B2N-RNG v1", oracle/philox.py): chain c of round r = ChainStream(seed,
chain0 + r*K + c); the round driver = ChainStream(seed, 2^62 + r): event 0 = K uniforms for the
start rows, event 1 = K uniforms for the ellipsoid picks (only consumed when the bound has more
than one ellipsoid -- the kernel addresses events by tick, so skipping is harmless).
"""
import math

import numpy as np

from . import philox, samplers as OS

DRIVER_CHAIN = 1 << 62
LOWL = -1e300


def logaddexp(a, b):
    hi, lo = (a, b) if a > b else (b, a)
    if lo == -math.inf:
        return hi
    return hi + math.log1p(math.exp(lo - hi))


class BatchNS:
    """State + one-round step of the batched-replacement loop.

    bound: dict(ctrs (Ke, nc), ams (Ke, nc, nc), axes (Ke, nc, nc), logvols (Ke,), strict bool)
    sampler: 'rwalk' | 'rslice' | 'slice'; steps = walks / slices."""

    def __init__(self, model, live_u, live_v, live_logl, batch, sampler, steps, seed, chain0=0, facc=0.5,
                 scale=1.0, logvol=0.0, logz=LOWL, loglstar=LOWL, ncall=0, update_interval=1 << 62,
                 dlogz=0.01, maxiter=1 << 62, maxcall=1 << 62, bound=None, dimflags=None, unit_cube_phase=False,
                 first_min_ncall=0, first_min_eff=100., it0=0, logl_max=math.inf):
        self.model = model
        self.live_u = np.array(live_u, dtype=float)
        self.live_v = np.array(live_v, dtype=float)
        self.live_logl = np.array(live_logl, dtype=float)
        self.N, self.n = self.live_u.shape
        self.K = int(batch)
        assert 1 <= self.K < self.N
        self.sampler, self.steps = sampler, int(steps)
        self.seed, self.chain0 = int(seed), int(chain0)
        self.facc = min(1., max(1. / max(self.steps, 2), facc)) if sampler == 'rwalk' else facc
        self.scale, self.logvol, self.logz, self.loglstar = float(scale), float(logvol), float(logz), float(loglstar)
        self.ncall = self.ncall_last_update = int(ncall)
        self.update_interval, self.dlogz, self.maxiter, self.maxcall = update_interval, dlogz, maxiter, maxcall
        self.bound = bound
        # per-dimension boundary flags (1 periodic, 2 reflective; utils.py:950-976) -> rwalk_chain arguments
        self.per = self.ref = self.nb = None
        if dimflags is not None:
            f = np.asarray(dimflags)
            self.per = np.nonzero(f & 1)[0] if (f & 1).any() else None
            self.ref = np.nonzero(f & 2)[0] if (f & 2).any() else None
            self.nb = f == 0
        # phase 0: the rounds before the first bound draw from the prior (UnitCubeSampler,
        # internal_samplers.py:343-441) until ncall >= first_min_ncall and eff < first_min_eff (sampler.py:640-647)
        self.phase = 0 if unit_cube_phase else 1
        self.first_min_ncall, self.first_min_eff, self.it0, self.logl_max = first_min_ncall, first_min_eff, it0, logl_max
        self.error = 0
        self.it = self.round = 0
        self.done = self.need_bound = 0
        self.doubling = False
        self.delta_logz = math.inf
        self.dead = dict(u=[], v=[], logl=[], logvol=[], ncall=[])
        self.last = None

    # ------------------------------------------------------------------ helpers
    def _contains(self, x):
        b = self.bound
        nc = b['ctrs'].shape[1]
        for k in range(b['ctrs'].shape[0]):
            d = x[:nc] - b['ctrs'][k]
            d2 = float(d @ b['ams'][k] @ d)
            if (d2 < 1.0) if b['strict'] else (d2 <= 1.0):
                return True
        return False

    def bound_updated(self, bound=None):
        if bound is not None:
            self.bound = bound
        self.need_bound = 0
        self.ncall_last_update = self.ncall
        self.phase = 1

    # ------------------------------------------------------------------ one round
    def step(self):
        """Returns True if the round ran, False if a stop flag is (or became) set."""
        if self.done or self.need_bound:
            return False
        N, K, n = self.N, self.K, self.n
        order = np.lexsort((np.arange(N), self.live_logl))           # (logl, row) ascending
        sl = self.live_logl[order]
        lmax = float(sl[-1])
        self.delta_logz = logaddexp(0.0, lmax + self.logvol - self.logz)
        if (self.delta_logz < self.dlogz or self.it >= self.maxiter or self.ncall >= self.maxcall or sl[0] == lmax
                or sl[0] > self.logl_max):
            self.done = 1
            return False
        thr = float(sl[K - 1])
        # start rows need logl STRICTLY above the threshold (sampler.py:471): first sorted position above it
        first = int(np.searchsorted(sl, thr, side='right'))
        if first >= N:
            self.done, self.error = 1, 13           # B2N_ERR_PLATEAU: no live point above the threshold
            return False
        if self.phase == 0:
            return self._step_unitcube(order, sl, thr)
        if self.sampler == 'unif':
            return self._step_unif(order, sl, thr)
        drv = philox.ChainStream(self.seed, DRIVER_CHAIN + self.round)
        U = philox.event_uniforms(self.seed, drv.chain, 0, K)
        nsurv = N - first
        starts = order[first + np.minimum((U * nsurv).astype(np.int64), nsurv - 1)]
        b = self.bound
        Ke = b['ctrs'].shape[0]
        ell = np.zeros(K, dtype=np.int64)
        if Ke > 1:
            lv = b['logvols']
            m = float(lv.max())
            tot = 0.0
            for k in range(Ke):
                tot += math.exp(lv[k] - m)
            lt = m + math.log(tot)
            cum = np.cumsum(np.exp(lv - lt))
            U2 = philox.event_uniforms(self.seed, drv.chain, 1, K)
            ell = np.minimum(np.searchsorted(cum, U2), Ke - 1)
        if not all(self._contains(self.live_u[s]) for s in starts):
            self.need_bound = 2
            return False
        # ---- chains
        out = []
        warned = False
        for c in range(K):
            st = philox.ChainStream(self.seed, self.chain0 + self.round * K + c)
            u0, ax = self.live_u[starts[c]], b['axes'][ell[c]]
            if self.sampler == 'rwalk':
                r = OS.rwalk_chain(u0, thr, ax, self.scale, self.model, st, self.steps, periodic=self.per,
                                   reflective=self.ref, nonbounded=self.nb)
            elif self.sampler == 'rslice':
                r = OS.rslice_chain(u0, thr, ax, self.scale, self.model, st, self.steps, doubling=self.doubling)
            else:
                r = OS.slice_chain(u0, thr, ax, self.scale, self.model, st, self.steps, doubling=self.doubling)
            out.append(r)
            warned = warned or bool(r.get('expansion_warning_set', False))
        self._commit(order, sl, thr, out)
        # ---- tune (update=True)
        b = self.bound
        if self.sampler == 'rwalk':
            a, r_ = sum(o['n_accept'] for o in out), sum(o['n_reject'] for o in out)
            nc = b['ctrs'].shape[1]
            # K serial updates of the reference (one per iteration, internal_samplers.py:486-493) at one scale are the
            # K-th power of one update; capped at nc (loop gain).  K = 1: the reference's rule.
            self.scale *= math.exp(min(self.K, nc) * (a / (a + r_) - self.facc) / nc / self.facc)
            self.last = dict(starts=starts, ell=ell, thr=thr, n_accept=a, n_reject=r_)
        else:
            ne, ncn = sum(o['n_expand'] for o in out), sum(o['n_contract'] for o in out)
            if warned:
                self.doubling = True
            ne = max(ne, 1)
            self.scale *= min(max(ne * 2. / (ne + ncn), 0.5), 2.)
            self.last = dict(starts=starts, ell=ell, thr=thr, n_expand=ne, n_contract=ncn)
        if self.ncall >= self.ncall_last_update + self.update_interval:
            self.need_bound = 1
        return True

    def _step_unitcube(self, order, sl, thr):
        """Round of the phase before the first bound: every chain draws from the prior until logl > thr."""
        out = [OS.unitcube_chain(thr, self.model, philox.ChainStream(self.seed, self.chain0 + self.round * self.K + c),
                                 self.n) for c in range(self.K)]
        self._commit(order, sl, thr, out)
        self.last = dict(thr=thr)
        eff = 100.0 * (self.it0 + self.it) / self.ncall
        if self.ncall >= self.first_min_ncall and eff < self.first_min_eff:
            self.need_bound = 4
        return True

    def _step_unif(self, order, sl, thr):
        """Round of the uniform sampler (UniformBoundSampler.sample, internal_samplers.py:243-340): every chain
        draws from the bound until logl > thr; no start rows, nothing to tune."""
        from . import bounding as OB
        b = self.bound
        Ke = b['ctrs'].shape[0]
        me = OB.MultiEll.__new__(OB.MultiEll)
        me.ells = []
        for k in range(Ke):
            e = OB.Ell.__new__(OB.Ell)
            e.ctr, e.am, e.axes = b['ctrs'][k], b['ams'][k], b['axes'][k]
            e.ndim = len(e.ctr)
            me.ells.append(e)
        me.nells, me.ctrs, me.ams, me.logvol_ells = Ke, b['ctrs'], b['ams'], b['logvols']
        lv = b['logvols']
        m = float(lv.max())
        me.logvol = m + math.log(float(np.exp(lv - m).sum()))
        out = [OS.unif_chain(thr, me, self.model, philox.ChainStream(self.seed, self.chain0 + self.round * self.K + c),
                             self.n, nonbounded=self.nb) for c in range(self.K)]
        self._commit(order, sl, thr, out)
        self.last = dict(thr=thr)
        if self.ncall >= self.ncall_last_update + self.update_interval:
            self.need_bound = 1
        return True

    def _commit(self, order, sl, thr, out):
        """dead records + evidence (live count N - j at the j-th removal), chain end points into the freed slots"""
        N, K = self.N, self.K
        ws = np.empty(K)
        for j in range(K):
            L, Lp = float(sl[j]), (float(sl[j - 1]) if j else self.loglstar)
            lv = self.logvol + math.log((N - j) / (N + 1.0))
            ws[j] = logaddexp(L, Lp) + lv + math.log(0.5 / (N - j))
            slot = order[j]
            self.dead['u'].append(self.live_u[slot].copy())
            self.dead['v'].append(self.live_v[slot].copy())
            self.dead['logl'].append(L)
            self.dead['logvol'].append(lv)
            self.dead['ncall'].append(out[j]['ncall'])
            self.live_u[slot], self.live_v[slot], self.live_logl[slot] = out[j]['u'], out[j]['v'], out[j]['logl']
        m = float(ws.max())
        self.logz = logaddexp(self.logz, m + math.log(float(np.exp(ws - m).sum())))
        self.logvol = self.logvol + math.log((N - K + 1) / (N + 1.0))
        self.loglstar = thr
        self.it += K
        self.ncall += sum(o['ncall'] for o in out)
        self.round += 1

    def dead_arrays(self):
        d = self.dead
        return (np.array(d['u']).reshape(-1, self.n), np.array(d['v']).reshape(-1, self.n), np.array(d['logl']),
                np.array(d['logvol']), np.array(d['ncall'], dtype=np.int64))
