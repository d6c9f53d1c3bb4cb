"""Concurrent replicas: many independent nested-sampling runs in flight on one GPU, spread over GPUs.

Why.  One device-resident run (``NestedSampler.run_nested(loop='device')``) is a chain of small launches: a round
of K = nlive/40 random-walk chains occupies K of the 2 x 148 resident-CTA slots of a B200 for `walks` strictly
sequential steps, and the single-CTA step / bound-update kernels in between leave the rest of the chip idle
(round-1 VERDICT: 19 % busy).  That latency chain cannot be shortened by more SMs or more GPUs -- but runs are
independent of each other, and real work comes in ensembles: the reference's own tests repeat a run over seeds
(tests/test_gau.py:104-140, tests/utils.py:12-20), a logZ error bar IS the scatter of repeats
(``dynesty.utils.resample_run`` / ``jitter_run`` exist to fake that scatter from one run), and model comparison
runs one sampler per model.  So the unit that fills the machine is the replica:

  * every replica owns a ``_lib.Context`` -- its own CUDA stream, scratch memory, resident bound and device
    run state; the library keeps no global mutable state, so contexts never synchronise with each other;
  * one host thread per replica drives ``b2n_ns_run`` / ``b2n_ns_update_bound`` (ctypes releases the GIL for
    the duration of every C call, and with the whole run on the device -- prior-draw phase included -- the
    Python work per call is microseconds);
  * the GPU interleaves the replicas' kernels: while one replica's single-CTA step kernel runs, the chain
    kernels of the others fill the SMs;
  * across GPUs replicas shard with no data-path collective at all (rank r takes replicas r, r + W, ...);
    only the per-run summaries are gathered on rank 0.  A fixed ensemble therefore scales STRONGLY with the
    number of GPUs as long as every GPU still has enough replicas in flight.
"""
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib, nested


def _one(pool, make_sampler, run_kwargs, seed, keep_results, errors='raise'):
    ctx = pool.get()                     # a context is reused by the replicas that follow each other on it: its device
    try:                                 # allocations (scratch, run state) are made once, not once per run
        t0 = time.perf_counter()
        s = make_sampler(seed, ctx)
        res = s.run_nested(loop='device', keep_samples=keep_results, **run_kwargs)
        wall = time.perf_counter() - t0
        out = dict(seed=int(seed), logz=float(res.logz[-1]), logzerr=float(res.logzerr[-1]), niter=int(res.niter),
                   ncall=int(res.ncall), nbound=int(res.nbound), rounds=int(s.device_rounds), wall_s=wall,
                   rounds_s=s.device_timing['rounds_s'], bound_s=s.device_timing['bound_s'])
        if keep_results:
            out['results'] = res
        return out
    except Exception as e:                # errors='record': one failed replica must not lose the ensemble
        if errors == 'raise':
            raise
        return dict(seed=int(seed), error=repr(e)[:300])
    finally:
        pool.put(ctx)


class ContextPool:
    """`size` contexts on one GPU, handed to the replica threads; ``chain_pack``: chains per CTA of their chain
    kernels (b2n_set_chain_pack) -- with many runs in flight a few chains per CTA leave room for everybody."""

    def __init__(self, device, size, chain_pack=1):
        import queue
        self.q = queue.Queue()
        self.all = [_lib.Context(device) for _ in range(size)]
        for c in self.all:
            if chain_pack > 1:
                c.set_chain_pack(chain_pack)
            self.q.put(c)

    def get(self):
        return self.q.get()

    def put(self, c):
        self.q.put(c)

    def close(self):
        for c in self.all:
            c.close()


def run_replicas(model, seeds, nlive=500, bound='multi', sample='rwalk', device=None, max_in_flight=16, comm=None,
                 keep_results=False, sampler_kwargs=None, chain_pack=1, pool=None, errors='raise', **run_kwargs):
    """Run one device-resident nested-sampling run per seed, `max_in_flight` at a time on this GPU.

    model / nlive / bound / sample / sampler_kwargs : as for ``nested.NestedSampler``.
    run_kwargs : passed to ``run_nested`` (dlogz, maxiter, maxcall, batch, ...).
    chain_pack : chains per CTA of the replicas' chain kernels (see ContextPool); pool: a ContextPool to reuse
           across calls (its size bounds the number of replicas in flight).
    comm : optional ``dist.Comm`` -- the seeds are dealt over the ranks (rank r: seeds[r::world]); rank 0
           returns the summaries of ALL replicas (in seed order), the other ranks their own.
    Returns (summaries, wall_seconds): one dict per replica (seed, logz, logzerr, niter, ncall, nbound, rounds,
    wall_s, rounds_s, bound_s [, results]) and the wall time of the whole ensemble on this rank."""
    seeds = [int(s) for s in seeds]
    mine = seeds if comm is None else seeds[comm.rank::comm.world]
    if device is None:
        import os
        device = int(os.environ.get('LOCAL_RANK', '0'))
    kw = dict(sampler_kwargs or {})

    def make(seed, ctx):
        return nested.NestedSampler(model, nlive=nlive, bound=bound, sample=sample, seed=seed, ctx=ctx, **kw)

    nthr = max(1, min(int(max_in_flight), len(mine))) if mine else 0
    own_pool = pool is None and nthr > 0
    if own_pool:
        pool = ContextPool(device, nthr, chain_pack)
    t0 = time.perf_counter()
    outs = []
    try:
        if mine:
            with ThreadPoolExecutor(max_workers=nthr) as ex:
                futs = [ex.submit(_one, pool, make, run_kwargs, s, keep_results, errors) for s in mine]
                outs = [f.result() for f in futs]
        wall = time.perf_counter() - t0
    finally:
        if own_pool:
            pool.close()
    if comm is not None:
        slim = [{k: v for k, v in o.items() if k != 'results'} for o in outs]
        parts = [None] * comm.world
        comm.dist.all_gather_object(parts, slim)
        if comm.rank == 0:
            by_seed = {o['seed']: o for p in parts for o in p}
            for o in outs:
                by_seed[o['seed']] = o
            outs = [by_seed[s] for s in seeds]
    return outs, wall


def summarize(outs, wall):
    """Ensemble statistics: logZ mean / scatter, aggregate calls per second over the ensemble's wall time."""
    failed = [o for o in outs if 'error' in o]
    outs = [o for o in outs if 'error' not in o]
    lz = np.array([o['logz'] for o in outs])
    ncall = int(sum(o['ncall'] for o in outs))
    return dict(replicas=len(outs), failed=len(failed), logz_mean=float(lz.mean()), logz_std=float(lz.std(ddof=1)) if len(lz) > 1 else None,
                ncall=ncall, niter=int(sum(o['niter'] for o in outs)), wall_s=wall, calls_per_s=ncall / wall,
                run_wall_s_mean=float(np.mean([o['wall_s'] for o in outs])))
