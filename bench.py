#!/usr/bin/env python
"""bench.py -- proposals/sec of the bounding-and-proposal hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A *step* is one pass of the hot path over one batch: one ``_fill_queue`` of the nested
sampler = propose `Q` start points + run `Q` random-walk chains of `walks` proposals each
inside ONE kernel against the resident multi-ellipsoid bound, at a fixed likelihood
threshold (reference sampler.py:676-717 -> internal_samplers.py:866-986).

Workload (BASELINE.json configs[1], "C2"): 50-D correlated Gaussian (rho 0.4, prior
U(-5,5)^50), nlive=2000, bound='multi', sample='rwalk' (walks = ndim+20 = 70), queue of
Q = nlive chains per step, synthetic mid-run live-point state (see make_state).

value  : proposals/s with inputs resident in HBM (device pointers, CUDA events per step).
e2e    : proposals/s through the plug-in call with HOST (pinned) buffers: H2D of the start
         points + D2H of (u, v, logl, counters) inside the timed region, every step.
roofline: algorithmic bytes B_rwalk(n) = 16 n^2 + 24 n per proposal
         x proposals per launch / kernel time (events on the launch stream) vs measured HBM peak.
cpu_baseline / --impl reference: the oracle port of the reference's pure-Python rwalk chain
         (oracle/samplers.py) on the host cores, bounded sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 56432
WORKLOADS = {
    # name: (ndim, nlive, sampler, steps-per-chain, bound)
    'c2': dict(ndim=50, nlive=2000, sample='rwalk', walks=70, bound='multi',
               desc="50-D correlated Gaussian, bound=multi, sample=rwalk, nlive=2000"),
}
METRIC = "rwalk proposals/sec (50-D correlated Gaussian, multi-ellipsoid bound)"


# ----------------------------------------------------------------------------- workload
def make_state(ndim, nlive, seed=SEED):
    """Synthetic mid-run nested-sampling state: `nlive` points uniform inside the iso-likelihood
    ellipsoid (v-mu)^T Cinv (v-mu) < r^2 with r^2 = ndim, so every live point has
    logl > loglstar = lnorm - r^2/2 (what the live set looks like when the run has
    compressed to the bulk of the posterior)."""
    rng = np.random.default_rng(seed)
    Cm = np.full((ndim, ndim), 0.4)
    np.fill_diagonal(Cm, 1.0)
    L = np.linalg.cholesky(Cm)
    z = rng.standard_normal((nlive, ndim))
    z *= (rng.random(nlive)**(1. / ndim) / np.linalg.norm(z, axis=1))[:, None]
    r = math.sqrt(ndim)
    v = r * z @ L.T
    u = (v + 5.0) / 10.0
    lnorm = -0.5 * (math.log(2 * math.pi) * ndim + np.linalg.slogdet(Cm)[1])
    loglstar = lnorm - 0.5 * r * r
    return np.ascontiguousarray(u), loglstar


def algorithmic_bytes(n):
    return 16 * n * n + 24 * n      # B_rwalk(n), GAUSS_PREC likelihood


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------- CPU arm
_REF = {}


def reference_kind():
    """"reference" when the unmodified dynesty is importable on this box (the git-ignored offline
    install baseline/_ref, or /root/reference in the build container), else "port" (the oracle)."""
    from oracle import refshim
    return 'reference' if refshim.available() else 'port'


def _cpu_worker_ref(args):
    """The UNMODIFIED reference: dynesty.internal_samplers.RWalkSampler.sample(SamplerArgument)
    per chain -- the static method dynesty's pool maps over the queue (sampler.py:717) -- with
    utils.LogLikelihood around the notebook's numpy likelihood and numpy's PCG64 generator."""
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    u0s, loglstar, axes, scale, walks, chain0, ndim = args
    if 'mod' not in _REF:
        from oracle import refshim
        dynesty = refshim.import_reference()
        from dynesty import internal_samplers as RIS, utils as RU
        Cm = np.full((ndim, ndim), 0.4)
        np.fill_diagonal(Cm, 1.0)
        Cinv = np.linalg.inv(Cm)
        lnorm = -0.5 * (math.log(2 * math.pi) * ndim + np.linalg.slogdet(Cm)[1])
        _REF.update(mod=RIS, ptform=lambda u: 10. * u - 5.,
                    logl=RU.LogLikelihood(lambda x: -0.5 * np.dot(x, np.dot(Cinv, x)) + lnorm, ndim))
    RIS = _REF['mod']
    kw = {'walks': walks, 'ncdim': ndim, 'nonbounded': None, 'periodic': None, 'reflective': None}
    nacc = 0
    for i, u0 in enumerate(u0s):
        a = RIS.SamplerArgument(u=u0, loglstar=loglstar, axes=axes, scale=scale, prior_transform=_REF['ptform'],
                                loglikelihood=_REF['logl'], rseed=SEED + chain0 + i, kwargs=kw)
        r = RIS.RWalkSampler.sample(a)
        nacc += r.proposal_stats['n_accept']
    return len(u0s) * walks, nacc


def _cpu_worker(args):
    """Oracle port of the reference's per-chain pure-Python loop (what dynesty.pool.Pool maps)."""
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from oracle import samplers as OS, philox, likelihoods as OL
    u0s, loglstar, axes, scale, walks, chain0, ndim = args
    m = OL.gauss_corr(ndim, 0.4, 5.0)
    nacc = 0
    for i, u0 in enumerate(u0s):
        r = OS.rwalk_chain(u0, loglstar, axes, scale, m, philox.NumpyStream(SEED, chain0 + i), walks)
        nacc += r['n_accept']
    return len(u0s) * walks, nacc


def cpu_sample(cfg, target_seconds, pool, cores, state=None, kind='port'):
    """Times the reference's (kind="reference") or the oracle's (kind="port") rwalk chains on
    `cores` processes for ~target_seconds."""
    from oracle import bounding as OB
    worker = _cpu_worker_ref if kind == 'reference' else _cpu_worker
    u, loglstar = state if state is not None else make_state(cfg['ndim'], cfg['nlive'])
    ell = OB.bounding_ellipsoid(u)
    ell.scale_to_logvol(ell.logvol + math.log(1.25))
    rng = np.random.default_rng(1)
    scale, walks, n = 0.15, cfg['walks'], cfg['ndim']
    # pilot on every process at once (import + contention included) to size the bounded sample
    pilot = [(u[:4], loglstar, ell.axes, scale, walks, c * 4, n) for c in range(cores)]
    _ = pool.map(worker, pilot) if pool is not None else [worker(t) for t in pilot]    # imports, untimed
    t0 = time.perf_counter()
    _ = pool.map(worker, pilot) if pool is not None else [worker(t) for t in pilot]
    per_chain = (time.perf_counter() - t0) / 4
    per_core = max(4, min(int(target_seconds / per_chain), 20000))
    tasks = []
    for c in range(cores):
        starts = u[rng.integers(len(u), size=per_core)]
        tasks.append((starts, loglstar, ell.axes, scale, walks, 10**6 + c * per_core, n))
    t0 = time.perf_counter()
    res = pool.map(worker, tasks) if pool is not None else [worker(t) for t in tasks]
    dt = time.perf_counter() - t0
    nprop = sum(r[0] for r in res)
    return nprop / dt, nprop, dt, per_core * cores


def run_reference(args, cfg):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = host_cores()
    os.environ['OMP_NUM_THREADS'] = '1'
    pool = mp.get_context('fork').Pool(cores) if cores > 1 else None
    state = make_state(cfg['ndim'], cfg['nlive'])
    kind = reference_kind()
    per_step = float(os.environ.get('B2N_BENCH_CPU_SECONDS', max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))))
    for _ in range(args.warmup):
        cpu_sample(cfg, per_step, pool, cores, state, kind)
    tot_p = tot_t = 0.0
    nchains = 0
    for _ in range(args.steps):
        _, p, t, nch = cpu_sample(cfg, per_step, pool, cores, state, kind)
        tot_p += p
        tot_t += t
        nchains = nch
    if pool is not None:
        pool.close()
    val = tot_p / tot_t
    who = ("dynesty RWalkSampler.sample (unmodified reference, baseline/_ref)" if kind == 'reference'
           else "oracle rwalk")
    sample = "%d %s chains x %d walks per step on %d processes" % (nchains, who, cfg['walks'], cores)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "proposals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": cfg['desc'], "queue_chains": nchains},
            "cpu_baseline": {"value": val, "unit": "proposals/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "proposals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(device), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '20'],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        out = self.p.communicate()[0]
        sm, mx, reasons = [], None, set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        load = sorted(sm)[len(sm) // 2:] if sm else []          # upper half = samples under load
        return {"sm_mhz": (sorted(load)[len(load) // 2] if load else None), "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- GPU arm
def run_b200(args, cfg):
    import torch
    import torch.distributed as dist
    from dynesty_b200 import _lib, ops, likelihoods as DL, bounding as B

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    # stdout carries exactly ONE line (the JSON): while the benchmark runs, file descriptor 1 points at stderr,
    # so that banners printed by native libraries (NCCL's "NCCL version ..." goes to stdout) cannot precede it
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    ctx = _lib.Context(local)
    n, nlive, walks = cfg['ndim'], cfg['nlive'], cfg['walks']
    Q = args.chains or nlive
    model = DL.gauss_corr(n, 0.4, 5.0)
    mid = model.model_id(ctx)
    u_live, loglstar = make_state(n, nlive)

    # ---- the bounding half of the path: build + enlarge the bound on the GPU (timed separately)
    bound = B.B200MultiEllipsoid(n, ctx=ctx)
    t0 = time.perf_counter()
    bound.update(u_live, rstate=np.random.default_rng(SEED))
    bound.scale_to_logvol(bound.logvol + math.log(1.25))
    bound_ms_first = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for _ in range(3):
        bound.update(u_live, rstate=np.random.default_rng(SEED))
        bound.scale_to_logvol(bound.logvol + math.log(1.25))
    bound_ms = 1e3 * (time.perf_counter() - t0) / 3
    bound.make_resident()
    rng = np.random.default_rng(SEED + rank)

    # ---- buffers
    pin = lambda *s, dt=torch.float64: torch.empty(*s, dtype=dt).pin_memory()
    h_u0 = pin(Q, n)
    h_out = dict(u=pin(Q, n), v=pin(Q, n), logl=pin(Q), n_accept=pin(Q, dt=torch.int32),
                 n_reject=pin(Q, dt=torch.int32), ncall=pin(Q, dt=torch.int32))
    h_np = {k: t.numpy() for k, t in h_out.items()}
    d_live = torch.from_numpy(u_live).to(dev)
    d_u0 = torch.empty(Q, n, dtype=torch.float64, device=dev)
    d_pack = torch.empty(Q * n + Q, dtype=torch.float64, device=dev)     # (u | logl): ONE all-gather per fill
    d_out = dict(u=d_pack[:Q * n].view(Q, n), v=torch.empty(Q, n, dtype=torch.float64, device=dev),
                 logl=d_pack[Q * n:],
                 n_accept=torch.empty(Q, dtype=torch.int32, device=dev),
                 n_reject=torch.empty(Q, dtype=torch.int32, device=dev),
                 ncall=torch.empty(Q, dtype=torch.int32, device=dev))
    fused = world > 1 and args.exchange == 'fused'
    if world > 1:
        g_pack = torch.empty(world * (Q * n + Q), dtype=torch.float64, device=dev)
    if fused:
        # the exchange step fused into the kernel: finished chains are stored into every rank's
        # window over NVLink and the kernel ends with a cross-GPU arrive/wait (csrc/b2n_peer.cu)
        from dynesty_b200.dist import Comm
        Comm(dev).attach_peer(ctx, world * Q, n)
        hg_out = dict(u=pin(world * Q, n), v=pin(world * Q, n), logl=pin(world * Q),
                      n_accept=pin(world * Q, dt=torch.int32), n_reject=pin(world * Q, dt=torch.int32),
                      ncall=pin(world * Q, dt=torch.int32))
        hg_np = {k: t.numpy() for k, t in hg_out.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    # one explicit (non-default) stream shared by torch and the library, so that torch's CUDA
    # events bracket the library's launches (a NULL handle would mean "library-owned stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    ctx.set_timing(True)

    state = dict(scale=0.2, chain=0)

    def propose():
        starts = rng.integers(nlive, size=Q)
        ell = bound.random_ells(rng, Q)
        return starts, ell

    def step_host():
        """The plug-in call with host buffers (what Sampler._fill_queue does per fill)."""
        starts, ell = propose()
        np.take(u_live, starts, axis=0, out=h_u0.numpy(), mode='clip')    # ('raise' buffers `out`: 3x slower)
        ctx.set_pointer_mode(_lib.PTR_HOST)
        c0 = state['chain']
        state['chain'] += Q * world
        if fused:         # every rank's host receives the whole queue (replicated sampler state)
            o = ops.rwalk_batch(mid, h_u0.numpy(), loglstar, state['scale'], walks, SEED, chain0=c0 + rank * Q,
                                ell=ell, ctx=ctx, out=hg_np, peer=(rank * Q, world * Q))
            return {k: v[rank * Q:(rank + 1) * Q] for k, v in o.items()}
        o = ops.rwalk_batch(mid, h_u0.numpy(), loglstar, state['scale'], walks, SEED, chain0=c0 + rank * Q,
                            ell=ell, ctx=ctx, out=h_np)
        return o

    def prep_dev():
        """Inputs of the next device-resident step: start points gathered from the live set in HBM
        (outside the timed events: `value` starts with its inputs resident)."""
        starts, ell = propose()
        torch.index_select(d_live, 0, torch.from_numpy(starts).to(dev, non_blocking=True), out=d_u0)
        return ell

    def step_dev(ell=None):
        """Same step with the inputs already resident in HBM (device pointers, async)."""
        if ell is None:
            ell = prep_dev()
        ctx.set_pointer_mode(_lib.PTR_DEVICE)
        c0 = state['chain']
        state['chain'] += Q * world
        if fused:           # the exchange step of the sharded path, inside the kernel
            ops.rwalk_batch(mid, d_u0, loglstar, state['scale'], walks, SEED, chain0=c0 + rank * Q, ell=ell,
                            ctx=ctx, out=ops.NO_OUT, peer=(rank * Q, world * Q))
            return
        ops.rwalk_batch(mid, d_u0, loglstar, state['scale'], walks, SEED, chain0=c0 + rank * Q, ell=ell,
                        ctx=ctx, out=d_out)
        if world > 1:       # --exchange nccl: one packed all-gather of (u | logl) after the kernel
            dist.all_gather_into_tensor(g_pack, d_pack)

    # ---- warm-up: also tunes the proposal scale with the reference's rule (internal_samplers.py:491)
    clocks = ClockSampler(local)         # samples every 20 ms from here to the end of the timed regions
    for _ in range(max(args.warmup, 3)):
        o = step_host()
        acc, rej = int(o['n_accept'].sum()), int(o['n_reject'].sum())
        state['scale'] *= math.exp((acc / (acc + rej) - 0.5) / n / 0.5)
    for _ in range(3):
        o = step_host()
    accept_frac = float(o['n_accept'].sum() / (o['n_accept'].sum() + o['n_reject'].sum()))
    # clock ramp: ~0.7 s of the same kernel before timing, so that the nvidia-smi sampler has
    # samples under load.  N>1: a FIXED step count (every rank must issue the same exchanges).
    if world > 1:
        for _ in range(2500):
            step_dev()
    else:
        t_ramp = time.perf_counter() + args.ramp
        while time.perf_counter() < t_ramp:
            step_dev()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed: device-resident (value + roofline)
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    barrier()
    for a, b in ev:
        ell = prep_dev()
        flush.zero_()                        # L2 flush between timed iterations (outside the events)
        a.record(stream)
        step_dev(ell)
        b.record(stream)
        b.synchronize()
        kern_ms.append(ctx.last_kernel_ms())
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    # ---- timed: end to end through the plug-in call with host buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    launches = ctx.launch_count() - launches0
    clk = clocks.stop()
    if fused:
        ctx.peer_check()                     # raises if any in-kernel exchange ever timed out

    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    props_per_step = Q * walks * world
    value = props_per_step * args.steps / (dev_ms * 1e-3)
    e2e = props_per_step * args.steps / e2e_s

    line = None
    if rank == 0:
        peaks, peak_src = None, "fallback 6650 GB/s (B200_PROFILING.md)"
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peaks = json.load(f)
            peak, peak_src = float(peaks['hbm_gbs']), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        except Exception:
            peak = 6650.0
        kms = float(np.mean(kern_ms))
        achieved = algorithmic_bytes(n) * Q * walks / (kms * 1e-3) / 1e9
        traffic = None          # measured DRAM bytes per launch of this kernel (ncu --set full)
        try:
            with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
                traffic = json.load(f)['rwalk_mma_kernel']['dram_bytes_per_launch'] if Q == 2000 else None
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "proposals/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg['desc'], "queue_chains_per_gpu": Q, "walks": walks, "nells": int(bound.nells),
                       "accept_fraction": round(accept_frac, 3), "scale": round(state['scale'], 4),
                       "l2": "flushed (256 MB memset) between timed iterations",
                       "exchange": ("fused into the kernel: NVLink peer stores + in-kernel arrive/wait" if fused else
                                    ("NCCL all-gather after the kernel" if world > 1 else "none (1 GPU)")),
                       "bound_update_ms": round(bound_ms, 3), "bound_update_first_ms": round(bound_ms_first, 2)},
            "e2e": {"value": e2e, "unit": "proposals/s",
                    "h2d_bytes_per_step": Q * n * 8 + Q * 4 + 16 * (Q // 8 + 1),
                    "d2h_bytes_per_step": (world if fused else 1) * (2 * Q * n * 8 + Q * 8 + 3 * Q * 4),
                    "bytes_are": "per rank"},
            "gpu_launches": int(launches),
            "accepted_proposals_per_s": value * accept_frac,      # rwalk n_accept / wall
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes/launch",
                         "algorithmic_bytes_per_launch": algorithmic_bytes(n) * Q * walks,
                         "kernel": "rwalk_mma_kernel<GAUSS_PREC, KT=13, 8 chains/CTA, ring 8>",
                         "note": ("frac > 1: SURVEY 8(d)'s byte model counts both 20 KB matrices per proposal (no reuse); the "
                                  "kernel keeps them in registers, so HBM is idle (traffic) and the kernel is latency-bound "
                                  "at the problem's parallelism (DESIGN.md 9.1)"),
                         "kernel_ms": kms, "algorithmic_bytes_per_proposal": algorithmic_bytes(n),
                         "peak_source": peak_src},
            # the byte model above assumes no reuse; the honest ceiling of this kernel is FP64 issue: SURVEY 8(d)'s
            # 4n^2 + 7n flop per proposal against the chip's nominal FP64 tensor (DMMA) rate
            "roofline_fp64": {"achieved_tflops": (4 * n * n + 7 * n) * Q * walks / (kms * 1e-3) / 1e12,
                              "peak_tflops": 45.0, "peak_source": "nominal B200 FP64 tensor rate (blackwell_cuda_programming.md:52)",
                              "frac": (4 * n * n + 7 * n) * Q * walks / (kms * 1e-3) / 45e12,
                              "flops_per_proposal": 4 * n * n + 7 * n},
            "clocks": clk,
        }
    if world > 1:
        dist.barrier()

    # ---- the second half of the metric: logZ error of full runs (not in the timed region)
    if rank == 0 and world == 1 and args.logz:
        from dynesty_b200 import nested
        ctx.set_timing(False)
        ctx.set_stream(None)
        ctx.set_pointer_mode(_lib.PTR_HOST)
        # Full C2 runs with the rounds paced on the device (b2n_ns_run): each round replaces the
        # `batch` lowest live points.  rwalk chains at 50-D use proposal shapes estimated from the live
        # points themselves and mix slowly along under-estimated directions (true of the reference too:
        # its own logZ is +0.4 off the analytic value at nlive = 2000); the resulting bias grows with the
        # fraction of the live set replaced per round, and batch = nlive/40 reproduces the reference's
        # serial (queue_size = 1) result -- see DESIGN.md section 9.4.  The throughput steps above use
        # Q = nlive chains per launch.
        batch = args.logz_batch or max(1, nlive // 40)
        runs = []
        for k in range(args.logz):
            t0 = time.perf_counter()
            ns = nested.NestedSampler(model, nlive=nlive, bound=cfg['bound'], sample=cfg['sample'], walks=walks,
                                      seed=SEED + k, ctx=ctx, queue_size=args.logz_queue)
            res = ns.run_nested(loop='device', batch=batch)
            wall = time.perf_counter() - t0
            runs.append({"seed": SEED + k, "logz": float(res.logz[-1]), "logzerr": float(res.logzerr[-1]),
                         "niter": int(res.niter), "ncall": int(res.ncall), "nbound": int(res.nbound),
                         "rounds": int(ns.device_rounds), "wall_s": round(wall, 3),
                         "rounds_s": round(ns.device_timing['rounds_s'], 3),
                         "bound_s": round(ns.device_timing['bound_s'], 3)})
        lz = np.array([r["logz"] for r in runs])
        best = min(r["wall_s"] for r in runs)
        line["logz"] = {"loop": "device rounds (b2n_ns_run), K-worst replacement", "batch": batch, "runs": runs,
                        "logz_mean": float(lz.mean()), "logz_std": float(lz.std(ddof=1)) if len(lz) > 1 else None,
                        "truth": model.logz_truth, "abs_err_mean": abs(float(lz.mean()) - model.logz_truth),
                        "wall_s_best": best, "calls_per_s": runs[-1]["ncall"] / runs[-1]["wall_s"],
                        "iterations_per_s": runs[-1]["niter"] / runs[-1]["wall_s"]}
        try:        # the UNMODIFIED reference on this config (CPU, serial), recorded by scripts/ref_c2_run.py
            with open(os.path.join(ROOT, 'profiles', 'ref_c2_rwalk_nlive2000.jsonl')) as f:
                ref = [json.loads(x) for x in f if x.strip()]
            rz = np.array([r["logz"] for r in ref])
            line["logz"]["reference"] = {"logz_mean": float(rz.mean()), "logz_std": float(rz.std(ddof=1)), "runs": len(rz),
                                         "wall_s_mean": float(np.mean([r["wall"] for r in ref])),
                                         "source": "profiles/ref_c2_rwalk_nlive2000.jsonl (dynesty, 1 CPU core, build container)"}
            line["logz"]["mean_minus_reference_mean"] = float(lz.mean() - rz.mean())
        except Exception:
            pass
    # ---- CPU baseline on the host cores (rank 0, N=1 only), bounded sample
    if rank == 0 and world == 1 and args.cpu_baseline:
        import multiprocessing as mp
        cores = host_cores()
        os.environ['OMP_NUM_THREADS'] = '1'
        pool = mp.get_context('fork').Pool(cores) if cores > 1 else None
        kind = reference_kind()
        v, p, tsec, nch = cpu_sample(cfg, 12.0, pool, cores, (u_live, loglstar), kind)
        if pool is not None:
            pool.close()
        v1, _, t1, nch1 = cpu_sample(cfg, 3.0, None, 1, (u_live, loglstar), kind)   # (i) one core, serial
        who = "dynesty RWalkSampler.sample (unmodified reference)" if kind == 'reference' else "oracle rwalk"
        line["cpu_baseline"] = {"value": v, "unit": "proposals/s", "cores": cores, "kind": kind,
                                "sample": "%d %s chains x %d walks (%.1f s) on %d processes" % (nch, who, walks, tsec, cores),
                                "one_core_value": v1, "one_core_sample": "%d chains (%.1f s), serial" % (nch1, t1)}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)               # (teardown messages of native libraries: not on stdout either)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--chains', type=int, default=0, help='chains per step per GPU (default nlive)')
    ap.add_argument('--logz', type=int, default=8, help='number of full C2 nested-sampling runs (seeds) for logZ; 0 = none')
    ap.add_argument('--logz-queue', type=int, default=200, help='queue_size of the host (unit-cube) phase of the logZ runs')
    ap.add_argument('--logz-batch', type=int, default=0, help='points replaced per device round (default nlive/40)')
    ap.add_argument('--cpu-baseline', type=int, default=1)
    ap.add_argument('--ramp', type=float, default=0.7, help='seconds of untimed steps before timing (clock ramp; 0 under ncu)')
    ap.add_argument('--exchange', default='fused', choices=['fused', 'nccl'],
                    help='N>1: how the finished chains reach every rank')
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == '__main__':
    main()
