"""A/B timing of the multi-ellipsoid bound update (GPU box): speculative root fit (B2N_BOUND_SPEC), k-means rows staged
in shared memory + thread-per-row iteration (B2N_KM_STAGE), the candidate fit as two stream-parallel halves
(B2N_CHOL_SPLIT), candidates' stats read once per update (B2N_BOUND_DEFER).
usage: python scripts/bound_ab.py > gpurun_out/bound_ab.jsonl"""
import json
import math
import os
import sys
import time

sys.path.insert(0, '.')
import numpy as np
import torch

from dynesty_b200 import _lib, ops, bounding as B
import bench

SEED = 56432


def clouds():
    rng = np.random.default_rng(SEED)
    u, _ = bench.make_state(50, 2000)
    yield 'C2 live set 2000x50 (bench.make_state)', u
    ctrs = 0.2 + 0.6 * rng.random((8, 25))
    yield '8 clusters 4000x25 (C3 shape)', np.concatenate([c + 0.01 * rng.standard_normal((500, 25)) for c in ctrs])
    yield '2 clusters 500x10 (C5 shape)', np.concatenate([0.3 + 0.02 * rng.standard_normal((250, 10)), 0.7 + 0.02 * rng.standard_normal((250, 10))])


def timed(fn, reps):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ctx = _lib.default_context()
    off = dict(B2N_BOUND_SPEC='0', B2N_KM_STAGE='0', B2N_CHOL_SPLIT='0', B2N_BOUND_DEFER='0')
    combos = [('all off', off), ('spec', dict(off, B2N_BOUND_SPEC='1')), ('km stage', dict(off, B2N_KM_STAGE='1')),
              ('chol split', dict(off, B2N_CHOL_SPLIT='1')), ('defer', dict(off, B2N_BOUND_DEFER='1')),
              ('all on', {k: '1' for k in off})]
    for tag, pts in clouds():
        N, n = pts.shape
        d = torch.from_numpy(pts).cuda()
        K = max(1, N // (2 * n))
        outs = [torch.empty(s, dtype=torch.float64, device='cuda') for s in ((K, n), (K, n, n), (K, n, n), (K, n, n), (K, n), (K,))]
        import ctypes as C
        nells, warn = C.c_int32(0), C.c_uint32(0)

        def dev():
            ctx.set_pointer_mode(_lib.PTR_DEVICE)
            try:
                ctx.check(ctx.lib.b2n_multi_decompose(ctx.h, d.data_ptr(), N, n, K, C.addressof(nells), None, outs[0].data_ptr(),
                                                      outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), outs[4].data_ptr(),
                                                      outs[5].data_ptr(), C.addressof(warn)))
            finally:
                ctx.set_pointer_mode(_lib.PTR_HOST)

        def host():
            return ops.multi_decompose(pts)

        bound = B.B200MultiEllipsoid(n, ctx=ctx)

        def plugin():
            bound.update(pts, rstate=np.random.default_rng(SEED))
            bound.scale_to_logvol(bound.logvol + math.log(1.25))

        for name, env in combos:
            os.environ.update(env)
            os.environ['B2N_BOUND_FAST'] = '1'
            o = host()
            rec = dict(cloud=tag, variant=name, nells=int(o['nells']), logvol0=float(o['logvols'][0]))
            rec['device_resident_ms_median'], rec['device_resident_ms_min'] = timed(dev, 30)
            rec['host_arrays_ms_median'], rec['host_arrays_ms_min'] = timed(host, 20)
            rec['plugin_update_plus_enlarge_ms_median'], rec['plugin_update_plus_enlarge_ms_min'] = timed(plugin, 20)
            print(json.dumps(rec), flush=True)


if __name__ == '__main__':
    main()
