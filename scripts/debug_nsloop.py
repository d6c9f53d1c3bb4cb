"""GPU: per-round deviation between b2n_ns_run and oracle.nsloop on one test case (diagnostic)."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from oracle import nsloop
from dynesty_b200 import ops
import test_gpu_nsloop as T

kind, n, N, K, sampler, steps, two, rounds = ('gauss', 20, 96, 24, 'rwalk', 12, False, 4)
dm, om = T._models(kind, n)
rng = np.random.default_rng(100 + n + K)
u, v, l, groups = T._live(om, n, N, rng, two)
b = T._bound(groups)
o = nsloop.BatchNS(om, u, v, l, K, sampler, steps, 56432, chain0=1000, scale=0.7, logvol=-2.5, logz=-40.0,
                   loglstar=float(l.min()) - 0.5, ncall=500, bound=b, dlogz=1e-6)
ops.ns_create(dm.model_id(), N, n, K, 0, steps, 56432, chain0=1000, dlogz=1e-6, dead_capacity=rounds * K + 5)
ops.ns_set_state(u, v, l, -2.5, -40.0, float(l.min()) - 0.5, 500, 0.7)
for r in range(rounds):
    lu = o.live_u
    b = T._bound([lu])
    o.bound = b
    ops.bound_set(b['axes'], b['ctrs'], b['ams'], b['logvols'])
    before = o.live_u.copy()
    assert o.step()
    st = ops.ns_run(1, 0)
    du, dv, dl = ops.ns_get_live(N, n)
    dev = np.abs(du - o.live_u).max(axis=1)
    bad = np.nonzero(dev > 1e-9)[0]
    print('round', r, 'scale', st['scale'], o.scale, 'ncall', st['ncall'], o.ncall, 'max dev', dev.max(), 'bad rows', bad.tolist(),
          'n_acc oracle', o.last['n_accept'])
    for i in bad[:4]:
        print('   row', i, 'dev logl', dl[i], o.live_logl[i], 'thr', o.last['thr'], 'moved(oracle)', np.abs(o.live_u[i] - before[i]).max())
