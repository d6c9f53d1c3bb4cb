"""GPU: full C2 runs (50-D correlated Gaussian, nlive 2000, multi, rwalk walks=70) with the rounds on the
device (run_nested(loop='device')) for several seeds / batch sizes; prints one JSON line per run.
usage: python scripts/c2_device_loop.py [batch ...]   (default 200)"""
import json
import sys
import time

sys.path.insert(0, '.')
import numpy as np
from dynesty_b200 import likelihoods as DL, nested

import os
batches = [int(x) for x in sys.argv[1:]] or [200]
SEEDS = [int(x) for x in os.environ.get('SEEDS', '1,2,3,4').split(',')]
m = DL.gauss_corr(50, 0.4, 5.0)
for K in batches:
    for seed in SEEDS:
        t0 = time.time()
        s = nested.NestedSampler(m, nlive=2000, bound='multi', sample='rwalk', seed=seed, queue_size=200)
        r = s.run_nested(loop='device', batch=K)
        print(json.dumps(dict(batch=K, seed=seed, logz=round(float(r.logz[-1]), 3), err=round(float(r.logzerr[-1]), 3),
                              truth=round(m.logz_truth, 3), niter=int(r.niter), ncall=int(r.ncall), nbound=int(r.nbound),
                              rounds=int(s.device_rounds), wall=round(time.time() - t0, 3),
                              rounds_s=round(s.device_timing['rounds_s'], 3), bound_s=round(s.device_timing['bound_s'], 3))), flush=True)
