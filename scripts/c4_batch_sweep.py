"""C4 (200-D iid normal, normal-ppf prior, single / rwalk, nlive 8000): logZ of whole device-resident runs as a
function of the round size (batch) and of the chain length (walks), next to the analytic value -253.102 and the
unmodified reference's -250.18 / -249.75 (profiles/ref_c4_rwalk_nlive8000.jsonl).  VERDICT r1 item 1(b).
usage (GPU box): python scripts/c4_batch_sweep.py [out.jsonl]"""
import json
import sys
import time

sys.path.insert(0, '.')
from dynesty_b200 import likelihoods as DL, replicas

out = open(sys.argv[1], 'a') if len(sys.argv) > 1 else sys.stdout
m = DL.iid_normal_ppf(200)
CASES = [(200, 220), (100, 220), (50, 220), (20, 220), (400, 220), (200, 440)]
for batch, walks in CASES:
    t0 = time.perf_counter()
    outs, wall = replicas.run_replicas(m, [11, 12, 13], nlive=8000, bound='single', sample='rwalk',
                                       sampler_kwargs=dict(walks=walks), max_in_flight=3, batch=batch)
    rec = dict(config='C4', nlive=8000, batch=batch, walks=walks, truth=m.logz_truth,
               logz=[round(o['logz'], 3) for o in outs], logzerr=[round(o['logzerr'], 3) for o in outs],
               niter=[o['niter'] for o in outs], ncall=[o['ncall'] for o in outs], nbound=[o['nbound'] for o in outs],
               wall_s=round(wall, 2))
    out.write(json.dumps(rec) + '\n')
    out.flush()
