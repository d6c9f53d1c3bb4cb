"""What does the launch path sustain?  empty kernels from 1 .. 32 threads, one context (stream) each."""
import ctypes as C, json, sys, threading, time
sys.path.insert(0, '.')
from dynesty_b200 import _lib
def rate(nthreads, n=20000):
    ctxs = [_lib.Context(0) for _ in range(nthreads)]
    out = [0.0] * nthreads
    def work(i):
        us = C.c_double(0.0)
        ctxs[i].check(ctxs[i].lib.b2n_debug_launch_rate(ctxs[i].h, n, C.byref(us)))
        out[i] = us.value
    t0 = time.perf_counter()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(nthreads)]
    [t.start() for t in ts]; [t.join() for t in ts]
    wall = time.perf_counter() - t0
    for c in ctxs: c.close()
    return dict(threads=nthreads, us_per_launch_per_thread=round(sum(out) / nthreads, 2), launches_per_s_total=round(nthreads * n / wall))
rate(1, 2000)
for k in (1, 2, 4, 8, 16, 32):
    print(json.dumps(rate(k)), flush=True)
