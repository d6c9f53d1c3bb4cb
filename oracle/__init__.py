"""oracle/ -- CPU restatement of dynesty's bounding-and-proposal hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import anything from this
package, and only as the *checker* (or as the timed CPU baseline).  Nothing
under ``dynesty_b200/`` imports it; the product path fails loudly when the
CUDA library is missing.

Parity pinning: every function here cites the reference file:line it follows
(reference = /root/reference = joshspeagle/dynesty @ 99451618, dynesty 3.0.0+).
The restatement is pinned against the reference itself, imported in the build
container by ``oracle/make_golden.py`` (which drives the reference's own
functions with a scripted ``numpy.random.Generator`` subclass that replays the
B2N Philox stream) -- the resulting vectors are committed under
``tests/golden/`` and checked by ``tests/test_oracle_golden.py``.

Third-party arithmetic the reference relies on and that is NOT in
/root/reference (pyproject.toml:27-31, versions unpinned; container has
numpy 2.3.5 / scipy 1.18.1):
  * LAPACK ``?syevr`` through ``scipy.linalg.eigh``  -> oracle calls
    ``numpy.linalg.eigh`` (same LAPACK family); eigen-internals are "parity
    unpinned" beyond the invariants tested (cov = V L V^T, am = cov^-1, ...).
  * ``scipy.cluster.vq.kmeans2(minit='matrix', iter=10)`` -> restated in
    ``oracle.bounding.kmeans2_matrix`` and pinned against scipy in the tests.
"""
