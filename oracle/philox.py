"""Philox4x32-10 + the B2N chain random stream, restated on the CPU (numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws its randomness from ``numpy.random.Generator(PCG64)``
(utils.py:993-999) -- a *sequential* generator that cannot be replayed on a
GPU.  The B200 kernels instead use the counter-based Philox4x32-10 (Salmon et
al. 2011, "Parallel random numbers: as easy as 1, 2, 3"; the same function
cuRAND exposes as ``curand_Philox4x32_10`` in curand_philox4x32_x.h).  This
module restates (a) the block function, pinned by the Random123 known-answer
vectors in tests/test_oracle_philox.py, and (b) the B2N stream layout, so that
the oracle *and the unmodified reference* (through ``ScriptedGenerator``) can
consume exactly the numbers a CUDA chain consumes.

B2N stream layout ("B2N-RNG v1")
--------------------------------
key      = (seed & 0xffffffff, seed >> 32)
counter  = (blk, tick, chain & 0xffffffff, chain >> 32)
   chain : global 64-bit chain id (one per proposal chain / queue slot)
   tick  : index of the draw *event* inside the chain (0, 1, 2, ...); every
           call the reference makes on its generator is one event
   blk   : 4-word block index inside a vector-valued event
out      = philox4x32_10(counter, key) = (r0, r1, r2, r3)
uniforms : U0 = u53(r0, r1), U1 = u53(r2, r3),
           u53(a, b) = ((a >> 6) * 2^26 + (b >> 6) + 0.5) * 2^-52   in (0, 1)
           (52 random bits + half-ulp offset: exactly representable, never 0 or 1)
   vector uniform event of size m: element e = block e // 2, slot e % 2
   scalar uniform event: block 0, slot 0
normals  : block b -> (z[2b], z[2b+1]) = R cos(2 pi U1), R sin(2 pi U1),
           R = sqrt(-2 ln U0)   (Box-Muller)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)
TWO_M52 = 2.0**-52


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32-valued array, key: (..., 2).  Returns (..., 4) uint32."""
    c = np.asarray(ctr, dtype=np.uint64) & MASK32
    k = np.asarray(key, dtype=np.uint64) & MASK32
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    k0, k1 = k[..., 0], k[..., 1]
    for _ in range(10):
        p0 = M0 * c0          # 64-bit products of 32-bit values: no overflow
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0), lo1, (hi0 ^ c3 ^ k1), lo0
        k0 = (k0 + np.uint64(W0)) & MASK32
        k1 = (k1 + np.uint64(W1)) & MASK32
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def u53(a, b):
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    return ((a >> np.uint64(6)).astype(np.float64) * 67108864.0 +
            (b >> np.uint64(6)).astype(np.float64) + 0.5) * TWO_M52


def event_blocks(seed, chain, tick, nblk):
    """Raw Philox output of the first `nblk` blocks of event (chain, tick)."""
    seed = int(seed)
    chain = int(chain)
    ctr = np.empty((nblk, 4), dtype=np.uint64)
    ctr[:, 0] = np.arange(nblk, dtype=np.uint64)
    ctr[:, 1] = int(tick) & 0xFFFFFFFF
    ctr[:, 2] = chain & 0xFFFFFFFF
    ctr[:, 3] = (chain >> 32) & 0xFFFFFFFF
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF],
                   dtype=np.uint64)
    return philox4x32_10(ctr, key)


def event_uniforms(seed, chain, tick, m):
    if m == 0:
        return np.empty(0)
    r = event_blocks(seed, chain, tick, (m + 1) // 2)
    u = np.stack([u53(r[:, 0], r[:, 1]), u53(r[:, 2], r[:, 3])],
                 axis=1).reshape(-1)
    return u[:m]


def event_normals(seed, chain, tick, m):
    if m == 0:
        return np.empty(0)
    r = event_blocks(seed, chain, tick, (m + 1) // 2)
    u0 = u53(r[:, 0], r[:, 1])
    u1 = u53(r[:, 2], r[:, 3])
    rad = np.sqrt(-2.0 * np.log(u0))
    ang = 2.0 * np.pi * u1
    z = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1).reshape(-1)
    return z[:m]


class ChainStream:
    """The draw events of one chain, in order (see module docstring)."""

    def __init__(self, seed, chain, tick=0):
        self.seed = int(seed)
        self.chain = int(chain)
        self.tick = int(tick)

    def uniform(self):
        v = event_uniforms(self.seed, self.chain, self.tick, 1)[0]
        self.tick += 1
        return float(v)

    def uniforms(self, m):
        if m == 0:          # a size-0 draw is not an event (numpy consumes nothing)
            return np.empty(0)
        v = event_uniforms(self.seed, self.chain, self.tick, m)
        self.tick += 1
        return v

    def normals(self, m):
        if m == 0:
            return np.empty(0)
        v = event_normals(self.seed, self.chain, self.tick, m)
        self.tick += 1
        return v

    def integers(self, n, m):
        """m integers in [0, n): floor(U * n) of one uniform vector event."""
        return np.minimum((self.uniforms(m) * n).astype(np.int64), n - 1)

    def permutation(self, m):
        """Random permutation of range(m): argsort of one uniform vector event
        (ties have probability ~2^-53 and are broken by index, stable sort)."""
        if m <= 1:
            return np.arange(m)
        v = self.uniforms(m)
        return np.argsort(v, kind='stable')


class ScriptedGenerator(np.random.Generator):
    """A real ``numpy.random.Generator`` (passes utils.py:997 isinstance test)
    whose draws replay a B2N ChainStream.  Passing it as ``rseed`` to the
    *unmodified* reference samplers (internal_samplers.py:558, 652, 802) makes
    the reference consume exactly the numbers a CUDA chain consumes.

    Only the methods the reference's hot path calls are scripted:
      random()/random(m)       bounding.py:1295, internal_samplers.py:1011,1099,1151,1173
      standard_normal(size=n)  bounding.py:1291, internal_samplers.py:820
      uniform(size=m)          internal_samplers.py:327
      shuffle(idxs)            internal_samplers.py:674
    """

    def __init__(self, seed, chain):
        super().__init__(np.random.PCG64(0))
        self._s = ChainStream(seed, chain)

    @property
    def tick(self):
        return self._s.tick

    def random(self, size=None, *a, **k):
        if size is None:
            return self._s.uniform()
        return self._s.uniforms(int(np.prod(size))).reshape(size)

    def uniform(self, low=0.0, high=1.0, size=None):
        """internal_samplers.py:327 (0, 1); bounding.py:1082 ``uniform(-1, 1, size=ndim)`` (SupFriends)."""
        r = self.random(size)
        return r if (low == 0.0 and high == 1.0) else low + (high - low) * r

    def standard_normal(self, size=None, *a, **k):
        if size is None:
            return float(self._s.normals(1)[0])
        return self._s.normals(int(np.prod(size))).reshape(size)

    def shuffle(self, x, axis=0):
        p = self._s.permutation(len(x))
        x[:] = np.asarray(x)[p]

    def integers(self, low, high=None, size=None, **k):
        """bounding.py:1603 ``rstate.integers(npoints, size=npoints)``:
        element e = floor(U_e * npoints) of one uniform vector event."""
        assert high is None
        if size is None:            # bounding.py:819, 1089 ``rstate.integers(nctrs)``: one uniform event
            return int(self._s.integers(int(low), 1)[0])
        return self._s.integers(int(low), int(np.prod(size))).reshape(size)

    def choice(self, *a, **k):  # pragma: no cover
        raise NotImplementedError("not scripted")


class NumpyStream:
    """Same interface as ChainStream but backed by numpy's native PCG64 Generator, i.e.
    exactly what the reference uses (utils.py:993-999).  Used by bench.py's cpu_baseline /
    --impl reference legs so that the timed CPU path has the reference's RNG cost, not
    the cost of emulating Philox in numpy."""

    def __init__(self, seed, chain):
        self.g = np.random.Generator(np.random.PCG64([int(seed), int(chain)]))
        self.tick = 0

    def uniform(self):
        self.tick += 1
        return self.g.random()

    def uniforms(self, m):
        if m == 0:
            return np.empty(0)
        self.tick += 1
        return self.g.random(m)

    def normals(self, m):
        if m == 0:
            return np.empty(0)
        self.tick += 1
        return self.g.standard_normal(m)

    def integers(self, n, m):
        self.tick += 1
        return self.g.integers(n, size=m)

    def permutation(self, m):
        if m <= 1:
            return np.arange(m)
        self.tick += 1
        return self.g.permutation(m)
