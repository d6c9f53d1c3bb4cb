// b2n_rwalk.cu -- batched random-walk proposal chains: a warp-per-chain kernel (this comment), and two
// lock-step FP64-tensor-core kernels further down (rwalk_mma_kernel: the default for 16 <= n <= 64,
// rwalk_mmas_kernel: n > 64).
//
// Replaces RWalkSampler.sample -> generic_random_walk -> propose_ball_point
// (reference internal_samplers.py:505-561, 866-986, 989-1035) for a whole queue
// of chains in ONE launch.  Per chain and per step:
//   1. fresh U(0,1) on the non-clustered dims (:1011-1013)       [vector uniform event]
//   2. dr = randsphere(ncdim) (bounding.py:1288-1297)            [normal event + uniform event]
//   3. u' = u + scale * axes @ dr on the clustered dims (:1020-1021)
//   4. periodic wrap / reflect (:1024-1029), unitcheck (:1032); an out-of-cube
//      proposal counts as a call and a reject WITHOUT a likelihood call (:951-954)
//   5. v = prior_transform(u'), logl = loglikelihood(v), accept iff logl > loglstar
// exactly `walks` steps; with zero accepts v/logl are recomputed at the start (:970-975).
//
// Mapping of rwalk_kernel (ncdim < n, n < 16, or B2N_RWALK_IMPL=warp).  The grid is persistent-sized: ~one CTA per SM, each CTA owns an equal share of
// the queue (<= 16 chains in flight, one warp per chain) and only chains of ONE ellipsoid,
// whose axes^T and the precision matrix of a GAUSS_PREC model are staged ONCE into shared
// memory (column stride padded to 128 B so every column read is bank-conflict free) and
// then streamed `walks` times by every warp: HBM sees each matrix once per CTA.  Lane i
// owns rows i, i+32 of each mat-vec, the proposal vector is a warp-private shared vector
// read as 16-byte broadcasts; all shared-memory traffic is explicit b2n_sm[] indexing (no
// generic-pointer fix-ups), the wrap/reflect/cube test and the prior transform are fused
// into the mat-vec epilogue, and the two draw events of a step share one Philox + one log.
// ncu (profiles/) shows this kernel is shared-memory-pipe bound (84 % of peak wavefronts),
// not HBM bound: DRAM traffic is ~0.9 MB per launch against 5.8 GB of algorithmic bytes.
#include "b2n_chain.cuh"
#include <algorithm>

struct RwalkParams {
    B2nModel m;
    int n, nc, walks;
    int ldA, ldP;          // leading dims of axes^T / precision as seen by the kernel
    const double* u0;
    const int* start;      // optional: chain q starts from row start[q] of u0 (b2n_set_start_rows); NULL: row q
    const int* order;      // chains grouped by ellipsoid
    const int3* cta;       // (first, count, ell) per CTA
    const double* axesT;   // K x nc x nc, transposed (column-major axes)
    const uint32_t* dimflags;
    double loglstar, scale;
    uint64_t seed, chain0;
    double *u, *v, *logl;
    int *nacc, *nrej, *ncall;
    PeerSet peer;          // fused multi-GPU gather of the outputs (b2n_peer.cu); world == 0: off
    const B2nDyn* dyn;     // device-paced launch (b2n_ns.cu): threshold / scale / chain ids / CTA count in HBM
};

// Per-launch scalars: kernel arguments, or -- device-paced -- the B2nDyn the previous kernel on
// the stream wrote (a skipped round or a CTA beyond the round's worklist returns at once).
#define B2N_DYN_PROLOGUE(p)                                                                  \
    double loglstar_ = (p).loglstar, scale_ = (p).scale;                                     \
    unsigned long long chain0_ = (p).chain0;                                                 \
    if ((p).dyn) {                                                                           \
        if ((p).dyn->skip || (int)blockIdx.x >= (p).dyn->ncta) return;                       \
        loglstar_ = (p).dyn->loglstar; scale_ = (p).dyn->scale; chain0_ = (p).dyn->chain0;   \
    }

template <int LIKE, bool AX_SMEM, bool PREC_SMEM>
__global__ void __launch_bounds__(512, 1) rwalk_kernel(const RwalkParams p) {
    const int n = p.n, nc = p.nc;
    const int npad = (n + 1) & ~1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    B2N_DYN_PROLOGUE(p)
    const int3 cd = p.cta[blockIdx.x];
    // ---- shared-memory plan (all offsets in doubles, all even)
    int off = 0;
    const double* Ag = p.axesT + (size_t)cd.z * nc * nc;
    int offA = 0, ldA = nc;
    if (AX_SMEM) {
        offA = off; ldA = p.ldA;
        for (int e = threadIdx.x; e < nc * nc; e += blockDim.x) {
            const int j = e / nc, i = e - j * nc;
            b2n_sm[offA + j * ldA + i] = Ag[e];
        }
        off += nc * ldA;
    }
    const double* Pg = p.m.lmat;
    int offP = 0, ldP = n;
    if (LIKE == B2N_LIKE_GAUSS_PREC && PREC_SMEM) {
        offP = off; ldP = p.ldP;
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const int j = e / n, i = e - j * n;
            b2n_sm[offP + j * ldP + i] = Pg[e];
        }
        off += n * ldP;
    }
    const ModelSm ms = stage_model(p.m, off, n, npad);     // prior p0/p1, likelihood vec0/vec1
    const int op0 = ms.op0, op1 = ms.op1, omu = ms.olv0;
    off += 4 * npad;
    uint32_t* fl = reinterpret_cast<uint32_t*>(&b2n_sm[off]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) fl[i] = p.dimflags ? p.dimflags[i] : 0u;
    off += ((n + 3) >> 2) << 1;
    __syncthreads();

    int oucur = off + warp * 6 * npad;
    int ouprop = oucur + npad;
    int ovcur = ouprop + npad;
    int ovprop = ovcur + npad;
    const int ox = ovprop + npad;       // direction vector
    const int od = ox + npad;           // v - mean (GAUSS_PREC) / likelihood scratch
    const double inv_nc = 1.0 / (double)nc;
    const int pk = p.m.prior_kind;

    for (int c = warp; c < cd.y; c += nwarps) {
        const int q = p.order[cd.x + c];
        ChainRng g;
        g.init(p.seed, chain0_ + (uint64_t)q);
        for (int i = lane; i < n; i += 32) b2n_sm[oucur + i] = p.u0[(size_t)(p.start ? p.start[q] : q) * n + i];
        __syncwarp();
        int nacc = 0, nrej = 0;
        double lcur = 0.0;
        for (int step = 0; step < p.walks; step++) {
            // (the previous step's likelihood READ the delta vector that the loops below rewrite: order the two --
            //  warp shuffles converge the lanes but are not a memory barrier; compute-sanitizer racecheck, round 2)
            __syncwarp();
            // (1) non-clustered dims: one vector uniform event (only if there are any)
            if (n > nc) {
                for (int e = lane; e < n - nc; e += 32) {
                    const double t = rng_uniform_elem(g, e);
                    const double vi = prior_sm(pk, op0, op1, nc + e, t);
                    b2n_sm[ouprop + nc + e] = t;
                    b2n_sm[ovprop + nc + e] = vi;
                    b2n_sm[od + nc + e] = vi - b2n_sm[omu + nc + e];
                }
                g.tick++;
            }
            // (2) uniform point in the unit nc-ball
            const double fac = scale_ * ball_direction(g, ox, nc, lane, inv_nc);
            __syncwarp();
            // (3) u' = u + fac * axes @ z on the clustered dims, (4) wrap / reflect / cube test,
            //     and (speculatively) the prior transform of the rows this lane owns
            bool ok = true;
            for (int base = 0; base < nc; base += 64) {
                double y0, y1;
                matvec2o<AX_SMEM>(Ag, offA, ldA, nc, ox, base + lane, nc, y0, y1);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = base + lane + 32 * h;
                    if (i < nc) {
                        double t = fma(fac, h ? y1 : y0, b2n_sm[oucur + i]);
                        const uint32_t f = fl[i];
                        if (f & B2N_DIM_PERIODIC) t = mod1(t);
                        if (f & B2N_DIM_REFLECTIVE) t = reflect1(t);
                        ok = ok && in_cube(t, f);
                        const double vi = prior_sm(pk, op0, op1, i, t);
                        b2n_sm[ouprop + i] = t;
                        b2n_sm[ovprop + i] = vi;
                        b2n_sm[od + i] = vi - b2n_sm[omu + i];
                    }
                }
            }
            ok = __all_sync(B2N_FULL, ok);       // also orders the shared-memory writes above
            if (!ok) { nrej++; continue; }
            // (5) likelihood
            double l;
            if (LIKE == B2N_LIKE_GAUSS_PREC) {
                l = fma(-0.5, quadform_full<PREC_SMEM>(Pg, offP, ldP, n, od, lane), p.m.s0);
            } else {
                l = loglike_sm<LIKE, PREC_SMEM>(p.m, ms, Pg, offP, ldP, n, ovprop, od, lane);
            }
            if (l > loglstar_) {
                int t = oucur; oucur = ouprop; ouprop = t;
                t = ovcur; ovcur = ovprop; ovprop = t;
                lcur = l;
                nacc++;
            } else {
                nrej++;
            }
        }
        if (nacc == 0) {       // recompute (v, logl) of the start point (:970-975)
            for (int i = lane; i < n; i += 32) {
                const double vi = prior_sm(pk, op0, op1, i, b2n_sm[oucur + i]);
                b2n_sm[ovcur + i] = vi;
                b2n_sm[od + i] = vi - b2n_sm[omu + i];
            }
            __syncwarp();
            if (LIKE == B2N_LIKE_GAUSS_PREC) {
                lcur = fma(-0.5, quadform_full<PREC_SMEM>(Pg, offP, ldP, n, od, lane), p.m.s0);
            } else {
                lcur = loglike_sm<LIKE, PREC_SMEM>(p.m, ms, Pg, offP, ldP, n, ovcur, od, lane);
            }
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            peer_put(p.peer, &p.u[(size_t)q * n + i], b2n_sm[oucur + i]);
            peer_put(p.peer, &p.v[(size_t)q * n + i], b2n_sm[ovcur + i]);
        }
        if (lane == 0) {
            peer_put(p.peer, &p.logl[q], lcur);
            peer_put(p.peer, &p.nacc[q], nacc);
            peer_put(p.peer, &p.nrej[q], nrej);
            peer_put(p.peer, &p.ncall[q], (int)p.walks);
        }
        __syncwarp();
    }
    peer_finish(p.peer);
}

// =====================================================================================
// rwalk_mma_kernel -- the same chains, 8 at a time per CTA in LOCK-STEP, with both mat-vecs
// done as FP64 tensor-core MMAs (mma.sync m8n8k4 f64 = DMMA) whose A operands -- 8-row slabs
// of axes and of the precision matrix -- live in REGISTERS for the whole kernel.
//
// Why: the warp-per-chain kernel above streams 40 KB of matrix per proposal out of shared
// memory and is bound by the shared-memory pipe (84 % of peak, profiles/r1c).  All chains of
// a CTA use the same matrices and take the same number of steps, so the per-step work of a
// CTA is Y[n x 8] = A[n x n] X[n x 8]: a small dense contraction.  Distributing A over the
// lanes as DMMA fragments (1 double per lane per 8x4 tile; 13 k-tiles for n = 50 -> 26
// registers per matrix per warp) removes the matrix traffic from shared memory entirely and
// replaces ~500 LDS/DFMA per proposal by ~4 DMMA.  Only the 8 direction vectors go through
// shared memory.  Warp w is (i) the owner of chain w (RNG, wrap/reflect/cube test, prior,
// accept/reject: phases 1,3,5) and (ii) the owner of work item (slab w % S, chain-tile w / S)
// of the two contractions (phases 2,4); the directions (phase 1) come from a ring generated 8 steps ahead
// by all warps (see DEPTH below); 3 CTA barriers per step.
// Used when ncdim == ndim, 16 <= n <= 64 (fragments fit in registers); otherwise the
// warp-per-chain kernel runs.  Results agree with it to round-off (different summation order).
// =====================================================================================
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// KT = k-tiles of 4 columns (n <= 4*KT); CH = chains in lock-step per CTA (8 -> 256 threads, two
// CTAs per SM overlap each other's barriers: measured 1.3x faster than one 16-chain CTA per SM).
// DEPTH = direction ring: the draws of a step do not depend on the chain state (the Philox counter
// is a function of (chain, tick = 2 * step) only), so the directions of the next DEPTH steps of all
// live chains of the CTA are generated up front and dealt over ALL warps of the CTA.  A full CTA
// (8 chains) gains a barrier per step; a CTA with few chains -- the small rounds of b2n_ns_run run
// one chain per CTA -- generates its directions on 8 warps in parallel instead of serially on one,
// which is the longest dependency chain of a step (Philox -> log -> sqrt -> sincospi).
// FAST = the draws use the branch-free math of b2n_fastmath.cuh, two ring items at a time per warp
// (the default; B2N_RWALK_DRAWS=libm selects libdevice math, results differ by a few ulp in the directions).
template <int LIKE, int KT, int CH, int DEPTH, bool FAST>
__global__ void __launch_bounds__(CH * 32, CH == 8 ? 2 : 1) rwalk_mma_kernel(const RwalkParams p) {
    constexpr int B2N_MMA_CH = CH;
    constexpr int RS = 8 * ((4 * KT + 7) / 8);                    // rows padded to whole 8-row slabs
    constexpr int XS = RS + ((RS % 16 == 4) ? 0 : ((20 - RS % 16) % 16));   // chain stride == 4 (mod 16):
                                                                  // the B-fragment loads are conflict free
    constexpr int YS = RS + 2;
    const int n = p.n;
    const int npad = (n + 1) & ~1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    B2N_DYN_PROLOGUE(p)
    const int3 cd = p.cta[blockIdx.x];
    const int S = (n + 7) >> 3;                     // 8-row slabs
    // ---- shared-memory plan
    int off = 0;
    const ModelSm ms = stage_model(p.m, off, n, npad);
    const int op0 = ms.op0, op1 = ms.op1, omu = ms.olv0;
    off += 4 * npad;
    uint32_t* fl = reinterpret_cast<uint32_t*>(&b2n_sm[off]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) fl[i] = p.dimflags ? p.dimflags[i] : 0u;
    off += ((n + 3) >> 2) << 1;
    constexpr int XB = B2N_MMA_CH * XS;             // one direction buffer (all chains of the CTA)
    const int oX = off;  off += DEPTH * XB;         // ring: direction z of step s, later delta = v - mean
    const int oY = off;  off += B2N_MMA_CH * YS;    // axes @ z (chain-major)
    const int oQ = off;  off += 8 * B2N_MMA_CH;     // per-slab partial quadratic forms
    const int oF = off;  off += DEPTH * B2N_MMA_CH; // step factors scale * U^(1/n) / |z|
    const int ost = off;                            // per-chain state: ucur, uprop, vcur, vprop
    for (int e = threadIdx.x; e < DEPTH * XB; e += blockDim.x) b2n_sm[oX + e] = 0.0;
    // ---- matrix fragments -> registers.  item (s, t): slab s of rows, chain tile t
    const int s_it = warp % S, t_it = warp / S;
    const bool has_item = warp < (CH / 8) * S && t_it < CH / 8;
    double fragA[KT], fragP[KT];
    {
        const double* Ag = p.axesT + (size_t)cd.z * n * n;      // axesT[j*n + i] = axes[i][j]
        const double* Pg = p.m.lmat;
        const int row = 8 * s_it + (lane >> 2);
#pragma unroll
        for (int kt = 0; kt < KT; kt++) {
            const int col = 4 * kt + (lane & 3);
            const bool in = has_item && row < n && col < n;
            fragA[kt] = in ? Ag[(size_t)col * n + row] : 0.0;
            fragP[kt] = (LIKE == B2N_LIKE_GAUSS_PREC && in) ? Pg[(size_t)row * n + col] : 0.0;
        }
    }
    __syncthreads();
    const double inv_n = 1.0 / (double)n;
    const int pk = p.m.prior_kind;

    for (int g0 = 0; g0 < cd.y; g0 += B2N_MMA_CH) {             // groups of CH chains
        const int c = warp;                                     // chain slot owned by this warp
        const int nlc = (cd.y - g0) < B2N_MMA_CH ? (cd.y - g0) : B2N_MMA_CH;   // live chains of the group
        const bool live = c < nlc;
        const int q = live ? p.order[cd.x + g0 + c] : 0;
        int oucur = ost + c * 4 * npad, ouprop = oucur + npad, ovcur = ouprop + npad, ovprop = ovcur + npad;
        const int oy = oY + c * YS;
        if (live)
            for (int i = lane; i < n; i += 32) b2n_sm[oucur + i] = p.u0[(size_t)(p.start ? p.start[q] : q) * n + i];
        int nacc = 0, nrej = 0;
        double lcur = 0.0;
        for (int step0 = 0; step0 < p.walks; step0 += DEPTH) {
            // ---- phase 1 (all warps): directions in the unit ball of the next DEPTH steps of every
            //      live chain -> X[s][chain], factors -> F[s][chain].  Item w = (step s, chain c2).
            const int nd = (p.walks - step0) < DEPTH ? (p.walks - step0) : DEPTH;
            if (FAST && n <= 62) {
                for (int w = warp; w < nd * nlc; w += 2 * B2N_MMA_CH) {       // items w and w + CH together
                    const int w2 = w + B2N_MMA_CH;
                    const bool two = w2 < nd * nlc;
                    const int sa = w / nlc, ca = w - sa * nlc;
                    const int sb = two ? w2 / nlc : sa, cb = two ? w2 - sb * nlc : ca;
                    ChainRng ga, gb;
                    ga.init(p.seed, chain0_ + (uint64_t)p.order[cd.x + g0 + ca]);
                    gb.init(p.seed, chain0_ + (uint64_t)p.order[cd.x + g0 + cb]);
                    ga.tick = 2u * (uint32_t)(step0 + sa);
                    gb.tick = 2u * (uint32_t)(step0 + sb);
                    double fa, fb;
                    ball_direction_pair_fast(ga, gb, oX + sa * XB + ca * XS, oX + sb * XB + cb * XS, two, n, lane, inv_n,
                                             fa, fb);
                    if (lane == 0) {
                        b2n_sm[oF + sa * B2N_MMA_CH + ca] = scale_ * fa;
                        if (two) b2n_sm[oF + sb * B2N_MMA_CH + cb] = scale_ * fb;
                    }
                }
            } else
            for (int w = warp; w < nd * nlc; w += B2N_MMA_CH) {
                const int s = w / nlc, c2 = w - s * nlc;
                ChainRng g;
                g.init(p.seed, chain0_ + (uint64_t)p.order[cd.x + g0 + c2]);
                g.tick = 2u * (uint32_t)(step0 + s);             // two draw events per step (:1011-1016)
                const double f = scale_ * ball_direction(g, oX + s * XB + c2 * XS, n, lane, inv_n);
                if (lane == 0) b2n_sm[oF + s * B2N_MMA_CH + c2] = f;
            }
            __syncthreads();
            for (int s = 0; s < nd; s++) {
            const int oXs = oX + s * XB, ox = oXs + c * XS;
            const double fac = live ? b2n_sm[oF + s * B2N_MMA_CH + c] : 0.0;
            // ---- phase 2 (item warp): Y[rows of slab][chains of tile] = A_slab @ X
            if (has_item) {
                // NACC independent accumulator pairs: the k-tiles form NACC short DMMA dependency chains instead
                // of one long one (a CTA with one chain -- b2n_ns_run's small rounds -- is bound by that latency)
                double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
                const int xb = oXs + (8 * t_it + (lane >> 2)) * XS + (lane & 3);
#pragma unroll
                for (int kt = 0; kt + 1 < KT; kt += 2) {
                    dmma884(d0, d1, fragA[kt], b2n_sm[xb + 4 * kt]);
                    dmma884(e0, e1, fragA[kt + 1], b2n_sm[xb + 4 * kt + 4]);
                }
                if (KT & 1) dmma884(d0, d1, fragA[KT - 1], b2n_sm[xb + 4 * (KT - 1)]);
                d0 += e0;
                d1 += e1;
                const int row = 8 * s_it + (lane >> 2), c0 = 8 * t_it + 2 * (lane & 3);
                b2n_sm[oY + c0 * YS + row] = d0;
                b2n_sm[oY + (c0 + 1) * YS + row] = d1;
            }
            __syncthreads();
            // ---- phase 3 (chain warp): u' = u + fac*y, wrap / reflect / cube test, prior, delta -> X[s][c]
            bool ok = true;
            if (live) {
                for (int i = lane; i < n; i += 32) {
                    double t = fma(fac, b2n_sm[oy + i], b2n_sm[oucur + i]);
                    const uint32_t f = fl[i];
                    if (f & B2N_DIM_PERIODIC) t = mod1(t);
                    if (f & B2N_DIM_REFLECTIVE) t = reflect1(t);
                    ok = ok && in_cube(t, f);
                    const double vi = prior_sm(pk, op0, op1, i, t);
                    b2n_sm[ouprop + i] = t;
                    b2n_sm[ovprop + i] = vi;
                    b2n_sm[ox + i] = vi - b2n_sm[omu + i];
                }
                ok = __all_sync(B2N_FULL, ok);
            }
            double l = 0.0;
            if (LIKE == B2N_LIKE_GAUSS_PREC) {
                __syncthreads();
                // ---- phase 4 (item warp): partial delta^T P delta over the rows of the slab
                if (has_item) {
                    double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
                    const int xb = oXs + (8 * t_it + (lane >> 2)) * XS + (lane & 3);
#pragma unroll
                    for (int kt = 0; kt + 1 < KT; kt += 2) {
                        dmma884(d0, d1, fragP[kt], b2n_sm[xb + 4 * kt]);
                        dmma884(e0, e1, fragP[kt + 1], b2n_sm[xb + 4 * kt + 4]);
                    }
                    if (KT & 1) dmma884(d0, d1, fragP[KT - 1], b2n_sm[xb + 4 * (KT - 1)]);
                    d0 += e0;
                    d1 += e1;
                    const int row = 8 * s_it + (lane >> 2), c0 = 8 * t_it + 2 * (lane & 3);
                    double q0 = d0 * b2n_sm[oXs + c0 * XS + row], q1 = d1 * b2n_sm[oXs + (c0 + 1) * XS + row];
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        q0 += __shfl_xor_sync(B2N_FULL, q0, o);
                        q1 += __shfl_xor_sync(B2N_FULL, q1, o);
                    }
                    if (lane < 4) {
                        b2n_sm[oQ + s_it * B2N_MMA_CH + c0] = q0;
                        b2n_sm[oQ + s_it * B2N_MMA_CH + c0 + 1] = q1;
                    }
                }
                __syncthreads();
                // ---- phase 5 (chain warp): logl
                double qf = 0.0;
                for (int s2 = 0; s2 < S; s2++) qf += b2n_sm[oQ + s2 * B2N_MMA_CH + c];
                l = fma(-0.5, qf, p.m.s0);
            } else {
                if (live && ok) {
                    __syncwarp();
                    l = loglike_sm<LIKE, false>(p.m, ms, nullptr, 0, n, n, ovprop, oy, lane);
                }
                __syncthreads();      // Y (scratch of the likelihood) is rewritten by the next step's phase 2
            }
            if (live) {
                if (!ok) {
                    nrej++;
                } else if (l > loglstar_) {
                    int t = oucur; oucur = ouprop; ouprop = t;
                    t = ovcur; ovcur = ovprop; ovprop = t;
                    lcur = l;
                    nacc++;
                } else {
                    nrej++;
                }
            }
            }   // s
            // the ring is regenerated only after every warp has finished reading it (phase 4 of the
            // last step sits before a barrier; phase 3 of non-GAUSS_PREC likelihoods does too)
        }
        if (live) {
            if (nacc == 0) {   // recompute (v, logl) of the start point (:970-975), warp-local
                for (int i = lane; i < n; i += 32) b2n_sm[ovcur + i] = prior_sm(pk, op0, op1, i, b2n_sm[oucur + i]);
                __syncwarp();
                lcur = loglike_sm<LIKE, false>(p.m, ms, p.m.lmat, 0, n, n, ovcur, oy, lane);
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                peer_put(p.peer, &p.u[(size_t)q * n + i], b2n_sm[oucur + i]);
                peer_put(p.peer, &p.v[(size_t)q * n + i], b2n_sm[ovcur + i]);
            }
            if (lane == 0) {
                peer_put(p.peer, &p.logl[q], lcur);
                peer_put(p.peer, &p.nacc[q], nacc);
                peer_put(p.peer, &p.nrej[q], nrej);
                peer_put(p.peer, &p.ncall[q], (int)p.walks);
            }
        }
        __syncthreads();
        // X rows of chains that are not live in the next group must read as zero
        for (int e = threadIdx.x; e < DEPTH * XB; e += blockDim.x) b2n_sm[oX + e] = 0.0;
        __syncthreads();
    }
    peer_finish(p.peer);
}

// =====================================================================================
// rwalk_mma16_kernel -- the lock-step kernel above re-cut for SIXTEEN warps per 8 chains (GAUSS_PREC models).
//
// Why (profiles/r1n, r2n): rwalk_mma_kernel is bound by dependent-instruction latency -- 2 CTAs x 8 warps per SM
// (128 registers) issue on 44 % of the cycles, each warp one instruction per ~8 cycles -- and the queue has only
// 13.5 chains per SM, so more CTAs cannot be made resident.  The same work is therefore dealt over twice the warps,
// and the instruction count per step is cut:
//   * contractions: item = (8-row slab, HALF of the k-tiles) -> 2 S <= 16 item warps with 7 instead of 13 dependent
//     DMMAs, and 7 + 7 instead of 13 + 13 fragment registers per thread (<= 64 registers: 2 CTAs x 512 threads);
//     the two partial sums meet in shared memory (phase 3 adds the halves of y, phase 5 the 2 S partial forms);
//   * chain phases: two warps per chain, lane = component, so a thread owns ONE component of its chain for the whole
//     kernel: the chain state (u, v, and the proposal) lives in registers, not in shared memory;
//   * every shared-memory offset is a compile-time constant (the layout depends on KT only, not on n);
//   * two CTA barriers per step instead of three: the item warps run phase 4 of step s and phase 2 of step s + 1
//     back to back, the chain warps phase 5 of step s and phase 3 of step s + 1 (phase 2 does not depend on the
//     accept / reject of the step before: the directions are in the ring);
//   * draws: the ring items are dealt over 16 warps, and the step factor scale * U^(1/n) / |z| (exp, divide, sqrt) is
//     no longer computed on all 32 lanes per item but for the whole ring at once with one LANE per item, by the two
//     warps that have no contraction item (the first two when S = 8) while the others run phase 2 of the first step.
// Draw events, ticks and the arithmetic of every draw are those of rwalk_mma_kernel; only the summation order of the
// two contractions differs (round-off).
// =====================================================================================
template <int KT>
struct Mma16Layout {
    static constexpr int CH = 8, NW = 16, DEPTH = 8;
    static constexpr int KH = (KT + 1) / 2;                              // k-tiles per half
    static constexpr int RS = 8 * ((4 * KT + 7) / 8);                    // rows padded to whole slabs (>= n)
    static constexpr int XS = RS + ((RS % 16 == 4) ? 0 : ((20 - RS % 16) % 16));
    static constexpr int YS = RS + 2;
    static constexpr int XB = CH * XS;
    static constexpr int O_P0 = 0, O_P1 = RS;                            // prior vectors (mean, flags: registers)
    static constexpr int O_X = 2 * RS;                                   // ring: direction z of step s, later delta = v - mean
    static constexpr int O_Y = O_X + DEPTH * XB;                         // the two k-halves of axes @ z (chain-major)
    static constexpr int O_Q = O_Y + 2 * CH * YS;                        // partial quadratic forms per (slab, half)
    static constexpr int O_F = O_Q + 16 * CH;                            // step factors per ring slot
    static constexpr int O_SS = O_F + DEPTH * CH;                        // |z|^2 per ring slot
    static constexpr int O_LG = O_SS + DEPTH * CH;                       // log U of the radius per ring slot
    static constexpr int O_OK = O_LG + DEPTH * CH;                       // in-cube flags [step parity][chain][half]
    static constexpr int TOTAL = O_OK + 16;                              // doubles
};

template <int KT>
__global__ void __launch_bounds__(512, 2) rwalk_mma16_kernel(const RwalkParams p) {
    using L = Mma16Layout<KT>;
    constexpr int CH = L::CH, NW = L::NW, DEPTH = L::DEPTH, KH = L::KH, XS = L::XS, YS = L::YS, XB = L::XB;
    const int n = p.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    B2N_DYN_PROLOGUE(p)
    const int3 cd = p.cta[blockIdx.x];
    const int S = (n + 7) >> 3;
    int* okf = reinterpret_cast<int*>(&b2n_sm[L::O_OK]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        b2n_sm[L::O_P0 + i] = p.m.pp0 ? p.m.pp0[i] : 0.0;
        b2n_sm[L::O_P1 + i] = p.m.pp1 ? p.m.pp1[i] : 1.0;
    }
    for (int e = threadIdx.x; e < DEPTH * XB; e += blockDim.x) b2n_sm[L::O_X + e] = 0.0;
    // ---- fragments: item (slab s_it, half h_it) of both matrices
    const int s_it = warp % S, h_it = warp / S;
    const bool has_item = warp < 2 * S;
    double fragA[KH], fragP[KH];
    {
        const double* Ag = p.axesT + (size_t)cd.z * n * n;
        const double* Pg = p.m.lmat;
        const int row = 8 * s_it + (lane >> 2);
#pragma unroll
        for (int j = 0; j < KH; j++) {
            const int col = 4 * (h_it * KH + j) + (lane & 3);
            const bool in = has_item && row < n && col < n;
            fragA[j] = in ? Ag[(size_t)col * n + row] : 0.0;
            fragP[j] = in ? Pg[(size_t)row * n + col] : 0.0;
        }
    }
    // item-phase addresses (relative to the ring slot of the step)
    const int xb_it = (lane >> 2) * XS + (lane & 3) + 4 * h_it * KH;           // B fragment of k-tile 0 of the half
    const int yst_it = L::O_Y + (h_it * CH + 2 * (lane & 3)) * YS + 8 * s_it + (lane >> 2);
    const int xr_it = 2 * (lane & 3) * XS + 8 * s_it + (lane >> 2);            // delta[row] of chain c0
    // the two warps that turn (|z|^2, log U) of the ring into step factors: those without an item, else the first two
    const int fw = (2 * S <= NW - 2) ? warp - (NW - 2) : warp;
    const double inv_n = 1.0 / (double)n;
    const int pk = p.m.prior_kind;
    const int c = warp >> 1, hh = warp & 1;                       // chain slot and component half of this warp
    const int ci = 32 * hh + lane;                                // the component this thread owns
    const bool cin = ci < n;
    const uint32_t myfl = cin ? (p.dimflags ? p.dimflags[ci] : 0u) : 0u;
    const double mymu = cin && p.m.lv0 ? p.m.lv0[ci] : 0.0;
    __syncthreads();

    for (int g0 = 0; g0 < cd.y; g0 += CH) {
        const int nlc = (cd.y - g0) < CH ? (cd.y - g0) : CH;
        const bool live = c < nlc;
        const int q = live ? p.order[cd.x + g0 + c] : 0;
        double ucur = (live && cin) ? p.u0[(size_t)(p.start ? p.start[q] : q) * n + ci] : 0.0, vcur = 0.0, uprop = 0.0, vprop = 0.0;
        int nacc = 0, nrej = 0;
        double lcur = 0.0;

        // phase 2 of ring slot s: Y_h[rows of slab][chains] = A_slab[:, half h] @ X[half h]
        auto phase2 = [&](int oXs) {
            double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
            const int xb = oXs + xb_it;
#pragma unroll
            for (int j = 0; j + 1 < KH; j += 2) {
                dmma884(d0, d1, fragA[j], b2n_sm[xb + 4 * j]);
                dmma884(e0, e1, fragA[j + 1], b2n_sm[xb + 4 * j + 4]);
            }
            if (KH & 1) dmma884(d0, d1, fragA[KH - 1], b2n_sm[xb + 4 * (KH - 1)]);
            b2n_sm[yst_it] = d0 + e0;
            b2n_sm[yst_it + YS] = d1 + e1;
        };
        // phase 3 of ring slot s: u' = u + fac*y, wrap / reflect / cube test, prior, delta -> X[s][c]
        auto phase3 = [&](int s) {
            bool ok = true;
            if (cin) {
                const double fac = b2n_sm[L::O_F + s * CH + c];
                const double y = b2n_sm[L::O_Y + c * YS + ci] + b2n_sm[L::O_Y + (CH + c) * YS + ci];
                double t = fma(fac, y, ucur);
                if (myfl & B2N_DIM_PERIODIC) t = mod1(t);
                if (myfl & B2N_DIM_REFLECTIVE) t = reflect1(t);
                ok = in_cube(t, myfl);
                uprop = t;
                vprop = prior_sm(pk, L::O_P0, L::O_P1, ci, t);
                b2n_sm[L::O_X + s * XB + c * XS + ci] = vprop - mymu;
            }
            ok = __all_sync(B2N_FULL, ok);
            if (lane == 0) okf[(s & 1) * 16 + warp] = ok ? 1 : 0;
        };
        // phase 4 of ring slot s: partial delta^T P delta over (rows of the slab) x (columns of the half)
        auto phase4 = [&](int oXs) {
            double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
            const int xb = oXs + xb_it;
#pragma unroll
            for (int j = 0; j + 1 < KH; j += 2) {
                dmma884(d0, d1, fragP[j], b2n_sm[xb + 4 * j]);
                dmma884(e0, e1, fragP[j + 1], b2n_sm[xb + 4 * j + 4]);
            }
            if (KH & 1) dmma884(d0, d1, fragP[KH - 1], b2n_sm[xb + 4 * (KH - 1)]);
            double q0 = (d0 + e0) * b2n_sm[oXs + xr_it], q1 = (d1 + e1) * b2n_sm[oXs + xr_it + XS];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                q0 += __shfl_xor_sync(B2N_FULL, q0, o);
                q1 += __shfl_xor_sync(B2N_FULL, q1, o);
            }
            if (lane < 4) {
                b2n_sm[L::O_Q + warp * CH + 2 * lane] = q0;
                b2n_sm[L::O_Q + warp * CH + 2 * lane + 1] = q1;
            }
        };
        // phase 5 of ring slot s (both warps of the chain, identically): logl, accept / reject
        auto phase5 = [&](int s) {
            double qf = (lane < 2 * S) ? b2n_sm[L::O_Q + lane * CH + c] : 0.0;     // 2 S <= 16 partials
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) qf += __shfl_xor_sync(B2N_FULL, qf, o);
            qf = __shfl_sync(B2N_FULL, qf, 0);
            const double l = fma(-0.5, qf, p.m.s0);
            const int* ok2 = &okf[(s & 1) * 16 + 2 * c];
            const bool ok = (ok2[0] & ok2[1]) != 0;
            if (ok && l > loglstar_) {
                ucur = uprop;
                vcur = vprop;
                lcur = l;
                nacc++;
            } else {
                nrej++;
            }
        };

        for (int step0 = 0; step0 < p.walks; step0 += DEPTH) {
            const int nd = (p.walks - step0) < DEPTH ? (p.walks - step0) : DEPTH;
            const int nit = nd * nlc;
            // ---- phase 1 (all warps): the directions of the next nd steps of every live chain
            if (nit > NW) {
                for (int w = warp; w < nit; w += 2 * NW) {        // two items side by side (their chains interleave)
                    const int w2 = w + NW;
                    const bool two = w2 < nit;
                    int sa, ca, sb, cb;
                    if (nlc == CH) { sa = w >> 3; ca = w & 7; sb = w2 >> 3; cb = w2 & 7; }
                    else { sa = w / nlc; ca = w - sa * nlc; sb = w2 / nlc; cb = w2 - sb * nlc; }
                    if (!two) { sb = sa; cb = ca; }
                    ChainRng ga, gb;
                    ga.init(p.seed, chain0_ + (uint64_t)p.order[cd.x + g0 + ca]);
                    gb.init(p.seed, chain0_ + (uint64_t)p.order[cd.x + g0 + cb]);
                    ga.tick = 2u * (uint32_t)(step0 + sa);
                    gb.tick = 2u * (uint32_t)(step0 + sb);
                    double ssa, lga, ssb, lgb;
                    ball_draw_fast(ga, L::O_X + sa * XB + ca * XS, true, n, lane, ssa, lga);
                    ball_draw_fast(gb, L::O_X + sb * XB + cb * XS, two, n, lane, ssb, lgb);
                    if (lane == 0) {
                        b2n_sm[L::O_SS + sa * CH + ca] = ssa;
                        b2n_sm[L::O_LG + sa * CH + ca] = lga;
                        if (two) {
                            b2n_sm[L::O_SS + sb * CH + cb] = ssb;
                            b2n_sm[L::O_LG + sb * CH + cb] = lgb;
                        }
                    }
                }
            } else if (warp < nit) {                              // few chains (b2n_ns_run's rounds): one item per warp
                const int sa = warp / nlc, ca = warp - sa * nlc;
                ChainRng ga;
                ga.init(p.seed, chain0_ + (uint64_t)p.order[cd.x + g0 + ca]);
                ga.tick = 2u * (uint32_t)(step0 + sa);
                double ssa, lga;
                ball_draw_fast(ga, L::O_X + sa * XB + ca * XS, true, n, lane, ssa, lga);
                if (lane == 0) {
                    b2n_sm[L::O_SS + sa * CH + ca] = ssa;
                    b2n_sm[L::O_LG + sa * CH + ca] = lga;
                }
            }
            __syncthreads();
            // ---- step factors of the whole ring, one lane per ring slot (step e / CH, chain e % CH)
            if (fw == 0 || fw == 1) {
                const int e = 32 * fw + lane;
                if (e < nd * CH && (e & (CH - 1)) < nlc)
                    b2n_sm[L::O_F + e] = scale_ * b2n_div(exp(b2n_sm[L::O_LG + e] * inv_n), b2n_sqrt(b2n_sm[L::O_SS + e]));
            }
            if (has_item) phase2(L::O_X);
            __syncthreads();
            if (live) phase3(0);
            __syncthreads();
            for (int s = 0; s < nd; s++) {
                if (has_item) {
                    phase4(L::O_X + s * XB);
                    if (s + 1 < nd) phase2(L::O_X + (s + 1) * XB);
                }
                __syncthreads();
                if (live) phase5(s);
                if (s + 1 < nd) {
                    if (live) phase3(s + 1);
                    __syncthreads();
                }
            }
            // (the next ring is drawn in the same barrier interval as phase 5 of the last step: every read of the ring
            //  -- phase 4 of that step -- sits before the barrier above)
        }
        // no accept: (v, logl) of the start point are recomputed (:970-975); every thread its own component of v
        if (live && nacc == 0 && cin) {
            vcur = prior_sm(pk, L::O_P0, L::O_P1, ci, ucur);
            b2n_sm[L::O_Y + c * YS + ci] = vcur - mymu;
        }
        __syncthreads();
        if (live) {
            if (nacc == 0 && hh == 0)
                lcur = fma(-0.5, quadform_full<false>(p.m.lmat, 0, n, n, L::O_Y + c * YS, lane), p.m.s0);
            if (cin) {
                peer_put(p.peer, &p.u[(size_t)q * n + ci], ucur);
                peer_put(p.peer, &p.v[(size_t)q * n + ci], vcur);
            }
            if (lane == 0 && hh == 0) {
                peer_put(p.peer, &p.logl[q], lcur);
                peer_put(p.peer, &p.nacc[q], nacc);
                peer_put(p.peer, &p.nrej[q], nrej);
                peer_put(p.peer, &p.ncall[q], (int)p.walks);
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < DEPTH * XB; e += blockDim.x) b2n_sm[L::O_X + e] = 0.0;
        __syncthreads();
    }
    peer_finish(p.peer);
}

// =====================================================================================
// rwalk_mmaws_kernel -- the lock-step kernel, WARP-SPECIALISED (GAUSS_PREC models): 8 step warps + 4 draw warps.
//
// Why (profiles/r2n): the draws are ~40 % of the instructions of rwalk_mma_kernel and the only part of it that is
// bound by instruction issue; the step phases are chains of dependent shared-memory loads, DMMAs, shuffles and
// barriers that leave the schedulers idle most of the time.  The draws of a step do not depend on the chain state, so
// they do not have to sit in the same instruction stream at all:
//   * warps 8..11 (one per scheduler) do nothing but draw: they fill a DOUBLE-BUFFERED ring of directions (8 steps x 8
//     chains per buffer) one buffer ahead of the step warps and issue into the cycles the step warps leave empty;
//     the step factors scale * U^(1/n) / |z| of a buffer are formed with one LANE per ring slot;
//   * warps 0..7 run the step phases of rwalk_mma_kernel on register-resident DMMA fragments, with every shared-memory
//     offset a compile-time constant and TWO barriers per step instead of three (phase 4 of step s and phase 2 of
//     step s + 1 run back to back, then phase 5 of step s and phase 3 of step s + 1);
//   * the two roles meet at named barriers only (FULL / EMPTY per ring buffer: bar.arrive on one side, bar.sync on
//     the other), the step warps synchronise among themselves on a 256-thread named barrier;
//   * setmaxnreg moves registers from the draw warps to the step warps (the fragments alone are 52 registers).
// Draw events, ticks and arithmetic are those of rwalk_mma_kernel (the step factor is the same expression); the
// contractions are summed in the same order: results are bit-identical to rwalk_mma_kernel with fast draws.
// =====================================================================================
template <int KT>
struct MmaWsLayout {
    static constexpr int CH = 8, DEPTH = 8, NSW = 8, NDW = 4;
    static constexpr int RS = 8 * ((4 * KT + 7) / 8);                    // rows padded to whole slabs (>= n)
    static constexpr int XS = RS + ((RS % 16 == 4) ? 0 : ((20 - RS % 16) % 16));
    static constexpr int YS = RS + 2;
    static constexpr int XB = CH * XS;
    static constexpr int RING = DEPTH * XB;
    static constexpr int SMAX = RS / 8;                                  // slabs at the largest n of this KT
    static constexpr int NTMAX = SMAX * KT - SMAX * (SMAX - 1);          // tiles on or above the diagonal: sum (KT - 2 s)
    static constexpr int TPW = (NTMAX + NSW - 1) / NSW;                  // of them per step warp
    static constexpr int O_P0 = 0, O_P1 = RS, O_MU = 2 * RS;            // prior vectors, likelihood mean
    static constexpr int O_FL = 3 * RS;                                  // dimension flags (RS uint32)
    static constexpr int O_X = O_FL + RS / 2;                            // two rings: direction z, later delta = v - mean
    static constexpr int O_Y = O_X + 2 * RING;                           // axes @ z (chain-major), two buffers (slot parity)
    static constexpr int O_Q = O_Y + 2 * CH * YS;                        // partial quadratic forms per step warp
    static constexpr int O_F = O_Q + 8 * CH;                             // step factors, two rings
    static constexpr int O_SS = O_F + 2 * DEPTH * CH;                    // |z|^2 per ring slot (draw warps' scratch)
    static constexpr int O_LG = O_SS + DEPTH * CH;                       // log U of the radius per ring slot
    static constexpr int O_ST = O_LG + DEPTH * CH;                       // chain state: ucur, uprop, vcur, vprop
    static constexpr int TOTAL = O_ST + CH * 4 * RS;                     // doubles
};
// Static schedule of the symmetric quadratic form (largest n of a KT: S = SMAX slabs): the tiles (slab s, k-tile k)
// with k >= 2 s, in (s, k) order, TPW consecutive ones per step warp.  Everything below is resolved at compile time
// (W = warp index through a switch), so a warp's phase 4 is straight-line code: its DMMAs with immediate offsets and
// one multiply by delta[rows of the slab] at the end of each run of tiles of one slab.
template <int KT>
struct SymSched {
    using L = MmaWsLayout<KT>;
    static constexpr int NT = L::NTMAX, TPW = L::TPW;
    static constexpr __host__ __device__ int slab(int t) { int s = 0; while (t >= KT - 2 * s) { t -= KT - 2 * s; s++; } return s; }
    static constexpr __host__ __device__ int ktile(int t) { int s = 0; while (t >= KT - 2 * s) { t -= KT - 2 * s; s++; } return 2 * s + t; }
    // position of tile t inside the run of tiles of its slab that belongs to warp t / TPW
    static constexpr __host__ __device__ int runpos(int t) {
        const int first = (t / TPW) * TPW;
        return slab(first) == slab(t) ? t - first : ktile(t) - 2 * slab(t);
    }
    static constexpr __host__ __device__ bool runend(int t) { return (t + 1) % TPW == 0 || t + 1 == NT || slab(t + 1) != slab(t); }
};
template <int KT, int W, int J>
__device__ __forceinline__ void sym_load(double (&fragU)[MmaWsLayout<KT>::TPW], const double* __restrict__ Pg, int n, int lane) {
    using Sch = SymSched<KT>;
    if constexpr (J < Sch::TPW) {
        constexpr int t = W * Sch::TPW + J;
        double u = 0.0;
        if constexpr (t < Sch::NT) {
            constexpr int ts = Sch::slab(t), tk = Sch::ktile(t);
            const int row = 8 * ts + (lane >> 2), col = 4 * tk + (lane & 3);
            if (row < n && col < n && col >= row) u = (col == row ? 0.5 : 1.0) * Pg[(size_t)row * n + col];
        }
        fragU[J] = u;
        sym_load<KT, W, J + 1>(fragU, Pg, n, lane);
    }
}
template <int KT, int W, int J>
__device__ __forceinline__ void sym_tiles(const double (&fragU)[MmaWsLayout<KT>::TPW], int xt, int xr, double& d0, double& d1,
                                          double& e0, double& e1, double& q0, double& q1) {
    using Sch = SymSched<KT>;
    constexpr int XS = MmaWsLayout<KT>::XS;
    constexpr int t = W * Sch::TPW + J;
    if constexpr (J < Sch::TPW && t < Sch::NT) {
        constexpr int s = Sch::slab(t), k = Sch::ktile(t), pos = Sch::runpos(t);
        if constexpr (pos & 1) dmma884(e0, e1, fragU[J], b2n_sm[xt + 4 * k]);
        else dmma884(d0, d1, fragU[J], b2n_sm[xt + 4 * k]);
        if constexpr (Sch::runend(t)) {       // multiply the slab's rows by delta[rows], start the next run from zero
            if constexpr (pos >= 1) {
                q0 = fma(d0 + e0, b2n_sm[xr + 8 * s], q0);
                q1 = fma(d1 + e1, b2n_sm[xr + XS + 8 * s], q1);
                e0 = 0.0; e1 = 0.0;
            } else {
                q0 = fma(d0, b2n_sm[xr + 8 * s], q0);
                q1 = fma(d1, b2n_sm[xr + XS + 8 * s], q1);
            }
            d0 = 0.0; d1 = 0.0;
        }
        sym_tiles<KT, W, J + 1>(fragU, xt, xr, d0, d1, e0, e1, q0, q1);
    }
}
#define B2N_WARP_SWITCH(W_, CALL)                                                                      \
    switch (W_) {                                                                                      \
        case 0: CALL(0); break; case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; \
        case 4: CALL(4); break; case 5: CALL(5); break; case 6: CALL(6); break; default: CALL(7); break; \
    }

__device__ __forceinline__ void nbar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void nbar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// SETREG: 0 = every warp keeps the 80 registers of the launch; 1 = 88 (step) / 64 (draw); 2 = 96 / 48
// (256 x step + 128 x draw must not exceed the 384 x 80 registers the CTA is launched with; measured at C2, profiles/r2o:
// 0.226 / 0.2045 / 0.2086 ms -- only 1 is instantiated).
// PLAIN: no periodic / reflective dimension and a prior that is affine per component (uniform, identity): phase 3 is
// then straight-line code for the (at most) two components of a lane.
// Ring: RB = 2 buffers of DB = 8 steps; the step warps start every buffer with a two-interval prologue.  Measured and not
// kept (profiles/r2r, r2t): ONE pipeline over all ring slots, the next buffer awaited where it is first touched (two steps
// early) -- on 2 x 8 slots 0.200 ms, on 4 x 4 slots 0.207 ms against 0.193 ms: the draw warps are ~80 % busy, and taking
// two steps of slack away from them makes both roles wait for each other.
template <int KT, int SETREG, bool PLAIN>
__global__ void __launch_bounds__(384, 2) rwalk_mmaws_kernel(const RwalkParams p) {
    using L = MmaWsLayout<KT>;
    constexpr int CH = L::CH, XS = L::XS, YS = L::YS, XB = L::XB, RS = L::RS;
    constexpr int DB = L::DEPTH, RB = 2;
    constexpr int BAR_STEP = 1, BAR_FULL = 2, BAR_EMPTY = 2 + RB;  // named barriers (0 = __syncthreads)
    const int n = p.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    B2N_DYN_PROLOGUE(p)
    const int3 cd = p.cta[blockIdx.x];
    uint32_t* fl = reinterpret_cast<uint32_t*>(&b2n_sm[L::O_FL]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        b2n_sm[L::O_P0 + i] = p.m.pp0 ? p.m.pp0[i] : 0.0;
        b2n_sm[L::O_P1 + i] = p.m.pp1 ? p.m.pp1[i] : 1.0;
        b2n_sm[L::O_MU + i] = p.m.lv0 ? p.m.lv0[i] : 0.0;
        fl[i] = p.dimflags ? p.dimflags[i] : 0u;
    }
    for (int e = threadIdx.x; e < 2 * L::RING; e += blockDim.x) b2n_sm[L::O_X + e] = 0.0;
    __syncthreads();
    const int NB = (p.walks + DB - 1) / DB;                       // ring buffers filled per group of chains

    if (warp >= L::NSW) {
        // =============================== draw warps ===============================
        if (SETREG == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (SETREG == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        const int dw = warp - L::NSW;
        const double inv_n = 1.0 / (double)n;
        for (int g0 = 0; g0 < cd.y; g0 += CH) {
            const int nlc = (cd.y - g0) < CH ? (cd.y - g0) : CH;
            for (int blk = 0; blk < NB; blk++) {
                const int b = blk & (RB - 1), step0 = blk * DB;
                const int nd = (p.walks - step0) < DB ? (p.walks - step0) : DB;
                const int nit = nd * nlc;
                const int oXb = L::O_X + b * DB * XB;
                if (blk >= RB) nbar_sync(BAR_EMPTY + b, 384);     // the step warps have finished with this buffer
                for (int w = dw; w < nit; w += 2 * L::NDW) {      // two items side by side (their chains interleave)
                    const int w2 = w + L::NDW;
                    const bool two = w2 < nit;
                    int sa, ca, sb, cb;
                    if (nlc == CH) { sa = w >> 3; ca = w & 7; sb = w2 >> 3; cb = w2 & 7; }
                    else { sa = w / nlc; ca = w - sa * nlc; sb = w2 / nlc; cb = w2 - sb * nlc; }
                    if (!two) { sb = sa; cb = ca; }
                    ChainRng ga, gb;
                    ga.init(p.seed, chain0_ + (uint64_t)p.order[cd.x + g0 + ca]);
                    gb.init(p.seed, chain0_ + (uint64_t)p.order[cd.x + g0 + cb]);
                    ga.tick = 2u * (uint32_t)(step0 + sa);
                    gb.tick = 2u * (uint32_t)(step0 + sb);
                    double ssa, lga, ssb, lgb;
                    ball_draw_fast(ga, oXb + sa * XB + ca * XS, true, n, lane, ssa, lga);
                    ball_draw_fast(gb, oXb + sb * XB + cb * XS, two, n, lane, ssb, lgb);
                    if (lane == 0) {
                        b2n_sm[L::O_SS + sa * CH + ca] = ssa;
                        b2n_sm[L::O_LG + sa * CH + ca] = lga;
                        if (two) {
                            b2n_sm[L::O_SS + sb * CH + cb] = ssb;
                            b2n_sm[L::O_LG + sb * CH + cb] = lgb;
                        }
                    }
                }
                __syncwarp();
                {   // step factors of this warp's items, one lane per item: item w = dw + NDW * lane
                    const int w = dw + L::NDW * lane;
                    if (lane < 2 * DB && w < nit) {
                        int sa, ca;
                        if (nlc == CH) { sa = w >> 3; ca = w & 7; }
                        else { sa = w / nlc; ca = w - sa * nlc; }
                        const int e = sa * CH + ca;
                        b2n_sm[L::O_F + b * DB * CH + e] =
                            scale_ * b2n_div(exp(b2n_sm[L::O_LG + e] * inv_n), b2n_sqrt(b2n_sm[L::O_SS + e]));
                    }
                }
                __syncwarp();             // lane 0 rewrites the scratch in the next buffer's first pass (racecheck, r2t)
                __threadfence_block();
                nbar_arrive(BAR_FULL + b, 384);
            }
            // group boundary: both roles meet, the rings are cleared (rows of chains that are not live in the next group)
            __syncthreads();
            for (int e = threadIdx.x; e < 2 * L::RING; e += blockDim.x) b2n_sm[L::O_X + e] = 0.0;
            __syncthreads();
        }
    } else {
        // =============================== step warps ===============================
        if (SETREG == 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
        if (SETREG == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        const int S = (n + 7) >> 3;
        const int s_it = warp;                                    // slab of this warp's item of axes @ z
        const bool has_item = warp < S;
        double fragA[KT];
        {
            const double* Ag = p.axesT + (size_t)cd.z * n * n;
            const int row = 8 * s_it + (lane >> 2);
#pragma unroll
            for (int kt = 0; kt < KT; kt++) {
                const int col = 4 * kt + (lane & 3);
                fragA[kt] = (has_item && row < n && col < n) ? Ag[(size_t)col * n + row] : 0.0;
            }
        }
        // The precision matrix is symmetric: delta^T P delta = 2 delta^T U delta with U = its upper triangle and HALF
        // its diagonal, so only the (slab, k-tile) tiles on or above the diagonal are contracted -- 49 of the 91 at
        // n = 50 -- dealt over all 8 warps by the static schedule SymSched (this kernel runs for S == SMAX only).
        double fragU[L::TPW];
#define B2N_SYM_LOAD(W_) sym_load<KT, W_, 0>(fragU, p.m.lmat, n, lane)
        B2N_WARP_SWITCH(warp, B2N_SYM_LOAD)
#undef B2N_SYM_LOAD
        const int xb_it = (lane >> 2) * XS + (lane & 3);                          // B fragment of k-tile 0
        const int yst_it = L::O_Y + 2 * (lane & 3) * YS + 8 * s_it + (lane >> 2);
        const int xr_it = 2 * (lane & 3) * XS + (lane >> 2);                      // delta[row 0 of a slab] of chain c0
        const int pk = p.m.prior_kind;
        const int c = warp;                                       // chain slot owned by this warp

        for (int g0 = 0; g0 < cd.y; g0 += CH) {
            const int nlc = (cd.y - g0) < CH ? (cd.y - g0) : CH;
            const bool live = c < nlc;
            const int q = live ? p.order[cd.x + g0 + c] : 0;
            // chain state in shared memory: [u_a | u_b | v_a | v_b]; `par` says which half holds the current point
            const int ost = L::O_ST + c * 4 * RS;
            int par = 0;
#define oucur (ost + par * RS)
#define ouprop (ost + (par ^ 1) * RS)
#define ovcur (ost + 2 * RS + par * RS)
#define ovprop (ost + 2 * RS + (par ^ 1) * RS)
            if (live)
                for (int i = lane; i < n; i += 32) b2n_sm[oucur + i] = p.u0[(size_t)(p.start ? p.start[q] : q) * n + i];
            int nacc = 0, nrej = 0;
            double lcur = 0.0;
            bool ok = true;

            // phase 2 of a ring slot: Y[rows of slab][chains] = A_slab @ X -- the DMMAs are issued here, their result
            // is stored by phase2_store AFTER the chain phases of the same barrier interval (the tensor pipe works on
            // the next-but-one step while the warp walks through the latency chains of phases 5 and 3)
            auto phase2_issue = [&](int oXs, double& d0, double& d1, double& e0, double& e1) {
                d0 = 0.0; d1 = 0.0; e0 = 0.0; e1 = 0.0;
                const int xb = oXs + xb_it;
#pragma unroll
                for (int kt = 0; kt + 1 < KT; kt += 2) {
                    dmma884(d0, d1, fragA[kt], b2n_sm[xb + 4 * kt]);
                    dmma884(e0, e1, fragA[kt + 1], b2n_sm[xb + 4 * kt + 4]);
                }
                if (KT & 1) dmma884(d0, d1, fragA[KT - 1], b2n_sm[xb + 4 * (KT - 1)]);
            };
            auto phase2_store = [&](int buf, double d0, double d1, double e0, double e1) {
                b2n_sm[yst_it + buf * CH * YS] = d0 + e0;
                b2n_sm[yst_it + buf * CH * YS + YS] = d1 + e1;
            };
            auto phase3 = [&](int oXs, double fac, int oy) {   // u' = u + fac*y, wrap / reflect / cube test, prior, delta -> X[s][c]
                if (PLAIN) {
                    const int i0 = lane, i1 = lane + 32;
                    const bool v0 = i0 < n, v1 = i1 < n;
                    double t0 = 0.5, t1 = 0.5;
                    if (v0) t0 = fma(fac, b2n_sm[oy + i0], b2n_sm[oucur + i0]);
                    if (v1) t1 = fma(fac, b2n_sm[oy + i1], b2n_sm[oucur + i1]);
                    const bool good = (t0 > 0.0 && t0 < 1.0) && (t1 > 0.0 && t1 < 1.0);
                    if (v0) {
                        const double vi = fma(b2n_sm[L::O_P1 + i0], t0, b2n_sm[L::O_P0 + i0]);
                        b2n_sm[ouprop + i0] = t0;
                        b2n_sm[ovprop + i0] = vi;
                        b2n_sm[oXs + c * XS + i0] = vi - b2n_sm[L::O_MU + i0];
                    }
                    if (v1) {
                        const double vi = fma(b2n_sm[L::O_P1 + i1], t1, b2n_sm[L::O_P0 + i1]);
                        b2n_sm[ouprop + i1] = t1;
                        b2n_sm[ovprop + i1] = vi;
                        b2n_sm[oXs + c * XS + i1] = vi - b2n_sm[L::O_MU + i1];
                    }
                    ok = __all_sync(B2N_FULL, good);
                    return;
                }
                bool good = true;
                for (int i = lane; i < n; i += 32) {
                    double t = fma(fac, b2n_sm[oy + i], b2n_sm[oucur + i]);
                    const uint32_t f = fl[i];
                    if (f & B2N_DIM_PERIODIC) t = mod1(t);
                    if (f & B2N_DIM_REFLECTIVE) t = reflect1(t);
                    good = good && in_cube(t, f);
                    const double vi = prior_sm(pk, L::O_P0, L::O_P1, i, t);
                    b2n_sm[ouprop + i] = t;
                    b2n_sm[ovprop + i] = vi;
                    b2n_sm[oXs + c * XS + i] = vi - b2n_sm[L::O_MU + i];
                }
                ok = __all_sync(B2N_FULL, good);
            };
            auto phase4 = [&](int oXs) {          // this warp's tiles of delta^T U delta
                double q0 = 0.0, q1 = 0.0, d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
                const int xt = oXs + xb_it, xr = oXs + xr_it;
#define B2N_SYM_TILES(W_) sym_tiles<KT, W_, 0>(fragU, xt, xr, d0, d1, e0, e1, q0, q1)
                B2N_WARP_SWITCH(warp, B2N_SYM_TILES)
#undef B2N_SYM_TILES
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    q0 += __shfl_xor_sync(B2N_FULL, q0, o);
                    q1 += __shfl_xor_sync(B2N_FULL, q1, o);
                }
                if (lane < 4) {
                    b2n_sm[L::O_Q + warp * CH + 2 * lane] = q0;
                    b2n_sm[L::O_Q + warp * CH + 2 * lane + 1] = q1;
                }
            };
            auto phase5 = [&]() {                 // logl = s0 - delta^T U delta, accept / reject
                double qf = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < L::NSW; w2++) qf += b2n_sm[L::O_Q + w2 * CH + c];
                const double l = p.m.s0 - qf;
                if (ok && l > loglstar_) {
                    par ^= 1;
                    lcur = l;
                    nacc++;
                } else {
                    nrej++;
                }
            };

            // Two barrier intervals per step,
            //   I(g) = phase 4 of slot g                          C(g) = [issue phase 2 of slot g + 2] phase 5 of slot g,
            //                                                            phase 3 of slot g + 1, [store phase 2]
            // per ring buffer (nd <= 8 slots), with a prologue
            for (int blk = 0; blk < NB; blk++) {
                const int b = blk & 1, step0 = blk * DB;
                const int nd = (p.walks - step0) < DB ? (p.walks - step0) : DB;
                const int oXb = L::O_X + b * DB * XB, oFb = L::O_F + b * DB * CH;
                nbar_sync(BAR_FULL + b, 384);                     // the draw warps have filled this buffer
                if (has_item) {
                    double d0, d1, e0, e1;
                    phase2_issue(oXb, d0, d1, e0, e1);
                    phase2_store(0, d0, d1, e0, e1);
                    if (nd > 1) {
                        phase2_issue(oXb + XB, d0, d1, e0, e1);
                        phase2_store(1, d0, d1, e0, e1);
                    }
                }
                nbar_sync(BAR_STEP, 256);
                if (live) phase3(oXb, b2n_sm[oFb + c], L::O_Y + c * YS);
                nbar_sync(BAR_STEP, 256);
                for (int s2 = 0; s2 < nd; s2++) {
                    phase4(oXb + s2 * XB);
                    nbar_sync(BAR_STEP, 256);
                    // every read of this ring buffer is done: hand it back to the draw warps (if they will ask for it)
                    if (s2 == nd - 1 && blk + 2 < NB) nbar_arrive(BAR_EMPTY + b, 384);
                    const bool p2 = has_item && s2 + 2 < nd;
                    double d0, d1, e0, e1;
                    if (p2) phase2_issue(oXb + (s2 + 2) * XB, d0, d1, e0, e1);
                    if (live) phase5();
                    if (s2 + 1 < nd) {
                        if (live) phase3(oXb + (s2 + 1) * XB, b2n_sm[oFb + (s2 + 1) * CH + c], L::O_Y + (((s2 + 1) & 1) * CH + c) * YS);
                        if (p2) phase2_store(s2 & 1, d0, d1, e0, e1);
                        nbar_sync(BAR_STEP, 256);
                    }
                }
            }
            if (live) {
                if (nacc == 0) {   // recompute (v, logl) of the start point (:970-975), warp-local
                    const int oy = L::O_Y + c * YS;
                    __syncwarp();
                    for (int i = lane; i < n; i += 32) {
                        const double vi = prior_sm(pk, L::O_P0, L::O_P1, i, b2n_sm[oucur + i]);
                        b2n_sm[ovcur + i] = vi;
                        b2n_sm[oy + i] = vi - b2n_sm[L::O_MU + i];
                    }
                    __syncwarp();
                    lcur = fma(-0.5, quadform_full<false>(p.m.lmat, 0, n, n, oy, lane), p.m.s0);
                }
                __syncwarp();
                for (int i = lane; i < n; i += 32) {
                    peer_put(p.peer, &p.u[(size_t)q * n + i], b2n_sm[oucur + i]);
                    peer_put(p.peer, &p.v[(size_t)q * n + i], b2n_sm[ovcur + i]);
                }
                if (lane == 0) {
                    peer_put(p.peer, &p.logl[q], lcur);
                    peer_put(p.peer, &p.nacc[q], nacc);
                    peer_put(p.peer, &p.nrej[q], nrej);
                    peer_put(p.peer, &p.ncall[q], (int)p.walks);
                }
            }
            __syncthreads();
            for (int e = threadIdx.x; e < 2 * L::RING; e += blockDim.x) b2n_sm[L::O_X + e] = 0.0;
            __syncthreads();
#undef oucur
#undef ouprop
#undef ovcur
#undef ovprop
        }
    }
    peer_finish(p.peer);
}

// =====================================================================================
// rwalk_mmas_kernel -- lock-step DMMA kernel for LARGE n (64 < n: the matrix fragments do not
// fit in registers, and at n = 200 the 320 KB axes matrix does not even fit in shared memory).
// Same phases as rwalk_mma_kernel with 16 chains per CTA, but the A-operand fragments are
// STREAMED from global memory (L2-resident) every step: each 8x4 fragment is loaded once per
// CTA-step and feeds two DMMAs (the two 8-chain tiles), so the L2 traffic per proposal is
// n^2*8/16 bytes instead of the n^2*8 of the warp-per-chain kernel (20 KB vs 320 KB at n=200).
// Warp w owns slabs w, w+16, ...  BASELINE config C4 (200-D, single ellipsoid, rwalk).
// =====================================================================================
template <int LIKE>
__global__ void __launch_bounds__(512, 1) rwalk_mmas_kernel(const RwalkParams p, int XS, int YS) {
    constexpr int CH = 16;
    const int n = p.n;
    const int npad = (n + 1) & ~1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    B2N_DYN_PROLOGUE(p)
    const int3 cd = p.cta[blockIdx.x];
    const int S = (n + 7) >> 3, KT = (n + 3) >> 2;
    int off = 0;
    const ModelSm ms = stage_model(p.m, off, n, npad);
    const int op0 = ms.op0, op1 = ms.op1, omu = ms.olv0;
    off += 4 * npad;
    uint32_t* fl = reinterpret_cast<uint32_t*>(&b2n_sm[off]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) fl[i] = p.dimflags ? p.dimflags[i] : 0u;
    off += ((n + 3) >> 2) << 1;
    const int oX = off;  off += CH * XS;
    const int oY = off;  off += CH * YS;
    const int oQ = off;  off += S * CH;
    const int ost = off;
    for (int e = threadIdx.x; e < CH * XS; e += blockDim.x) b2n_sm[oX + e] = 0.0;
    const double* __restrict__ Ag = p.axesT + (size_t)cd.z * n * n;     // axesT[col*n + row] = axes[row][col]
    const double* __restrict__ Pg = p.m.lmat;                             // symmetric
    __syncthreads();
    const double inv_n = 1.0 / (double)n;
    const int pk = p.m.prior_kind;
    const int lr = lane >> 2, lc = lane & 3;

    // ---- helper scratch: step factor, U^(1/n), which state buffer is current, per-helper cube flags
    const int oH = ost + CH * 4 * npad;
    double* facbuf = &b2n_sm[oH];
    double* pwbuf = &b2n_sm[oH + CH];
    int* selbuf = reinterpret_cast<int*>(&b2n_sm[oH + 2 * CH]);
    int* okbuf = reinterpret_cast<int*>(&b2n_sm[oH + 3 * CH]);            // CH x CH

    for (int g0 = 0; g0 < cd.y; g0 += CH) {
        // L chains are live in this pass.  A pass with few chains (the rounds of b2n_ns_run put 1-2 chains on a CTA)
        // used to leave 14 of the 16 warps idle outside the contraction while the chain's own warp generated its
        // 200 normals (4 Philox / log / sincos rounds), U^(1/n) and 200 ndtri's one after the other: H = 16 / L
        // warps now serve a chain -- the owner (h = 0: chain state, accept test, counters) and H - 1 helpers that
        // take their share of the ELEMENTWISE work (normal blocks, the radius power, wrap / cube test / prior / delta
        // of their elements).  Every element is computed by the same instructions as before and every sum is still
        // formed by the owner in the old order (|z|^2 from the stored z), so a chain's result does not depend on
        // how many warps worked on it -- nor, therefore, on the batch it is part of.
        const int L = min(CH, cd.y - g0);
        const int H = CH / L;
        const int c = warp % L, h = warp / L;
        const bool owner = h == 0, helper = h < H;
        const int q = p.order[cd.x + g0 + c];
        const int sbase = ost + c * 4 * npad;
        int oucur = sbase, ouprop = sbase + npad, ovcur = ouprop + npad, ovprop = ovcur + npad;
        const int ox = oX + c * XS, oy = oY + c * YS;
        const bool two = L > 8;                   // chains 8..15 exist: second column tile of the contractions
        ChainRng g;
        g.init(p.seed, chain0_ + (uint64_t)q);
        if (owner) {
            for (int i = lane; i < n; i += 32) b2n_sm[oucur + i] = p.u0[(size_t)(p.start ? p.start[q] : q) * n + i];
            if (lane == 0) selbuf[c] = 0;
        }
        int nacc = 0, nrej = 0;
        double lcur = 0.0;
        const int nb = (n + 1) >> 1;
        __syncthreads();
        for (int step = 0; step < p.walks; step++) {
            // ---- draw events of the step: normal vector (tick 2 step), radius uniform (tick 2 step + 1)
            if (helper) {
                g.tick = 2u * (uint32_t)step;
                for (int b = 32 * h + lane; b < nb; b += 32 * H) {
                    double z0, z1;
                    rng_normal_pair(g, b, z0, z1);
                    if (2 * b + 1 < n) *reinterpret_cast<double2*>(&b2n_sm[ox + 2 * b]) = make_double2(z0, z1);
                    else b2n_sm[ox + 2 * b] = z0;
                }
                if (h == H - 1) {
                    g.tick = 2u * (uint32_t)step + 1u;
                    const double U = rng_uniform(g);
                    const double pw = pow(U, inv_n);
                    if (lane == 0) pwbuf[c] = pw;
                }
            }
            __syncthreads();
            if (owner) {                              // |z|^2 in normals_sm's order, then the step factor
                double ss = 0.0;
                for (int b = lane; b < nb; b += 32) {
                    const double z0 = b2n_sm[ox + 2 * b];
                    ss = fma(z0, z0, ss);
                    if (2 * b + 1 < n) { const double z1 = b2n_sm[ox + 2 * b + 1]; ss = fma(z1, z1, ss); }
                }
                ss = warp_sum(ss);
                if (lane == 0) facbuf[c] = scale_ * (pwbuf[c] / sqrt(ss));
            }
            // ---- Y = A X, fragments of A streamed from L2 (slabs dealt from the LAST warp down: the owners -- warps
            //      0 .. L-1, which have just formed the step factors -- get the fewest)
            for (int s = CH - 1 - warp; s < S; s += CH) {
                const int row = 8 * s + lr;
                const bool rv = row < n;
                const double* ap = Ag + (rv ? row : 0) + (size_t)lc * n;
                const int xb0 = oX + lr * XS + lc, xb1 = xb0 + 8 * XS;
                double d00 = 0, d01 = 0, d10 = 0, d11 = 0;
                if (two) {
#pragma unroll 8
                    for (int kt = 0; kt < KT; kt++) {
                        const bool in = rv && (4 * kt + lc) < n;
                        const double a = in ? __ldg(ap + (size_t)(4 * kt) * n) : 0.0;
                        dmma884(d00, d01, a, b2n_sm[xb0 + 4 * kt]);
                        dmma884(d10, d11, a, b2n_sm[xb1 + 4 * kt]);
                    }
                } else {
                    // (one column tile: the loop is a load and a DMMA -- 25 loads in flight per warp instead of 8,
                    //  the phase is bound by the L2 round trips of its fragment loads)
#pragma unroll 25
                    for (int kt = 0; kt < KT; kt++) {
                        const bool in = rv && (4 * kt + lc) < n;
                        const double a = in ? __ldg(ap + (size_t)(4 * kt) * n) : 0.0;
                        dmma884(d00, d01, a, b2n_sm[xb0 + 4 * kt]);
                    }
                }
                const int c0 = 2 * lc;
                b2n_sm[oY + c0 * YS + row] = d00;
                b2n_sm[oY + (c0 + 1) * YS + row] = d01;
                if (two) {
                    b2n_sm[oY + (8 + c0) * YS + row] = d10;
                    b2n_sm[oY + (9 + c0) * YS + row] = d11;
                }
            }
            __syncthreads();
            // ---- proposal u' = u + fac y, wrap / reflect / cube test, prior, delta: elements dealt over the H warps
            if (helper) {
                const double fac = facbuf[c];
                const int sel = selbuf[c];
                const int ucur_ = sel ? sbase + npad : sbase, uprop_ = sel ? sbase : sbase + npad;
                const int vprop_ = sel ? sbase + 2 * npad : sbase + 3 * npad;
                bool okp = true;
                for (int i = 32 * h + lane; i < n; i += 32 * H) {
                    double t = fma(fac, b2n_sm[oy + i], b2n_sm[ucur_ + i]);
                    const uint32_t f = fl[i];
                    if (f & B2N_DIM_PERIODIC) t = mod1(t);
                    if (f & B2N_DIM_REFLECTIVE) t = reflect1(t);
                    okp = okp && in_cube(t, f);
                    const double vi = prior_sm(pk, op0, op1, i, t);
                    b2n_sm[uprop_ + i] = t;
                    b2n_sm[vprop_ + i] = vi;
                    b2n_sm[ox + i] = vi - b2n_sm[omu + i];
                }
                okp = __all_sync(B2N_FULL, okp);
                if (lane == 0) okbuf[c * CH + h] = okp ? 1 : 0;
            }
            __syncthreads();
            bool ok = true;
            if (owner)
                for (int k = 0; k < H; k++) ok = ok && (okbuf[c * CH + k] != 0);
            double l = 0.0;
            if (LIKE == B2N_LIKE_GAUSS_PREC) {
                for (int s = warp; s < S; s += CH) {
                    const int row = 8 * s + lr;
                    const bool rv = row < n;
                    const double* pp = Pg + (rv ? row : 0) + (size_t)lc * n;
                    const int xb0 = oX + lr * XS + lc, xb1 = xb0 + 8 * XS;
                    double d00 = 0, d01 = 0, d10 = 0, d11 = 0;
#pragma unroll 8
                    for (int kt = 0; kt < KT; kt++) {
                        const bool in = rv && (4 * kt + lc) < n;
                        const double a = in ? __ldg(pp + (size_t)(4 * kt) * n) : 0.0;
                        dmma884(d00, d01, a, b2n_sm[xb0 + 4 * kt]);
                        if (two) dmma884(d10, d11, a, b2n_sm[xb1 + 4 * kt]);
                    }
                    const int c0 = 2 * lc;
                    double q00 = d00 * b2n_sm[oX + c0 * XS + row], q01 = d01 * b2n_sm[oX + (c0 + 1) * XS + row];
                    double q10 = d10 * b2n_sm[oX + (8 + c0) * XS + row], q11 = d11 * b2n_sm[oX + (9 + c0) * XS + row];
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        q00 += __shfl_xor_sync(B2N_FULL, q00, o);
                        q01 += __shfl_xor_sync(B2N_FULL, q01, o);
                        q10 += __shfl_xor_sync(B2N_FULL, q10, o);
                        q11 += __shfl_xor_sync(B2N_FULL, q11, o);
                    }
                    if (lane < 4) {
                        b2n_sm[oQ + s * CH + c0] = q00;
                        b2n_sm[oQ + s * CH + c0 + 1] = q01;
                        b2n_sm[oQ + s * CH + 8 + c0] = q10;
                        b2n_sm[oQ + s * CH + 9 + c0] = q11;
                    }
                }
                __syncthreads();
                double qf = 0.0;
                for (int s = 0; s < S; s++) qf += b2n_sm[oQ + s * CH + c];
                l = fma(-0.5, qf, p.m.s0);
            } else if (owner && ok) {
                l = loglike_sm<LIKE, false>(p.m, ms, nullptr, 0, n, n, ovprop, oy, lane);
            }
            if (owner) {
                if (!ok) {
                    nrej++;
                } else if (l > loglstar_) {
                    int t = oucur; oucur = ouprop; ouprop = t;
                    t = ovcur; ovcur = ovprop; ovprop = t;
                    lcur = l;
                    nacc++;
                    if (lane == 0) selbuf[c] ^= 1;
                } else {
                    nrej++;
                }
            }
        }
        if (owner) {
            if (nacc == 0) {
                for (int i = lane; i < n; i += 32) b2n_sm[ovcur + i] = prior_sm(pk, op0, op1, i, b2n_sm[oucur + i]);
                __syncwarp();
                lcur = loglike_sm<LIKE, false>(p.m, ms, p.m.lmat, 0, n, n, ovcur, oy, lane);
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                peer_put(p.peer, &p.u[(size_t)q * n + i], b2n_sm[oucur + i]);
                peer_put(p.peer, &p.v[(size_t)q * n + i], b2n_sm[ovcur + i]);
            }
            if (lane == 0) {
                peer_put(p.peer, &p.logl[q], lcur);
                peer_put(p.peer, &p.nacc[q], nacc);
                peer_put(p.peer, &p.nrej[q], nrej);
                peer_put(p.peer, &p.ncall[q], (int)p.walks);
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < CH * XS; e += blockDim.x) b2n_sm[oX + e] = 0.0;
        __syncthreads();
    }
    peer_finish(p.peer);
}

// Host-side grouping of chains by ellipsoid -> per-CTA work descriptors.
// (shared with the slice kernels)
int b2n_build_worklist(b2n_ctx* ctx, int64_t Q, const int32_t* ell, int K, int chains_per_cta,
                       std::vector<int>& order, std::vector<int3>& cta) {
    order.resize(Q);
    cta.clear();
    std::vector<int64_t> count(K + 1, 0);
    if (ell) {
        for (int64_t q = 0; q < Q; q++) {
            if (ell[q] < 0 || ell[q] >= K) return b2n_fail(ctx, B2N_ERR_ARG, "chain ellipsoid index out of range");
            count[ell[q] + 1]++;
        }
    } else {
        count[1] = Q;
    }
    for (int k = 0; k < K; k++) count[k + 1] += count[k];
    std::vector<int64_t> pos(count.begin(), count.end() - 1);
    for (int64_t q = 0; q < Q; q++) order[pos[ell ? ell[q] : 0]++] = (int)q;
    for (int k = 0; k < K; k++) {
        // split the group into equal CTAs (no short tail CTA)
        const int64_t c = count[k + 1] - count[k];
        if (c == 0) continue;
        const int64_t parts = (c + chains_per_cta - 1) / chains_per_cta;
        for (int64_t i = 0; i < parts; i++) {
            const int64_t lo = count[k] + c * i / parts, hi = count[k] + c * (i + 1) / parts;
            cta.push_back(make_int3((int)lo, (int)(hi - lo), k));
        }
    }
    return B2N_OK;
}

int b2n_worklist_dev(b2n_ctx* ctx, int64_t Q, const int32_t* ell, int K, int chains_per_cta, const void** dorder,
                     const void** dcta, unsigned* ncta) {
    const bool trivial = (ell == nullptr || K == 1);
    if (trivial && ell)
        for (int64_t q = 0; q < Q; q++)
            if (ell[q] != 0) return b2n_fail(ctx, B2N_ERR_ARG, "chain ellipsoid index out of range");
    if (trivial && ctx->wl_Q == Q && ctx->wl_cpc == chains_per_cta && ctx->wl_order.p && ctx->wl_cta.p) {
        *dorder = ctx->wl_order.p; *dcta = ctx->wl_cta.p; *ncta = (unsigned)ctx->wl_ncta;
        return B2N_OK;
    }
    std::vector<int> order;
    std::vector<int3> cta;
    B2N_TRY(b2n_build_worklist(ctx, Q, trivial ? nullptr : ell, K, chains_per_cta, order, cta));
    *ncta = (unsigned)cta.size();
    if (trivial) {
        // (the buffers may still be read by an enqueued kernel of a device-pointer caller)
        B2N_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        B2N_CUDA(ctx, ctx->wl_order.ensure(order.size() * sizeof(int)));
        B2N_CUDA(ctx, ctx->wl_cta.ensure(cta.size() * sizeof(int3)));
        B2N_CUDA(ctx, cudaMemcpy(ctx->wl_order.p, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
        B2N_CUDA(ctx, cudaMemcpy(ctx->wl_cta.p, cta.data(), cta.size() * sizeof(int3), cudaMemcpyHostToDevice));
        ctx->wl_Q = Q; ctx->wl_cpc = chains_per_cta; ctx->wl_ncta = (int)cta.size();
        *dorder = ctx->wl_order.p; *dcta = ctx->wl_cta.p;
        return B2N_OK;
    }
    B2N_TRY(b2n_in_host(ctx, ctx->work0, order.data(), order.size() * sizeof(int), dorder));
    B2N_TRY(b2n_in_host(ctx, ctx->work1, cta.data(), cta.size() * sizeof(int3), dcta));
    return B2N_OK;
}

// Persistent-sized grid: one CTA per SM when every chain can have its own warp
// (Q <= 16 x SMs), two per SM beyond that; warps loop over their CTA's chains.
void b2n_chain_grid(const b2n_ctx* ctx, int64_t Q, int max_warps, int& chains_per_cta, int& warps) {
    const int64_t sms = ctx->sm_count;
    const int64_t ctas = (Q <= 16 * sms) ? sms : 2 * sms;
    chains_per_cta = (int)std::max<int64_t>(std::min(ctx->min_cpc, 16), (Q + ctas - 1) / ctas);
    warps = std::max(1, std::min(max_warps, std::min(16, chains_per_cta)));
}

extern "C" int b2n_rwalk_batch(b2n_ctx* ctx, const b2n_chain_args* a, int32_t walks, double* u,
                               double* v, double* logl, int32_t* n_accept, int32_t* n_reject,
                               int32_t* ncall) {
    if (!ctx || !a) return B2N_ERR_ARG;
    // start points by index (b2n_set_start_rows): consumed by THIS call, however it ends
    const int32_t* sidx = ctx->start_idx;
    const int64_t srows = ctx->start_nrows;
    ctx->start_idx = nullptr; ctx->start_nrows = 0;
    const bool gather = ctx->peer.total > 0;      // outputs may be NULL in gather mode (b2n_peer_result)
    if (!gather && (!u || !v || !logl || !n_accept || !n_reject || !ncall)) return B2N_ERR_ARG;
    if (a->model_id < 0 || a->model_id >= (int)ctx->models.size()) return B2N_ERR_ARG;
    const B2nModel m = ctx->models[a->model_id];
    const int n = a->ndim, nc = a->ncdim;
    const int64_t Q = a->nchain;
    if (n != m.ndim || nc < 1 || nc > n || walks < 1 || Q < 0 || !a->u0) return B2N_ERR_ARG;
    if (ctx->bK < 1 || ctx->bn != nc) return b2n_fail(ctx, B2N_ERR_ARG, "resident bound missing or of wrong dimension (b2n_bound_set)");
    if (Q == 0) return gather ? b2n_fail(ctx, B2N_ERR_ARG, "gather mode: every rank must run at least one chain") : B2N_OK;
    B2N_CUDA(ctx, cudaSetDevice(ctx->device));
    ZcScope zc(ctx);          // pinned caller buffers are read / written in place (host-pointer mode)

    // shared-memory plan: per-warp state always; matrices (128-byte padded columns) when they fit
    const int npad = (n + 1) & ~1;
    const size_t per_warp = (size_t)6 * npad * sizeof(double);
    const size_t flags_b = (size_t)((((n + 3) >> 2) << 1) + 4 * npad) * sizeof(double);
    const size_t limit = (size_t)ctx->max_smem_optin;
    const int max_warps = (int)std::min<size_t>(16, (limit - flags_b) / per_warp);
    if (max_warps < 1) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "ndim too large for the rwalk kernel");
    int chains_per_cta, warps;
    b2n_chain_grid(ctx, Q, max_warps, chains_per_cta, warps);
    // lock-step DMMA kernel: matrices as register fragments (needs ncdim == ndim, 16 <= n <= 64
    // and enough chains to fill CTAs).  B2N_RWALK_IMPL=warp|mma forces one of the two.
    const char* impl = getenv("B2N_RWALK_IMPL");
    // (the choice depends on the problem shape only, never on the queue size: a chain's result
    // must not depend on which batch it is part of -- sharded multi-GPU runs rely on that)
    bool use_mma = nc == n && n >= 16 && n <= 64;
    if (impl && !strcmp(impl, "warp")) use_mma = false;
    if (impl && !strcmp(impl, "mma")) {
        if (!(nc == n && n >= 4 && n <= 64)) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "B2N_RWALK_IMPL=mma needs ncdim == ndim <= 64");
        use_mma = true;
    }
    const int KT = n <= 32 ? 8 : (n <= 52 ? 13 : 16);
    // direction ring depth of the lock-step kernel (B2N_RWALK_DEPTH=1: one step at a time, as before the
    // ring; results identical)
    int DU = 8;
    if (const char* e = getenv("B2N_RWALK_DEPTH")) DU = (atoi(e) == 1) ? 1 : 8;
    // draws: branch-free math, two ring items per warp at a time (default); B2N_RWALK_DRAWS=libm = libdevice
    // log / sqrt / sincospi, one item at a time (measured 3.6 % slower per C2 launch, results differ by a few ulp)
    const char* denv = getenv("B2N_RWALK_DRAWS");
    const bool fast_draws = !(denv && !strcmp(denv, "libm")) && DU == 8;
    const int occ = 2;                                      // 8-chain CTAs resident per SM
    // Lock-step variants for the precision-matrix Gaussian (B2N_RWALK_WARPS forces one; measured at C2, profiles/r2n-r2q):
    //   12 = rwalk_mmaws_kernel, 8 step + 4 draw warps -- the default where it applies: its static schedule of the
    //        symmetric quadratic form is laid out for the largest slab count of a KT, and the draws need an idle lane 31
    //        for the radius (n <= 62): n in 25..32, 49..52, 57..62;
    //    8 = rwalk_mma_kernel (every other shape and likelihood);
    //   16 = rwalk_mma16_kernel, sixteen warps per 8 chains (slower: kept as the measured counter-example).
    int want = 12;
    if (const char* e = getenv("B2N_RWALK_WARPS")) want = atoi(e);
    const bool ws_base = use_mma && n <= 62 && m.like_kind == B2N_LIKE_GAUSS_PREC && fast_draws && occ == 2;
    const bool full_slabs = ((n + 7) >> 3) == (8 * ((4 * KT + 7) / 8)) / 8;
    const int use_ws = (want == 12 && ws_base && full_slabs) ? 12 : 0;
    const bool use_mma16 = want == 16 && ws_base;
    size_t mma_smem = 0;
    if (use_mma) {
        const int ctas = occ * ctx->sm_count;               // 8-chain CTAs, two (three) resident per SM
        chains_per_cta = (int)std::max<int64_t>(std::min(ctx->min_cpc, 8), (Q + ctas - 1) / ctas);
        warps = 8;
        const int RS = 8 * ((4 * KT + 7) / 8);
        const int XS = RS + ((RS % 16 == 4) ? 0 : ((20 - RS % 16) % 16)), YS = RS + 2;
        mma_smem = (size_t)(4 * npad + (((n + 3) >> 2) << 1) + DU * 8 * XS + 8 * YS + 8 * 8 + DU * 8 + 8 * 4 * npad) * sizeof(double);
        if (use_ws)
            mma_smem = (size_t)(KT == 8 ? MmaWsLayout<8>::TOTAL : (KT == 13 ? MmaWsLayout<13>::TOTAL : MmaWsLayout<16>::TOTAL)) * sizeof(double);
        if (use_mma16)
            mma_smem = (size_t)(KT == 8 ? Mma16Layout<8>::TOTAL : (KT == 13 ? Mma16Layout<13>::TOTAL : Mma16Layout<16>::TOTAL)) * sizeof(double);
    }
    // large n: lock-step kernel with matrix fragments streamed from L2 (16 chains share each load)
    bool use_mmas = false;
    int sXS = 0, sYS = 0;
    if (!use_mma && nc == n && n > 64 && !(impl && !strcmp(impl, "warp"))) {
        const int RS = 8 * ((n + 7) / 8);
        sXS = RS + ((RS % 16 == 4) ? 0 : ((20 - RS % 16) % 16));
        sYS = RS + 2;
        const size_t need = (size_t)(4 * npad + (((n + 3) >> 2) << 1) + 16 * sXS + 16 * sYS + (RS / 8) * 16 +
                                     16 * 4 * npad + 3 * 16 + 16 * 16 / 2) * sizeof(double);      // + the helper scratch
        if (need <= limit) {
            use_mmas = true;
            mma_smem = need;
            const int ctas = ctx->sm_count;
            chains_per_cta = (int)std::max<int64_t>(std::min(ctx->min_cpc, 16), (Q + ctas - 1) / ctas);
            warps = 16;
        }
    }
    const size_t fixed = per_warp * warps + flags_b;
    const int ldA = (nc + 15) & ~15, ldP = (n + 15) & ~15;
    const size_t ax_b = (size_t)nc * ldA * sizeof(double);
    const size_t pr_b = (m.like_kind == B2N_LIKE_GAUSS_PREC) ? (size_t)n * ldP * sizeof(double) : 0;
    const bool ax_s = fixed + ax_b <= limit;
    const bool pr_s = pr_b > 0 && fixed + (ax_s ? ax_b : 0) + pr_b <= limit;
    const size_t smem = (use_mma || use_mmas) ? mma_smem : fixed + (ax_s ? ax_b : 0) + (pr_s ? pr_b : 0);

    const bool dyn = ctx->dyn.active;        // device-paced launch (b2n_ns.cu): worklist + scalars in HBM
    if (dyn) {
        ctx->dyn.cpc = chains_per_cta;
        if (ctx->dyn.plan_only) return B2N_OK;
        if (gather || ctx->ptr_mode != B2N_PTR_DEVICE) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "device-paced launch needs device pointers and no gather mode");
    }
    RwalkParams p;
    p.dyn = dyn ? ctx->dyn.dev : nullptr;
    p.m = m; p.n = n; p.nc = nc; p.walks = walks; p.ldA = ldA; p.ldP = ldP;
    p.loglstar = a->loglstar; p.scale = a->scale; p.seed = a->seed; p.chain0 = a->chain0;
    p.axesT = ctx->b_axesT.as<double>();
    const void *du0, *dorder, *dcta, *dfl = nullptr, *dstart = nullptr;
    // start points by index: u0 is then the whole live set
    if (sidx) {
        if (dyn) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "start rows by index: not in a device-paced launch");
        if (ctx->ptr_mode == B2N_PTR_DEVICE) dstart = sidx;
        else {
            for (int64_t i = 0; i < Q; i++)
                if (sidx[i] < 0 || sidx[i] >= srows) return b2n_fail(ctx, B2N_ERR_ARG, "start row index out of range");
            B2N_TRY(b2n_in_host(ctx, ctx->in1, sidx, (size_t)Q * sizeof(int32_t), &dstart));
        }
    }
    B2N_TRY(b2n_in(ctx, ctx->in0, a->u0, (size_t)(sidx ? srows : Q) * n * sizeof(double), &du0));
    unsigned ncta = 0;
    if (dyn) {
        dorder = ctx->dyn.order; dcta = ctx->dyn.cta;
    } else {
        B2N_TRY(b2n_worklist_dev(ctx, Q, a->ell, ctx->bK, chains_per_cta, &dorder, &dcta, &ncta));
    }
    std::vector<uint32_t> fl;
    if (a->dimflags) {
        fl.assign(a->dimflags, a->dimflags + n);
        B2N_TRY(b2n_in_host(ctx, ctx->in3, fl.data(), fl.size() * sizeof(uint32_t), &dfl));
    }
    void *du, *dv, *dl, *dna, *dnr, *dncl;
    void* gdev[7];
    bool peer_on = false;
    B2N_TRY(b2n_peer_begin(ctx, n, &p.peer, gdev, &peer_on));
    if (peer_on) {
        if (ctx->peer.row0 + Q > ctx->peer.total) return b2n_fail(ctx, B2N_ERR_ARG, "gather rows out of range (b2n_peer_rows)");
        du = gdev[0]; dv = gdev[1]; dl = gdev[2]; dna = gdev[3]; dnr = gdev[4]; dncl = gdev[5];
    } else {
        B2N_TRY(b2n_out(ctx, ctx->out0, u, (size_t)Q * n * sizeof(double), &du));
        B2N_TRY(b2n_out(ctx, ctx->out1, v, (size_t)Q * n * sizeof(double), &dv));
        B2N_TRY(b2n_out(ctx, ctx->out2, logl, (size_t)Q * sizeof(double), &dl));
        B2N_TRY(b2n_out(ctx, ctx->out3, n_accept, (size_t)Q * sizeof(int), &dna));
        B2N_TRY(b2n_out(ctx, ctx->out4, n_reject, (size_t)Q * sizeof(int), &dnr));
        B2N_TRY(b2n_out(ctx, ctx->out5, ncall, (size_t)Q * sizeof(int), &dncl));
    }
    p.u0 = (const double*)du0; p.start = (const int*)dstart; p.order = (const int*)dorder; p.cta = (const int3*)dcta;
    p.dimflags = (const uint32_t*)dfl;
    p.u = (double*)du; p.v = (double*)dv; p.logl = (double*)dl;
    p.nacc = (int*)dna; p.nrej = (int*)dnr; p.ncall = (int*)dncl;

    const unsigned grid = dyn ? (unsigned)ctx->dyn.max_cta : ncta;
#define LAUNCH(L, AXS, PRS)                                                                       \
    do {                                                                                          \
        B2N_TRY(b2n_func_smem(ctx, (const void*)(rwalk_kernel<L, AXS, PRS>), (size_t)(smem))); \
        rwalk_kernel<L, AXS, PRS><<<grid, warps * 32, smem, ctx->stream>>>(p);                    \
    } while (0)
#define CALL(L)                                                        \
    if (ax_s && pr_s) LAUNCH(L, true, true);                           \
    else if (ax_s) LAUNCH(L, true, false);                             \
    else if (pr_s) LAUNCH(L, false, true);                             \
    else LAUNCH(L, false, false);
#define LAUNCH_MMA2(L, K, D, F)                                                                      \
    do {                                                                                            \
        B2N_TRY(b2n_func_smem(ctx, (const void*)(rwalk_mma_kernel<L, K, 8, D, F>), (size_t)(smem))); \
        rwalk_mma_kernel<L, K, 8, D, F><<<grid, 256, smem, ctx->stream>>>(p);                        \
    } while (0)
#define LAUNCH_MMA(L, K)                        \
    if (DU == 1) LAUNCH_MMA2(L, K, 1, false);   \
    else if (fast_draws) LAUNCH_MMA2(L, K, 8, true); \
    else LAUNCH_MMA2(L, K, 8, false);
#define CALL_MMA(L)                      \
    if (KT == 8) { LAUNCH_MMA(L, 8) }    \
    else if (KT == 13) { LAUNCH_MMA(L, 13) } \
    else { LAUNCH_MMA(L, 16) }
#define CALL_MMAS(L)                                                                                  \
    B2N_TRY(b2n_func_smem(ctx, (const void*)(rwalk_mmas_kernel<L>), (size_t)(smem)));                                                  \
    rwalk_mmas_kernel<L><<<grid, 512, smem, ctx->stream>>>(p, sXS, sYS);
#define LAUNCH_MMA16(K)                                                                             \
    do {                                                                                            \
        B2N_TRY(b2n_func_smem(ctx, (const void*)(rwalk_mma16_kernel<K>), (size_t)(smem)));          \
        rwalk_mma16_kernel<K><<<grid, 512, smem, ctx->stream>>>(p);                                 \
    } while (0)
#define LAUNCH_MMAWS(K, PL)                                                                        \
    do {                                                                                            \
        B2N_TRY(b2n_func_smem(ctx, (const void*)(rwalk_mmaws_kernel<K, 1, PL>), (size_t)(smem)));   \
        rwalk_mmaws_kernel<K, 1, PL><<<grid, 384, smem, ctx->stream>>>(p);                          \
    } while (0)
    B2N_TIME_BEGIN(ctx);
    if (use_ws) {
        // plain = no wrapped dimension and an affine prior: straight-line phase 3
        bool plain = m.prior_kind != B2N_PRIOR_NORMAL_PPF;
        if (a->dimflags)
            for (int i = 0; i < n; i++) plain = plain && a->dimflags[i] == 0u;
        if (KT == 8) { if (plain) LAUNCH_MMAWS(8, true); else LAUNCH_MMAWS(8, false); }
        else if (KT == 13) { if (plain) LAUNCH_MMAWS(13, true); else LAUNCH_MMAWS(13, false); }
        else { if (plain) LAUNCH_MMAWS(16, true); else LAUNCH_MMAWS(16, false); }
    } else if (use_mma16) {
        if (KT == 8) LAUNCH_MMA16(8);
        else if (KT == 13) LAUNCH_MMA16(13);
        else LAUNCH_MMA16(16);
    } else if (use_mma) {
        B2N_DISPATCH_LIKE(m.like_kind, CALL_MMA)
    } else if (use_mmas) {
        B2N_DISPATCH_LIKE(m.like_kind, CALL_MMAS)
    } else {
        B2N_DISPATCH_LIKE(m.like_kind, CALL)
    }
    B2N_TIME_END(ctx);
#undef LAUNCH_MMAWS
#undef LAUNCH_MMA16
#undef CALL_MMAS
#undef CALL_MMA
#undef LAUNCH_MMA
#undef LAUNCH_MMA2
#undef CALL
#undef LAUNCH
    B2N_LAUNCH_CHECK(ctx);
    if (peer_on) {
        void* const user7[7] = {u, v, logl, n_accept, n_reject, ncall, nullptr};
        B2N_TRY(b2n_peer_end(ctx, n, user7));
        return b2n_peer_finish(ctx, true);
    }
    B2N_TRY(b2n_out_done(ctx, u, du, (size_t)Q * n * sizeof(double)));
    B2N_TRY(b2n_out_done(ctx, v, dv, (size_t)Q * n * sizeof(double)));
    B2N_TRY(b2n_out_done(ctx, logl, dl, (size_t)Q * sizeof(double)));
    B2N_TRY(b2n_out_done(ctx, n_accept, dna, (size_t)Q * sizeof(int)));
    B2N_TRY(b2n_out_done(ctx, n_reject, dnr, (size_t)Q * sizeof(int)));
    B2N_TRY(b2n_out_done(ctx, ncall, dncl, (size_t)Q * sizeof(int)));
    return b2n_finish(ctx);
}
