// b2n_unif.cu -- batched uniform sampling within the resident (multi-)ellipsoid bound.
//
// Replaces UniformBoundSampler.sample (reference internal_samplers.py:243-340) whose
// bound draw is MultiEllipsoid.sample (bounding.py:525-590) / Ellipsoid.sample
// (:307-319): pick an ellipsoid with probability proportional to its volume (rand_choice,
// :1300-1308), draw uniformly inside it (randsphere, :1288-1297), count the q ellipsoids
// containing the draw (strict <1, with the 1e-3 slack retry of :565-579) and accept with
// probability 1/q; reject draws outside the unit cube (internal_samplers.py:314-322);
// append fresh U(0,1) for the non-clustered dims (:325-327); evaluate; repeat until
// logl > loglstar.  One warp per chain.
#include "b2n_device.cuh"
#include <vector>

#define B2N_UNIF_MAX_DRAWS 20000000

struct UnifParams {
    B2nModel m;
    int n, nc, K, draw_only;
    const double* ctrs;     // K x nc
    const double* ams;      // K x nc x nc
    const double* axesT;    // K x nc x nc (transposed)
    const double* cum;      // K cumulative volume fractions
    const uint32_t* dimflags;
    double loglstar;
    uint64_t seed, chain0;
    int64_t Q;
    double *u, *v, *logl;
    int *ncall, *nprop;
    uint32_t* flags;
    PeerSet peer;          // fused multi-GPU gather of the outputs (b2n_peer.cu)
    const B2nDyn* dyn;     // device-paced launch (b2n_ns.cu): threshold / chain ids in HBM
};

template <int LIKE>
__global__ void __launch_bounds__(256) unif_kernel(const UnifParams p) {
    extern __shared__ double sm[];
    const int n = p.n, nc = p.nc, K = p.K;
    double loglstar_ = p.loglstar;
    unsigned long long chain0_ = p.chain0;
    if (p.dyn) {
        if (p.dyn->skip) return;
        loglstar_ = p.dyn->loglstar; chain0_ = p.dyn->chain0;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double* uu = sm + (size_t)warp * 5 * n;   // candidate point (n)
    double* z = uu + n;                        // unit-ball draw (nc)
    double* dl = z + n;                        // delta (nc)
    double* vv = dl + n;                       // v (n)
    double* work = vv + n;                     // likelihood scratch (n)
    const double inv_nc = 1.0 / (double)nc;
    for (int64_t q = (int64_t)blockIdx.x * wpb + warp; q < p.Q; q += (int64_t)gridDim.x * wpb) {
        ChainRng g;
        g.init(p.seed, chain0_ + (uint64_t)q);
        int ncall = 0, nprop = 0;
        uint32_t fl = 0;
        double lcur = 0.0;
        bool done = false;
        while (!done) {
            if (nprop >= B2N_UNIF_MAX_DRAWS) { fl |= 0x80000000u | B2N_WARN_UNIF_INEFFICIENT; break; }
            if (nprop == 10000) fl |= B2N_WARN_UNIF_INEFFICIENT;        // :316-320
            // ---- bound.samples(1): a point uniform in the union of ellipsoids
            int idx = 0;
            for (;;) {
                if (K > 1) {                                             // rand_choice
                    const double xr = rng_uniform(g);
                    int lo = 0;
                    while (lo < K - 1 && p.cum[lo] < xr) lo++;           // searchsorted(left), clamped
                    idx = lo;
                }
                const double ss = rng_normals_to(g, z, nc, lane);
                const double U = rng_uniform(g);
                const double fac = pow(U, inv_nc) / sqrt(ss);
                __syncwarp();
                const double* A = p.axesT + (size_t)idx * nc * nc;
                for (int base = 0; base < nc; base += 64) {
                    double y0, y1;
                    warp_matvec2(A, nc, nc, z, base + lane, nc, y0, y1);
                    const int i0 = base + lane, i1 = i0 + 32;
                    if (i0 < nc) uu[i0] = fma(fac, y0, p.ctrs[(size_t)idx * nc + i0]);
                    if (i1 < nc) uu[i1] = fma(fac, y1, p.ctrs[(size_t)idx * nc + i1]);
                }
                __syncwarp();
                if (K == 1) { if (p.draw_only & 2) ncall = 1; break; }   // bounding.py:543-550
                int qn = 0, qslack = 0;
                for (int k = 0; k < K; k++) {
                    for (int i = lane; i < nc; i += 32) dl[i] = uu[i] - p.ctrs[(size_t)k * nc + i];
                    __syncwarp();
                    const double* AM = p.ams + (size_t)k * nc * nc;
                    double s = 0.0;
                    for (int base = 0; base < nc; base += 64) {
                        double y0, y1;
                        warp_matvec2(AM, nc, nc, dl, base + lane, nc, y0, y1);
                        if (base + lane < nc) s = fma(dl[base + lane], y0, s);
                        if (base + lane + 32 < nc) s = fma(dl[base + lane + 32], y1, s);
                    }
                    s = warp_sum(s);
                    qn += (s < 1.0) ? 1 : 0;
                    qslack += (s <= 1.0 + 1e-3) ? 1 : 0;
                    __syncwarp();
                }
                if (qn == 0) {                                           // :565-579
                    qn = qslack;
                    if (qn == 0) { fl |= 0x40000000u; done = true; break; }
                    fl |= B2N_WARN_Q0_SLACK;
                }
                if (p.draw_only & 2) { ncall = qn; break; }                // sample(return_q=True): no 1/q test
                if (qn == 1) break;
                if (rng_uniform(g) < 1.0 / (double)qn) break;            // :589
            }
            if (done) break;
            nprop++;
            if (p.draw_only) {           // Bound.samples(): no cube test, no likelihood (bounding.py:592-606)
                for (int i = lane; i < n; i += 32) vv[i] = uu[i];
                break;
            }
            // ---- unit-cube check on the clustered dims (internal_samplers.py:314)
            bool ok = true;
            for (int i = lane; i < nc; i += 32) ok = ok && in_cube(uu[i], p.dimflags ? p.dimflags[i] : 0u);
            ok = __all_sync(B2N_FULL, ok);
            if (!ok) continue;
            if (n > nc) {                                                // :325-327
                for (int e = lane; e < n - nc; e += 32) uu[nc + e] = rng_uniform_elem(g, e);
                g.tick++;
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) vv[i] = prior_1d(p.m, i, uu[i]);
            __syncwarp();
            lcur = warp_loglike<LIKE>(p.m, p.m.lmat, vv, work, lane);
            ncall++;
            if (lcur > loglstar_) done = true;
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            peer_put(p.peer, &p.u[q * n + i], uu[i]);
            peer_put(p.peer, &p.v[q * n + i], vv[i]);
        }
        if (lane == 0) {
            peer_put(p.peer, &p.logl[q], lcur);
            peer_put(p.peer, &p.ncall[q], ncall);
            peer_put(p.peer, &p.nprop[q], nprop);
            peer_put(p.peer, &p.flags[q], fl);
        }
        __syncwarp();
    }
    peer_finish(p.peer);
}

// ---- UnitCubeSampler.sample (internal_samplers.py:343-441) for a queue of chains: draw u ~ U(0,1)^n (one
// uniform vector event per draw), v = prior_transform(u), until loglikelihood(v) > loglstar.  This is what the
// reference runs before the first bound exists (sampler.py:407-409, 625-674).  One warp per chain.
struct CubeParams {
    B2nModel m;
    int n;
    double loglstar;
    uint64_t seed, chain0;
    int64_t Q;
    double *u, *v, *logl;
    int* ncall;
    uint32_t* flags;
    PeerSet peer;
    const B2nDyn* dyn;
};

template <int LIKE>
__global__ void __launch_bounds__(128) unitcube_kernel(const CubeParams p) {
    extern __shared__ double sm[];
    const int n = p.n;
    double loglstar_ = p.loglstar;
    unsigned long long chain0_ = p.chain0;
    if (p.dyn) {
        if (p.dyn->skip) return;
        loglstar_ = p.dyn->loglstar; chain0_ = p.dyn->chain0;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double* uu = sm + (size_t)warp * 3 * n;
    double* vv = uu + n;
    double* work = vv + n;
    for (int64_t q = (int64_t)blockIdx.x * wpb + warp; q < p.Q; q += (int64_t)gridDim.x * wpb) {
        ChainRng g;
        g.init(p.seed, chain0_ + (uint64_t)q);
        int ncall = 0;
        uint32_t fl = 0;
        double lcur = 0.0;
        for (;;) {
            if (ncall >= B2N_UNIF_MAX_DRAWS) { fl |= 0x80000000u; break; }
            for (int e = lane; e < n; e += 32) {
                const double t = rng_uniform_elem(g, e);
                uu[e] = t;
                vv[e] = prior_1d(p.m, e, t);
            }
            g.tick++;
            __syncwarp();
            lcur = warp_loglike<LIKE>(p.m, p.m.lmat, vv, work, lane);
            ncall++;
            if (lcur > loglstar_) break;
            __syncwarp();
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            peer_put(p.peer, &p.u[q * n + i], uu[i]);
            peer_put(p.peer, &p.v[q * n + i], vv[i]);
        }
        if (lane == 0) {
            peer_put(p.peer, &p.logl[q], lcur);
            peer_put(p.peer, &p.ncall[q], ncall);
            if (p.flags) peer_put(p.peer, &p.flags[q], fl);
        }
        __syncwarp();
    }
    peer_finish(p.peer);
}

__global__ void unif_error_kernel(const uint32_t* flags, int64_t Q, int* out) {
    int bad = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Q; i += (int64_t)gridDim.x * blockDim.x) {
        if (flags[i] & 0x40000000u) bad |= 1;
        if (flags[i] & 0x80000000u) bad |= 2;
    }
    if (bad) atomicOr(out, bad);
}

extern "C" int b2n_unif_batch(b2n_ctx* ctx, const b2n_chain_args* a, double* u, double* v, double* logl,
                              int32_t* ncall, int32_t* nprop, uint32_t* flags) {
    if (!ctx || !a) return B2N_ERR_ARG;
    if (ctx->start_idx) {       // b2n_set_start_rows is for the next b2n_rwalk_batch only: do not let it linger
        ctx->start_idx = nullptr; ctx->start_nrows = 0;
        return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "start rows by index (b2n_set_start_rows) are read by b2n_rwalk_batch only");
    }
    const bool gather = ctx->peer.total > 0;      // outputs may be NULL in gather mode (b2n_peer_result)
    if (!gather && (!u || !v || !logl || !ncall || !nprop || !flags)) return B2N_ERR_ARG;
    const int draw_only = (a->reserved & B2N_OPT_DRAW_ONLY) ? ((a->reserved & B2N_OPT_DRAW_MIXTURE) ? 3 : 1) : 0;
    B2nModel m;
    memset(&m, 0, sizeof(m));
    m.ndim = a->ndim;
    m.like_kind = B2N_LIKE_EGGBOX;
    if (!draw_only) {
        if (a->model_id < 0 || a->model_id >= (int)ctx->models.size()) return B2N_ERR_ARG;
        m = ctx->models[a->model_id];
    }
    const int n = a->ndim, nc = a->ncdim;
    const int64_t Q = a->nchain;
    if (n != m.ndim || nc < 1 || nc > n || Q < 0 || (draw_only && n != nc)) return B2N_ERR_ARG;
    if (ctx->bK < 1 || ctx->bn != nc || ctx->h_logvols.empty())
        return b2n_fail(ctx, B2N_ERR_ARG, "resident bound (with ctrs/ams/logvols) missing or of wrong dimension");
    if (Q == 0) return gather ? b2n_fail(ctx, B2N_ERR_ARG, "gather mode: every rank must run at least one chain") : B2N_OK;
    B2N_CUDA(ctx, cudaSetDevice(ctx->device));
    ZcScope zc(ctx);          // pinned caller buffers are written in place (host-pointer mode)
    const int K = ctx->bK;
    // probs = exp(logvol_ells - logsumexp(logvol_ells)) ; cumsum (bounding.py:552, 1305)
    std::vector<double> cum(K);
    double hi = -INFINITY;
    for (double x : ctx->h_logvols) hi = std::max(hi, x);
    double se = 0.0;
    for (double x : ctx->h_logvols) se += exp(x - hi);
    const double lse = hi + log(se);
    double run = 0.0;
    for (int k = 0; k < K; k++) { run += exp(ctx->h_logvols[k] - lse); cum[k] = run; }
    const void *dcum, *dfl_in = nullptr;
    B2N_TRY(b2n_in_host(ctx, ctx->work0, cum.data(), cum.size() * sizeof(double), &dcum));
    std::vector<uint32_t> fl;
    if (a->dimflags) {
        fl.assign(a->dimflags, a->dimflags + n);
        B2N_TRY(b2n_in_host(ctx, ctx->in3, fl.data(), fl.size() * sizeof(uint32_t), &dfl_in));
    }
    const bool dyn = ctx->dyn.active;        // device-paced launch (b2n_ns.cu)
    if (dyn && (gather || ctx->ptr_mode != B2N_PTR_DEVICE || draw_only))
        return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "device-paced launch needs device pointers and no gather mode");
    UnifParams p;
    p.dyn = dyn ? ctx->dyn.dev : nullptr;
    p.m = m; p.n = n; p.nc = nc; p.K = K; p.Q = Q; p.draw_only = draw_only;
    p.ctrs = ctx->b_ctrs.as<double>(); p.ams = ctx->b_ams.as<double>(); p.axesT = ctx->b_axesT.as<double>();
    p.cum = (const double*)dcum; p.dimflags = (const uint32_t*)dfl_in;
    p.loglstar = a->loglstar; p.seed = a->seed; p.chain0 = a->chain0;
    void *du, *dv, *dl, *dnc, *dnp, *dfl;
    void* gdev[7];
    bool peer_on = false;
    B2N_TRY(b2n_peer_begin(ctx, n, &p.peer, gdev, &peer_on));
    if (peer_on) {
        if (ctx->peer.row0 + Q > ctx->peer.total) return b2n_fail(ctx, B2N_ERR_ARG, "gather rows out of range (b2n_peer_rows)");
        du = gdev[0]; dv = gdev[1]; dl = gdev[2]; dnc = gdev[3]; dnp = gdev[4]; dfl = gdev[6];
    } else {
        B2N_TRY(b2n_out(ctx, ctx->out0, u, (size_t)Q * n * sizeof(double), &du));
        B2N_TRY(b2n_out(ctx, ctx->out1, v, (size_t)Q * n * sizeof(double), &dv));
        B2N_TRY(b2n_out(ctx, ctx->out2, logl, (size_t)Q * sizeof(double), &dl));
        B2N_TRY(b2n_out(ctx, ctx->out3, ncall, (size_t)Q * sizeof(int), &dnc));
        B2N_TRY(b2n_out(ctx, ctx->out4, nprop, (size_t)Q * sizeof(int), &dnp));
        B2N_TRY(b2n_out(ctx, ctx->out6, flags, (size_t)Q * sizeof(uint32_t), &dfl));
    }
    p.u = (double*)du; p.v = (double*)dv; p.logl = (double*)dl;
    p.ncall = (int*)dnc; p.nprop = (int*)dnp; p.flags = (uint32_t*)dfl;
    const int threads = 128, wpb = threads / 32;
    const size_t smem = (size_t)wpb * 5 * n * sizeof(double);
    if (smem > (size_t)ctx->max_smem_optin) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "ndim too large for the unif kernel");
    int64_t blocks = (Q + wpb - 1) / wpb;
#define CALL(L)                                                                                         \
    if (smem > 48 * 1024)                                                                               \
        B2N_TRY(b2n_func_smem(ctx, (const void*)(unif_kernel<L>), (size_t)(smem))); \
    unif_kernel<L><<<(unsigned)blocks, threads, smem, ctx->stream>>>(p);
    B2N_TIME_BEGIN(ctx);
    B2N_DISPATCH_LIKE(m.like_kind, CALL)
    B2N_TIME_END(ctx);
#undef CALL
    B2N_LAUNCH_CHECK(ctx);
    if (dyn) return B2N_OK;      // device-paced: the commit kernel of the round folds the flags
    int* herr = reinterpret_cast<int*>(ctx->pinned);
    *herr = 0;
    B2N_CUDA(ctx, ctx->out7.ensure(64));
    B2N_CUDA(ctx, cudaMemsetAsync(ctx->out7.p, 0, sizeof(int), ctx->stream));
    const uint32_t* eflags = peer_on ? (const uint32_t*)(ctx->peer.win + ctx->peer.off[6]) : (const uint32_t*)dfl;
    unif_error_kernel<<<64, 256, 0, ctx->stream>>>(eflags, peer_on ? ctx->peer.total : Q, ctx->out7.as<int>());
    B2N_LAUNCH_CHECK(ctx);
    B2N_CUDA(ctx, cudaMemcpyAsync(herr, ctx->out7.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (peer_on) {
        void* const user7[7] = {u, v, logl, ncall, nprop, nullptr, flags};
        B2N_TRY(b2n_peer_end(ctx, n, user7));
        B2N_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->ptr_mode == B2N_PTR_HOST && *ctx->peer.err_host)
            return b2n_fail(ctx, B2N_ERR_PEER, "a peer never arrived at the exchange (timeout in the kernel)");
        if (*herr & 1) return B2N_ERR_Q0;
        if (*herr & 2) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "uniform sampling did not find a point (bound draw limit)");
        return B2N_OK;
    }
    B2N_TRY(b2n_out_done(ctx, u, du, (size_t)Q * n * sizeof(double)));
    B2N_TRY(b2n_out_done(ctx, v, dv, (size_t)Q * n * sizeof(double)));
    B2N_TRY(b2n_out_done(ctx, logl, dl, (size_t)Q * sizeof(double)));
    B2N_TRY(b2n_out_done(ctx, ncall, dnc, (size_t)Q * sizeof(int)));
    B2N_TRY(b2n_out_done(ctx, nprop, dnp, (size_t)Q * sizeof(int)));
    B2N_TRY(b2n_out_done(ctx, flags, dfl, (size_t)Q * sizeof(uint32_t)));
    B2N_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*herr & 1) return B2N_ERR_Q0;
    if (*herr & 2) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "uniform sampling did not find a point (bound draw limit)");
    return B2N_OK;
}


extern "C" int b2n_unitcube_batch(b2n_ctx* ctx, const b2n_chain_args* a, double* u, double* v, double* logl,
                                  int32_t* ncall, uint32_t* flags) {
    if (!ctx || !a) return B2N_ERR_ARG;
    if (ctx->start_idx) {       // b2n_set_start_rows is for the next b2n_rwalk_batch only: do not let it linger
        ctx->start_idx = nullptr; ctx->start_nrows = 0;
        return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "start rows by index (b2n_set_start_rows) are read by b2n_rwalk_batch only");
    }
    const bool gather = ctx->peer.total > 0;
    if (!gather && (!u || !v || !logl || !ncall)) return B2N_ERR_ARG;
    if (a->model_id < 0 || a->model_id >= (int)ctx->models.size()) return B2N_ERR_ARG;
    const B2nModel m = ctx->models[a->model_id];
    const int n = a->ndim;
    const int64_t Q = a->nchain;
    if (n != m.ndim || Q < 0) return B2N_ERR_ARG;
    if (Q == 0) return gather ? b2n_fail(ctx, B2N_ERR_ARG, "gather mode: every rank must run at least one chain") : B2N_OK;
    B2N_CUDA(ctx, cudaSetDevice(ctx->device));
    ZcScope zc(ctx);
    const bool dyn = ctx->dyn.active;
    if (dyn) {
        ctx->dyn.cpc = 1;
        if (ctx->dyn.plan_only) return B2N_OK;
        if (gather || ctx->ptr_mode != B2N_PTR_DEVICE) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "device-paced launch needs device pointers and no gather mode");
    }
    CubeParams p;
    p.dyn = dyn ? ctx->dyn.dev : nullptr;
    p.m = m; p.n = n; p.Q = Q; p.loglstar = a->loglstar; p.seed = a->seed; p.chain0 = a->chain0;
    void *du, *dv, *dl, *dnc, *dfl = nullptr;
    void* gdev[7];
    bool peer_on = false;
    B2N_TRY(b2n_peer_begin(ctx, n, &p.peer, gdev, &peer_on));
    if (peer_on) {
        if (ctx->peer.row0 + Q > ctx->peer.total) return b2n_fail(ctx, B2N_ERR_ARG, "gather rows out of range (b2n_peer_rows)");
        du = gdev[0]; dv = gdev[1]; dl = gdev[2]; dnc = gdev[3]; dfl = gdev[6];
    } else {
        B2N_TRY(b2n_out(ctx, ctx->out0, u, (size_t)Q * n * sizeof(double), &du));
        B2N_TRY(b2n_out(ctx, ctx->out1, v, (size_t)Q * n * sizeof(double), &dv));
        B2N_TRY(b2n_out(ctx, ctx->out2, logl, (size_t)Q * sizeof(double), &dl));
        B2N_TRY(b2n_out(ctx, ctx->out3, ncall, (size_t)Q * sizeof(int), &dnc));
        if (dyn) dfl = flags;
        else { B2N_CUDA(ctx, ctx->out6.ensure((size_t)Q * sizeof(uint32_t))); dfl = ctx->out6.p; }
    }
    p.u = (double*)du; p.v = (double*)dv; p.logl = (double*)dl; p.ncall = (int*)dnc; p.flags = (uint32_t*)dfl;
    const int threads = 128, wpb = threads / 32;
    const size_t smem = (size_t)wpb * 3 * n * sizeof(double);
    if (smem > (size_t)ctx->max_smem_optin) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "ndim too large for the unit-cube kernel");
    const int64_t blocks = (Q + wpb - 1) / wpb;
#define CALL(L)                                                                                             \
    if (smem > 48 * 1024)                                                                                   \
        B2N_TRY(b2n_func_smem(ctx, (const void*)(unitcube_kernel<L>), (size_t)(smem))); \
    unitcube_kernel<L><<<(unsigned)blocks, threads, smem, ctx->stream>>>(p);
    B2N_TIME_BEGIN(ctx);
    B2N_DISPATCH_LIKE(m.like_kind, CALL)
    B2N_TIME_END(ctx);
#undef CALL
    B2N_LAUNCH_CHECK(ctx);
    if (dyn) return B2N_OK;
    int* herr = reinterpret_cast<int*>(ctx->pinned);
    *herr = 0;
    B2N_CUDA(ctx, ctx->out7.ensure(64));
    B2N_CUDA(ctx, cudaMemsetAsync(ctx->out7.p, 0, sizeof(int), ctx->stream));
    const uint32_t* eflags = peer_on ? (const uint32_t*)(ctx->peer.win + ctx->peer.off[6]) : (const uint32_t*)dfl;
    unif_error_kernel<<<64, 256, 0, ctx->stream>>>(eflags, peer_on ? ctx->peer.total : Q, ctx->out7.as<int>());
    B2N_LAUNCH_CHECK(ctx);
    B2N_CUDA(ctx, cudaMemcpyAsync(herr, ctx->out7.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (peer_on) {
        void* const user7[7] = {u, v, logl, ncall, nullptr, nullptr, flags};
        B2N_TRY(b2n_peer_end(ctx, n, user7));
        B2N_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->ptr_mode == B2N_PTR_HOST && *ctx->peer.err_host)
            return b2n_fail(ctx, B2N_ERR_PEER, "a peer never arrived at the exchange (timeout in the kernel)");
    } else {
        B2N_TRY(b2n_out_done(ctx, u, du, (size_t)Q * n * sizeof(double)));
        B2N_TRY(b2n_out_done(ctx, v, dv, (size_t)Q * n * sizeof(double)));
        B2N_TRY(b2n_out_done(ctx, logl, dl, (size_t)Q * sizeof(double)));
        B2N_TRY(b2n_out_done(ctx, ncall, dnc, (size_t)Q * sizeof(int)));
        if (flags && ctx->ptr_mode == B2N_PTR_HOST)
            B2N_CUDA(ctx, cudaMemcpyAsync(flags, dfl, (size_t)Q * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        B2N_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (*herr & 2) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "unit-cube sampling did not find a point above the threshold (draw limit)");
    return B2N_OK;
}
