// b2n_slice.cu -- batched slice-sampling chains (one warp per chain).
//
// Replaces RSliceSampler.sample (reference internal_samplers.py:745-855) and
// SliceSampler.sample (:593-709), both built on generic_slice_step (:1075-1206) and the
// Neal (2003) doubling acceptance test _slice_doubling_accept (:1038-1072).
//
// All control flow of a slice step (stepping out, doubling, shrinking) depends only on
// warp-uniform scalars (log-likelihood values, uniforms that every lane derives from the
// same Philox counter), so the 32 lanes of a chain stay converged while they cooperate on
// the vector work: u + x*d, the unit-cube test, the prior transform and the likelihood.
// Same layout as the rwalk kernel (b2n_chain.cuh): persistent-sized grid, one ellipsoid per
// CTA, axes^T / precision matrix staged once in shared memory with 128-byte padded columns,
// all per-chain vectors addressed as b2n_sm[offset].
// NOTE (reference behaviour kept): the slice samplers read kwargs['nonperiodic'], which
// the 3.0 sampler never sets (:654, 804), so every dimension is hard-bounded to (0, 1).
#include "b2n_chain.cuh"
#include <algorithm>
#include <vector>

#define B2N_MAX_EXPAND 4000000      // hard stop against a runaway stepping-out loop

struct SliceParams {
    B2nModel m;
    int n, slices, doubling;
    int ldA, ldP;
    const double* u0;
    const int* order;
    const int3* cta;
    const double* axesT;
    double loglstar, scale;
    uint64_t seed, chain0;
    double *u, *v, *logl;
    int *nexp, *ncon, *ncall;
    uint32_t* flags;
    PeerSet peer;          // fused multi-GPU gather of the outputs (b2n_peer.cu)
    const B2nDyn* dyn;     // device-paced launch (b2n_ns.cu)
};

// F(x) of generic_slice_step (:1112-1123): logl(u + x d) or -inf outside the unit cube.
template <int LIKE, bool PREC_SMEM>
struct SliceEval {
    const B2nModel& m;
    const ModelSm& ms;
    const double* Pg;
    int offP, ldP;
    int ou, odir, oun, ovn, owork;
    int lane, n, pk;
    int nc;
    __device__ __forceinline__ double operator()(double x) {
        bool ok = true;
        for (int i = lane; i < n; i += 32) {
            const double t = fma(x, b2n_sm[odir + i], b2n_sm[ou + i]);
            b2n_sm[oun + i] = t;
            b2n_sm[ovn + i] = prior_sm(pk, ms.op0, ms.op1, i, t);
            ok = ok && (t > 0.0 && t < 1.0);
        }
        nc++;
        ok = __all_sync(B2N_FULL, ok);          // also orders the writes above
        if (!ok) return -INFINITY;
        return loglike_sm<LIKE, PREC_SMEM>(m, ms, Pg, offP, ldP, n, ovn, owork, lane);
    }
};

template <class EVAL>
__device__ bool doubling_accept(EVAL& F, double x1, double loglstar, double L, double R, double fL, double fR) {
    double lhat = L, rhat = R, fl = fL, fr = fR;
    bool D = false;
    while (rhat - lhat > 1.1) {
        const double M = (lhat + rhat) / 2.0;
        if ((0.0 < M && M <= x1) || (x1 < M && M <= 0.0)) D = true;
        if (x1 < M) { rhat = M; fr = F(rhat); }
        else { lhat = M; fl = F(lhat); }
        if (D && loglstar >= fl && loglstar >= fr) return false;
    }
    return true;
}

// one generic_slice_step along b2n_sm[odir..] (already scaled, not yet length-capped).  On
// success the new point is left in b2n_sm[F.oun..] and its logl returned.
template <class EVAL>
__device__ double slice_step(EVAL& F, ChainRng& g, double loglstar, bool doubling, int& n_expand, int& n_contract,
                             bool& expansion_warning, int& err) {
    const int n = F.n, lane = F.lane, odir = F.odir;
    const double rand0 = rng_uniform(g);                        // :1099
    double ss = 0.0;
    for (int i = lane; i < n; i += 32) ss = fma(b2n_sm[odir + i], b2n_sm[odir + i], ss);
    const double dirlen = sqrt(warp_sum(ss));
    const double maxlen = sqrt((double)n) / 2.0;
    if (dirlen > maxlen) {                                      // :1103-1108
        const double dn = dirlen / maxlen;
        for (int i = lane; i < n; i += 32) b2n_sm[odir + i] = b2n_sm[odir + i] / dn;
    }
    __syncwarp();
    double xl = -rand0, xr = 1.0 - rand0;                       // :1126-1127
    double fl = F(xl), fr = F(xr);
    double L = 0, R = 0, fL = 0, fR = 0;
    int nexp = 0;
    expansion_warning = false;
    if (!doubling) {
        while (fl > loglstar) {                                 // :1134-1137
            xl -= 1.0; fl = F(xl); nexp++;
            if (nexp > B2N_MAX_EXPAND) { err = B2N_ERR_SLICE_FAIL; break; }
        }
        while (fr > loglstar && !err) {
            xr += 1.0; fr = F(xr); nexp++;
            if (nexp > B2N_MAX_EXPAND) { err = B2N_ERR_SLICE_FAIL; break; }
        }
        if (nexp > 1000) expansion_warning = true;              // :1142-1145
    } else {
        int K = 1;                                              // :1149-1163
        while (fl > loglstar || fr > loglstar) {
            const double V = rng_uniform(g);
            if (V < 0.5) { xl -= (xr - xl); fl = F(xl); }
            else { xr += (xr - xl); fr = F(xr); }
            nexp += K;
            if (K < (1 << 28)) K *= 2;
        }
        L = xl; R = xr; fL = fl; fR = fr;
    }
    n_expand += nexp;
    double lp = -INFINITY;
    for (int it = 0; !err; it++) {                              // :1168-1203
        const double xp = xl + rng_uniform(g) * (xr - xl);
        lp = F(xp);
        n_contract++;
        if (lp > loglstar && (!doubling || doubling_accept(F, xp, loglstar, L, R, fL, fR))) {
            if (doubling) {   // the acceptance test moved F's scratch point: restore the accepted one
                lp = F(xp);
                F.nc--;
            }
            break;
        }
        if (xp < 0.0) xl = xp;
        else if (xp > 0.0) xr = xp;
        else err = B2N_ERR_SLICE_FAIL;                          // :1191-1203
        if (it > 100000) err = B2N_ERR_SLICE_FAIL;
    }
    return lp;
}

template <int LIKE, bool RANDOM_DIR, bool AX_SMEM, bool PREC_SMEM>
__global__ void __launch_bounds__(512, 1) slice_kernel(const SliceParams p) {
    const int n = p.n;
    const int npad = (n + 1) & ~1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    double loglstar_ = p.loglstar, scale_ = p.scale;
    unsigned long long chain0_ = p.chain0;
    int doubling_ = p.doubling;
    if (p.dyn) {      // device-paced: scalars written by the previous kernel on the stream
        if (p.dyn->skip || (int)blockIdx.x >= p.dyn->ncta) return;
        loglstar_ = p.dyn->loglstar; scale_ = p.dyn->scale; chain0_ = p.dyn->chain0; doubling_ = p.dyn->doubling;
    }
    const int3 cd = p.cta[blockIdx.x];
    int off = 0;
    const double* Ag = p.axesT + (size_t)cd.z * n * n;
    int offA = 0, ldA = n;
    if (AX_SMEM) {
        offA = off; ldA = p.ldA;
        stage_matrix(Ag, offA, n, ldA);
        off += n * ldA;
    }
    const double* Pg = p.m.lmat;
    int offP = 0, ldP = n;
    if (LIKE == B2N_LIKE_GAUSS_PREC && PREC_SMEM) {
        offP = off; ldP = p.ldP;
        stage_matrix(Pg, offP, n, ldP);
        off += n * ldP;
    }
    const ModelSm ms = stage_model(p.m, off, n, npad);
    off += 4 * npad;
    __syncthreads();
    const int ou = off + warp * 6 * npad;
    const int odir = ou + npad, oun = odir + npad, ovn = oun + npad, owork = ovn + npad;
    int* idxs = reinterpret_cast<int*>(&b2n_sm[owork + npad]);   // permutation (n ints)
    const int pk = p.m.prior_kind;

    for (int c = warp; c < cd.y; c += nwarps) {
        const int q = p.order[cd.x + c];
        ChainRng g;
        g.init(p.seed, chain0_ + (uint64_t)q);
        for (int i = lane; i < n; i += 32) b2n_sm[ou + i] = p.u0[(size_t)q * n + i];
        __syncwarp();
        SliceEval<LIKE, PREC_SMEM> F{p.m, ms, Pg, offP, ldP, ou, odir, oun, ovn, owork, lane, n, pk, 0};
        int nexp = 0, ncon = 0, err = 0;
        bool doubling = doubling_ != 0, warned = false;
        double lcur = 0.0;
        for (int sl = 0; sl < p.slices && !err; sl++) {
            const int nsub = RANDOM_DIR ? 1 : n;
            if (!RANDOM_DIR && n > 1) {
                // rstate.shuffle(idxs) (:673-674): argsort (stable) of one uniform vector event
                for (int e = lane; e < n; e += 32) b2n_sm[owork + e] = rng_uniform_elem(g, e);
                g.tick++;
                __syncwarp();
                for (int e = lane; e < n; e += 32) {
                    const double ve = b2n_sm[owork + e];
                    int rk = 0;
                    for (int f = 0; f < n; f++) {
                        const double vf = b2n_sm[owork + f];
                        rk += (vf < ve || (vf == ve && f < e)) ? 1 : 0;
                    }
                    idxs[rk] = e;
                }
                __syncwarp();
            } else if (!RANDOM_DIR) {
                if (lane == 0) idxs[0] = 0;
                __syncwarp();
            }
            for (int sub = 0; sub < nsub && !err; sub++) {
                if (RANDOM_DIR) {
                    // drhat = z / |z| ; direction = axes @ drhat * scale (:820-824)
                    const double ssq = normals_sm(g, owork, n, lane);
                    const double fac = scale_ / sqrt(ssq);
                    __syncwarp();
                    for (int base = 0; base < n; base += 64) {
                        double y0, y1;
                        matvec2o<AX_SMEM>(Ag, offA, ldA, n, owork, base + lane, n, y0, y1);
                        if (base + lane < n) b2n_sm[odir + base + lane] = y0 * fac;
                        if (base + lane + 32 < n) b2n_sm[odir + base + lane + 32] = y1 * fac;
                    }
                } else {
                    // axes = scale * axes.T ; axis = axes[idx] (:665, 680) = column idx of the axes matrix
                    const int idx = idxs[sub];
                    for (int i = lane; i < n; i += 32) b2n_sm[odir + i] = scale_ * mat_ld<AX_SMEM>(Ag, offA + idx * ldA + i);
                }
                __syncwarp();
                bool ew = false;
                const double l = slice_step(F, g, loglstar_, doubling, nexp, ncon, ew, err);
                if (err) break;
                lcur = l;
                for (int i = lane; i < n; i += 32) b2n_sm[ou + i] = b2n_sm[oun + i];     // u = u_prop
                __syncwarp();
                if (ew && !doubling) { doubling = true; warned = true; }   // :689-693, 836-838
            }
        }
        // v_prop = prior_transform(u_prop) (:1204)
        for (int i = lane; i < n; i += 32) {
            const double ui = b2n_sm[ou + i];
            peer_put(p.peer, &p.u[(size_t)q * n + i], ui);
            peer_put(p.peer, &p.v[(size_t)q * n + i], prior_sm(pk, ms.op0, ms.op1, i, ui));
        }
        if (lane == 0) {
            peer_put(p.peer, &p.logl[q], lcur);
            peer_put(p.peer, &p.nexp[q], nexp);
            peer_put(p.peer, &p.ncon[q], ncon);
            peer_put(p.peer, &p.ncall[q], (int)F.nc);
            peer_put(p.peer, &p.flags[q], (warned ? B2N_WARN_DOUBLING : 0u) | (err ? 0x80000000u : 0u));
        }
        __syncwarp();
    }
    peer_finish(p.peer);
}

__global__ void any_error_kernel(const uint32_t* flags, int64_t Q, int* out) {
    int bad = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Q; i += (int64_t)gridDim.x * blockDim.x)
        bad |= (flags[i] & 0x80000000u) ? 1 : 0;
    if (bad) atomicOr(out, 1);
}

template <bool RANDOM_DIR>
static int slice_batch_impl(b2n_ctx* ctx, const b2n_chain_args* a, int32_t slices, int32_t doubling, double* u,
                            double* v, double* logl, int32_t* n_expand, int32_t* n_contract, int32_t* ncall,
                            uint32_t* flags) {
    if (!ctx || !a) return B2N_ERR_ARG;
    if (ctx->start_idx) {       // b2n_set_start_rows is for the next b2n_rwalk_batch only: do not let it linger
        ctx->start_idx = nullptr; ctx->start_nrows = 0;
        return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "start rows by index (b2n_set_start_rows) are read by b2n_rwalk_batch only");
    }
    const bool gather = ctx->peer.total > 0;      // outputs may be NULL in gather mode (b2n_peer_result)
    if (!gather && (!u || !v || !logl || !n_expand || !n_contract || !ncall || !flags)) return B2N_ERR_ARG;
    if (a->model_id < 0 || a->model_id >= (int)ctx->models.size()) return B2N_ERR_ARG;
    const B2nModel m = ctx->models[a->model_id];
    const int n = a->ndim;
    const int64_t Q = a->nchain;
    if (n != m.ndim || a->ncdim != n || slices < 1 || Q < 0 || !a->u0)
        return b2n_fail(ctx, B2N_ERR_ARG, "slice samplers need ncdim == ndim (internal_samplers.py:658, 809)");
    if (ctx->bK < 1 || ctx->bn != n) return b2n_fail(ctx, B2N_ERR_ARG, "resident bound missing or of wrong dimension");
    if (Q == 0) return gather ? b2n_fail(ctx, B2N_ERR_ARG, "gather mode: every rank must run at least one chain") : B2N_OK;
    B2N_CUDA(ctx, cudaSetDevice(ctx->device));
    ZcScope zc(ctx);          // pinned caller buffers are read / written in place (host-pointer mode)
    const int npad = (n + 1) & ~1;
    const size_t per_warp = (size_t)6 * npad * sizeof(double);           // u, d, un, vn, work, idxs
    const size_t model_b = (size_t)4 * npad * sizeof(double);
    const size_t limit = (size_t)ctx->max_smem_optin;
    const int max_warps = (int)std::min<size_t>(16, (limit - model_b) / per_warp);
    if (max_warps < 1) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "ndim too large for the slice kernel");
    int chains_per_cta, warps;
    b2n_chain_grid(ctx, Q, max_warps, chains_per_cta, warps);
    const size_t fixed = per_warp * warps + model_b;
    const int ldA = (n + 15) & ~15, ldP = ldA;
    const size_t ax_b = (size_t)n * ldA * sizeof(double);
    const size_t pr_b = (m.like_kind == B2N_LIKE_GAUSS_PREC) ? ax_b : 0;
    const bool ax_s = fixed + ax_b <= limit;
    const bool pr_s = pr_b > 0 && fixed + (ax_s ? ax_b : 0) + pr_b <= limit;
    const size_t smem = fixed + (ax_s ? ax_b : 0) + (pr_s ? pr_b : 0);
    const bool dyn = ctx->dyn.active;        // device-paced launch (b2n_ns.cu)
    if (dyn) {
        ctx->dyn.cpc = chains_per_cta;
        if (ctx->dyn.plan_only) return B2N_OK;
        if (gather || ctx->ptr_mode != B2N_PTR_DEVICE) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "device-paced launch needs device pointers and no gather mode");
    }
    SliceParams p;
    p.dyn = dyn ? ctx->dyn.dev : nullptr;
    p.m = m; p.n = n; p.slices = slices; p.doubling = doubling; p.ldA = ldA; p.ldP = ldP;
    p.loglstar = a->loglstar; p.scale = a->scale; p.seed = a->seed; p.chain0 = a->chain0;
    p.axesT = ctx->b_axesT.as<double>();
    const void *du0, *dorder, *dcta;
    B2N_TRY(b2n_in(ctx, ctx->in0, a->u0, (size_t)Q * n * sizeof(double), &du0));
    unsigned ncta = 0;
    if (dyn) {
        dorder = ctx->dyn.order; dcta = ctx->dyn.cta;
    } else {
        B2N_TRY(b2n_worklist_dev(ctx, Q, a->ell, ctx->bK, chains_per_cta, &dorder, &dcta, &ncta));
    }
    void *du, *dv, *dl, *dne, *dnc, *dncl, *dfl;
    void* gdev[7];
    bool peer_on = false;
    B2N_TRY(b2n_peer_begin(ctx, n, &p.peer, gdev, &peer_on));
    if (peer_on) {
        if (ctx->peer.row0 + Q > ctx->peer.total) return b2n_fail(ctx, B2N_ERR_ARG, "gather rows out of range (b2n_peer_rows)");
        du = gdev[0]; dv = gdev[1]; dl = gdev[2]; dne = gdev[3]; dnc = gdev[4]; dncl = gdev[5]; dfl = gdev[6];
    } else {
        B2N_TRY(b2n_out(ctx, ctx->out0, u, (size_t)Q * n * sizeof(double), &du));
        B2N_TRY(b2n_out(ctx, ctx->out1, v, (size_t)Q * n * sizeof(double), &dv));
        B2N_TRY(b2n_out(ctx, ctx->out2, logl, (size_t)Q * sizeof(double), &dl));
        B2N_TRY(b2n_out(ctx, ctx->out3, n_expand, (size_t)Q * sizeof(int), &dne));
        B2N_TRY(b2n_out(ctx, ctx->out4, n_contract, (size_t)Q * sizeof(int), &dnc));
        B2N_TRY(b2n_out(ctx, ctx->out5, ncall, (size_t)Q * sizeof(int), &dncl));
        B2N_TRY(b2n_out(ctx, ctx->out6, flags, (size_t)Q * sizeof(uint32_t), &dfl));
    }
    p.u0 = (const double*)du0; p.order = (const int*)dorder; p.cta = (const int3*)dcta;
    p.u = (double*)du; p.v = (double*)dv; p.logl = (double*)dl;
    p.nexp = (int*)dne; p.ncon = (int*)dnc; p.ncall = (int*)dncl; p.flags = (uint32_t*)dfl;
    const unsigned grid = dyn ? (unsigned)ctx->dyn.max_cta : ncta;
#define LAUNCH(L, AXS, PRS)                                                                          \
    do {                                                                                             \
        B2N_TRY(b2n_func_smem(ctx, (const void*)(slice_kernel<L, RANDOM_DIR, AXS, PRS>), (size_t)(smem))); \
        slice_kernel<L, RANDOM_DIR, AXS, PRS><<<grid, warps * 32, smem, ctx->stream>>>(p);            \
    } while (0)
#define CALL(L)                                \
    if (ax_s && pr_s) LAUNCH(L, true, true);   \
    else if (ax_s) LAUNCH(L, true, false);     \
    else if (pr_s) LAUNCH(L, false, true);     \
    else LAUNCH(L, false, false);
    B2N_TIME_BEGIN(ctx);
    B2N_DISPATCH_LIKE(m.like_kind, CALL)
    B2N_TIME_END(ctx);
#undef CALL
#undef LAUNCH
    B2N_LAUNCH_CHECK(ctx);
    if (dyn) return B2N_OK;      // device-paced: the commit kernel of the round folds the flags
    // error summary (a collapsed interval anywhere = RuntimeError in the reference)
    int* derr = reinterpret_cast<int*>(ctx->pinned);
    *derr = 0;
    B2N_CUDA(ctx, ctx->out7.ensure(64));
    B2N_CUDA(ctx, cudaMemsetAsync(ctx->out7.p, 0, sizeof(int), ctx->stream));
    // (gather mode: over the rows of ALL ranks, so that every rank raises the same error)
    const uint32_t* eflags = peer_on ? (const uint32_t*)(ctx->peer.win + ctx->peer.off[6]) : (const uint32_t*)dfl;
    any_error_kernel<<<64, 256, 0, ctx->stream>>>(eflags, peer_on ? ctx->peer.total : Q, ctx->out7.as<int>());
    B2N_LAUNCH_CHECK(ctx);
    B2N_CUDA(ctx, cudaMemcpyAsync(derr, ctx->out7.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (peer_on) {
        void* const user7[7] = {u, v, logl, n_expand, n_contract, ncall, flags};
        B2N_TRY(b2n_peer_end(ctx, n, user7));
        B2N_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->ptr_mode == B2N_PTR_HOST && *ctx->peer.err_host)
            return b2n_fail(ctx, B2N_ERR_PEER, "a peer never arrived at the exchange (timeout in the kernel)");
        if (*derr) return B2N_ERR_SLICE_FAIL;
        return B2N_OK;
    }
    B2N_TRY(b2n_out_done(ctx, u, du, (size_t)Q * n * sizeof(double)));
    B2N_TRY(b2n_out_done(ctx, v, dv, (size_t)Q * n * sizeof(double)));
    B2N_TRY(b2n_out_done(ctx, logl, dl, (size_t)Q * sizeof(double)));
    B2N_TRY(b2n_out_done(ctx, n_expand, dne, (size_t)Q * sizeof(int)));
    B2N_TRY(b2n_out_done(ctx, n_contract, dnc, (size_t)Q * sizeof(int)));
    B2N_TRY(b2n_out_done(ctx, ncall, dncl, (size_t)Q * sizeof(int)));
    B2N_TRY(b2n_out_done(ctx, flags, dfl, (size_t)Q * sizeof(uint32_t)));
    B2N_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the error word is a host result
    if (*derr) return B2N_ERR_SLICE_FAIL;
    return B2N_OK;
}

extern "C" int b2n_rslice_batch(b2n_ctx* ctx, const b2n_chain_args* a, int32_t slices, int32_t doubling,
                                double* u, double* v, double* logl, int32_t* n_expand, int32_t* n_contract,
                                int32_t* ncall, uint32_t* flags) {
    return slice_batch_impl<true>(ctx, a, slices, doubling, u, v, logl, n_expand, n_contract, ncall, flags);
}
extern "C" int b2n_slice_batch(b2n_ctx* ctx, const b2n_chain_args* a, int32_t slices, int32_t doubling, double* u,
                               double* v, double* logl, int32_t* n_expand, int32_t* n_contract, int32_t* ncall,
                               uint32_t* flags) {
    return slice_batch_impl<false>(ctx, a, slices, doubling, u, v, logl, n_expand, n_contract, ncall, flags);
}
