// b2n_membership.cu -- batched Mahalanobis membership of M points in K ellipsoids.
//
// Replaces MultiEllipsoid.within / overlap / contains (reference bounding.py:502-523:
//   delt = x - ctrs ; mask = einsum('ai,aij,aj->a', delt, ams, delt) < 1   -- strict)
// and Ellipsoid.distance_many / contains (:286-305: sqrt(d2) <= 1.0 -- non-strict).
//
// One warp per point, looping over the K ellipsoids; delta lives in a warp-private
// shared vector, each A_k streams from L2 (K*n*n*8 bytes total, shared by all warps)
// with lanes reading consecutive rows of a column (A is symmetric).  The summation
// order is fixed (sequential over columns inside a lane, xor-butterfly across
// lanes) so the result is bitwise reproducible run to run.
#include "b2n_device.cuh"

__global__ void __launch_bounds__(256) membership_kernel(const double* __restrict__ x, int64_t M, int n,
                                                         const double* __restrict__ ctrs,
                                                         const double* __restrict__ ams, int K, int strict,
                                                         uint8_t* __restrict__ mask, int* __restrict__ q,
                                                         double* __restrict__ d2) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double* xs = sm + (size_t)warp * 2 * n;
    double* dl = xs + n;
    for (int64_t p = (int64_t)blockIdx.x * wpb + warp; p < M; p += (int64_t)gridDim.x * wpb) {
        for (int i = lane; i < n; i += 32) xs[i] = x[p * n + i];
        int cnt = 0;
        for (int k = 0; k < K; k++) {
            __syncwarp();
            for (int i = lane; i < n; i += 32) dl[i] = xs[i] - ctrs[(size_t)k * n + i];
            __syncwarp();
            const double* A = ams + (size_t)k * n * n;
            double s = 0.0;
            for (int base = 0; base < n; base += 64) {
                double y0, y1;
                warp_matvec2(A, n, n, dl, base + lane, n, y0, y1);
                if (base + lane < n) s = fma(dl[base + lane], y0, s);
                if (base + lane + 32 < n) s = fma(dl[base + lane + 32], y1, s);
            }
            s = warp_sum(s);
            const bool in = strict ? (s < 1.0) : (s <= 1.0);
            cnt += in ? 1 : 0;
            if (lane == 0) {
                if (mask) mask[p * K + k] = in ? 1 : 0;
                if (d2) d2[p * K + k] = s;
            }
        }
        if (lane == 0 && q) q[p] = cnt;
        __syncwarp();
    }
}

// device-pointer entry used internally (bounding update, uniform sampler checks)
int b2n_membership_dev(b2n_ctx* ctx, const double* x, int64_t M, int n, const double* ctrs,
                       const double* ams, int K, int strict, uint8_t* mask, int* q, double* d2) {
    if (M == 0) return B2N_OK;
    const int threads = 256, wpb = threads / 32;
    const size_t smem = (size_t)wpb * 2 * n * sizeof(double);
    if (smem > (size_t)ctx->max_smem_optin) return b2n_fail(ctx, B2N_ERR_UNSUPPORTED, "ndim too large");
    int64_t blocks = (M + wpb - 1) / wpb;
    if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
    if (smem > 48 * 1024)
        B2N_TRY(b2n_func_smem(ctx, (const void*)(membership_kernel), (size_t)(smem)));
    membership_kernel<<<(unsigned)blocks, threads, smem, ctx->stream>>>(x, M, n, ctrs, ams, K, strict, mask, q, d2);
    B2N_LAUNCH_CHECK(ctx);
    return B2N_OK;
}

extern "C" int b2n_membership(b2n_ctx* ctx, const double* x, int64_t M, int32_t n, const double* ctrs,
                              const double* ams, int32_t K, int32_t strict, uint8_t* mask, int32_t* q,
                              double* d2) {
    if (!ctx || !x || !ctrs || !ams || n < 1 || K < 1 || M < 0) return B2N_ERR_ARG;
    if (M == 0) return B2N_OK;
    B2N_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *dx, *dc, *da;
    void *dm, *dq, *dd;
    B2N_TRY(b2n_in(ctx, ctx->in0, x, (size_t)M * n * sizeof(double), &dx));
    B2N_TRY(b2n_in(ctx, ctx->in1, ctrs, (size_t)K * n * sizeof(double), &dc));
    B2N_TRY(b2n_in(ctx, ctx->in2, ams, (size_t)K * n * n * sizeof(double), &da));
    B2N_TRY(b2n_out(ctx, ctx->out0, mask, (size_t)M * K, &dm));
    B2N_TRY(b2n_out(ctx, ctx->out1, q, (size_t)M * sizeof(int), &dq));
    B2N_TRY(b2n_out(ctx, ctx->out2, d2, (size_t)M * K * sizeof(double), &dd));
    B2N_TRY(b2n_membership_dev(ctx, (const double*)dx, M, n, (const double*)dc, (const double*)da, K, strict,
                               (uint8_t*)dm, (int*)dq, (double*)dd));
    B2N_TRY(b2n_out_done(ctx, mask, dm, (size_t)M * K));
    B2N_TRY(b2n_out_done(ctx, q, dq, (size_t)M * sizeof(int)));
    B2N_TRY(b2n_out_done(ctx, d2, dd, (size_t)M * K * sizeof(double)));
    return b2n_finish(ctx);
}
