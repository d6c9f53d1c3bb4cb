// b2n_eig_sliced.cu -- symmetric eigen-decomposition for covariances too large for the
// single-CTA kernel (n > ~117: A and V no longer fit in one CTA's shared memory), e.g. the
// 200-D bound of BASELINE config C4.
//
// Same cyclic Jacobi, reorganised so that nothing streams from L2 during the sweeps:
//  * A is kept as a PACKED upper triangle in shared memory (n(n+1)/2 doubles: 161 KB at n=200)
//    and each round of n/2 disjoint rotations is applied in the "2x2 block" form: for every
//    pair of rotation pairs (k1 < k2) one thread rewrites the 2x2 block B <- J1^T B J2, and the
//    diagonal blocks use the closed form a_pp -= t a_pq, a_qq += t a_pq, a_pq = 0.  One phase,
//    no rows-then-columns pass, no strided column access, half the flops.
//  * V^T is split by COLUMNS over S CTAs (grid.y).  Every slice CTA holds its own copy of A,
//    makes the identical (deterministic) decisions and rotates only its columns of V^T, so the
//    S CTAs never communicate; each writes its slice of V^T (and slice 0 the eigenvalues).
//  * The improve_covar_mat ladder (reference bounding.py:1311-1384) is then finished by small
//    per-node kernels: check -> (ok) am = V diag(1/l) V^T, axes = V sqrt(l)  |  (bad) repair
//    the covariance and flag the node for another decomposition (host loop, rare).
#include "b2n_bounding.cuh"
#include <algorithm>
#include <vector>

#define EIG_SLICES 8

__device__ __forceinline__ int pidx(int i, int j, int n) {      // i <= j, packed row-major upper triangle
    return i * n - ((i * (i - 1)) >> 1) + (j - i);
}
__device__ __forceinline__ int sidx(int x, int y, int n) { return x < y ? pidx(x, y, n) : pidx(y, x, n); }

__device__ __forceinline__ void rr_pair2(int m, int r, int k, int& p, int& q) {
    int a, b;
    if (k == 0) { a = m - 1; b = r; }
    else {
        a = r + k;
        if (a >= m - 1) a -= m - 1;
        b = r - k;
        if (b < 0) b += m - 1;
    }
    p = min(a, b);
    q = max(a, b);
}

__device__ double block_sum2(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; i++) t += red[i];
    return t;
}

__global__ void eig_init_kernel(NodeArrays na, const int* __restrict__ nodelist, int pass) {
    const int node = nodelist[blockIdx.x];
    const size_t nn = (size_t)na.n * na.n;
    if (pass == 0) {
        const double* src = na.covraw + (size_t)node * nn;
        double* dst = na.cov + (size_t)node * nn;
        for (size_t e = threadIdx.x; e < nn; e += blockDim.x) dst[e] = src[e];
    }
    if (threadIdx.x == 0) { na.stat[node].trial = 0; na.stat[node].retry = 0; }
}

__global__ void __launch_bounds__(1024) eig_sliced_kernel(NodeArrays na, const int* __restrict__ nodelist, int retry_only,
                                                          double* __restrict__ gVT, double* __restrict__ gLam,
                                                          int* __restrict__ gSweeps) {
    extern __shared__ double sm[];
    const int n = na.n, tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
    const int node = nodelist[blockIdx.x], slice = blockIdx.y, S = gridDim.y;
    if (retry_only && !na.stat[node].retry) return;
    const int m = (n + 1) & ~1, half = m >> 1;
    const int w = (n + S - 1) / S, j0 = slice * w, wj = max(0, min(n, j0 + w) - j0);
    const int np = n * (n + 1) / 2;
    const size_t nn = (size_t)n * n;
    double* A = sm;
    double* VTs = A + ((np + 1) & ~1);                // n x w, row k = eigenvector index
    double* cc = VTs + (size_t)n * w;
    double* ss = cc + half;
    double* tt = ss + half;
    double* red = tt + half;
    int* pp = reinterpret_cast<int*>(red + 32);
    int* qq = pp + half;
    const double* Cm = na.cov + (size_t)node * nn;
    for (int i = warp; i < n; i += nw)
        for (int j = i + lane; j < n; j += 32) A[pidx(i, j, n)] = Cm[(size_t)i * n + j];
    for (int k = warp; k < n; k += nw)
        for (int jj = lane; jj < w; jj += 32) VTs[(size_t)k * w + jj] = (k == j0 + jj) ? 1.0 : 0.0;
    __syncthreads();
    int sweep = 0;
    for (; sweep < 40; sweep++) {
        // off-diagonal mass summed DIRECTLY (as a difference "all - diagonal" it drowns in round-off as soon as it
        // is below eps * total, i.e. at a relative off-norm of 1e-8, and the convergence test fired there: eigenvectors
        // good to 1e-11 for well-conditioned covariances, useless for repaired ones with condition 1e11)
        double offh = 0.0, dg = 0.0;
        for (int i = warp; i < n; i += nw)
            for (int j = i + 1 + lane; j < n; j += 32) { const double a = A[pidx(i, j, n)]; offh = fma(a, a, offh); }
        for (int i = tid; i < n; i += T) { const double a = A[pidx(i, i, n)]; dg = fma(a, a, dg); }
        offh = block_sum2(offh, red);
        dg = block_sum2(dg, red);
        const double off = 2.0 * offh, tot = off + dg;
        if (!(tot < INFINITY) || tot == 0.0) break;
        if (off <= (double)n * (double)n * 2.5e-32 * tot) break;
        for (int r = 0; r < m - 1; r++) {
            // phase 1: the n/2 rotations of this round
            for (int k = tid; k < half; k += T) {
                int p, q;
                rr_pair2(m, r, k, p, q);
                double c = 1.0, s = 0.0, t = 0.0;
                if (q < n) {
                    const double app = A[pidx(p, p, n)], aqq = A[pidx(q, q, n)], apq = A[pidx(p, q, n)];
                    if (apq != 0.0 && apq * apq > 1e-34 * fabs(app * aqq)) {
                        const double d = aqq - app, b2 = 2.0 * apq;
                        const double x = fma(d, d, b2 * b2);
                        if (x > 1e-250 && x < 1e250) {
                            const double h = x * rsqrt(x);
                            t = b2 / (d + (d >= 0.0 ? h : -h));
                        } else {
                            const double tau = d / b2;
                            t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(fma(tau, tau, 1.0)));
                        }
                        c = rsqrt(fma(t, t, 1.0));
                        s = t * c;
                    }
                }
                cc[k] = c; ss[k] = s; tt[k] = t; pp[k] = p; qq[k] = q;
            }
            __syncthreads();
            // phase 2a: off-diagonal 2x2 blocks  B <- J1^T B J2  (k1 < k2)
            for (int k1 = warp; k1 < half; k1 += nw) {
                const int p1 = pp[k1], q1 = qq[k1];
                const double c1 = cc[k1], s1 = ss[k1];
                const bool hq1 = q1 < n;
                for (int k2 = k1 + 1 + lane; k2 < half; k2 += 32) {
                    const double c2 = cc[k2], s2 = ss[k2];
                    if (s1 == 0.0 && s2 == 0.0) continue;
                    const int p2 = pp[k2], q2 = qq[k2];
                    const bool hq2 = q2 < n;
                    const int i00 = sidx(p1, p2, n);
                    const int i01 = hq2 ? sidx(p1, q2, n) : i00;
                    const int i10 = hq1 ? sidx(q1, p2, n) : i00;
                    const int i11 = (hq1 && hq2) ? sidx(q1, q2, n) : i00;
                    const double b00 = A[i00], b01 = hq2 ? A[i01] : 0.0, b10 = hq1 ? A[i10] : 0.0,
                                 b11 = (hq1 && hq2) ? A[i11] : 0.0;
                    const double t00 = c1 * b00 - s1 * b10, t01 = c1 * b01 - s1 * b11;
                    const double t10 = s1 * b00 + c1 * b10, t11 = s1 * b01 + c1 * b11;
                    A[i00] = c2 * t00 - s2 * t01;
                    if (hq2) A[i01] = s2 * t00 + c2 * t01;
                    if (hq1) A[i10] = c2 * t10 - s2 * t11;
                    if (hq1 && hq2) A[i11] = s2 * t10 + c2 * t11;
                }
            }
            // phase 2b: diagonal blocks (closed form)
            for (int k = tid; k < half; k += T) {
                if (ss[k] != 0.0) {
                    const int p = pp[k], q = qq[k];
                    const int ipp = pidx(p, p, n), iqq = pidx(q, q, n), ipq = pidx(p, q, n);
                    const double apq = A[ipq], t = tt[k];
                    A[ipp] -= t * apq;
                    A[iqq] += t * apq;
                    A[ipq] = 0.0;
                }
            }
            // phase 2c: rows p,q of this slice of V^T
            for (int k = warp; k < half; k += nw) {
                const double s = ss[k];
                if (s != 0.0) {
                    const double c = cc[k];
                    double* Vp = VTs + (size_t)pp[k] * w;
                    double* Vq = VTs + (size_t)qq[k] * w;
                    for (int jj = lane; jj < wj; jj += 32) {
                        const double a = Vp[jj], b = Vq[jj];
                        Vp[jj] = c * a - s * b;
                        Vq[jj] = s * a + c * b;
                    }
                }
            }
            __syncthreads();
        }
    }
    double* VT = gVT + (size_t)blockIdx.x * nn;
    for (int k = warp; k < n; k += nw)
        for (int jj = lane; jj < wj; jj += 32) VT[(size_t)k * n + j0 + jj] = VTs[(size_t)k * w + jj];
    if (slice == 0) {
        for (int k = tid; k < n; k += T) gLam[(size_t)blockIdx.x * n + k] = A[pidx(k, k, n)];
        if (tid == 0) gSweeps[blockIdx.x] = sweep;
    }
}

// per node: ladder decision; on success the sort ranks + reciprocal eigenvalues, else the
// repaired covariance recipe.  mode[b]: 0 ok, 1 clamp-and-rebuild, 2 blended (done here), 3 fallback
__global__ void __launch_bounds__(1024) eig_check_kernel(NodeArrays na, const int* __restrict__ nodelist, int pass,
                                                         int retry_only, const double* __restrict__ gLam,
                                                         const int* __restrict__ gSweeps, double* __restrict__ gScale,
                                                         int* __restrict__ gRank, int* __restrict__ gMode) {
    __shared__ int s_failed;
    __shared__ double s_mx;
    const int n = na.n, tid = threadIdx.x, T = blockDim.x, b = blockIdx.x;
    const int node = nodelist[b];
    NodeStat* st = na.stat + node;
    if (retry_only && !st->retry) { if (tid == 0) gMode[b] = -1; return; }
    const double* lam = gLam + (size_t)b * n;
    const size_t nn = (size_t)n * n;
    const int trial = st->trial;        // read by every thread BEFORE the barrier: thread 0 rewrites it below
    if (tid == 0) {
        bool fin = true;
        double mx = -INFINITY, mn = INFINITY;
        for (int k = 0; k < n; k++) {
            const double l = lam[k];
            fin = fin && (l == l) && (fabs(l) < INFINITY);
            mx = fmax(mx, l);
            mn = fmin(mn, l);
        }
        int f = 0;
        if (!fin) f = 2;
        else if (mx <= 0) f = 2;
        else if (mn < mx / 1e12) f = 1;
        s_failed = f;
        s_mx = mx;
    }
    __syncthreads();
    const int failed = s_failed;
    double* Cm = na.cov + (size_t)node * nn;
    if (failed == 0) {
        for (int k = tid; k < n; k += T) {
            const double l = lam[k];
            int rk = 0;
            for (int j = 0; j < n; j++) rk += (lam[j] < l || (lam[j] == l && j < k)) ? 1 : 0;
            gRank[(size_t)b * n + k] = rk;
            gScale[(size_t)b * n + k] = 1.0 / l;
            na.lam[(size_t)node * n + rk] = l;
        }
        if (tid == 0) {
            gMode[b] = 0;
            if (pass == 0) st->good = (trial == 0) ? 1 : 0;
            st->fallback = 0; st->retry = 0; st->sweeps = gSweeps[b];
        }
    } else if (trial + 1 >= 100) {           // identity fallback (:1373-1378)
        for (size_t e = tid; e < nn; e += T) {
            const double v = (e / n == e % n) ? 1.0 : 0.0;
            Cm[e] = v; na.am[(size_t)node * nn + e] = v; na.axes[(size_t)node * nn + e] = v;
        }
        for (int k = tid; k < n; k += T) na.lam[(size_t)node * n + k] = 1.0;
        if (tid == 0) { gMode[b] = 3; st->good = 0; st->fallback = 1; st->retry = 0; st->sweeps = gSweeps[b]; }
    } else if (failed == 1) {                // clamp the small eigenvalues, rebuild (:1363-1366)
        const double floorv = 10.0 * s_mx / 1e12;
        for (int k = tid; k < n; k += T) gScale[(size_t)b * n + k] = fmax(lam[k], floorv);
        if (tid == 0) { gMode[b] = 1; st->trial = trial + 1; st->retry = 1; if (pass == 0) st->good = 0; }
    } else {                                 // blend with the identity (:1367-1371)
        const double coeff = 1e-10 * pow(1e10, (double)trial / 99.0);
        for (size_t e = tid; e < nn; e += T) Cm[e] = (1.0 - coeff) * Cm[e] + ((e / n == e % n) ? coeff : 0.0);
        if (tid == 0) { gMode[b] = 2; st->trial = trial + 1; st->retry = 1; if (pass == 0) st->good = 0; }
    }
}

// OUT = sum_k scale_k v_k v_k^T on 64x64 tiles (mode 0 -> am, mode 1 -> cov), V^T from global.
__global__ void __launch_bounds__(256) vdv_kernel(NodeArrays na, const int* __restrict__ nodelist,
                                                  const double* __restrict__ gVT, const double* __restrict__ gScale,
                                                  const int* __restrict__ gMode, int ntile) {
    __shared__ double As[16][65];
    __shared__ double Bs[16][65];
    const int b = blockIdx.x, mode = gMode[b];
    if (mode != 0 && mode != 1) return;
    const int n = na.n, node = nodelist[b];
    const size_t nn = (size_t)n * n;
    const double* VT = gVT + (size_t)b * nn;
    const double* sc = gScale + (size_t)b * n;
    int ib = 0, jt = blockIdx.y;
    while (jt >= ntile - ib) { jt -= ntile - ib; ib++; }
    const int jb = ib + jt;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < 4; y++) acc[x][y] = 0.0;
    for (int k0 = 0; k0 < n; k0 += 16) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int idx = threadIdx.x + e * 256;
            const int kk = idx >> 6, c = idx & 63, k = k0 + kk;
            double a = 0.0, bb = 0.0;
            if (k < n) {
                const int ca = ib * 64 + c, cb = jb * 64 + c;
                if (ca < n) a = VT[(size_t)k * n + ca] * sc[k];
                if (cb < n) bb = VT[(size_t)k * n + cb];
            }
            As[kk][c] = a;
            Bs[kk][c] = bb;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            double a[4], bq[4];
#pragma unroll
            for (int t = 0; t < 4; t++) { a[t] = As[kk][ty * 4 + t]; bq[t] = Bs[kk][tx * 4 + t]; }
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) acc[x][y] = fma(a[x], bq[y], acc[x][y]);
        }
        __syncthreads();
    }
    double* out = (mode == 0 ? na.am : na.cov) + (size_t)node * nn;
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const int i = ib * 64 + ty * 4 + x, j = jb * 64 + tx * 4 + y;
            if (i < n && j < n) {
                out[(size_t)i * n + j] = acc[x][y];
                if (ib != jb) out[(size_t)j * n + i] = acc[x][y];
            }
        }
}

__global__ void axes_kernel(NodeArrays na, const int* __restrict__ nodelist, const double* __restrict__ gVT,
                            const double* __restrict__ gLam, const int* __restrict__ gRank,
                            const int* __restrict__ gMode) {
    const int b = blockIdx.x;
    if (gMode[b] != 0) return;
    const int n = na.n, node = nodelist[b];
    const size_t nn = (size_t)n * n;
    const double* VT = gVT + (size_t)b * nn;
    double* AX = na.axes + (size_t)node * nn;
    for (size_t e = (size_t)blockIdx.y * blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.y * blockDim.x) {
        const int k = (int)(e / n), i = (int)(e - (size_t)k * n);      // eigenpair k, component i (coalesced read)
        AX[(size_t)i * n + gRank[(size_t)b * n + k]] = VT[e] * sqrt(gLam[(size_t)b * n + k]);
    }
}

// Returns B2N_OK with *used = 1 when the sliced path handled the batch, *used = 0 when the matrix
// does not fit (caller falls back to the single-CTA L2 path).
int b2n_eig_sliced(BoundWork& w, const int* dlist, int pn, int pass, int retry_only, int* used) {
    b2n_ctx* ctx = w.ctx;
    const int n = w.n, half = ((n + 1) & ~1) / 2;
    const int wslice = (n + EIG_SLICES - 1) / EIG_SLICES;
    const size_t np = (size_t)n * (n + 1) / 2;
    const size_t smem = (((np + 1) & ~(size_t)1) + (size_t)n * wslice + 3 * half + 32) * sizeof(double) + 2 * half * sizeof(int);
    *used = 0;
    if (smem > (size_t)ctx->max_smem_optin) return B2N_OK;
    *used = 1;
    const size_t nn = (size_t)n * n;
    // workspace: V^T, eigenvalues, scale vector, ranks, modes, sweeps per batch entry
    const size_t bytes = (size_t)pn * (nn + 2 * n) * sizeof(double) + (size_t)pn * (n + 2) * sizeof(int) + 64;
    B2N_CUDA(ctx, ctx->scratch2.ensure(bytes));
    double* gVT = ctx->scratch2.as<double>();
    double* gLam = gVT + (size_t)pn * nn;
    double* gScale = gLam + (size_t)pn * n;
    int* gRank = reinterpret_cast<int*>(gScale + (size_t)pn * n);
    int* gMode = gRank + (size_t)pn * n;
    int* gSweeps = gMode + pn;
    cudaStream_t st = ctx->stream;
    if (!retry_only) {
        eig_init_kernel<<<pn, 256, 0, st>>>(w.na, dlist, pass);
        B2N_LAUNCH_CHECK(ctx);
    }
    B2N_TRY(b2n_func_smem(ctx, (const void*)(eig_sliced_kernel), (size_t)(smem)));
    const int threads = 32 * std::max(8, std::min(32, half));
    eig_sliced_kernel<<<dim3(pn, EIG_SLICES), threads, smem, st>>>(w.na, dlist, retry_only, gVT, gLam, gSweeps);
    B2N_LAUNCH_CHECK(ctx);
    eig_check_kernel<<<pn, 1024, 0, st>>>(w.na, dlist, pass, retry_only, gLam, gSweeps, gScale, gRank, gMode);
    B2N_LAUNCH_CHECK(ctx);
    const int ntile = (n + 63) / 64;
    vdv_kernel<<<dim3(pn, ntile * (ntile + 1) / 2), 256, 0, st>>>(w.na, dlist, gVT, gScale, gMode, ntile);
    B2N_LAUNCH_CHECK(ctx);
    axes_kernel<<<dim3(pn, (unsigned)std::min<size_t>((nn + 255) / 256, 64)), 256, 0, st>>>(w.na, dlist, gVT, gLam, gRank, gMode);
    B2N_LAUNCH_CHECK(ctx);
    return B2N_OK;
}
