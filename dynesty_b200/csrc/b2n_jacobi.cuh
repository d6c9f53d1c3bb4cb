// b2n_jacobi.cuh -- one-CTA parallel cyclic Jacobi eigensolver for symmetric matrices held in shared (or
// global) memory: shared by the ellipsoid fit (b2n_bounding.cu) and the RadFriends / SupFriends metric
// (b2n_friends.cu).
#pragma once
#include "b2n_device.cuh"

// ------------------------------------------------------------------ Jacobi eigensolver
// round-robin pairing: m players (m even), round r in [0, m-1), slot k in [0, m/2)
static __device__ __forceinline__ void rr_pair(int m, int r, int k, int& p, int& q) {
    int a, b;
    if (k == 0) { a = m - 1; b = r; }
    else {          // (r + k) mod (m-1), (r - k) mod (m-1) without integer division: r < m-1, k < m/2
        a = r + k;
        if (a >= m - 1) a -= m - 1;
        b = r - k;
        if (b < 0) b += m - 1;
    }
    p = min(a, b);
    q = max(a, b);
}

static __device__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; i++) t += red[i];
    return t;
}

// In-place: A (n x n, ld) -> diagonal ; VT rows = eigenvectors.  Returns sweeps used.
// (force-inlined: called with pointers into the CTA's shared memory the loads / stores become LDS / STS with 32-bit
//  addresses; through a generic pointer a third of the instructions of a round were 64-bit address arithmetic)
static __device__ __forceinline__ int jacobi_eig(double* A, double* VT, int n, int ld, double* cc, double* ss, double* red) {
    const int T = blockDim.x, tid = threadIdx.x;
    const int m = (n + 1) & ~1, half = m >> 1;
    int sweep = 0;
    for (; sweep < 40; sweep++) {
        double off = 0.0, dg = 0.0;
        for (int e = tid; e < n * n; e += T) {
            const int i = e / n, j = e - i * n;
            const double a = A[i * ld + j];
            if (i == j) dg = fma(a, a, dg); else off = fma(a, a, off);
        }
        off = block_sum(off, red);
        dg = block_sum(dg, red);
        const double tot = off + dg;
        if (!(tot < INFINITY) || tot == 0.0) break;      // NaN/Inf or all-zero matrix
        // converged at the round-off floor of the off-diagonal mass (n^2 entries of size
        // ~eps*||A||): the same absolute accuracy LAPACK's eigh delivers
        if (off <= (double)n * (double)n * 2.5e-32 * tot) break;
        // One round = m/2 disjoint rotations.  A HALF-WARP owns a pair: it derives the rotation from its own three
        // matrix entries (uniform in the half-warp, no staging / no barrier), rotates rows p,q of A and of V^T with
        // its 16 lanes across the columns; after one barrier the same half-warp rotates columns p,q of A with its
        // lanes down the rows.  Two barriers per round.  (A full warp per pair spent a third of its ~350
        // instructions per round on the rotation's scalar arithmetic -- a division and two reciprocal square
        // roots, the same in all 32 lanes; two pairs per warp issue that sequence once for both.  The kernel is
        // issue bound: 3.0e6 warp instructions at n = 50.  Element by element the arithmetic is unchanged.)
        const int lane = tid & 31, sub = lane & 15, grp = tid >> 4, ng = T >> 4;
        for (int r = 0; r < m - 1; r++) {
            for (int k0 = (tid >> 5) * 2; k0 < half; k0 += (T >> 5) * 2) {       // warp-uniform trip count
                const int k = k0 + (lane >> 4);
                const bool act = k < half;
                int p = 0, q = 0;
                if (act) rr_pair(m, r, k, p, q);
                double c = 1.0, s = 0.0;
                if (act && q < n) {
                    const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
                    // skip test |apq| <= 1e-17 sqrt(|app aqq|) without a square root
                    if (apq != 0.0 && apq * apq > 1e-34 * fabs(app * aqq)) {
                        // t = tan(theta) = sgn(tau) / (|tau| + sqrt(tau^2 + 1)), tau = (aqq-app)/(2 apq),
                        // rewritten as t = 2 apq / (d + sgn(d) h), h = hypot(d, 2 apq): one division and
                        // two reciprocal square roots instead of three divisions and two square roots
                        // (FP64 div/sqrt are ~350-cycle software sequences and sit on the critical path)
                        const double d = aqq - app, b2 = 2.0 * apq;
                        const double x = fma(d, d, b2 * b2);
                        if (x > 1e-250 && x < 1e250) {
                            const double h = x * rsqrt(x);
                            const double t = b2 / (d + (d >= 0.0 ? h : -h));
                            c = rsqrt(fma(t, t, 1.0));
                            s = t * c;
                        } else {        // out of the safe range of d^2: robust (slow) form
                            const double tau = d / b2;
                            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(fma(tau, tau, 1.0)));
                            c = 1.0 / sqrt(fma(t, t, 1.0));
                            s = t * c;
                        }
                    }
                }
                __syncwarp();                      // every lane has read app/aqq/apq before rows change
                if (act && sub == 0) { cc[k] = c; ss[k] = s; }
                if (s != 0.0) {
                    double* Ap = A + p * ld;
                    double* Aq = A + q * ld;
                    double* Vp = VT + p * ld;
                    double* Vq = VT + q * ld;
                    for (int j = sub; j < n; j += 16) {
                        double a = Ap[j], b = Aq[j];
                        Ap[j] = c * a - s * b;
                        Aq[j] = s * a + c * b;
                        a = Vp[j];
                        b = Vq[j];
                        Vp[j] = c * a - s * b;
                        Vq[j] = s * a + c * b;
                    }
                }
            }
            __syncthreads();
            for (int k = grp; k < half; k += ng) {
                const double s = ss[k];
                if (s != 0.0) {
                    int p, q;
                    rr_pair(m, r, k, p, q);
                    const double c = cc[k];
                    for (int i = sub; i < n; i += 16) {
                        const double a = A[i * ld + p], b = A[i * ld + q];
                        A[i * ld + p] = c * a - s * b;
                        A[i * ld + q] = s * a + c * b;
                    }
                }
            }
            __syncthreads();
        }
    }
    return sweep;
}

