#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <random>
#include "b2n_fastmath.cuh"
static double ulp_err(double got, long double ref) {
    if (ref == 0) return fabs(got) > 0 ? 1e9 : 0;
    double r = (double)ref;
    double u = nextafter(fabs(r), INFINITY) - fabs(r);
    return (double)(fabsl((long double)got - ref) / u);
}
int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 4000000;
    std::mt19937_64 g(12345);
    double e_log = 0, e_sqrt = 0, e_sin = 0, e_cos = 0, e_div = 0, e_rad = 0;
    const long double PI2 = 6.283185307179586476925286766559005768L;
    for (long i = 0; i < n; i++) {
        uint64_t a = g(), b = g();
        // the B2N uniforms: ((a>>6)*2^26 + (b>>6) + 0.5) * 2^-52 built from two 32-bit words
        uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
        double u0 = ((double)(a0 >> 6) * 67108864.0 + (double)(a1 >> 6) + 0.5) * 0x1p-52;
        double u1 = ((double)(b0 >> 6) * 67108864.0 + (double)(b1 >> 6) + 0.5) * 0x1p-52;
        if (i % 7 == 0) u0 = ldexp(u0, -(int)(a % 50));         // small arguments too
        if (i % 11 == 0) u0 = 1.0 - ldexp(u0, -(int)(b % 50));  // arguments next to 1
        if (u0 <= 0 || u0 >= 1) continue;
        double lg = b2n_log(u0);
        e_log = fmax(e_log, ulp_err(lg, logl((long double)u0)));
        double arg = -2.0 * lg;
        double sq = b2n_sqrt(arg);
        e_sqrt = fmax(e_sqrt, ulp_err(sq, sqrtl((long double)arg)));
        e_rad = fmax(e_rad, ulp_err(sq, sqrtl(-2.0L * logl((long double)u0))));
        double s, c;
        b2n_sincos2pi(u1, &s, &c);
        long double th = PI2 * (long double)u1;
        e_sin = fmax(e_sin, fabs(sinl(th)) > 1e-3L ? ulp_err(s, sinl(th)) : (double)(fabsl(s - sinl(th)) / 2.2e-19L / 4096));
        e_cos = fmax(e_cos, fabs(cosl(th)) > 1e-3L ? ulp_err(c, cosl(th)) : (double)(fabsl(c - cosl(th)) / 2.2e-19L / 4096));
        double num = u0 - 0.5, den = 2.0 + u1;
        e_div = fmax(e_div, ulp_err(b2n_div(num, den), (long double)num / (long double)den));
    }
    printf("{\"n\": %ld, \"log_ulp\": %.3f, \"sqrt_ulp\": %.3f, \"rad_ulp\": %.3f, \"sin_ulp\": %.3f, \"cos_ulp\": %.3f, \"div_ulp\": %.3f}\n",
           n, e_log, e_sqrt, e_rad, e_sin, e_cos, e_div);
    return 0;
}
