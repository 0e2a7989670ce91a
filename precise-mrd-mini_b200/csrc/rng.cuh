// rng.cuh -- distribution transforms on top of Philox4x32-10 (common.cuh).
// All samplers are inversion samplers of ONE uniform: a draw is a pure function of
// (seed, stream, cell, family, read) and never depends on how many words another draw used --
// the property the reference's shared PCG64 stream lacks.
#pragma once
#include "common.cuh"
#include "stats.cuh"

// Poisson(lambda) by sequential inversion (lambda small: family sizes, lambda <= ~30)
__device__ __forceinline__ int sample_poisson_small(double u, double lambda) {
    double p = exp(-lambda), F = p;
    int k = 0;
    while (u > F && k < 4096) {
        ++k;
        p *= lambda / (double)k;
        F += p;
    }
    return k;
}

// NegativeBinomial(r = 2, p): pmf(k) = (k+1) p^2 (1-p)^k   (io.py:71-77 uses r = 2.0)
__device__ __forceinline__ int sample_negbin_r2(double u, double p) {
    const double q = 1.0 - p;
    double f = p * p, F = f;
    int k = 0;
    while (u > F && k < 65535) {
        ++k;
        f *= q * (double)(k + 1) / (double)k;
        F += f;
    }
    return k;
}

// Binomial(n, p): exact inversion, chopping outward from the mode so that the expected
// number of steps is O(sqrt(n p q)) and nothing underflows for n p >> 745.
__device__ __forceinline__ int64_t sample_binomial(double u, int64_t n_, double p) {
    if (n_ <= 0 || p <= 0.0) return 0;
    if (p >= 1.0) return n_;
    const double n = (double)n_;
    const double q = 1.0 - p;
    double m = floor((n + 1.0) * p);
    if (m > n) m = n;
    const double fm = pmrd_dbinom(m, n, p);
    u -= fm;
    if (u <= 0.0) return (int64_t)m;
    double lo = m, hi = m, flo = fm, fhi = fm;
    const double r_up = p / q, r_dn = q / p;
    for (;;) {
        bool moved = false;
        if (hi < n) {
            fhi *= ((n - hi) / (hi + 1.0)) * r_up;
            hi += 1.0;
            u -= fhi;
            if (u <= 0.0) return (int64_t)hi;
            moved = true;
        }
        if (lo > 0.0) {
            flo *= (lo / (n - lo + 1.0)) * r_dn;
            lo -= 1.0;
            u -= flo;
            if (u <= 0.0) return (int64_t)lo;
            moved = true;
        }
        if (!moved || (fhi < 1e-300 && flo < 1e-300)) return (int64_t)m;  // rounding leftovers
    }
}

// standard normal from two uniforms in (0,1) (Box-Muller, fp64)
__device__ __forceinline__ double sample_normal(double u1, double u2) {
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}
// fp32 pair
__device__ __forceinline__ float2 sample_normal2f(float u1, float u2) {
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    return make_float2(r * c, r * s);
}
