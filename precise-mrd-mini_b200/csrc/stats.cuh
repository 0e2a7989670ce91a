// stats.cuh -- fp64 Poisson / binomial point masses and tails (kernel K8 device functions).
//
// Replaces scipy.stats.poisson.cdf/sf and scipy.stats.binomtest as called from
// src/precise_mrd/call.py:14-38 and src/precise_mrd/stats.py:47-133.  SciPy evaluates the
// tails through regularised incomplete gamma/beta functions (Cephes/Boost); here every
// count is an integer, so a tail is a sum of point masses: the boundary mass is computed
// with Loader's saddle-point form (C. Loader, "Fast and accurate computation of binomial
// probabilities", 2000 -- the dpois_raw/dbinom_raw scheme) and the remaining masses by the
// exact ratio recurrence, summed AWAY from the mode so that all terms are positive and
// decreasing.  Relative error ~1e-15 x sqrt(#terms); parity gate is 1e-12.
//
// Everything is __host__ __device__ so that the same source can be exercised on the host
// by the development harness; the product only ever calls it from kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

#define PMRD_LN_SQRT_2PI 0.91893853320467274178
#define PMRD_LN_2PI 1.8378770664093454836
#define PMRD_2PI 6.283185307179586476925

__host__ __device__ static inline double pmrd_stirlerr(double n) {
    // log(n!) - log(sqrt(2 pi n) (n/e)^n), integer n >= 0
    const double tab[16] = {0.0,
                            0.08106146679532725822,
                            0.041340695955409294094,
                            0.027677925684998339149,
                            0.020790672103765093112,
                            0.016644691189821192163,
                            0.013876128823070747999,
                            0.011896709945891770095,
                            0.010411265261972096497,
                            0.0092554621827127329177,
                            0.0083305634333628712565,
                            0.007573675487951840795,
                            0.0069428401072095298657,
                            0.0064089941880042070684,
                            0.0059513701127588477356,
                            0.005554733551962801371};
    if (n <= 15.0) return tab[(int)n];
    const double S0 = 1.0 / 12.0, S1 = 1.0 / 360.0, S2 = 1.0 / 1260.0, S3 = 1.0 / 1680.0, S4 = 1.0 / 1188.0;
    const double nn = n * n;
    if (n > 500.0) return (S0 - S1 / nn) / n;
    if (n > 80.0) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35.0) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

// log(1+z) - z, accurate for small |z|
__host__ __device__ static inline double pmrd_log1pmx(double z) {
    if (fabs(z) > 0.35) return log1p(z) - z;
    // -z^2/2 + z^3/3 - z^4/4 ...
    double term = -z;  // (-z)^1 sign carrier: t_n = -(-z)^n / n
    double sum = 0.0;
    double mz = -z, pw = mz;  // pw = (-z)^n
    for (int n = 2; n < 400; ++n) {
        pw *= mz;
        term = -pw / (double)n;
        double s1 = sum + term;
        if (s1 == sum) break;
        sum = s1;
    }
    return sum;
}

// bd0(x, np) = x log(x/np) + np - x   (x > 0, np > 0)
__host__ __device__ static inline double pmrd_bd0(double x, double np) {
    const double mu = (np - x) / x;
    if (fabs(mu) <= 0.35) return -x * pmrd_log1pmx(mu);
    return x * log(x / np) + np - x;
}

// Poisson point mass P(X = x), integer x >= 0
__host__ __device__ static inline double pmrd_dpois(double x, double lambda) {
    if (lambda <= 0.0) return x == 0.0 ? 1.0 : 0.0;
    if (x < 0.0) return 0.0;
    if (x == 0.0) return exp(-lambda);
    return exp(-pmrd_stirlerr(x) - pmrd_bd0(x, lambda)) / sqrt(PMRD_2PI * x);
}

// P(X <= k) by downward summation from k.  Intended for k <= lambda-ish (small tail) but
// correct for any k (just slower / ~1 when k >> lambda).
__host__ __device__ static inline double pmrd_pois_lower(double k, double lambda, double pk) {
    // pk = dpois(k)
    double sum = pk, term = pk;
    for (double j = k; j > 0.0; j -= 1.0) {
        term *= j / lambda;  // pmf(j-1) = pmf(j) * j / lambda
        double s1 = sum + term;
        if (s1 == sum && j < lambda) break;
        sum = s1;
    }
    return sum;
}
// P(X >= k) by upward summation from k
__host__ __device__ static inline double pmrd_pois_upper(double k, double lambda, double pk) {
    double sum = pk, term = pk;
    for (double j = k + 1.0;; j += 1.0) {
        term *= lambda / j;  // pmf(j) = pmf(j-1) * lambda / j
        double s1 = sum + term;
        if (s1 == sum && j > lambda) break;
        sum = s1;
    }
    return sum;
}

// cdf(k) and sf(k-1) = P(X >= k) together
__host__ __device__ static inline void pmrd_pois_tails(double k, double lambda, double* cdf_k, double* sf_km1) {
    const double pk = pmrd_dpois(k, lambda);
    if (k < lambda) {
        const double lo = pmrd_pois_lower(k, lambda, pk);
        *cdf_k = lo;
        double hi = 1.0 - lo + pk;
        *sf_km1 = hi > 1.0 ? 1.0 : hi;
    } else {
        const double hi = pmrd_pois_upper(k, lambda, pk);
        *sf_km1 = hi;
        double lo = 1.0 - hi + pk;
        *cdf_k = lo > 1.0 ? 1.0 : lo;
    }
}

// call.py:14-25 poisson_test
__host__ __device__ static inline double pmrd_poisson_two_sided(int64_t observed, double expected) {
    if (!(expected > 0.0)) return 1.0;  // expected <= 0 (NaN also lands here)
    if (observed < 0) {                 // cdf(<0) = 0
        return 0.0;
    }
    double lo, hi;
    pmrd_pois_tails((double)observed, expected, &lo, &hi);
    double p = 2.0 * (lo < hi ? lo : hi);
    return p < 1.0 ? p : 1.0;
}
// stats.py:65-74: sf(k-1) = P(X >= k)
__host__ __device__ static inline double pmrd_poisson_greater(int64_t observed, double expected) {
    if (observed <= 0) return 1.0;
    if (!(expected > 0.0)) return 0.0;
    double lo, hi;
    pmrd_pois_tails((double)observed, expected, &lo, &hi);
    return hi;
}

// stats.py:61-63 alternative == "less": cdf(k)
__host__ __device__ static inline double pmrd_poisson_less(int64_t observed, double expected) {
    if (observed < 0) return 0.0;
    if (!(expected > 0.0)) return 1.0;
    double lo, hi;
    pmrd_pois_tails((double)observed, expected, &lo, &hi);
    return lo;
}

// Binomial point mass P(X = x | n, p), integers 0 <= x <= n, 0 < p < 1
__host__ __device__ static inline double pmrd_dbinom(double x, double n, double p) {
    const double q = 1.0 - p;
    if (x < 0.0 || x > n) return 0.0;
    if (x == 0.0) return exp(n * log1p(-p));
    if (x == n) return exp(n * log(p));
    const double lc = pmrd_stirlerr(n) - pmrd_stirlerr(x) - pmrd_stirlerr(n - x) - pmrd_bd0(x, n * p) - pmrd_bd0(n - x, n * q);
    const double lf = PMRD_LN_2PI + log(x) + log1p(-x / n);
    return exp(lc - 0.5 * lf);
}
// P(X <= k) downward from k
__host__ __device__ static inline double pmrd_binom_lower(double k, double n, double p) {
    if (k < 0.0) return 0.0;
    if (k >= n) return 1.0;
    const double pk = pmrd_dbinom(k, n, p);
    const double r = (1.0 - p) / p;
    const double mode = (n + 1.0) * p;
    double sum = pk, term = pk;
    for (double j = k; j > 0.0; j -= 1.0) {
        term *= (j / (n - j + 1.0)) * r;  // pmf(j-1) = pmf(j) * j/(n-j+1) * q/p
        double s1 = sum + term;
        if (s1 == sum && j < mode) break;
        sum = s1;
    }
    return sum < 1.0 ? sum : 1.0;
}
// P(X >= k) upward from k
__host__ __device__ static inline double pmrd_binom_upper(double k, double n, double p) {
    if (k <= 0.0) return 1.0;
    if (k > n) return 0.0;
    const double pk = pmrd_dbinom(k, n, p);
    const double r = p / (1.0 - p);
    const double mode = (n + 1.0) * p;
    double sum = pk, term = pk;
    for (double j = k; j < n; j += 1.0) {
        term *= ((n - j) / (j + 1.0)) * r;  // pmf(j+1) = pmf(j) * (n-j)/(j+1) * p/q
        double s1 = sum + term;
        if (s1 == sum && j > mode) break;
        sum = s1;
    }
    return sum < 1.0 ? sum : 1.0;
}
// cdf(k): use whichever side of the mode makes the direct sum the small one
__host__ __device__ static inline double pmrd_binom_cdf(double k, double n, double p) {
    if (k < 0.0) return 0.0;
    if (k >= n) return 1.0;
    if (k < n * p) return pmrd_binom_lower(k, n, p);
    double v = 1.0 - pmrd_binom_upper(k + 1.0, n, p);
    return v < 0.0 ? 0.0 : v;
}
// sf(k) = P(X > k)
__host__ __device__ static inline double pmrd_binom_sf(double k, double n, double p) {
    if (k < 0.0) return 1.0;
    if (k >= n) return 0.0;
    if (k + 1.0 > n * p) return pmrd_binom_upper(k + 1.0, n, p);
    double v = 1.0 - pmrd_binom_lower(k, n, p);
    return v < 0.0 ? 0.0 : v;
}

// scipy.stats._binomtest._binary_search_for_binom_tst (scipy 1.14+ form): returns i in
// [lo, hi] with a(i) <= d < a(i+1) for ascending a; sign = -1 searches a = -pmf.
__host__ __device__ static inline double pmrd_binom_search(double sign, double d, double lo, double hi, double n, double p) {
    while (lo < hi) {
        const double mid = lo + floor((hi - lo) * 0.5);
        const double midval = sign * pmrd_dbinom(mid, n, p);
        if (midval < d) lo = mid + 1.0;
        else if (midval > d) hi = mid - 1.0;
        else { lo = mid; hi = mid; }
    }
    return (sign * pmrd_dbinom(lo, n, p) <= d) ? lo : lo - 1.0;
}

// call.py:28-38 binomial_test -> scipy binomtest(k, n, p, 'two-sided') ("minlike")
__host__ __device__ static inline double pmrd_binomial_two_sided(int64_t k_, int64_t n_, double p) {
    if (n_ == 0) return 1.0;
    if (p <= 0.0) return k_ > 0 ? 0.0 : 1.0;
    if (p >= 1.0) return k_ < n_ ? 0.0 : 1.0;
    const double k = (double)k_, n = (double)n_;
    if (k < 0.0 || k > n) return NAN;  // scipy: invalid -> nan
    const double d = pmrd_dbinom(k, n, p);
    const double rerr = 1.0 + 1e-7;
    double pval;
    if (k < p * n) {
        const double ix = pmrd_binom_search(-1.0, -d * rerr, ceil(p * n), n, n, p);
        const double y = n - ix + ((d * rerr == pmrd_dbinom(ix, n, p)) ? 1.0 : 0.0);
        pval = pmrd_binom_cdf(k, n, p) + pmrd_binom_sf(n - y, n, p);
    } else {
        const double ix = pmrd_binom_search(1.0, d * rerr, 0.0, floor(p * n), n, p);
        const double y = ix + 1.0;
        pval = pmrd_binom_cdf(y - 1.0, n, p) + pmrd_binom_sf(k - 1.0, n, p);
    }
    return pval < 1.0 ? pval : 1.0;
}
// stats.py:109-114 binomtest(alternative='less') = cdf(k)
__host__ __device__ static inline double pmrd_binomial_less(int64_t k_, int64_t n_, double p) {
    if (n_ == 0) return 1.0;
    if (p <= 0.0) return 1.0;
    if (p >= 1.0) return k_ < n_ ? 0.0 : 1.0;
    return pmrd_binom_cdf((double)k_, (double)n_, p);
}
// stats.py:109-114 binomtest(alternative='greater') = sf(k-1)
__host__ __device__ static inline double pmrd_binomial_greater(int64_t k_, int64_t n_, double p) {
    if (n_ == 0) return 1.0;
    if (p <= 0.0) return k_ > 0 ? 0.0 : 1.0;
    if (p >= 1.0) return 1.0;
    return pmrd_binom_sf((double)k_ - 1.0, (double)n_, p);
}

// ----------------------------------------------------------------------------------------
// Hypergeometric point mass and tails; Fisher's exact test (filters.py:50-100 -> scipy fisher_exact)
// ----------------------------------------------------------------------------------------
// P(X = x) for x white balls in n draws from r white + b black (R's dhyper: three binomial
// point masses at p = n / (r + b), each through the saddle-point pmrd_dbinom)
__host__ __device__ static inline double pmrd_dhyper(double x, double r, double b, double n) {
    if (x < 0.0 || x > r || n - x > b || x > n) return 0.0;
    if (n == 0.0) return x == 0.0 ? 1.0 : 0.0;
    if (n == r + b) return x == r ? 1.0 : 0.0;
    const double p = n / (r + b);
    return pmrd_dbinom(x, r, p) * pmrd_dbinom(n - x, b, p) / pmrd_dbinom(n, r + b, p);
}
// P(X <= x): point masses summed downward from x (ratio recurrence), or 1 - upward sum past the mode
__host__ __device__ static inline double pmrd_hyper_lower(double x, double r, double b, double n) {
    const double lo = fmax(0.0, n - b);
    if (x < lo) return 0.0;
    double term = pmrd_dhyper(x, r, b, n), sum = term;
    for (double k = x; k > lo; k -= 1.0) {
        term *= (k * (b - n + k)) / ((r - k + 1.0) * (n - k + 1.0));   // pmf(k-1) / pmf(k)
        const double s1 = sum + term;
        if (s1 == sum) break;
        sum = s1;
    }
    return sum;
}
__host__ __device__ static inline double pmrd_hyper_upper(double x, double r, double b, double n) {   // P(X >= x)
    const double hi = fmin(r, n);
    if (x > hi) return 0.0;
    double term = pmrd_dhyper(x, r, b, n), sum = term;
    for (double k = x; k < hi; k += 1.0) {
        term *= ((r - k) * (n - k)) / ((k + 1.0) * (b - n + k + 1.0));   // pmf(k+1) / pmf(k)
        const double s1 = sum + term;
        if (s1 == sum) break;
        sum = s1;
    }
    return sum;
}
__host__ __device__ static inline double pmrd_hyper_cdf(double x, double r, double b, double n, double mode) {
    if (x >= fmin(r, n)) return 1.0;
    if (x < mode) return pmrd_hyper_lower(x, r, b, n);
    const double v = 1.0 - pmrd_hyper_upper(x + 1.0, r, b, n);
    return v < 0.0 ? 0.0 : v;
}
__host__ __device__ static inline double pmrd_hyper_sf(double x, double r, double b, double n, double mode) {   // P(X > x)
    if (x < fmax(0.0, n - b)) return 1.0;
    if (x + 1.0 > mode) return pmrd_hyper_upper(x + 1.0, r, b, n);
    const double v = 1.0 - pmrd_hyper_lower(x, r, b, n);
    return v < 0.0 ? 0.0 : v;
}
// scipy's _binary_search_for_binom_tst on sign * pmf: i in [lo, hi] with a(i) <= d < a(i+1)
__host__ __device__ static inline double pmrd_hyper_search(double sign, double d, double lo, double hi, double r, double b, double n) {
    while (lo < hi) {
        const double mid = lo + floor((hi - lo) * 0.5);
        const double v = sign * pmrd_dhyper(mid, r, b, n);
        if (v < d) lo = mid + 1.0;
        else if (v > d) hi = mid - 1.0;
        else { lo = mid; hi = mid; }
    }
    return (sign * pmrd_dhyper(lo, r, b, n) <= d) ? lo : lo - 1.0;
}
// scipy.stats.fisher_exact(table, 'two-sided') on [[c00, c01], [c10, c11]]: sample odds ratio and
// the sum of the tables no more likely than the observed one (relative slack 1e-14)
__host__ __device__ static inline double pmrd_fisher_two_sided(int64_t c00_, int64_t c01_, int64_t c10_, int64_t c11_, double* odds) {
    const double c00 = (double)c00_, c01 = (double)c01_, c10 = (double)c10_, c11 = (double)c11_;
    if (c00 + c10 == 0.0 || c01 + c11 == 0.0 || c00 + c01 == 0.0 || c10 + c11 == 0.0) {
        *odds = NAN;
        return 1.0;
    }
    *odds = (c10 > 0.0 && c01 > 0.0) ? (c00 * c11) / (c10 * c01) : INFINITY;
    const double n1 = c00 + c01, n2 = c10 + c11, n = c00 + c10;
    const double mode = floor((n + 1.0) * (n1 + 1.0) / (n1 + n2 + 2.0));
    const double pexact = pmrd_dhyper(c00, n1, n2, n), pmode = pmrd_dhyper(mode, n1, n2, n);
    const double gamma = 1.0 + 1e-14;
    if (fabs(pexact - pmode) / fmax(pexact, pmode) <= 1e-14) return 1.0;
    double pv;
    if (c00 < mode) {
        const double plower = pmrd_hyper_cdf(c00, n1, n2, n, mode);
        if (pmrd_dhyper(n, n1, n2, n) > pexact * gamma) return plower < 1.0 ? plower : 1.0;
        const double g = pmrd_hyper_search(-1.0, -pexact * gamma, mode, n, n1, n2, n);
        pv = plower + pmrd_hyper_sf(g, n1, n2, n, mode);
    } else {
        const double pupper = pmrd_hyper_sf(c00 - 1.0, n1, n2, n, mode);
        if (pmrd_dhyper(0.0, n1, n2, n) > pexact * gamma) return pupper < 1.0 ? pupper : 1.0;
        const double g = pmrd_hyper_search(1.0, pexact * gamma, 0.0, mode, n1, n2, n);
        pv = pupper + pmrd_hyper_cdf(g, n1, n2, n, mode);
    }
    return pv < 1.0 ? pv : 1.0;
}

