// Development harness ONLY (never loaded by the product or the tests): compiles the
// __host__ __device__ statistics in csrc/stats.cuh with g++ so their arithmetic can be
// compared against SciPy on a box without a GPU.  g++ -O2 -shared -fPIC -ffp-contract=off
#include "../precise-mrd-mini_b200/csrc/stats.cuh"
extern "C" {
void dev_poisson2(const int64_t* k, const double* e, int64_t m, double* p) { for (int64_t i = 0; i < m; ++i) p[i] = pmrd_poisson_two_sided(k[i], e[i]); }
void dev_poisson_gt(const int64_t* k, const double* e, int64_t m, double* p) { for (int64_t i = 0; i < m; ++i) p[i] = pmrd_poisson_greater(k[i], e[i]); }
void dev_binom2(const int64_t* k, const int64_t* n, const double* pr, int64_t m, double* p) { for (int64_t i = 0; i < m; ++i) p[i] = pmrd_binomial_two_sided(k[i], n[i], pr[i]); }
void dev_binom_gt(const int64_t* k, const int64_t* n, const double* pr, int64_t m, double* p) { for (int64_t i = 0; i < m; ++i) p[i] = pmrd_binomial_greater(k[i], n[i], pr[i]); }
void dev_dbinom(const int64_t* k, const int64_t* n, const double* pr, int64_t m, double* p) { for (int64_t i = 0; i < m; ++i) p[i] = pmrd_dbinom((double)k[i], (double)n[i], pr[i]); }
}
extern "C" void dev_fisher(const int64_t* t, int64_t m, double* odds, double* p) {
    for (int64_t i = 0; i < m; ++i) p[i] = pmrd_fisher_two_sided(t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3], &odds[i]);
}
