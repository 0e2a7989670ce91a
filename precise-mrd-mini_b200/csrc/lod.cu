// lod.cu -- K11: detection-curve fit and its bootstrap CI on the device.
//
// Replaces LODAnalyzer._fit_detection_curve / _bootstrap_lod_ci
// (src/precise_mrd/eval/lod.py:351-408): scipy.optimize.curve_fit(logistic, log10(af), hit,
// maxfev=2000) == MINPACK lmdif (forward-difference Jacobian, ftol = xtol = 1.49012e-8,
// gtol = 0, factor = 100, automatic column scaling) from p0 = (1, 1); LoD = 10^((logit(t)-b)/a).
// curve_fit raises unless lmdif ends with info in 1..4 and the reference then falls back to
// scipy.interpolate.interp1d(hit, af, fill_value='extrapolate')(t) -- which is NaN for the
// all-zero hit vectors the reference's own grids produce.  To get
// the same convergence / fallback decisions this is a statement of the published lmdif
// algorithm (More, Garbow, Hillstrom, "User Guide for MINPACK-1", ANL-80-74: lmdif, lmpar,
// qrfac, qrsolv, fdjac2, enorm) specialised to n = 2 parameters, one thread per fit.
// The 1 + n_boot fits (point estimate + Philox-resampled pairs) run in parallel.
#include "kernels.cuh"
#include "rng.cuh"
#include <stdlib.h>

#define LOD_MAX_M 64
#define LM_N 2

struct LmProblem {
    int m;
    double x[LOD_MAX_M];  // log10(af)
    double y[LOD_MAX_M];  // hit rate
};

__host__ __device__ static double lm_enorm(int n, const double* x) {
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
    const double agiant = rgiant / (double)n;
    for (int i = 0; i < n; ++i) {
        const double xabs = fabs(x[i]);
        if (xabs > rdwarf && xabs < agiant) {
            s2 += xabs * xabs;
        } else if (xabs <= rdwarf) {
            if (xabs > x3max) {
                const double t = x3max / xabs;
                s3 = 1.0 + s3 * (t * t);
                x3max = xabs;
            } else if (xabs != 0.0) {
                const double t = xabs / x3max;
                s3 += t * t;
            }
        } else {
            if (xabs > x1max) {
                const double t = x1max / xabs;
                s1 = 1.0 + s1 * (t * t);
                x1max = xabs;
            } else {
                const double t = xabs / x1max;
                s1 += t * t;
            }
        }
    }
    if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.0) {
        if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

// residuals: logistic(x; a, b) - y     (lod.py:360-361 through curve_fit's wrapper)
__host__ __device__ static void lm_fcn(const LmProblem& P, const double* p, double* f) {
#if defined(LM_COUNT_FEV) && !defined(__CUDA_ARCH__)
    ++g_fev;   // host development harness only
#endif
    // (the residuals are independent: unrolled so that the fp64 exp / divide chains of several points overlap --
    // a refit is ONE thread, bound by dependent-instruction latency)
#pragma unroll 5
    for (int i = 0; i < P.m; ++i) f[i] = 1.0 / (1.0 + exp(-(p[0] * P.x[i] + p[1]))) - P.y[i];
}

// column-major fjac[j*m + i]
__host__ __device__ static void lm_qrfac(int m, double* a, int* ipvt, double* rdiag, double* acnorm, double* wa) {
    const double epsmch = 2.220446049250313e-16;
    for (int j = 0; j < LM_N; ++j) {
        acnorm[j] = lm_enorm(m, a + j * m);
        rdiag[j] = acnorm[j];
        wa[j] = rdiag[j];
        ipvt[j] = j;
    }
    for (int j = 0; j < LM_N; ++j) {
        int kmax = j;
        for (int k = j; k < LM_N; ++k)
            if (rdiag[k] > rdiag[kmax]) kmax = k;
        if (kmax != j) {
            for (int i = 0; i < m; ++i) {
                const double t = a[j * m + i];
                a[j * m + i] = a[kmax * m + i];
                a[kmax * m + i] = t;
            }
            rdiag[kmax] = rdiag[j];
            wa[kmax] = wa[j];
            const int k = ipvt[j];
            ipvt[j] = ipvt[kmax];
            ipvt[kmax] = k;
        }
        double ajnorm = lm_enorm(m - j, a + j * m + j);
        if (ajnorm != 0.0) {
            if (a[j * m + j] < 0.0) ajnorm = -ajnorm;
            for (int i = j; i < m; ++i) a[j * m + i] /= ajnorm;
            a[j * m + j] += 1.0;
            for (int k = j + 1; k < LM_N; ++k) {
                double sum = 0.0;
                for (int i = j; i < m; ++i) sum += a[j * m + i] * a[k * m + i];
                const double temp = sum / a[j * m + j];
                for (int i = j; i < m; ++i) a[k * m + i] -= temp * a[j * m + i];
                if (rdiag[k] != 0.0) {
                    double t = a[k * m + j] / rdiag[k];
                    rdiag[k] *= sqrt(fmax(0.0, 1.0 - t * t));
                    t = rdiag[k] / wa[k];
                    if (0.05 * (t * t) <= epsmch) {
                        rdiag[k] = lm_enorm(m - j - 1, a + k * m + j + 1);
                        wa[k] = rdiag[k];
                    }
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
}

// r: n x n, r[j*ldr + i] column-major with ldr = m (the top of fjac)
__host__ __device__ static void lm_qrsolv(double* r, int ldr, const int* ipvt, const double* diag, const double* qtb, double* x,
                                 double* sdiag, double* wa) {
#if defined(LM_COUNT_QRSOLV) && !defined(__CUDA_ARCH__)
    ++g_qrsolv;   // host development harness only
#endif
    for (int j = 0; j < LM_N; ++j) {
        for (int i = j; i < LM_N; ++i) r[j * ldr + i] = r[i * ldr + j];
        x[j] = r[j * ldr + j];
        wa[j] = qtb[j];
    }
    for (int j = 0; j < LM_N; ++j) {
        const int l = ipvt[j];
        if (diag[l] != 0.0) {
            for (int k = j; k < LM_N; ++k) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
            for (int k = j; k < LM_N; ++k) {
                if (sdiag[k] == 0.0) continue;
                double sn, cs;
                if (fabs(r[k * ldr + k]) < fabs(sdiag[k])) {
                    const double cotan = r[k * ldr + k] / sdiag[k];
                    sn = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                    cs = sn * cotan;
                } else {
                    const double tn = sdiag[k] / r[k * ldr + k];
                    cs = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
                    sn = cs * tn;
                }
                r[k * ldr + k] = cs * r[k * ldr + k] + sn * sdiag[k];
                const double temp = cs * wa[k] + sn * qtbpj;
                qtbpj = -sn * wa[k] + cs * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < LM_N; ++i) {
                    const double t2 = cs * r[k * ldr + i] + sn * sdiag[i];
                    sdiag[i] = -sn * r[k * ldr + i] + cs * sdiag[i];
                    r[k * ldr + i] = t2;
                }
            }
        }
        sdiag[j] = r[j * ldr + j];
        r[j * ldr + j] = x[j];
    }
    int nsing = LM_N;
    for (int j = 0; j < LM_N; ++j) {
        if (sdiag[j] == 0.0 && nsing == LM_N) nsing = j;
        if (nsing < LM_N) wa[j] = 0.0;
    }
    for (int k = 1; k <= nsing; ++k) {
        const int j = nsing - k;
        double sum = 0.0;
        for (int i = j + 1; i < nsing; ++i) sum += r[j * ldr + i] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < LM_N; ++j) x[ipvt[j]] = wa[j];
}

__host__ __device__ static void lm_lmpar(double* r, int ldr, const int* ipvt, const double* diag, const double* qtb, double delta,
                                double* par, double* x, double* sdiag, double* wa1, double* wa2) {
    const double dwarf = 2.2250738585072014e-308;
    int nsing = LM_N;
    for (int j = 0; j < LM_N; ++j) {
        wa1[j] = qtb[j];
        if (r[j * ldr + j] == 0.0 && nsing == LM_N) nsing = j;
        if (nsing < LM_N) wa1[j] = 0.0;
    }
    for (int k = 1; k <= nsing; ++k) {
        const int j = nsing - k;
        wa1[j] /= r[j * ldr + j];
        const double temp = wa1[j];
        for (int i = 0; i < j; ++i) wa1[i] -= r[j * ldr + i] * temp;
    }
    for (int j = 0; j < LM_N; ++j) x[ipvt[j]] = wa1[j];
    int iter = 0;
    for (int j = 0; j < LM_N; ++j) wa2[j] = diag[j] * x[j];
    double dxnorm = lm_enorm(LM_N, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) {
        *par = 0.0;
        return;
    }
    double parl = 0.0;
    if (nsing >= LM_N) {
        for (int j = 0; j < LM_N; ++j) {
            const int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        for (int j = 0; j < LM_N; ++j) {
            double sum = 0.0;
            for (int i = 0; i < j; ++i) sum += r[j * ldr + i] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j * ldr + j];
        }
        const double temp = lm_enorm(LM_N, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
    for (int j = 0; j < LM_N; ++j) {
        double sum = 0.0;
        for (int i = 0; i <= j; ++i) sum += r[j * ldr + i] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    const double gnorm = lm_enorm(LM_N, wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = dwarf / fmin(delta, 0.1);
    *par = fmax(*par, parl);
    *par = fmin(*par, paru);
    if (*par == 0.0) *par = gnorm / dxnorm;
    for (;;) {
        ++iter;
        if (*par == 0.0) *par = fmax(dwarf, 0.001 * paru);
        double temp = sqrt(*par);
        for (int j = 0; j < LM_N; ++j) wa1[j] = temp * diag[j];
        lm_qrsolv(r, ldr, ipvt, wa1, qtb, x, sdiag, wa2);
        for (int j = 0; j < LM_N; ++j) wa2[j] = diag[j] * x[j];
        dxnorm = lm_enorm(LM_N, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        for (int j = 0; j < LM_N; ++j) {
            const int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        for (int j = 0; j < LM_N; ++j) {
            wa1[j] /= sdiag[j];
            const double t = wa1[j];
            for (int i = j + 1; i < LM_N; ++i) wa1[i] -= r[j * ldr + i] * t;
        }
        temp = lm_enorm(LM_N, wa1);
        const double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = fmax(parl, *par);
        if (fp < 0.0) paru = fmin(paru, *par);
        *par = fmax(parl, *par + parc);
    }
}

// returns MINPACK info; p holds the final parameters
__host__ __device__ static int lm_lmdif(const LmProblem& P, double* x, int maxfev) {
    const double epsmch = 2.220446049250313e-16;
    const double ftol = 1.49012e-8, xtol = 1.49012e-8, gtol = 0.0, factor = 100.0;
    const int m = P.m;
    double fvec[LOD_MAX_M], wa4[LOD_MAX_M], fjac[LM_N * LOD_MAX_M];
    double diag[LM_N], qtf[LM_N], wa1[LM_N], wa2[LM_N], wa3[LM_N];
    int ipvt[LM_N];
    int info = 0, nfev = 0;
    lm_fcn(P, x, fvec);
    nfev = 1;
    double fnorm = lm_enorm(m, fvec);
    double par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;
    int iter = 1;
    for (;;) {
        // fdjac2: forward differences, epsfcn = machine epsilon
        {
            const double eps = sqrt(epsmch);
            for (int j = 0; j < LM_N; ++j) {
                const double temp = x[j];
                double h = eps * fabs(temp);
                if (h == 0.0) h = eps;
                x[j] = temp + h;
                lm_fcn(P, x, wa4);
                x[j] = temp;
                for (int i = 0; i < m; ++i) fjac[j * m + i] = (wa4[i] - fvec[i]) / h;
            }
            nfev += LM_N;
        }
        lm_qrfac(m, fjac, ipvt, wa1, wa2, wa3);
        if (iter == 1) {
            for (int j = 0; j < LM_N; ++j) {
                diag[j] = wa2[j];
                if (wa2[j] == 0.0) diag[j] = 1.0;
            }
            for (int j = 0; j < LM_N; ++j) wa3[j] = diag[j] * x[j];
            xnorm = lm_enorm(LM_N, wa3);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        for (int i = 0; i < m; ++i) wa4[i] = fvec[i];
        for (int j = 0; j < LM_N; ++j) {
            if (fjac[j * m + j] != 0.0) {
                double sum = 0.0;
                for (int i = j; i < m; ++i) sum += fjac[j * m + i] * wa4[i];
                const double temp = -sum / fjac[j * m + j];
                for (int i = j; i < m; ++i) wa4[i] += fjac[j * m + i] * temp;
            }
            fjac[j * m + j] = wa1[j];
            qtf[j] = wa4[j];
        }
        gnorm = 0.0;
        if (fnorm != 0.0) {
            for (int j = 0; j < LM_N; ++j) {
                const int l = ipvt[j];
                if (wa2[l] != 0.0) {
                    double sum = 0.0;
                    for (int i = 0; i <= j; ++i) sum += fjac[j * m + i] * (qtf[i] / fnorm);
                    gnorm = fmax(gnorm, fabs(sum / wa2[l]));
                }
            }
        }
        if (gnorm <= gtol) info = 4;
        if (info != 0) break;
        for (int j = 0; j < LM_N; ++j) diag[j] = fmax(diag[j], wa2[j]);
        double ratio;
        do {
            lm_lmpar(fjac, m, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3, wa4);
            for (int j = 0; j < LM_N; ++j) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            const double pnorm = lm_enorm(LM_N, wa3);
            if (iter == 1) delta = fmin(delta, pnorm);
            lm_fcn(P, wa2, wa4);
            ++nfev;
            const double fnorm1 = lm_enorm(m, wa4);
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) {
                const double t = fnorm1 / fnorm;
                actred = 1.0 - t * t;
            }
            for (int j = 0; j < LM_N; ++j) {
                wa3[j] = 0.0;
                const double temp = wa1[ipvt[j]];
                for (int i = 0; i <= j; ++i) wa3[i] += fjac[j * m + i] * temp;
            }
            const double temp1 = lm_enorm(LM_N, wa3) / fnorm;
            const double temp2 = (sqrt(par) * pnorm) / fnorm;
            const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
            const double dirder = -(temp1 * temp1 + temp2 * temp2);
            ratio = 0.0;
            if (prered != 0.0) ratio = actred / prered;
            if (ratio <= 0.25) {
                double temp;
                if (actred >= 0.0) temp = 0.5;
                else temp = 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * fmin(delta, pnorm / 0.1);
                par /= temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par *= 0.5;
            }
            if (ratio >= 1e-4) {
                for (int j = 0; j < LM_N; ++j) {
                    x[j] = wa2[j];
                    wa2[j] = diag[j] * x[j];
                }
                for (int i = 0; i < m; ++i) fvec[i] = wa4[i];
                xnorm = lm_enorm(LM_N, wa2);
                fnorm = fnorm1;
                ++iter;
            }
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) info = 1;
            if (delta <= xtol * xnorm) info = 2;
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0 && info == 2) info = 3;
            if (info != 0) break;
            if (nfev >= maxfev) info = 5;
            if (fabs(actred) <= epsmch && prered <= epsmch && 0.5 * ratio <= 1.0) info = 6;
            if (delta <= epsmch * xnorm) info = 7;
            if (gnorm <= epsmch) info = 8;
            if (info != 0) break;
        } while (ratio < 1e-4);
        if (info != 0) break;
    }
    return info;
}

// ---------------------------------------------------------------- one fit per WARP ------------------
// A refit is a strictly sequential iteration and the kernel lasts as long as its slowest refit -- the resamples
// whose parameters run away until maxfev = 2000 (about 5 % of them: all-zero or perfectly separated hit rates, 667
// iterations of three evaluations; the others stop after ~30 evaluations).  One thread per fit spends that time in
// ~150 dependent fp64 exp / divide chains per iteration (measured: 45 ms for five curves x 201 refits).  Here a fit
// owns a warp and lane i owns curve point i: residuals, Jacobian columns and Householder updates are one operation
// per lane, every sum over the points is taken by ALL lanes in the index order of the loops above, and the 2 x 2
// algebra (lmpar, qrsolv) is written out in registers, redundantly on every lane.  The arithmetic -- and with it
// every convergence / fallback decision -- is that of lm_lmdif: the two kernels return the same bits.
// What a runaway iteration costs was found with ncu's source view (profiles/r02_ncu_source_lod_fit_warp_kernel.txt):
// its residuals are ~1e-290, so (1) every norm goes through enorm's "small" branch, fifteen dependent divisions by
// the running maximum, (2) those and a dozen more divisions per iteration take the SLOW path of the fp64 division
// routine (operands at extreme exponents), (3) pivot-indexed work vectors lived in local memory.  25 -> 13.4 ms.
#ifdef __CUDACC__
#define LMW_FULL 0xffffffffu
__device__ __forceinline__ double lmw_get(double v, int lane) { return __shfl_sync(LMW_FULL, v, lane); }

// Sums over the curve points are taken by ALL lanes in index order.  The operands cross the warp through a 32-double
// row of shared memory per vector (one store per lane, then broadcast reads in batches of 16 so that the loads are
// in flight together and only the DFMA chain itself is sequential); start is 0 or 1, start + n <= 32.
struct LmwRow {
    double a[32];
    double b[32];
};

#define LMW_LOAD16(dst, src, base)                                              \
    _Pragma("unroll") for (int k_ = 0; k_ < 8; ++k_) {                          \
        const double2 p_ = reinterpret_cast<const double2*>(src)[(base) / 2 + k_]; \
        dst[2 * k_] = p_.x;                                                     \
        dst[2 * k_ + 1] = p_.y;                                                 \
    }

// enorm over points [start, start + n) of a lane-distributed vector.
// The usual case (every |x| in enorm's middle range) is the plain sum of squares in index order.  The refits that
// run away to maxfev are the other case: their residuals shrink to ~1e-290, enorm's "small" branch divides every
// component by the running maximum.  The running maximum is a prefix maximum: a shuffle scan gives every lane the
// maximum of the components before it, every lane then divides ITS component (all divisions of a norm at once), and
// only the accumulation s = 1 + s t^2 / s += t^2 stays sequential -- the same operations on the same operands as the
// loop form.
__device__ __noinline__ double lmw_enorm(double v, int start, int n, double agiant /* rgiant / n */, LmwRow* row) {
    const double rdwarf = 3.834e-20;
    double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
    const int lane = threadIdx.x & 31, end = start + n;
    const bool in = lane >= start && lane < end;
    const double xa = fabs(v);
    const bool mid = xa > rdwarf && xa < agiant;
    double x[16];
    if (__all_sync(LMW_FULL, mid || !in)) {
        row->a[lane] = xa;
        __syncwarp();
        LMW_LOAD16(x, row->a, 0)
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i >= start && i < end) s2 = fma(x[i], x[i], s2);
        if (end > 16) {
            LMW_LOAD16(x, row->a, 16)
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (16 + i < end) s2 = fma(x[i], x[i], s2);
        }
        __syncwarp();
        return s2 != 0.0 ? sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3))) : 0.0;
    }
    const bool sml = in && !mid && xa <= rdwarf;
    const bool big = in && !mid && !sml;
    const bool any_big = __any_sync(LMW_FULL, big);
    // inclusive prefix maxima of the small and of the large components (NaN never becomes a maximum: `xabs > max`)
    double m3 = sml ? xa : 0.0, m1 = (big && xa == xa) ? xa : 0.0;
    for (int d = 1; d < end; d <<= 1) {
        const double t3 = __shfl_up_sync(LMW_FULL, m3, d);
        if (lane >= d) m3 = fmax(m3, t3);
        if (any_big) {
            const double t1 = __shfl_up_sync(LMW_FULL, m1, d);
            if (lane >= d) m1 = fmax(m1, t1);
        }
    }
    double e3 = __shfl_up_sync(LMW_FULL, m3, 1), e1 = __shfl_up_sync(LMW_FULL, m1, 1);   // maxima BEFORE this component
    if (lane == 0) e3 = e1 = 0.0;
    // this lane's operand of the accumulation and which accumulation it is
    double t = xa;
    int op = 0;                                   // 1: s2 += t t   2: s3 = 1 + s3 t t   3: s3 += t t   4, 5: the same on s1
    if (in) {
        if (mid) {
            op = 1;
        } else if (sml) {
            if (xa > e3) {
                t = e3 / xa;
                op = 2;
            } else if (xa != 0.0) {
                t = xa / e3;
                op = 3;
            }
        } else {
            if (xa > e1) {
                t = e1 / xa;
                op = 4;
            } else {
                t = xa / e1;
                op = 5;
            }
        }
    }
    row->a[lane] = t;
    const unsigned b1 = __ballot_sync(LMW_FULL, op == 1), b2 = __ballot_sync(LMW_FULL, op == 2), b3 = __ballot_sync(LMW_FULL, op == 3),
                   b4 = __ballot_sync(LMW_FULL, op == 4), b5 = __ballot_sync(LMW_FULL, op == 5);
    __syncwarp();
    // (the products are written as the fused multiply-adds the loop form compiles to)
    if ((b1 | b4 | b5) == 0u) {                   // small components only: every norm of a runaway refit
#pragma unroll
        for (int h = 0; h < 32; h += 16) {
            if (h < end) {
                LMW_LOAD16(x, row->a, h)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const unsigned bit = 1u << (h + i);
                    const double ti = x[i];
                    const double r2 = fma(s3, ti * ti, 1.0), r3 = fma(ti, ti, s3);
                    s3 = (b2 & bit) ? r2 : (b3 & bit) ? r3 : s3;
                }
            }
        }
    } else {
#pragma unroll
        for (int h = 0; h < 32; h += 16) {
            if (h < end) {
                LMW_LOAD16(x, row->a, h)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const unsigned bit = 1u << (h + i);
                    const double ti = x[i];
                    if (b1 & bit) s2 = fma(ti, ti, s2);
                    if (b2 & bit) s3 = fma(s3, ti * ti, 1.0);
                    if (b3 & bit) s3 = fma(ti, ti, s3);
                    if (b4 & bit) s1 = fma(s1, ti * ti, 1.0);
                    if (b5 & bit) s1 = fma(ti, ti, s1);
                }
            }
        }
    }
    __syncwarp();
    if (n > 0) {
        x3max = lmw_get(m3, end - 1);
        x1max = lmw_get(m1, end - 1);
    }
    if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.0) {
        if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

// TWO norms for the price of one: a curve has at most 16 points here, so half of the warp idles.  Lanes 0-15 carry
// vector A, lanes 16-31 vector B (point i on lane i of either half), every shuffle is confined to its half (width 16) and
// each half folds ITS 16 operands -- one instruction stream, two norms.  The path (plain sum of squares / small components
// only / general) is chosen for the warp as a whole, the more general one when the halves differ: all three give the same
// bits where they apply.  Returns the norm of the caller's half.  m <= 16.
__device__ __noinline__ double lmw_enorm_pair(double v, int start, int n, double agiant, LmwRow* row) {
    const double rdwarf = 3.834e-20;
    double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
    const int lane = threadIdx.x & 31, sl = lane & 15, hb = lane & 16, end = start + n;
    const bool in = sl >= start && sl < end;
    const double xa = fabs(v);
    const bool mid = xa > rdwarf && xa < agiant;
    double x[16];
    if (__all_sync(LMW_FULL, mid || !in)) {
        row->a[lane] = xa;
        __syncwarp();
        LMW_LOAD16(x, row->a, hb)
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i >= start && i < end) s2 = fma(x[i], x[i], s2);
        __syncwarp();
        return s2 != 0.0 ? sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3))) : 0.0;
    }
    const bool sml = in && !mid && xa <= rdwarf;
    const bool big = in && !mid && !sml;
    double m3 = sml ? xa : 0.0, m1 = (big && xa == xa) ? xa : 0.0;
    for (int d = 1; d < end; d <<= 1) {
        const double t3 = __shfl_up_sync(LMW_FULL, m3, d, 16), t1 = __shfl_up_sync(LMW_FULL, m1, d, 16);
        if (sl >= d) {
            m3 = fmax(m3, t3);
            m1 = fmax(m1, t1);
        }
    }
    double e3 = __shfl_up_sync(LMW_FULL, m3, 1, 16), e1 = __shfl_up_sync(LMW_FULL, m1, 1, 16);
    if (sl == 0) e3 = e1 = 0.0;
    double t = xa;
    int op = 0;
    if (in) {
        if (mid) {
            op = 1;
        } else if (sml) {
            if (xa > e3) {
                t = e3 / xa;
                op = 2;
            } else if (xa != 0.0) {
                t = xa / e3;
                op = 3;
            }
        } else {
            if (xa > e1) {
                t = e1 / xa;
                op = 4;
            } else {
                t = xa / e1;
                op = 5;
            }
        }
    }
    row->a[lane] = t;
    const unsigned a1 = __ballot_sync(LMW_FULL, op == 1), a2 = __ballot_sync(LMW_FULL, op == 2), a3 = __ballot_sync(LMW_FULL, op == 3),
                   a4 = __ballot_sync(LMW_FULL, op == 4), a5 = __ballot_sync(LMW_FULL, op == 5);
    const unsigned b1 = (a1 >> hb) & 0xffffu, b2 = (a2 >> hb) & 0xffffu, b3 = (a3 >> hb) & 0xffffu, b4 = (a4 >> hb) & 0xffffu,
                   b5 = (a5 >> hb) & 0xffffu;
    __syncwarp();
    LMW_LOAD16(x, row->a, hb)
    if ((a1 | a4 | a5) == 0u) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const unsigned bit = 1u << i;
            const double ti = x[i];
            const double r2 = fma(s3, ti * ti, 1.0), r3 = fma(ti, ti, s3);
            s3 = (b2 & bit) ? r2 : (b3 & bit) ? r3 : s3;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const unsigned bit = 1u << i;
            const double ti = x[i];
            if (b1 & bit) s2 = fma(ti, ti, s2);
            if (b2 & bit) s3 = fma(s3, ti * ti, 1.0);
            if (b3 & bit) s3 = fma(ti, ti, s3);
            if (b4 & bit) s1 = fma(s1, ti * ti, 1.0);
            if (b5 & bit) s1 = fma(ti, ti, s1);
        }
    }
    __syncwarp();
    if (n > 0) {
        x3max = __shfl_sync(LMW_FULL, m3, end - 1, 16);
        x1max = __shfl_sync(LMW_FULL, m1, end - 1, 16);
    }
    if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.0) {
        if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

// sum_{i = start}^{m-1} a[i] * b[i], in index order
__device__ __noinline__ double lmw_dot(double a, double b, int start, int m, LmwRow* row) {
    const int lane = threadIdx.x & 31;
    row->a[lane] = a;
    row->b[lane] = b;
    __syncwarp();
    double sum = 0.0, x[16], y[16];
    LMW_LOAD16(x, row->a, 0)
    LMW_LOAD16(y, row->b, 0)
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (i >= start && i < m) sum = fma(x[i], y[i], sum);
    if (m > 16) {
        LMW_LOAD16(x, row->a, 16)
        LMW_LOAD16(y, row->b, 16)
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (16 + i < m) sum = fma(x[i], y[i], sum);
    }
    __syncwarp();
    return sum;
}

// ---- the 2 x 2 algebra in REGISTERS ------------------------------------------------------------------------
// The general lm_lmpar / lm_qrsolv index their work vectors by the pivot order (x[ipvt[j]], diag[ipvt[j]]): ptxas
// keeps every such array in LOCAL memory and each dependent step waits for an L1 round trip.  Below, n = 2 is written
// out: the pivot order is one bit (`sw`: the columns were exchanged), every vector is two scalars, the loops are
// unrolled by hand.  Every expression is the one of the loop form, in its order, so the fits are unchanged.
__device__ __forceinline__ void lm2_enorm_step(double xabs, double agiant, double& s1, double& s2, double& s3, double& x1max, double& x3max) {
    const double rdwarf = 3.834e-20;
    if (xabs > rdwarf && xabs < agiant) {
        s2 += xabs * xabs;
    } else if (xabs <= rdwarf) {
        if (xabs > x3max) {
            const double t = x3max / xabs;
            s3 = 1.0 + s3 * (t * t);
            x3max = xabs;
        } else if (xabs != 0.0) {
            const double t = xabs / x3max;
            s3 += t * t;
        }
    } else {
        if (xabs > x1max) {
            const double t = x1max / xabs;
            s1 = 1.0 + s1 * (t * t);
            x1max = xabs;
        } else {
            const double t = xabs / x1max;
            s1 += t * t;
        }
    }
}
__device__ __noinline__ double lm2_enorm(double a, double b) {          // lm_enorm(2, {a, b})
    const double rgiant = 1.304e19;
    double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
    const double agiant = rgiant / 2.0;
    lm2_enorm_step(fabs(a), agiant, s1, s2, s3, x1max, x3max);
    lm2_enorm_step(fabs(b), agiant, s1, s2, s3, x1max, x3max);
    if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.0) {
        if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}
__device__ __forceinline__ double lm2_sel(bool sw, double a0, double a1) { return sw ? a1 : a0; }   // a[ipvt[0]]
__device__ __forceinline__ void lm2_givens(double rkk, double sd, double& cs, double& sn) {
    if (fabs(rkk) < fabs(sd)) {
        const double cotan = rkk / sd;
        sn = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
        cs = sn * cotan;
    } else {
        const double tn = sd / rkk;
        cs = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
        sn = cs * tn;
    }
}

// lm_qrsolv for n = 2.  (r00, r01, r11): the upper triangle, left untouched (the general form saves and restores
// its diagonal); t10: the strict lower triangle it leaves behind, which lmpar's Newton step reads.
__device__ __forceinline__ void lm2_qrsolv(double r00, double r01, double r11, bool sw, double dg0, double dg1, double qtb0, double qtb1,
                                           double& x0, double& x1, double& sdiag0, double& sdiag1, double& t10) {
    double d00 = r00, d11 = r11, wa0 = qtb0, wa1 = qtb1;
    t10 = r01;
    {   // j = 0
        const double dl = lm2_sel(sw, dg0, dg1);
        if (dl != 0.0) {
            double sd0 = dl, sd1 = 0.0, qtbpj = 0.0, cs, sn;
            lm2_givens(d00, sd0, cs, sn);                    // k = 0
            d00 = cs * d00 + sn * sd0;
            {
                const double temp = cs * wa0 + sn * qtbpj;
                qtbpj = -sn * wa0 + cs * qtbpj;
                wa0 = temp;
            }
            {
                const double t2 = cs * t10 + sn * sd1;
                sd1 = -sn * t10 + cs * sd1;
                t10 = t2;
            }
            if (sd1 != 0.0) {                                // k = 1
                lm2_givens(d11, sd1, cs, sn);
                d11 = cs * d11 + sn * sd1;
                const double temp = cs * wa1 + sn * qtbpj;
                qtbpj = -sn * wa1 + cs * qtbpj;
                wa1 = temp;
            }
        }
        sdiag0 = d00;
    }
    {   // j = 1
        const double dl = lm2_sel(sw, dg1, dg0);
        if (dl != 0.0) {
            double sd1 = dl, qtbpj = 0.0, cs, sn;
            lm2_givens(d11, sd1, cs, sn);
            d11 = cs * d11 + sn * sd1;
            const double temp = cs * wa1 + sn * qtbpj;
            wa1 = temp;
        }
        sdiag1 = d11;
    }
    if (sdiag0 == 0.0) {
        wa0 = 0.0;
        wa1 = 0.0;
    } else if (sdiag1 == 0.0) {
        wa1 = 0.0;
        double sum = 0.0;
        wa0 = (wa0 - sum) / sdiag0;
    } else {
        {
            double sum = 0.0;
            wa1 = (wa1 - sum) / sdiag1;
        }
        double sum = 0.0;
        sum += t10 * wa1;
        wa0 = (wa0 - sum) / sdiag0;
    }
    x0 = sw ? wa1 : wa0;                                     // x[ipvt[j]] = wa[j]
    x1 = sw ? wa0 : wa1;
}

// lm_lmpar for n = 2 (x0, x1: the step, indexed by parameter)
__device__ __forceinline__ void lm2_lmpar(double r00, double r01, double r11, bool sw, double diag0, double diag1, double qtb0, double qtb1,
                                          double delta, double& par, double& x0, double& x1) {
    const double dwarf = 2.2250738585072014e-308;
    int nsing = 2;
    double w0 = qtb0, w1 = qtb1;
    if (r00 == 0.0) {
        nsing = 0;
        w0 = 0.0;
        w1 = 0.0;
    } else if (r11 == 0.0) {
        nsing = 1;
        w1 = 0.0;
    }
    if (nsing == 2) {
        w1 /= r11;
        const double temp = w1;
        w0 -= r01 * temp;
        w0 /= r00;
    } else if (nsing == 1) {
        w0 /= r00;
    }
    x0 = sw ? w1 : w0;
    x1 = sw ? w0 : w1;
    int iter = 0;
    double v0 = diag0 * x0, v1 = diag1 * x1;                 // wa2
    double dxnorm = lm2_enorm(v0, v1);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) {
        par = 0.0;
        return;
    }
    double parl = 0.0;
    if (nsing >= 2) {
        w0 = lm2_sel(sw, diag0, diag1) * (lm2_sel(sw, v0, v1) / dxnorm);
        w1 = lm2_sel(sw, diag1, diag0) * (lm2_sel(sw, v1, v0) / dxnorm);
        {
            double sum = 0.0;
            w0 = (w0 - sum) / r00;
        }
        {
            double sum = 0.0;
            sum += r01 * w0;
            w1 = (w1 - sum) / r11;
        }
        const double temp = lm2_enorm(w0, w1);
        parl = ((fp / delta) / temp) / temp;
    }
    {
        double sum = 0.0;
        sum += r00 * qtb0;
        w0 = sum / lm2_sel(sw, diag0, diag1);
    }
    {
        double sum = 0.0;
        sum += r01 * qtb0;
        sum += r11 * qtb1;
        w1 = sum / lm2_sel(sw, diag1, diag0);
    }
    const double gnorm = lm2_enorm(w0, w1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = dwarf / fmin(delta, 0.1);
    par = fmax(par, parl);
    par = fmin(par, paru);
    if (par == 0.0) par = gnorm / dxnorm;
    for (;;) {
        ++iter;
        if (par == 0.0) par = fmax(dwarf, 0.001 * paru);
        double temp = sqrt(par);
        w0 = temp * diag0;
        w1 = temp * diag1;
        double sdiag0, sdiag1, t10;
        lm2_qrsolv(r00, r01, r11, sw, w0, w1, qtb0, qtb1, x0, x1, sdiag0, sdiag1, t10);
        v0 = diag0 * x0;
        v1 = diag1 * x1;
        dxnorm = lm2_enorm(v0, v1);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        w0 = lm2_sel(sw, diag0, diag1) * (lm2_sel(sw, v0, v1) / dxnorm);
        w1 = lm2_sel(sw, diag1, diag0) * (lm2_sel(sw, v1, v0) / dxnorm);
        w0 /= sdiag0;
        {
            const double t = w0;
            w1 -= t10 * t;
        }
        w1 /= sdiag1;
        temp = lm2_enorm(w0, w1);
        const double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = fmax(parl, par);
        if (fp < 0.0) paru = fmin(paru, par);
        par = fmax(parl, par + parc);
    }
}

// two dot products over the points at once: lanes 0-15 take sum a[i] b[i] of the values on THEIR half, lanes 16-31 of theirs
__device__ __noinline__ double lmw_dot_pair(double a, double b, int start, int m, LmwRow* row) {
    const int lane = threadIdx.x & 31, hb = lane & 16;
    row->a[lane] = a;
    row->b[lane] = b;
    __syncwarp();
    double sum = 0.0, x[16], y[16];
    LMW_LOAD16(x, row->a, hb)
    LMW_LOAD16(y, row->b, hb)
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (i >= start && i < m) sum = fma(x[i], y[i], sum);
    __syncwarp();
    return sum;
}

__device__ __noinline__ double lmw_fcn2(double px, double py, double a, double b) {
    return 1.0 / (1.0 + exp(-(a * px + b))) - py;
}

__device__ static int lmw_lmdif2(int m, int lane, double px, double py, double* xio, int maxfev, LmwRow* row) {
    const double epsmch = 2.220446049250313e-16;
    const double ftol = 1.49012e-8, xtol = 1.49012e-8, gtol = 0.0, factor = 100.0;
    double x0 = xio[0], x1 = xio[1];
    double fvec, wa4v, c0, c1;
    double diag0 = 0.0, diag1 = 0.0;
    int info = 0, nfev = 0;
    const double ag_m = 1.304e19 / (double)m, ag_m1 = 1.304e19 / (double)(m - 1);      // enorm's agiant for n = m, m - 1
    const double px2 = __shfl_sync(LMW_FULL, px, lane & 15), py2 = __shfl_sync(LMW_FULL, py, lane & 15);   // the curve, on both halves
    fvec = lmw_fcn2(px, py, x0, x1);
    nfev = 1;
    double fnorm = lmw_enorm(fvec, 0, m, ag_m, row);
    double par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;
    int iter = 1;
    for (;;) {
        double acnorm0, acnorm1;
        if (m <= 16) {
            // both forward differences at once, column 0 on lanes 0-15 and column 1 on lanes 16-31 (which hold a copy of
            // the curve), and both column norms by one pass of lmw_enorm_pair
            const double eps = sqrt(epsmch);
            const bool up = lane >= 16;
            const double temp = up ? x1 : x0;
            double h = eps * fabs(temp);
            if (h == 0.0) h = eps;
            const double f2 = __shfl_sync(LMW_FULL, fvec, lane & 15);
            const double val = lmw_fcn2(px2, py2, up ? x0 : temp + h, up ? temp + h : x1);
            const double d = (val - f2) / h;
            const double nrm = lmw_enorm_pair(d, 0, m, ag_m, row);
            c0 = __shfl_sync(LMW_FULL, d, lane & 15);
            c1 = __shfl_sync(LMW_FULL, d, (lane & 15) | 16);
            acnorm0 = __shfl_sync(LMW_FULL, nrm, 0);
            acnorm1 = __shfl_sync(LMW_FULL, nrm, 16);
            nfev += LM_N;
        } else {
            const double eps = sqrt(epsmch);
            {
                const double temp = x0;
                double h = eps * fabs(temp);
                if (h == 0.0) h = eps;
                wa4v = lmw_fcn2(px, py, temp + h, x1);
                c0 = (wa4v - fvec) / h;
            }
            {
                const double temp = x1;
                double h = eps * fabs(temp);
                if (h == 0.0) h = eps;
                wa4v = lmw_fcn2(px, py, x0, temp + h);
                c1 = (wa4v - fvec) / h;
            }
            nfev += LM_N;
            acnorm0 = lmw_enorm(c0, 0, m, ag_m, row);
            acnorm1 = lmw_enorm(c1, 0, m, ag_m, row);
        }
        // qrfac (lmw_qrfac with the pivot order as one bit)
        double rdiag0, rdiag1, qw1, sum_qf = 0.0;
        bool sw = false, have_qf = false, have_aj = false;
        {
            rdiag0 = acnorm0;
            rdiag1 = qw1 = acnorm1;
            if (rdiag1 > rdiag0) {
                const double t = c0; c0 = c1; c1 = t;
                rdiag1 = rdiag0;
                qw1 = acnorm0;
                sw = true;
            }
            double ajnorm = sw ? acnorm1 : acnorm0;         // enorm of the pivot column: the norm already taken
            if (ajnorm != 0.0) {
                if (lmw_get(c0, 0) < 0.0) ajnorm = -ajnorm;
                if (lane < m) c0 /= ajnorm;
                if (lane == 0) c0 += 1.0;
                double sum;
                if (m <= 16) {
                    // this dot product and the one Q^T f needs of the same column (below), one per half-warp
                    const double c0m = __shfl_sync(LMW_FULL, c0, lane & 15), fm = __shfl_sync(LMW_FULL, fvec, lane & 15);   // (all lanes shuffle)
                    const double pr = lmw_dot_pair(c0m, lane < 16 ? c1 : fm, 0, m, row);
                    sum = __shfl_sync(LMW_FULL, pr, 0);
                    sum_qf = __shfl_sync(LMW_FULL, pr, 16);
                    have_qf = true;
                } else {
                    sum = lmw_dot(c0, c1, 0, m, row);
                }
                const double temp = sum / lmw_get(c0, 0);
                if (lane < m) c1 -= temp * c0;
                if (rdiag1 != 0.0) {
                    double t = lmw_get(c1, 0) / rdiag1;
                    rdiag1 *= sqrt(fmax(0.0, 1.0 - t * t));
                    t = rdiag1 / qw1;
                    if (0.05 * (t * t) <= epsmch) {
                        rdiag1 = lmw_enorm(c1, 1, m - 1, ag_m1, row);
                        qw1 = rdiag1;
                        have_aj = true;                             // the very norm the next step asks for
                    }
                }
            }
            rdiag0 = -ajnorm;
            ajnorm = have_aj ? rdiag1 : lmw_enorm(c1, 1, m - 1, ag_m1, row);
            if (ajnorm != 0.0) {
                if (lmw_get(c1, 1) < 0.0) ajnorm = -ajnorm;
                if (lane >= 1 && lane < m) c1 /= ajnorm;
                if (lane == 1) c1 += 1.0;
            }
            rdiag1 = -ajnorm;
        }
        if (iter == 1) {
            diag0 = acnorm0;
            if (acnorm0 == 0.0) diag0 = 1.0;
            diag1 = acnorm1;
            if (acnorm1 == 0.0) diag1 = 1.0;
            xnorm = lm2_enorm(diag0 * x0, diag1 * x1);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        wa4v = fvec;
        double qtf0, qtf1;
        {   // j = 0
            const double ajj = lmw_get(c0, 0);
            if (ajj != 0.0) {
                const double sum = have_qf ? sum_qf : lmw_dot(c0, wa4v, 0, m, row);
                const double temp = -sum / ajj;
                if (lane < m) wa4v += c0 * temp;
            }
            if (lane == 0) c0 = rdiag0;
            qtf0 = lmw_get(wa4v, 0);
        }
        {   // j = 1
            const double ajj = lmw_get(c1, 1);
            if (ajj != 0.0) {
                const double sum = lmw_dot(c1, wa4v, 1, m, row);
                const double temp = -sum / ajj;
                if (lane >= 1 && lane < m) wa4v += c1 * temp;
            }
            if (lane == 1) c1 = rdiag1;
            qtf1 = lmw_get(wa4v, 1);
        }
        const double r00 = lmw_get(c0, 0), r01 = lmw_get(c1, 0), r11 = lmw_get(c1, 1);
        gnorm = 0.0;
        if (fnorm != 0.0) {
            {
                const double acn = lm2_sel(sw, acnorm0, acnorm1);
                if (acn != 0.0) {
                    double sum = 0.0;
                    sum += r00 * (qtf0 / fnorm);
                    gnorm = fmax(gnorm, fabs(sum / acn));
                }
            }
            {
                const double acn = lm2_sel(sw, acnorm1, acnorm0);
                if (acn != 0.0) {
                    double sum = 0.0;
                    sum += r01 * (qtf0 / fnorm);
                    sum += r11 * (qtf1 / fnorm);
                    gnorm = fmax(gnorm, fabs(sum / acn));
                }
            }
        }
        if (gnorm <= gtol) info = 4;
        if (info != 0) break;
        diag0 = fmax(diag0, acnorm0);
        diag1 = fmax(diag1, acnorm1);
        double ratio;
        do {
            double p0, p1;
            lm2_lmpar(r00, r01, r11, sw, diag0, diag1, qtf0, qtf1, delta, par, p0, p1);
            p0 = -p0;
            p1 = -p1;
            const double n0 = x0 + p0, n1 = x1 + p1;           // wa2
            const double pnorm = lm2_enorm(diag0 * p0, diag1 * p1);
            if (iter == 1) delta = fmin(delta, pnorm);
            wa4v = lmw_fcn2(px, py, n0, n1);
            ++nfev;
            const double fnorm1 = lmw_enorm(wa4v, 0, m, ag_m, row);
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) {
                const double t = fnorm1 / fnorm;
                actred = 1.0 - t * t;
            }
            double u0 = 0.0, u1;                                // wa3 = R * P^T p
            {
                const double temp = lm2_sel(sw, p0, p1);
                u0 += r00 * temp;
            }
            {
                u1 = 0.0;
                const double temp = lm2_sel(sw, p1, p0);
                u0 += r01 * temp;
                u1 += r11 * temp;
            }
            const double temp1 = lm2_enorm(u0, u1) / fnorm;
            const double temp2 = (sqrt(par) * pnorm) / fnorm;
            const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
            const double dirder = -(temp1 * temp1 + temp2 * temp2);
            ratio = 0.0;
            if (prered != 0.0) ratio = actred / prered;
            if (ratio <= 0.25) {
                double temp;
                if (actred >= 0.0) temp = 0.5;
                else temp = 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * fmin(delta, pnorm / 0.1);
                par /= temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par *= 0.5;
            }
            if (ratio >= 1e-4) {
                x0 = n0;
                x1 = n1;
                fvec = wa4v;
                xnorm = lm2_enorm(diag0 * x0, diag1 * x1);
                fnorm = fnorm1;
                ++iter;
            }
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) info = 1;
            if (delta <= xtol * xnorm) info = 2;
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0 && info == 2) info = 3;
            if (info != 0) break;
            if (nfev >= maxfev) info = 5;
            if (fabs(actred) <= epsmch && prered <= epsmch && 0.5 * ratio <= 1.0) info = 6;
            if (delta <= epsmch * xnorm) info = 7;
            if (gnorm <= epsmch) info = 8;
            if (info != 0) break;
        } while (ratio < 1e-4);
        if (info != 0) break;
    }
    xio[0] = x0;
    xio[1] = x1;
    return info;
}
#endif  // __CUDACC__

// scipy.interpolate.interp1d(hit, af, bounds_error=False, fill_value='extrapolate')(target)
__host__ __device__ static double lod_interp_fallback(const LmProblem& P, const double* af, double target) {
    const int m = P.m;
    if (m <= 1) return af[0];
    int ord[LOD_MAX_M];
    for (int i = 0; i < m; ++i) ord[i] = i;
    for (int i = 1; i < m; ++i) {  // stable insertion sort by hit rate (interp1d: mergesort argsort)
        const int v = ord[i];
        int j = i - 1;
        while (j >= 0 && P.y[ord[j]] > P.y[v]) { ord[j + 1] = ord[j]; --j; }
        ord[j + 1] = v;
    }
    int idx = 0;  // searchsorted(x, target) side='left'
    while (idx < m && P.y[ord[idx]] < target) ++idx;
    idx = idx < 1 ? 1 : (idx > m - 1 ? m - 1 : idx);
    const double xlo = P.y[ord[idx - 1]], xhi = P.y[ord[idx]];
    const double ylo = af[ord[idx - 1]], yhi = af[ord[idx]];
    // scipy _call_linear: ((x - x_lo)/(x_hi - x_lo)) * y_hi + ((x_hi - x)/(x_hi - x_lo)) * y_lo
    // (inf - inf = NaN when the bracketing hit rates coincide -- the reference's all-zero case)
    return ((target - xlo) / (xhi - xlo)) * yhi + ((xhi - target) / (xhi - xlo)) * ylo;
}

#define LOD_MAXFEV 2000
__host__ __device__ static double lod_fit_one(const LmProblem& P, const double* af, double target) {
    double p[2] = {1.0, 1.0};
    bool finite_in = true;
    for (int i = 0; i < P.m; ++i) finite_in &= isfinite(P.x[i]) && isfinite(P.y[i]);
    int info = finite_in ? lm_lmdif(P, p, LOD_MAXFEV) : 0;
    if (info >= 1 && info <= 4) {
        // lod.py:365-371
        const double logit_target = log(target / (1.0 - target));
        return pow(10.0, (logit_target - p[1]) / p[0]);
    }
    return lod_interp_fallback(P, af, target);
}

// point i of resample t of a series (t = 0: the curve itself): lod.py:389-399 on a Philox stream
__device__ __forceinline__ int lod_source(const Philox& ph, uint64_t stream_id, int t, int i, int m) {
    if (t == 0) return i;
    const uint4 r = ph((uint32_t)i, (uint32_t)t, (uint32_t)stream_id, (PMRD_STREAM_LODBOOT << 24) | (uint32_t)((stream_id >> 32) & 0xffffffu));
    const int src = (int)(u53(r.x, r.y) * (double)m);
    return src >= m ? m - 1 : src;
}

// thread 0 of a series: point estimate; threads 1..n_boot: resampled pairs (lod.py:389-399).
// blockIdx.y = series (one detection curve per depth): all curves of an analysis are fitted by ONE launch.
// One THREAD per fit: the general form (any m <= LOD_MAX_M); curves of <= 32 points take the warp kernel below.
__global__ void __launch_bounds__(64)
lod_fit_kernel(uint64_t seed, const uint64_t* __restrict__ stream_ids, const double* __restrict__ af,
               const double* __restrict__ hits /*[S][m]*/, int m, double target, int n_boot, double* __restrict__ fits /*[S][1+n_boot]*/) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_boot) return;
    const uint64_t stream_id = stream_ids[blockIdx.y];
    const double* __restrict__ hit = hits + (size_t)blockIdx.y * m;
    LmProblem P;
    double afs[LOD_MAX_M];
    P.m = m;
    const Philox ph(seed);
    for (int i = 0; i < m; ++i) {
        const int src = lod_source(ph, stream_id, t, i, m);
        afs[i] = af[src];
        P.x[i] = log10(af[src]);
        P.y[i] = hit[src];
    }
    fits[(size_t)blockIdx.y * (n_boot + 1) + t] = lod_fit_one(P, afs, target);
}

// One WARP per fit, lane i = curve point i (m <= 32): see lmw_lmdif.
#define LODW_FITS_PER_BLOCK 4
__global__ void __launch_bounds__(32 * LODW_FITS_PER_BLOCK)
lod_fit_warp_kernel(uint64_t seed, const uint64_t* __restrict__ stream_ids, const double* __restrict__ af,
                    const double* __restrict__ hits /*[S][m]*/, int m, double target, int n_boot, double* __restrict__ fits /*[S][1+n_boot]*/) {
    const int t = blockIdx.x * LODW_FITS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t > n_boot) return;                          // (whole warps leave together)
    const uint64_t stream_id = stream_ids[blockIdx.y];
    const double* __restrict__ hit = hits + (size_t)blockIdx.y * m;
    const Philox ph(seed);
    double a = 1.0, px = 0.0, py = 0.0;
    if (lane < m) {
        const int src = lod_source(ph, stream_id, t, lane, m);
        a = af[src];
        px = log10(a);
        py = hit[src];
    }
    const bool finite_in = __all_sync(0xffffffffu, lane >= m || (isfinite(px) && isfinite(py)));
    double p[2] = {1.0, 1.0};
    __shared__ __align__(16) LmwRow s_row[LODW_FITS_PER_BLOCK];
    const int info = finite_in ? lmw_lmdif2(m, lane, px, py, p, LOD_MAXFEV, &s_row[threadIdx.x >> 5]) : 0;
    double lod;
    if (info >= 1 && info <= 4) {
        const double logit_target = log(target / (1.0 - target));          // lod.py:365-371
        lod = pow(10.0, (logit_target - p[1]) / p[0]);
    } else {
        // interpolation fallback on the gathered curve (every lane runs it: the shuffles need the whole warp)
        LmProblem P;
        double afs[32];
        P.m = m;
        for (int i = 0; i < m; ++i) {
            P.x[i] = __shfl_sync(0xffffffffu, px, i);
            P.y[i] = __shfl_sync(0xffffffffu, py, i);
            afs[i] = __shfl_sync(0xffffffffu, a, i);
        }
        lod = lod_interp_fallback(P, afs, target);
    }
    if (lane == 0) fits[(size_t)blockIdx.y * (n_boot + 1) + t] = lod;
}

// numpy.percentile(method='linear') incl. its lerp form; NaN in -> NaN out
__host__ __device__ static double np_percentile_sorted(const double* s, int n, double q) {
    const double pos = q / 100.0 * (double)(n - 1);
    const int lo = (int)floor(pos);
    const int hi = lo + 1 < n ? lo + 1 : n - 1;
    const double t = pos - (double)lo;
    const double a = s[lo], b = s[hi], d = b - a;
    return t >= 0.5 ? b - d * (1.0 - t) : a + d * t;
}

__global__ void __launch_bounds__(256) lod_ci_kernel(const double* __restrict__ fits_all, int n_boot, double* __restrict__ out3_all) {
    extern __shared__ double s_v[];
    __shared__ int s_nan;
    const double* __restrict__ fits = fits_all + (size_t)blockIdx.x * (n_boot + 1);   // one block per series
    double* __restrict__ out3 = out3_all + (size_t)blockIdx.x * 3;
    if (threadIdx.x == 0) s_nan = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n_boot; i += blockDim.x) {
        const double v = fits[1 + i];
        if (isnan(v)) s_nan = 1;
        // rank by counting (ties broken by index): O(n^2) but n_boot is a few hundred
        int rank = 0;
        for (int j = 0; j < n_boot; ++j) {
            const double w = fits[1 + j];
            rank += (w < v) || (w == v && j < i);
        }
        if (!isnan(v)) s_v[rank] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        out3[0] = fits[0];
        if (n_boot <= 0) {
            out3[1] = out3[2] = NAN;
        } else if (s_nan) {
            out3[1] = out3[2] = NAN;  // np.percentile propagates NaN
        } else {
            out3[1] = np_percentile_sorted(s_v, n_boot, 2.5);
            out3[2] = np_percentile_sorted(s_v, n_boot, 97.5);
        }
    }
}

int32_t k11_fit_lod_batch(pmrd_ctx* ctx, const double* d_af, const double* d_hits, int32_t n, int32_t n_series, double target,
                          int32_t n_boot, const uint64_t* d_stream_ids, double* d_out3 /*[S][3]*/) {
    PMRD_ARG(ctx, n >= 1 && n <= LOD_MAX_M, "1 <= n <= 64 curve points");
    PMRD_ARG(ctx, n_series >= 1 && n_series <= 65535, "1 <= n_series <= 65535");
    PMRD_ARG(ctx, n_boot >= 0 && n_boot <= 4096, "0 <= n_boot <= 4096");
    PMRD_ARG(ctx, target > 0.0 && target < 1.0, "0 < target < 1");
    DevBuf fits;
    PMRD_CUDA(ctx, fits.alloc(ctx->stream, (size_t)n_series * (size_t)(n_boot + 1) * 8));
    const char* e = getenv("PMRD_LOD_WARP");               // 0: one thread per fit (the general kernel) whatever n is
    const int warp_fit = e ? atoi(e) : 1;
    if (warp_fit && n <= 32) {
        PMRD_LAUNCH(ctx, lod_fit_warp_kernel, dim3(pmrd_blocks(n_boot + 1, LODW_FITS_PER_BLOCK), n_series), 32 * LODW_FITS_PER_BLOCK, 0,
                    ctx->seed, d_stream_ids, d_af, d_hits, (int)n, target, (int)n_boot, fits.as<double>());
    } else {
        PMRD_LAUNCH(ctx, lod_fit_kernel, dim3(pmrd_blocks(n_boot + 1, 64), n_series), 64, 0, ctx->seed, d_stream_ids, d_af, d_hits, (int)n,
                    target, (int)n_boot, fits.as<double>());
    }
    PMRD_LAUNCH(ctx, lod_ci_kernel, n_series, 256, (size_t)(n_boot > 0 ? n_boot : 1) * 8, (const double*)fits.as<double>(), (int)n_boot, d_out3);
    return PMRD_OK;
}
