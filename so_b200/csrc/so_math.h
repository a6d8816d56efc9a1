/* so_math.h — the exact-arithmetic contract of the SO hot path, shared by host and device code.
 *
 * Every expression here reproduces, operation for operation, what the reference computes:
 *   fDist2      smooth2.c:89-92, image choice kd2.h:165-252 (per particle)
 *   rhoEnclosed kd2.c:588-593 (fp32 in, fp64 sqrt/mul/div, fp32 out, constant 1.33333333*M_PI)
 *   schedule    kd2.c:745,765-768 (fBall *= 1.2 in double, rounded to float)
 *   mass        kd2.c:787,807 — a SEQUENTIAL fp32 running sum; for equal particle masses it is a
 *               function of the count only, S[k] = fl(S[k-1] + m), tabulated here in closed form.
 */
#ifndef SO_MATH_H
#define SO_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SO_HD __host__ __device__ __forceinline__
#else
#define SO_HD static inline
#endif

/* 1.33333333*M_PI and (4./3.)*M_PI as the C compiler folds them (IEEE double products). */
#define SO_C133PI 4.1887901943144152   /* 1.33333333 * M_PI = 0x1.0c15237798b29p+2 (gcc-folded) */
#define SO_C43PI  4.1887902047863905   /* (4./3.)    * M_PI = 0x1.0c152382d7365p+2 (gcc-folded) */

/* ---- sequential fp32 sum of k equal masses, compressed ---------------------------------------
 * Within one binade the increment of S[k] = fl(S[k-1]+m) is constant after at most one
 * transient step (round-to-nearest-even), so S is piecewise linear in k with a few segments per
 * binade.  Entry e covers k in [k0[e], k0[e+1]):  S[k] = s0[e] + (k-k0[e])*inc[e], exact. */
#define SO_MT_MAX 224

typedef struct {
    int32_t n;                    /* number of entries */
    float m;                      /* the particle mass */
    uint32_t k0[SO_MT_MAX + 1];   /* k0[n] = sentinel (max) */
    float s0[SO_MT_MAX];
    float inc[SO_MT_MAX];
} so_mass_table;

#ifdef __CUDACC__
#define SO_HD_FN __host__ __device__ inline
#else
#define SO_HD_FN static inline
#endif

SO_HD float so_f32_add(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;   /* force fp32 rounding on any host */
    return r;
#endif
}

/* Build the table for k = 0..kmax.  Returns 0, or -1 if it would need more than SO_MT_MAX
 * entries (does not happen for normal, positive m). */
SO_HD_FN int so_mass_table_build(so_mass_table *t, float m, uint64_t kmax)
{
    uint64_t k = 0;
    float S = 0.0f;
    t->n = 0;
    t->m = m;
    if (!(m > 0.0f) || !(m < INFINITY)) return -1;
    while (k <= kmax) {
        float S1 = so_f32_add(S, m), S2 = so_f32_add(S1, m);
        float d1 = S1 - S, d2 = S2 - S1;   /* exact: both operands are multiples of ulp(S) */
        int e0, e2;
        if (t->n >= SO_MT_MAX) return -1;
        t->k0[t->n] = (uint32_t)k;
        t->s0[t->n] = S;
        if (!(S1 < INFINITY)) return -1;
        frexpf(S, &e0);
        frexpf(S2, &e2);
        if (S > 0.0f && d1 == d2 && e0 == e2) {
            /* stable increment inside this binade: jump to its end */
            if (d1 == 0.0f) {          /* m below half an ulp: the sum has saturated */
                t->inc[t->n++] = 0.0f;
                k = kmax + 1;
                break;
            } else {
                double top = ldexp(1.0, e0) - ldexp(1.0, e0 - 24);   /* largest float in binade */
                double steps = floor((top - (double)S) / (double)d1);
                uint64_t ns = (uint64_t)steps;
                if (ns < 2) ns = 2;    /* S1,S2 are known to follow the rule */
                if (k + ns > kmax + 1) ns = kmax + 1 - k;
                t->inc[t->n++] = d1;
                S = (float)((double)S + (double)ns * (double)d1);
                k += ns;
            }
        } else {
            t->inc[t->n++] = d1;       /* single literal step */
            S = S1;
            k += 1;
        }
    }
    t->k0[t->n] = 0xFFFFFFFFu;
    return 0;
}

/* S[k] from table arrays (works on host arrays and on device shared-memory copies). */
SO_HD float so_mass_prefix_eval(const uint32_t *k0, const float *s0, const float *inc, int n,
                                uint32_t k)
{
    int lo = 0, hi = n - 1;
    while (lo < hi) {              /* last entry with k0 <= k */
        int mid = (lo + hi + 1) >> 1;
        if (k0[mid] <= k) lo = mid; else hi = mid - 1;
    }
    return (float)((double)s0[lo] + (double)(k - k0[lo]) * (double)inc[lo]);
}

/* ---- rhoEnclosed(mass, r2) < thr, kd2.c:588-593 and the comparisons at kd2.c:791-792,814-815 */
SO_HD int so_rho_below(float mass, float r2, float thr)
{
    float r3 = (float)((double)r2 * sqrt((double)r2));
    float rho = (float)((double)mass / (SO_C133PI * (double)r3));
    return rho < thr;
}

/* Certificate that NO particle with enclosed mass >= mass_lo and r^2 <= r2_hi can satisfy
 * so_rho_below(): true density exceeds thr by more than every rounding in so_rho_below. */
SO_HD int so_surely_not_below(float mass_lo, float r2_hi, float thr)
{
    double r3 = (double)r2_hi * sqrt((double)r2_hi);
    return (double)mass_lo >= (double)thr * SO_C133PI * r3 * (1.0 + 1.0e-6);
}

/* ---- kd2.c:817-818 (host only: libm pow) ---------------------------------------------------- */
static inline float so_rdelta_host(float mvir, float thr)
{
    float r3 = (float)((double)mvir / (SO_C43PI * (double)thr));
    return (float)pow((double)r3, 0.3333333333);
}

/* ---- kd2.c:765: fRootPeriod = sqrt(sqr(Lx)+sqr(Ly)+sqr(Lz)) (float sums, double sqrt) -------- */
SO_HD float so_root_period(float lx, float ly, float lz)
{
#ifdef __CUDA_ARCH__
    float s = __fadd_rn(__fadd_rn(__fmul_rn(lx, lx), __fmul_rn(ly, ly)), __fmul_rn(lz, lz));
#else
    volatile float xx = lx * lx, yy = ly * ly, zz = lz * lz;
    volatile float s1 = xx + yy;
    volatile float s = s1 + zz;
#endif
    return (float)sqrt((double)s);
}

/* kd2.c:767: fBall *= 1.2  (float * double -> float) */
SO_HD float so_next_ball(float ball) { return (float)((double)ball * 1.2); }

#endif /* SO_MATH_H */
