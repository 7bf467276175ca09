// kp_math.cuh — device arithmetic of the pattern score, bit-compatible with the reference's CPU path.
//
// The reference computes the self-score in float64 inside numba-compiled code
// (src/kmerpapa/algorithms/bottum_up_array_w_numba.py:54-61, ..._CV.py:56-78): numba lowers math.log /
// np.log to the C library, i.e. glibc 2.39 __log_fma (sysdeps/ieee754/dbl-64/e_log.c), which is < 1 ulp
// but not correctly rounded.  The float32 score table decides the partition through exact ties, so
// kp_log() below is the same IEEE-754 operation sequence with glibc's own coefficient table
// (kp_log_data.h, extracted by tools/extract_glibc_log_data.py): every fma/add/mul is spelled with an
// explicit round-to-nearest intrinsic so the compiler can neither fuse nor reassociate.
// Level 0 uses scipy's xlogy / xlog1py (..._w_numba.py:26-29, ..._CV.py:15-20): x*log(y) and
// x*cephes_log1p(y); kp_log1p() restates cephes' rational approximation (scipy xsf/cephes/unity.h).
#pragma once
#include <stdint.h>

#include "kp_log_data.h"

__constant__ double kpc_logA[5] = KP_LOG_A_INIT;
__constant__ double kpc_logB[11] = KP_LOG_B_INIT;
__constant__ double kpc_logTab[256] = KP_LOG_TAB_INIT;  // {invc, logc} x 128; CTAs stage it in shared memory

#define KP_FMA(a, b, c) __fma_rn((a), (b), (c))
#define KP_ADD(a, b) __dadd_rn((a), (b))
#define KP_SUB(a, b) __dsub_rn((a), (b))
#define KP_MUL(a, b) __dmul_rn((a), (b))

// Coefficients that enter an fma next to another constant: an fma takes one constant-bank operand for
// free, the second one costs a load.  Pinning these few in registers (the empty asm hides them from
// rematerialisation) removes ~12 constant loads per scored pattern.
struct KpLogK { double a1, a3, b1, b4, b7; };
__device__ __forceinline__ KpLogK kp_logk_load()
{
    KpLogK k;
    k.a1 = kpc_logA[1]; k.a3 = kpc_logA[3]; k.b1 = kpc_logB[1]; k.b4 = kpc_logB[4]; k.b7 = kpc_logB[7];
    asm volatile("" : "+d"(k.a1), "+d"(k.a3), "+d"(k.b1), "+d"(k.b4), "+d"(k.b7));
    return k;
}

// window tests on the high word (the window bounds have zero low words, so this is exact)
__device__ __forceinline__ bool kp_log_is_near1(int hi) { return (unsigned)(hi - 0x3fee0000) < 0x00030900u; }
__device__ __forceinline__ bool kp_log_is_plain(int hi) { return (((unsigned)hi >> 16) - 0x10u) <= 0x7fdfu; }  // +normal finite

// 1 - 2^-4 <= x < 1 + 0x1.09p-4, x != 1: polynomial in r = x - 1 with a split hi/lo head
__device__ __forceinline__ double kp_log_near1(double x, const KpLogK &K)
{
    const double *B = kpc_logB;
    double r = KP_SUB(x, 1.0);
    double r2 = KP_MUL(r, r);
    double r3 = KP_MUL(r, r2);
    double t2 = KP_FMA(r, B[2], K.b1);
    double t3 = KP_FMA(r, B[5], K.b4);
    double t5 = KP_FMA(r, B[8], K.b7);
    t2 = KP_FMA(r2, B[3], t2);
    t3 = KP_FMA(r2, B[6], t3);
    double t1 = KP_FMA(r2, B[9], t5);
    t1 = KP_FMA(r3, B[10], t1);
    t1 = KP_FMA(t1, r3, t3);
    t1 = KP_FMA(t1, r3, t2);
    double rw = KP_FMA(r, 0x1p27, r);
    double rhi = KP_FMA(-0x1p27, r, rw);
    double rlo = KP_SUB(r, rhi);
    double rhi2 = KP_MUL(rhi, rhi);
    double hi = KP_FMA(rhi2, -0.5, r);       // B[0] == -0.5 exactly
    double t8 = KP_SUB(r, hi);
    double rr = KP_ADD(r, rhi);
    double lo = KP_FMA(rhi2, -0.5, t8);
    double t = KP_MUL(-0.5, rlo);
    lo = KP_FMA(t, rr, lo);
    double y = KP_FMA(t1, r3, lo);
    return KP_ADD(hi, y);
}

// positive, normal, finite x outside the near-1 window; (hi, lo) are the words of x
__device__ __forceinline__ double kp_log_main(int hi, int lo, const double2 *__restrict__ tab, const KpLogK &K)
{
    const double *A = kpc_logA;
    unsigned thi = (unsigned)hi - 0x3fe60000u;            // high word of ix - OFF (OFF has a zero low word)
    int i = (int)((thi >> 13) & 127u);
    int k = (int)thi >> 20;
    double z = __hiloint2double(hi - (int)(thi & 0xfff00000u), lo);
    double2 e = tab[i];  // invc, logc
    double kd = (double)k;
    double w = KP_FMA(kd, KP_LOG_LN2HI, e.y);
    double r = KP_FMA(z, e.x, -1.0);
    double q5 = KP_FMA(r, A[2], K.a1);
    double hi_ = KP_ADD(r, w);
    double r2 = KP_MUL(r, r);
    double lo_ = KP_SUB(w, hi_);
    lo_ = KP_ADD(lo_, r);
    lo_ = KP_FMA(kd, KP_LOG_LN2LO, lo_);
    double r3 = KP_MUL(r, r2);
    double q1 = KP_FMA(r, A[4], K.a3);
    lo_ = KP_FMA(r2, A[0], lo_);
    q1 = KP_FMA(q1, r2, q5);
    double y = KP_FMA(r3, q1, lo_);
    return KP_ADD(y, hi_);
}

// tab: 128 x {invc, logc}, in shared memory (data-dependent index; constant memory would serialise)
__device__ __forceinline__ double kp_log(double x, const double2 *__restrict__ tab)
{
    KpLogK K;
    K.a1 = kpc_logA[1]; K.a3 = kpc_logA[3]; K.b1 = kpc_logB[1]; K.b4 = kpc_logB[4]; K.b7 = kpc_logB[7];
    int hi = __double2hiint(x), lo = __double2loint(x);
    if (kp_log_is_near1(hi)) {
        if (hi == 0x3ff00000 && lo == 0) return 0.0;
        return kp_log_near1(x, K);
    }
    if (!kp_log_is_plain(hi)) {
        unsigned long long ix = (unsigned long long)__double_as_longlong(x);
        unsigned top = (unsigned)(ix >> 48);
        if (ix * 2 == 0) return -__longlong_as_double(0x7ff0000000000000LL);  // log(0) = -inf
        if (ix == 0x7ff0000000000000ULL) return x;
        if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return __longlong_as_double(0x7ff8000000000000LL);
        ix = (unsigned long long)__double_as_longlong(KP_MUL(x, 0x1p52));   // subnormal
        ix -= 52ULL << 52;
        hi = (int)(ix >> 32);
        lo = (int)(unsigned)ix;
    }
    return kp_log_main(hi, lo, tab, K);
}

// out-of-line copy for the rare arguments the scoring kernel does not handle inline
__device__ __noinline__ double kp_log_slow(double x, const double2 *tab) { return kp_log(x, tab); }

// cephes log1p (scipy.special): log(1+x) by a 6/6 rational on [sqrt(1/2)-1, sqrt(2)-1], no fused ops
__device__ __forceinline__ double kp_log1p(double x, const double2 *__restrict__ tab)
{
    double z = KP_ADD(1.0, x);
    if (z < 0.70710678118654752440 || z > 1.41421356237309504880) return kp_log(z, tab);
    const double LP[7] = {4.5270000862445199635215E-5, 4.9854102823193375972212E-1, 6.5787325942061044846969E0,
                          2.9911919328553073277375E1,  6.0949667980987787057556E1,  5.7112963590585538103336E1,
                          2.0039553499201281259648E1};
    const double LQ[6] = {1.5062909083469192043167E1, 8.3047565967967209469434E1, 2.2176239823732856465394E2,
                          3.0909872225312059774938E2, 2.1642788614495947685003E2, 6.0118660497603843919306E1};
    z = KP_MUL(x, x);
    double num = LP[0];
#pragma unroll
    for (int i = 1; i <= 6; i++) num = KP_ADD(KP_MUL(num, x), LP[i]);
    double den = KP_ADD(x, LQ[0]);
#pragma unroll
    for (int i = 1; i < 6; i++) den = KP_ADD(KP_MUL(den, x), LQ[i]);
    z = KP_ADD(KP_MUL(-0.5, z), KP_MUL(x, __ddiv_rn(KP_MUL(z, num), den)));
    return KP_ADD(x, z);
}

__device__ __forceinline__ double kp_xlogy(double x, double y, const double2 *tab)
{
    return (x == 0.0 && !(y != y)) ? 0.0 : KP_MUL(x, kp_log(y, tab));
}
__device__ __forceinline__ double kp_xlog1py(double x, double y, const double2 *tab)
{
    return (x == 0.0 && !(y != y)) ? 0.0 : KP_MUL(x, kp_log1p(y, tab));
}

// p = (M + alpha) / (M + U + alpha + beta), float64, left to right (w_numba.py:56, _CV.py:60)
__device__ __forceinline__ double kp_rate(unsigned long long M, unsigned long long U, double alpha, double beta)
{
    double num = KP_ADD((double)M, alpha);
    double den = KP_ADD(KP_ADD((double)(M + U), alpha), beta);
    return __ddiv_rn(num, den);
}

// level >= 1 self-score: s = penalty (+ (-2M) log p) (+ (-2U) log(1-p))   (w_numba.py:57-61)
// also returns log p and log(1-p) for the CV held-out term (_CV.py:61-78)
__device__ __forceinline__ double kp_self_score(unsigned long long M, unsigned long long U, double alpha, double beta,
                                                double penalty, const double2 *tab, double &logp, double &log1mp)
{
    double p = kp_rate(M, U, alpha, beta);
    logp = kp_log(p, tab);
    log1mp = kp_log(KP_SUB(1.0, p), tab);
    double s = penalty;
    if (M > 0) s = KP_ADD(s, KP_MUL(KP_MUL(-2.0, (double)M), logp));
    if (U > 0) s = KP_ADD(s, KP_MUL(KP_MUL(-2.0, (double)U), log1mp));
    return s;
}

// held-out -2 log-lik of a pattern kept whole (_CV.py:73-78)
__device__ __forceinline__ double kp_test_ll(unsigned long long Mt, unsigned long long Ut, double logp, double log1mp)
{
    double t = 0.0;
    if (Mt > 0) t = KP_ADD(t, KP_MUL(KP_MUL(-2.0, (double)Mt), logp));
    if (Ut > 0) t = KP_ADD(t, KP_MUL(KP_MUL(-2.0, (double)Ut), log1mp));
    return t;
}

// level 0 (k-mers): -2 (xlogy(M,p) + xlog1py(U,-p)) + penalty   (w_numba.py:26-29)
__device__ __forceinline__ double kp_leaf_score(unsigned long long M, unsigned long long U, double alpha, double beta,
                                                double penalty, const double2 *tab)
{
    double p = kp_rate(M, U, alpha, beta);
    double a = kp_xlogy((double)M, p, tab);
    double b = kp_xlog1py((double)U, -p, tab);
    return KP_ADD(KP_MUL(-2.0, KP_ADD(a, b)), penalty);
}

// level 0 of a CV fold (_CV.py:15-20): train score from train counts, held-out LL from held-out counts
__device__ __forceinline__ void kp_leaf_cv(unsigned long long Mtr, unsigned long long Utr, unsigned long long Mte,
                                           unsigned long long Ute, double alpha, double beta, double penalty,
                                           const double2 *tab, double &train, double &test)
{
    double p = kp_rate(Mtr, Utr, alpha, beta);
    double np_ = -p;
    train = KP_ADD(KP_MUL(-2.0, KP_ADD(kp_xlogy((double)Mtr, p, tab), kp_xlog1py((double)Utr, np_, tab))), penalty);
    test = KP_MUL(-2.0, KP_ADD(kp_xlogy((double)Mte, p, tab), kp_xlog1py((double)Ute, np_, tab)));
}
