// Per-marker scalar sampling: the part of BayesRBase::processColumn between computeNum and the
// residual update (reference src/BayesRBase.cpp:306-366), restated for one GPU thread.
//
// Everything that depends only on (group, component) and the sweep's sigmaE / sigmaG / pi is
// hoisted into GroupConsts once per sweep (the reference recomputes it per marker, :312-317,
// :329-333); per marker only muk, logL and the exp() sums remain.
#pragma once
#include "common.cuh"
#include <cstring>

namespace brr {

// shared-memory table, index [g * K + k]
struct GroupConsts {
    double logPi[kMaxGK];    // log(pi[g][k])                                   (:330)
    double denom[kMaxGK];    // (N-1) + sigmaE/sigmaG[g] * (1/cva[g][k-1])  k>=1 (:316-317)
    double sdev[kMaxGK];     // sqrt(sigmaE / denom)                        k>=1 (:352)
    double halfLog[kMaxGK];  // 0.5 * log(sigmaG/sigmaE*(N-1)*cva + 1)      k>=1 (:329,332)
    double invDenom[kMaxGK]; // 1 / denom: the walk multiplies instead of dividing (1 ulp, see sample_marker_fast_t)
    double halfInvSigmaE;    // 0.5 / sigmaE
};

// called by all threads of a CTA; every CTA computes the identical table
__device__ inline void fill_group_consts(GroupConsts &gc, int nG, int K, double NM1, double sigmaE,
                                         const double *__restrict__ pi, const double *__restrict__ sigmaG,
                                         const double *__restrict__ cva)
{
    for (int i = threadIdx.x; i < nG * K; i += blockDim.x) {
        const int g = i / K, k = i % K;
        gc.logPi[i] = log(pi[i]);
        if (k > 0) {
            const double sG = sigmaG[g];
            const double c = cva[g * (K - 1) + (k - 1)];
            const double sigmaEOverSigmaG = sigmaE / sG;
            const double denom = NM1 + sigmaEOverSigmaG * (1.0 / c);
            const double logLScale = sG / sigmaE * NM1;
            gc.denom[i] = denom;
            gc.invDenom[i] = 1.0 / denom;
            gc.sdev[i] = sqrt(sigmaE / denom);
            gc.halfLog[i] = 0.5 * log(logLScale * c + 1.0);
        } else {
            gc.denom[i] = 1.0; gc.invDenom[i] = 1.0; gc.sdev[i] = 0.0; gc.halfLog[i] = 0.0;
        }
    }
    if (threadIdx.x == 0) gc.halfInvSigmaE = 0.5 / sigmaE;
}

struct Draw {
    double beta_new;
    double acum0;   // P(component 0) as stored in m_acum (:341)
    int comp;       // chosen component, -1 = walk fell through (nothing assigned, :346-366)
};

// src/BayesRBase.cpp:325-366, the literal restatement (up to K*K exponentials).  This is the rarely taken general path
// (K > 5, or a |logL| difference beyond the reference's 700 guard): one out-of-line copy, loops not unrolled, so that it
// costs no instruction-cache space on the walkers' critical path.
__device__ __noinline__ Draw sample_marker_general(const GroupConsts &gc, int g, int K, double num, double sigmaE,
                                                   double beta_old, double u, double z)
{
    const double *logPi = gc.logPi + g * K;
    const double *invDenom = gc.invDenom + g * K;
    const double *sdev = gc.sdev + g * K;
    const double *halfLog = gc.halfLog + g * K;
    double logL[kMaxK], muk[kMaxK];
    muk[0] = 0.0;
    logL[0] = logPi[0];
    bool big = false;
#pragma unroll 1
    for (int k = 1; k < K; ++k) {
        muk[k] = num * invDenom[k];                                         // :325 (same muk as the fast path and the walk's solve)
        logL[k] = logPi[k] - halfLog[k] + 0.5 * (muk[k] * num) / sigmaE;   // :331-333
        big |= fabs(logL[k] - logL[0]) > 700.0;                             // :335
    }
    double acum = 0.0;
    if (!big) {
        double s = 0.0;
#pragma unroll 1
        for (int l = 0; l < K; ++l) s += exp(logL[l] - logL[0]);
        acum = 1.0 / s;                                                     // :338
    }
    Draw d;
    d.acum0 = acum;
    d.beta_new = beta_old;
    d.comp = -1;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        if (u <= acum) {                                                    // :347
            d.beta_new = (k == 0) ? 0.0 : __fma_rn(z, sdev[k], muk[k]);     // :350-352
            d.comp = k;
            break;
        }
        if (k + 1 == K) break;     // reference reads logL[K] out of bounds; result unused
        bool b2 = false;
#pragma unroll 1
        for (int l = 1; l < K; ++l) b2 |= fabs(logL[l] - logL[k + 1]) > 700.0;   // :360
        if (!b2) {
            double s = 0.0;
#pragma unroll 1
            for (int l = 0; l < K; ++l) s += exp(logL[l] - logL[k + 1]);
            acum += 1.0 / s;                                                // :363
        }
    }
    return d;
}

__device__ inline Draw sample_marker(const GroupConsts &gc, int g, int K, double num, double sigmaE,
                                     double beta_old, double u, double z)
{
    return sample_marker_general(gc, g, K, num, sigmaE, beta_old, u, z);
}

// Branch-free exp and reciprocal for the walk (arguments within the reference's 700 guard): the library exp() carries a
// special-case branch and a 13-deep Horner chain, which keeps the exponentials of one marker -- and of the two markers a walker
// lane samples -- from overlapping; here the polynomial is evaluated by Estrin's scheme (depth 4) in straight-line code, so six
// exponentials cost about as much latency as one.  |relative error| < 4e-16 (host-checked against libm in tests/test_exp_nb.py).
__host__ __device__ __forceinline__ double exp_nb(double x)
{
    const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);        // rint(x log2 e) in the low mantissa bits
    const double fn = t - 6755399441055744.0;
    double r = fma(fn, -6.93147180369123816490e-01, x);
    r = fma(fn, -1.90821492927058770002e-10, r);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = 1.0 + r;
    const double a1 = fma(r, 1.0 / 6.0, 0.5);
    const double a2 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    const double a3 = fma(r, 1.0 / 5040.0, 1.0 / 720.0);
    const double a4 = fma(r, 1.0 / 362880.0, 1.0 / 40320.0);
    const double a5 = fma(r, 1.0 / 39916800.0, 1.0 / 3628800.0);
    const double a6 = fma(r, 1.0 / 6227020800.0, 1.0 / 479001600.0);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    const double d0 = fma(b1, r4, b0), d1 = fma(a6, r4, b2);
    const double pz = fma(d1, r8, d0);
#ifdef __CUDA_ARCH__
    const int n = __double2loint(t);
    return __hiloint2double(__double2hiint(pz) + (n << 20), __double2loint(pz));
#else
    long long bits, tb;
    memcpy(&tb, &t, 8); memcpy(&bits, &pz, 8);
    bits += (long long)(int)(unsigned)(tb & 0xffffffffLL) << 52;
    double out; memcpy(&out, &bits, 8);
    return out;
#endif
}
__device__ __forceinline__ double rcp_nb(double s)      // 1 / s for a normal positive s, two Newton steps on the hardware approximation
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    double e = fma(-s, y, 1.0);
    y = fma(y, e, y);
    e = fma(-s, y, 1.0);
    y = fma(y, e, y);
    e = fma(-s, y, 1.0);
    return fma(y, e, y);
}

// The same draw with K exponentials instead of up to K*K: when no |logL_a - logL_b| can exceed the
// reference's 700 guard (:335, :360), 1/sum_l exp(logL_l - logL_k) == exp(logL_k - logL_0) * acum0 up to
// rounding, so the cumulative probabilities of components k >= 1 reuse the first pass.  acum0 (what the
// reference stores in m_acum) is computed by the reference's own expression and is bit-identical; the
// guarded (huge |num|) case falls back to the literal restatement above.  The two divisions per
// component (num/denom, ./sigmaE) are multiplications by reciprocals hoisted into GroupConsts: the chain is
// a latency-bound serial section, and a 1-ulp difference in muk / logL is far inside the 1e-8 replay gate.
template <int KT>
__device__ __forceinline__ Draw sample_marker_fast_t(const GroupConsts &gc, int g, int Krt, double num, double sigmaE,
                                                     double beta_old, double u, double z)
{
    static_assert(KT >= 2, "compile-time component count required");
    constexpr int K = KT;
    const double *logPi = gc.logPi + g * K;
    const double *invDenom = gc.invDenom + g * K;
    const double *sdev = gc.sdev + g * K;
    const double *halfLog = gc.halfLog + g * K;
    const double hise = gc.halfInvSigmaE;
    double logL[K], muk[K], ex[K];
    muk[0] = 0.0;
    logL[0] = logPi[0];
    double mx = logL[0], mn = logL[0];
#pragma unroll
    for (int k = 1; k < K; ++k) {
        muk[k] = num * invDenom[k];                                         // :325 (num / denom, to 1 ulp)
        logL[k] = logPi[k] - halfLog[k] + (muk[k] * num) * hise;           // :331-333
        mx = fmax(mx, logL[k]); mn = fmin(mn, logL[k]);
    }
    if (!(mx - mn <= 700.0)) return sample_marker_general(gc, g, Krt, num, sigmaE, beta_old, u, z);
    double s = 1.0;                                                          // exp(logL[0] - logL[0])
    ex[0] = 1.0;
#pragma unroll
    for (int l = 1; l < K; ++l) ex[l] = exp_nb(logL[l] - logL[0]);
#pragma unroll
    for (int l = 1; l < K; ++l) s += ex[l];
    double acum = rcp_nb(s);                                                // :338
    Draw d;
    d.acum0 = acum;
    d.beta_new = beta_old;
    d.comp = -1;
    const double p0 = acum;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (u <= acum) {                                                    // :347
            d.beta_new = (k == 0) ? 0.0 : __fma_rn(z, sdev[k], muk[k]);     // :350-352
            d.comp = k;
            break;
        }
        if (k + 1 == K) break;
        acum += ex[k + 1] * p0;                                             // :363
    }
    return d;
}

// dispatch on the common component counts so that muk/logL live in registers; ONE out-of-line copy for all call sites
__device__ __noinline__ Draw sample_marker_k(const GroupConsts &gc, int g, int K, double num, double sigmaE,
                                                double beta_old, double u, double z)
{
    switch (K) {
    case 2: return sample_marker_fast_t<2>(gc, g, K, num, sigmaE, beta_old, u, z);
    case 3: return sample_marker_fast_t<3>(gc, g, K, num, sigmaE, beta_old, u, z);
    case 4: return sample_marker_fast_t<4>(gc, g, K, num, sigmaE, beta_old, u, z);
    case 5: return sample_marker_fast_t<5>(gc, g, K, num, sigmaE, beta_old, u, z);
    default: return sample_marker_general(gc, g, K, num, sigmaE, beta_old, u, z);
    }
}

// Two markers sampled by ONE thread with their dependent chains interleaved (the walk is latency-bound: two independent fp64
// chains cost little more than one).  Same arithmetic, operation by operation, as sample_marker_fast_t; a marker whose |logL|
// spread exceeds the reference's 700 guard is redone by the literal restatement.
template <int KT>
__device__ __forceinline__ void sample_pair_fast_t(const GroupConsts &gc, const int (&g)[2], int Krt, const double (&num)[2], double sigmaE,
                                                   const double (&beta_old)[2], const double (&u)[2], const double (&z)[2], Draw (&d)[2])
{
    constexpr int K = KT;
    const double hise = gc.halfInvSigmaE;
    double logL[2][K], muk[2][K], ex[2][K], mx[2], mn[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) { muk[m][0] = 0.0; logL[m][0] = gc.logPi[g[m] * K]; mx[m] = logL[m][0]; mn[m] = logL[m][0]; }
#pragma unroll
    for (int k = 1; k < K; ++k)
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            muk[m][k] = num[m] * gc.invDenom[g[m] * K + k];
            logL[m][k] = gc.logPi[g[m] * K + k] - gc.halfLog[g[m] * K + k] + (muk[m][k] * num[m]) * hise;
            mx[m] = fmax(mx[m], logL[m][k]); mn[m] = fmin(mn[m], logL[m][k]);
        }
#pragma unroll
    for (int l = 1; l < K; ++l)
#pragma unroll
        for (int m = 0; m < 2; ++m) ex[m][l] = exp_nb(logL[m][l] - logL[m][0]);
    double s[2] = {1.0, 1.0};
#pragma unroll
    for (int l = 1; l < K; ++l)
#pragma unroll
        for (int m = 0; m < 2; ++m) s[m] += ex[m][l];
    double p0[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) p0[m] = rcp_nb(s[m]);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        double acum = p0[m];
        d[m].acum0 = acum; d[m].beta_new = beta_old[m]; d[m].comp = -1;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (u[m] <= acum) {
                d[m].beta_new = (k == 0) ? 0.0 : __fma_rn(z[m], gc.sdev[g[m] * K + k], muk[m][k]);
                d[m].comp = k;
                break;
            }
            if (k + 1 == K) break;
            acum += ex[m][k + 1] * p0[m];
        }
        if (!(mx[m] - mn[m] <= 700.0)) d[m] = sample_marker_general(gc, g[m], Krt, num[m], sigmaE, beta_old[m], u[m], z[m]);
    }
}

// ONE out-of-line copy for all call sites (code size: the sweep kernels are instruction-cache bound), scalars in registers
struct DrawPair { double beta0, beta1, acum0, acum1; int comp0, comp1; };
__device__ __noinline__ DrawPair sample_pair_k(const GroupConsts &gc, int g0, int g1, int K, double num0, double num1, double sigmaE,
                                               double bo0, double bo1, double u0, double u1, double z0, double z1)
{
    const int g[2] = {g0, g1};
    const double num[2] = {num0, num1}, beta_old[2] = {bo0, bo1}, u[2] = {u0, u1}, z[2] = {z0, z1};
    Draw d[2];
    switch (K) {
    case 2: sample_pair_fast_t<2>(gc, g, K, num, sigmaE, beta_old, u, z, d); break;
    case 3: sample_pair_fast_t<3>(gc, g, K, num, sigmaE, beta_old, u, z, d); break;
    case 4: sample_pair_fast_t<4>(gc, g, K, num, sigmaE, beta_old, u, z, d); break;
    case 5: sample_pair_fast_t<5>(gc, g, K, num, sigmaE, beta_old, u, z, d); break;
    default:
        d[0] = sample_marker_general(gc, g[0], K, num[0], sigmaE, beta_old[0], u[0], z[0]);
        d[1] = sample_marker_general(gc, g[1], K, num[1], sigmaE, beta_old[1], u[1], z[1]);
    }
    DrawPair r;
    r.beta0 = d[0].beta_new; r.beta1 = d[1].beta_new; r.acum0 = d[0].acum0; r.acum1 = d[1].acum0; r.comp0 = d[0].comp; r.comp1 = d[1].comp;
    return r;
}

// beta of a marker whose component is already known (the walk's triangular solve), :350-352
__device__ __forceinline__ double beta_given_comp(const GroupConsts &gc, int g, int K, int comp, double num,
                                                  double beta_old, double z)
{
    if (comp < 0) return beta_old;
    if (comp == 0) return 0.0;
    return __fma_rn(z, gc.sdev[g * K + comp], num * gc.invDenom[g * K + comp]);
}

} // namespace brr
