// Per-marker scalar sampling: the part of BayesRBase::processColumn between computeNum and the
// residual update (reference src/BayesRBase.cpp:306-366), restated for one GPU thread.
//
// Everything that depends only on (group, component) and the sweep's sigmaE / sigmaG / pi is
// hoisted into GroupConsts once per sweep (the reference recomputes it per marker, :312-317,
// :329-333); per marker only muk, logL and the exp() sums remain.
#pragma once
#include "common.cuh"

namespace brr {

// shared-memory table, index [g * K + k]
struct GroupConsts {
    double logPi[kMaxGK];    // log(pi[g][k])                                   (:330)
    double denom[kMaxGK];    // (N-1) + sigmaE/sigmaG[g] * (1/cva[g][k-1])  k>=1 (:316-317)
    double sdev[kMaxGK];     // sqrt(sigmaE / denom)                        k>=1 (:352)
    double halfLog[kMaxGK];  // 0.5 * log(sigmaG/sigmaE*(N-1)*cva + 1)      k>=1 (:329,332)
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
            gc.sdev[i] = sqrt(sigmaE / denom);
            gc.halfLog[i] = 0.5 * log(logLScale * c + 1.0);
        } else {
            gc.denom[i] = 1.0; gc.sdev[i] = 0.0; gc.halfLog[i] = 0.0;
        }
    }
}

struct Draw {
    double beta_new;
    double acum0;   // P(component 0) as stored in m_acum (:341)
    int comp;       // chosen component, -1 = walk fell through (nothing assigned, :346-366)
};

// src/BayesRBase.cpp:325-366.  KT > 0: compile-time component count (arrays stay in registers).
template <int KT>
__device__ __forceinline__ Draw sample_marker_t(const GroupConsts &gc, int g, int Krt, double num, double sigmaE,
                                                double beta_old, double u, double z)
{
    const int K = KT > 0 ? KT : Krt;
    const double *logPi = gc.logPi + g * K;
    const double *denom = gc.denom + g * K;
    const double *sdev = gc.sdev + g * K;
    const double *halfLog = gc.halfLog + g * K;
    double logL[KT > 0 ? KT : kMaxK], muk[KT > 0 ? KT : kMaxK];
    muk[0] = 0.0;
    logL[0] = logPi[0];
    bool big = false;
#pragma unroll
    for (int k = 1; k < K; ++k) {
        muk[k] = num / denom[k];                                            // :325
        logL[k] = logPi[k] - halfLog[k] + 0.5 * (muk[k] * num) / sigmaE;   // :331-333
        big |= fabs(logL[k] - logL[0]) > 700.0;                             // :335
    }
    double acum = 0.0;
    if (!big) {
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < K; ++l) s += exp(logL[l] - logL[0]);
        acum = 1.0 / s;                                                     // :338
    }
    Draw d;
    d.acum0 = acum;
    d.beta_new = beta_old;
    d.comp = -1;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (u <= acum) {                                                    // :347
            d.beta_new = (k == 0) ? 0.0 : z * sdev[k] + muk[k];             // :350-352
            d.comp = k;
            break;
        }
        if (k + 1 == K) break;     // reference reads logL[K] out of bounds; result unused
        bool b2 = false;
#pragma unroll
        for (int l = 1; l < K; ++l) b2 |= fabs(logL[l] - logL[k + 1]) > 700.0;   // :360
        if (!b2) {
            double s = 0.0;
#pragma unroll
            for (int l = 0; l < K; ++l) s += exp(logL[l] - logL[k + 1]);
            acum += 1.0 / s;                                                // :363
        }
    }
    return d;
}

__device__ inline Draw sample_marker(const GroupConsts &gc, int g, int K, double num, double sigmaE,
                                     double beta_old, double u, double z)
{
    return sample_marker_t<0>(gc, g, K, num, sigmaE, beta_old, u, z);
}

// dispatch on the common component counts so that muk/logL live in registers
__device__ __forceinline__ Draw sample_marker_k(const GroupConsts &gc, int g, int K, double num, double sigmaE,
                                                double beta_old, double u, double z)
{
    switch (K) {
    case 2: return sample_marker_t<2>(gc, g, K, num, sigmaE, beta_old, u, z);
    case 3: return sample_marker_t<3>(gc, g, K, num, sigmaE, beta_old, u, z);
    case 4: return sample_marker_t<4>(gc, g, K, num, sigmaE, beta_old, u, z);
    case 5: return sample_marker_t<5>(gc, g, K, num, sigmaE, beta_old, u, z);
    default: return sample_marker_t<0>(gc, g, K, num, sigmaE, beta_old, u, z);
    }
}

} // namespace brr
