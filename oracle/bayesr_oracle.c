/* TEST INFRASTRUCTURE ONLY -- see bayesr_oracle.h.  Plain C restatement of the reference's BayesR
 * hot path; every function cites the reference file:line it follows.  Nothing here is shipped,
 * measured as the product, or reachable from the product path. */
#include "bayesr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* src/preprocessgraph.cpp:44-53: bitset<8> b = ch; allele1 = !b[k++]; allele2 = !b[k++];
 * src/densemarkerbuilder.cpp:30, src/sparsemarker.cpp:10: (allele1,allele2)==(0,1) is missing. */
void orc_decode_bed_column(const uint8_t *bedcol, uint32_t N, uint8_t *cls)
{
    for (uint32_t i = 0; i < N; ++i) {
        const unsigned byte = bedcol[i >> 2];
        const unsigned k = (i & 3u) * 2u;
        const unsigned a1 = !((byte >> k) & 1u);
        const unsigned a2 = !((byte >> (k + 1u)) & 1u);
        cls[i] = (a1 == 0 && a2 == 1) ? ORC_MISSING : (uint8_t)(a1 + a2);
    }
}

/* src/densemarkerbuilder.cpp:23-60 */
void orc_dense_column(const uint8_t *bedcol, uint32_t N, double *cx)
{
    uint8_t *cls = (uint8_t *)malloc(N);
    orc_decode_bed_column(bedcol, N, cls);
    double sum = 0.0;      /* m_sum, :39 */
    uint32_t nmiss = 0;
    for (uint32_t i = 0; i < N; ++i) {
        if (cls[i] == ORC_MISSING) { ++nmiss; cx[i] = 0.0; }
        else { cx[i] = (double)cls[i]; sum += (double)cls[i]; }
    }
    const double mean = sum / ((double)N - (double)nmiss);            /* :50 */
    for (uint32_t i = 0; i < N; ++i) if (cls[i] == ORC_MISSING) cx[i] = mean;   /* :51-53 */
    double s = 0.0;                                                    /* snpData.mean(), :56 */
    for (uint32_t i = 0; i < N; ++i) s += cx[i];
    const double mean2 = s / (double)N;
    double sqn = 0.0;
    for (uint32_t i = 0; i < N; ++i) { cx[i] -= mean2; sqn += cx[i] * cx[i]; } /* :56-57 */
    const double sigma = 1.0 / sqrt(sqn / ((double)N - 1.0));          /* :58 */
    for (uint32_t i = 0; i < N; ++i) cx[i] *= sigma;                   /* :59 */
    free(cls);
}

void orc_dense_lut(const uint8_t *bedcol, uint32_t N, double lut[4])
{
    double *cx = (double *)malloc(sizeof(double) * N);
    uint8_t *cls = (uint8_t *)malloc(N);
    orc_dense_column(bedcol, N, cx);
    orc_decode_bed_column(bedcol, N, cls);
    int seen[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < N; ++i) if (!seen[cls[i]]) { lut[cls[i]] = cx[i]; seen[cls[i]] = 1; }
    /* classes that do not occur: give them the value the builder WOULD have produced, from one that does */
    int ref = -1;
    for (int c = 0; c < 3; ++c) if (seen[c]) { ref = c; break; }
    if (ref >= 0) {
        /* values are affine in the class for non-missing: lut[c] = (c - mean2)*sigma */
        double sigma = 0.0; int a = -1, b = -1;
        for (int c = 0; c < 3; ++c) if (seen[c]) { if (a < 0) a = c; else if (b < 0) b = c; }
        if (b >= 0) sigma = (lut[b] - lut[a]) / (double)(b - a);
        for (int c = 0; c < 3; ++c) if (!seen[c]) lut[c] = lut[ref] + sigma * (double)(c - ref);
        if (!seen[3]) lut[3] = 0.0;
    } else {
        for (int c = 0; c < 4; ++c) if (!seen[c]) lut[c] = 0.0;
    }
    free(cx); free(cls);
}

/* src/raggedsparsemarkerbuilder.cpp:25-55, src/sparsemarker.cpp:8-19 */
void orc_ragged_marker(const uint8_t *bedcol, uint32_t N, double stats[4], int32_t *ones, uint32_t *n1,
                       int32_t *twos, uint32_t *n2, int32_t *miss, uint32_t *nm)
{
    uint8_t *cls = (uint8_t *)malloc(N);
    orc_decode_bed_column(bedcol, N, cls);
    double mean = 0.0, sqrdZ = 0.0, Zsum = 0.0;
    *n1 = *n2 = *nm = 0;
    for (uint32_t i = 0; i < N; ++i) {
        const unsigned c = cls[i];
        if (c == 1 || c == 2) {                      /* updateStatistics, sparsemarker.cpp:12-18 */
            const double value = (double)c;
            mean += value; sqrdZ += value * value; Zsum += value;
        }
        if (c == ORC_MISSING) miss[(*nm)++] = (int32_t)i;        /* :34-35 */
        else if (c == 1) ones[(*n1)++] = (int32_t)i;            /* :36-37 */
        else if (c == 2) twos[(*n2)++] = (int32_t)i;            /* :38-39 */
    }
    const double n = (double)N;
    mean /= n;                                                               /* :49 */
    const double sd = sqrt((sqrdZ - 2.0 * mean * Zsum + n * mean * mean) / (n - 1.0));   /* :53-54 */
    stats[0] = mean; stats[1] = sd; stats[2] = sqrdZ; stats[3] = Zsum;
    free(cls);
}

/* src/eigensparsemarkerbuilder.cpp:21-57 */
void orc_eigen_stats(const uint8_t *bedcol, uint32_t N, double stats[4])
{
    uint8_t *cls = (uint8_t *)malloc(N);
    orc_decode_bed_column(bedcol, N, cls);
    double mean = 0.0, sqrdZ = 0.0, Zsum = 0.0, nmiss = 0.0;
    for (uint32_t i = 0; i < N; ++i) {
        const unsigned c = cls[i];
        if (c == 1 || c == 2) { const double v = (double)c; mean += v; sqrdZ += v * v; Zsum += v; }
        if (c == ORC_MISSING) nmiss += 1.0;
    }
    const double n = (double)N;
    mean /= n;                                     /* :46 */
    sqrdZ += (mean * mean) * nmiss;                /* :50 */
    Zsum += mean * nmiss;                          /* :53 */
    const double sd = sqrt((sqrdZ - 2.0 * mean * Zsum + n * mean * mean) / (n - 1.0));  /* :56-57 */
    stats[0] = mean; stats[1] = sd; stats[2] = sqrdZ; stats[3] = Zsum;
    free(cls);
}

/* src/densebayesrkernel.cpp:10-15 */
double orc_dense_compute_num(const double *cx, const double *eps, uint32_t N, double beta_old)
{
    double dot = 0.0;
    for (uint32_t i = 0; i < N; ++i) dot += cx[i] * eps[i];
    return dot + beta_old * (double)(N - 1);
}

/* src/densebayesrkernel.cpp:17-20, src/BayesRBase.cpp:376 */
void orc_dense_eps_update(const double *cx, double *eps, uint32_t N, double beta_old, double beta)
{
    const double d = beta_old - beta;
    for (uint32_t i = 0; i < N; ++i) eps[i] += d * cx[i];
}

/* src/sparsebayesrkernel.cpp:24-27, src/raggedbayesrkernel.cpp:32-35 */
double orc_ragged_compute_num(const double stats[4], const int32_t *ones, uint32_t n1, const int32_t *twos,
                              uint32_t n2, const double *eps, uint32_t N, double beta_old, double eps_sum)
{
    const double mean = stats[0], sd = stats[1];
    double s1 = 0.0, s2 = 0.0;
    for (uint32_t i = 0; i < n1; ++i) s1 += eps[ones[i]];
    for (uint32_t i = 0; i < n2; ++i) s2 += eps[twos[i]];
    const double dot = (s1 + 2 * s2) / sd;
    return beta_old * ((double)N - 1.0) - mean * eps_sum / sd + dot;
}

/* src/raggedbayesrkernel.cpp:10-30, src/sparsebayesrkernel.cpp:29-34, src/SparseBayesRRG.cpp:84-92 */
void orc_ragged_eps_update(const double stats[4], const int32_t *ones, uint32_t n1, const int32_t *twos,
                           uint32_t n2, const int32_t *miss, uint32_t nm, double *eps, uint32_t N,
                           double beta_old, double beta, double *eps_sum)
{
    const double mean = stats[0], sd = stats[1], Zsum = stats[3];
    const double dBeta = beta_old - beta;
    const double meanAdjustment = dBeta * mean / sd;
    const double oneAdjustment = dBeta / sd;
    /* the reference builds delta = Constant(-meanAdjustment), scatters into it, then eps += delta */
    double *delta = (double *)malloc(sizeof(double) * N);
    for (uint32_t i = 0; i < N; ++i) delta[i] = -meanAdjustment;
    for (uint32_t i = 0; i < n1; ++i) delta[ones[i]] += oneAdjustment;
    for (uint32_t i = 0; i < n2; ++i) delta[twos[i]] += 2 * oneAdjustment;
    for (uint32_t i = 0; i < nm; ++i) delta[miss[i]] += meanAdjustment;
    for (uint32_t i = 0; i < N; ++i) eps[i] += delta[i];
    free(delta);
    *eps_sum += dBeta * Zsum / sd - dBeta * mean * (double)N / sd;
}

/* src/eigenbayesrkernel.cpp:17-20 with src/sparsebayesrkernel.cpp:24-27 */
double orc_eigen_compute_num(const double stats[4], const uint8_t *cls, const double *eps, uint32_t N,
                             double beta_old, double eps_sum)
{
    const double mean = stats[0], sd = stats[1];
    double dot = 0.0;
    for (uint32_t i = 0; i < N; ++i) {
        if (cls[i] == 1) dot += eps[i];
        else if (cls[i] == 2) dot += 2.0 * eps[i];
        else if (cls[i] == ORC_MISSING) dot += mean * eps[i];
    }
    return beta_old * ((double)N - 1.0) - mean * eps_sum / sd + dot / sd;
}

/* src/eigenbayesrkernel.cpp:10-15 */
void orc_eigen_eps_update(const double stats[4], const uint8_t *cls, double *eps, uint32_t N, double beta_old,
                          double beta, double *eps_sum)
{
    const double mean = stats[0], sd = stats[1], Zsum = stats[3];
    const double dBeta = beta_old - beta;
    for (uint32_t i = 0; i < N; ++i) {
        double zg = 0.0;
        if (cls[i] == 1) zg = 1.0; else if (cls[i] == 2) zg = 2.0; else if (cls[i] == ORC_MISSING) zg = mean;
        eps[i] += dBeta * zg / sd - dBeta * mean / sd * 1.0;
    }
    *eps_sum += dBeta * Zsum / sd - dBeta * mean * (double)N / sd;
}

#define ORC_MAXK 64

/* src/BayesRBase.cpp:306-366 */
int orc_sample_marker(const orc_model *m, int group, double num, double beta_old, double u, double z,
                      double *beta_new, double *acum_out)
{
    const int K = (int)m->K, km1 = K - 1;
    const double NM1 = (double)(m->N - 1);
    const double sigmaG = m->sigmaG[group], sigmaE = m->sigmaE;
    const double *cva = m->cva + (size_t)group * km1;
    const double *pi = m->pi + (size_t)group * K;
    double denom[ORC_MAXK], muk[ORC_MAXK], logL[ORC_MAXK + 1];

    const double sigmaEOverSigmaG = sigmaE / sigmaG;                        /* :316 */
    for (int k = 0; k < km1; ++k) denom[k] = NM1 + sigmaEOverSigmaG * (1.0 / cva[k]);   /* :313,317 */
    muk[0] = 0.0;
    for (int k = 1; k < K; ++k) muk[k] = num / denom[k - 1];                /* :325 */
    const double logLScale = sigmaG / sigmaE * NM1;                         /* :329 */
    logL[0] = log(pi[0]);                                                   /* :330 */
    for (int k = 1; k < K; ++k)                                             /* :331-333 */
        logL[k] = log(pi[k]) - 0.5 * log(logLScale * cva[k - 1] + 1.0) + 0.5 * (muk[k] * num) / sigmaE;

    double acum;
    int big = 0;
    for (int k = 1; k < K; ++k) if (fabs(logL[k] - logL[0]) > 700.0) big = 1;    /* :335 */
    if (big) acum = 0.0;
    else { double s = 0.0; for (int l = 0; l < K; ++l) s += exp(logL[l] - logL[0]); acum = 1.0 / s; }  /* :338 */
    *acum_out = acum;                                                       /* :341 */

    *beta_new = beta_old;
    for (int k = 0; k < K; ++k) {                                           /* :346-366 */
        if (u <= acum) {
            if (k == 0) *beta_new = 0.0;
            else *beta_new = z * sqrt(sigmaE / denom[k - 1]) + muk[k];
            return k;
        }
        if (k + 1 == K) break;   /* the reference reads logL[K] out of bounds here; its value only feeds
                                    an acum that is never looked at again, so nothing is assigned */
        big = 0;
        for (int l = 1; l < K; ++l) if (fabs(logL[l] - logL[k + 1]) > 700.0) big = 1;   /* :360 */
        if (!big) { double s = 0.0; for (int l = 0; l < K; ++l) s += exp(logL[l] - logL[k + 1]); acum += 1.0 / s; }
    }
    return -1;
}

typedef struct { double beta_old, beta_new; int comp; double acum; } orc_draw;

static void orc_commit(const orc_model *m, int g, uint32_t j, const orc_draw *d, int write_acum,
                       double *beta, double *comp, double *acum, double *v, double *betasq)
{
    if (write_acum) acum[j] = d->acum;                                      /* src/BayesRBase.cpp:341 */
    if (d->comp >= 0) {
        beta[j] = d->beta_new;                                              /* :350-352 */
        if (d->comp > 0) betasq[g] += d->beta_new * d->beta_new;            /* :353 (pow(beta,2)) */
        v[(size_t)g * m->K + d->comp] += 1.0;                               /* :355 */
        comp[j] = (double)d->comp;                                          /* :356 */
    }
}

int orc_sweep_dense(const uint8_t *bed, uint32_t M, const uint32_t *order, uint32_t n_order,
                    const double *rand_uz, const orc_model *m, const int32_t *groups, uint32_t tokens,
                    double *eps, double *beta, double *comp, double *acum, double *v, double *betasq)
{
    const uint32_t N = m->N;
    const size_t cb = ((size_t)N + 3) / 4;
    const uint32_t T = tokens ? tokens : 1;
    double *cx = (double *)malloc(sizeof(double) * N * T);
    orc_draw *draws = (orc_draw *)malloc(sizeof(orc_draw) * T);
    if (!cx || !draws) return 1;
    for (uint32_t p = 0; p < n_order; p += T) {
        const uint32_t nb = (n_order - p < T) ? n_order - p : T;
        for (uint32_t t = 0; t < nb; ++t) {               /* sample against the batch-start residual */
            const uint32_t j = order[p + t];
            if (j >= M) { free(cx); free(draws); return 2; }
            const int g = groups ? groups[j] : 0;
            orc_dense_column(bed + cb * j, N, cx + (size_t)t * N);
            draws[t].beta_old = beta[j];                                            /* :300 */
            const double num = orc_dense_compute_num(cx + (size_t)t * N, eps, N, beta[j]);   /* :320 */
            draws[t].comp = orc_sample_marker(m, g, num, beta[j], rand_uz[2 * (size_t)j], rand_uz[2 * (size_t)j + 1],
                                              &draws[t].beta_new, &draws[t].acum);
            if (tokens == 0) {                            /* synchronous: apply before the next marker */
                orc_commit(m, g, j, &draws[t], 1, beta, comp, acum, v, betasq);
                if (!(draws[t].beta_old == 0.0 && draws[t].beta_new == 0.0))        /* :372 */
                    orc_dense_eps_update(cx, eps, N, draws[t].beta_old, draws[t].beta_new);
            }
        }
        if (tokens) {
            for (uint32_t t = 0; t < nb; ++t) {
                const uint32_t j = order[p + t];
                const int g = groups ? groups[j] : 0;
                orc_commit(m, g, j, &draws[t], 0, beta, comp, acum, v, betasq);     /* :492-493,511,522 */
                if (!(draws[t].beta_old == 0.0 && draws[t].beta_new == 0.0))        /* :482,514-523 */
                    orc_dense_eps_update(cx + (size_t)t * N, eps, N, draws[t].beta_old, draws[t].beta_new);
            }
        }
    }
    free(cx); free(draws);
    return 0;
}

static int orc_sweep_sparse(int eigen, const uint8_t *bed, uint32_t M, const uint32_t *order, uint32_t n_order,
                            const double *rand_uz, const orc_model *m, const int32_t *groups, uint32_t tokens,
                            double *eps, double *eps_sum, double *beta, double *comp, double *acum, double *v,
                            double *betasq)
{
    const uint32_t N = m->N;
    const size_t cb = ((size_t)N + 3) / 4;
    const uint32_t T = tokens ? tokens : 1;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)N * T);
    uint8_t *cls = (uint8_t *)malloc((size_t)N * T);
    double *stats = (double *)malloc(sizeof(double) * 4 * T);
    uint32_t *cnt = (uint32_t *)malloc(sizeof(uint32_t) * 3 * T);
    orc_draw *draws = (orc_draw *)malloc(sizeof(orc_draw) * T);
    if (!idx || !cls || !stats || !cnt || !draws) return 1;
    for (uint32_t p = 0; p < n_order; p += T) {
        const uint32_t nb = (n_order - p < T) ? n_order - p : T;
        const double eps_sum_batch = *eps_sum;            /* readWithSharedLock, src/SparseBayesRRG.cpp:76-82 */
        for (uint32_t t = 0; t < nb; ++t) {
            const uint32_t j = order[p + t];
            if (j >= M) { free(idx); free(cls); free(stats); free(cnt); free(draws); return 2; }
            const int g = groups ? groups[j] : 0;
            int32_t *ones = idx + (size_t)3 * N * t, *twos = ones + N, *miss = twos + N;
            uint8_t *c = cls + (size_t)N * t;
            orc_ragged_marker(bed + cb * j, N, stats + 4 * t, ones, &cnt[3 * t], twos, &cnt[3 * t + 1], miss,
                              &cnt[3 * t + 2]);
            if (eigen) { orc_eigen_stats(bed + cb * j, N, stats + 4 * t); orc_decode_bed_column(bed + cb * j, N, c); }
            draws[t].beta_old = beta[j];
            const double es = tokens ? eps_sum_batch : *eps_sum;
            const double num = eigen ? orc_eigen_compute_num(stats + 4 * t, c, eps, N, beta[j], es)
                                     : orc_ragged_compute_num(stats + 4 * t, ones, cnt[3 * t], twos, cnt[3 * t + 1], eps, N,
                                                              beta[j], es);
            draws[t].comp = orc_sample_marker(m, g, num, beta[j], rand_uz[2 * (size_t)j], rand_uz[2 * (size_t)j + 1],
                                              &draws[t].beta_new, &draws[t].acum);
            if (tokens == 0) {
                orc_commit(m, g, j, &draws[t], 1, beta, comp, acum, v, betasq);
                if (!(draws[t].beta_old == 0.0 && draws[t].beta_new == 0.0)) {
                    if (eigen) orc_eigen_eps_update(stats, c, eps, N, draws[t].beta_old, draws[t].beta_new, eps_sum);
                    else orc_ragged_eps_update(stats, ones, cnt[0], twos, cnt[1], miss, cnt[2], eps, N,
                                               draws[t].beta_old, draws[t].beta_new, eps_sum);
                }
            }
        }
        if (tokens) {
            for (uint32_t t = 0; t < nb; ++t) {
                const uint32_t j = order[p + t];
                const int g = groups ? groups[j] : 0;
                int32_t *ones = idx + (size_t)3 * N * t, *twos = ones + N, *miss = twos + N;
                orc_commit(m, g, j, &draws[t], 0, beta, comp, acum, v, betasq);
                if (!(draws[t].beta_old == 0.0 && draws[t].beta_new == 0.0)) {
                    if (eigen) orc_eigen_eps_update(stats + 4 * t, cls + (size_t)N * t, eps, N, draws[t].beta_old,
                                                    draws[t].beta_new, eps_sum);
                    else orc_ragged_eps_update(stats + 4 * t, ones, cnt[3 * t], twos, cnt[3 * t + 1], miss, cnt[3 * t + 2],
                                               eps, N, draws[t].beta_old, draws[t].beta_new, eps_sum);
                }
            }
        }
    }
    free(idx); free(cls); free(stats); free(cnt); free(draws);
    return 0;
}

int orc_sweep_ragged(const uint8_t *bed, uint32_t M, const uint32_t *order, uint32_t n_order,
                     const double *rand_uz, const orc_model *m, const int32_t *groups, uint32_t tokens,
                     double *eps, double *eps_sum, double *beta, double *comp, double *acum, double *v,
                     double *betasq)
{
    return orc_sweep_sparse(0, bed, M, order, n_order, rand_uz, m, groups, tokens, eps, eps_sum, beta, comp, acum, v, betasq);
}

int orc_sweep_eigen(const uint8_t *bed, uint32_t M, const uint32_t *order, uint32_t n_order,
                    const double *rand_uz, const orc_model *m, const int32_t *groups, uint32_t tokens,
                    double *eps, double *eps_sum, double *beta, double *comp, double *acum, double *v,
                    double *betasq)
{
    return orc_sweep_sparse(1, bed, M, order, n_order, rand_uz, m, groups, tokens, eps, eps_sum, beta, comp, acum, v, betasq);
}

/* src/BayesRBase.cpp:85-89 */
double orc_init_epsilon(const float *y, uint32_t N, double *eps)
{
    double s = 0.0;
    for (uint32_t i = 0; i < N; ++i) s += (double)y[i];
    const double mean = s / (double)N;
    double sq = 0.0;
    for (uint32_t i = 0; i < N; ++i) { eps[i] = (double)y[i] - mean; sq += eps[i] * eps[i]; }
    const double sc = sqrt(sq / (double)(N - 1));
    double sq2 = 0.0;
    for (uint32_t i = 0; i < N; ++i) { eps[i] /= sc; sq2 += eps[i] * eps[i]; }
    return sq2 / (double)N * 0.5;
}

/* src/DenseBayesRRmz.cpp:38-39: mean of the draw = sum(eps + mu_old)/N */
double orc_mu_mean_dense(const double *eps, uint32_t N, double mu_old)
{
    double s = 0.0;
    for (uint32_t i = 0; i < N; ++i) s += eps[i] + mu_old;
    return s / (double)N;
}

/* src/DenseBayesRRmz.cpp:38,40 */
void orc_mu_apply(double *eps, uint32_t N, double mu_old, double mu_new)
{
    for (uint32_t i = 0; i < N; ++i) eps[i] = (eps[i] + mu_old) - mu_new;
}
