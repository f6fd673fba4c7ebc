/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's BayesR hot path in plain C.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this; the product
 * path (bayesrrcmd_b200/, include/brr_gpu.h) never links, imports or executes it.
 *
 * Parity status: PINNED against the reference's own object code (oracle/_ref/libbrr_ref.so, built
 * by oracle/Makefile from /root/reference) by tests/test_oracle_vs_ref.py, and against the golden
 * vectors that library produced (tests/golden/, generator: oracle/make_golden.py).  The reference's
 * own test-suite holds no golden vectors for this path.
 * One boundary stays unpinned: seed -> random variates (Boost.Random is absent and unversioned
 * upstream); the sweep consumes the random table as an INPUT, so it is not on this path.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef BAYESR_ORACLE_H
#define BAYESR_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* genotype classes used throughout: 0,1,2 = allele count, 3 = missing */
enum { ORC_MISSING = 3 };

/* src/preprocessgraph.cpp:44-53 + src/densemarkerbuilder.cpp:30: PLINK 2-bit -> class.
 * allele1 = !bit[2k], allele2 = !bit[2k+1]; (0,1) = missing, else count = allele1+allele2. */
void orc_decode_bed_column(const uint8_t *bedcol, uint32_t N, uint8_t *cls);

/* src/densemarkerbuilder.cpp:23-60: standardised dense column, N doubles. */
void orc_dense_column(const uint8_t *bedcol, uint32_t N, double *cx);

/* The <=4 distinct values of that column: lut[c] for class c (src/densemarkerbuilder.cpp:50-59). */
void orc_dense_lut(const uint8_t *bedcol, uint32_t N, double lut[4]);

/* src/raggedsparsemarkerbuilder.cpp:25-55 + src/sparsemarker.cpp:8-19.
 * stats = {mean, sd, sqrdZ, Zsum}; index lists ascending; each list needs room for N entries. */
void orc_ragged_marker(const uint8_t *bedcol, uint32_t N, double stats[4], int32_t *ones, uint32_t *n1,
                       int32_t *twos, uint32_t *n2, int32_t *miss, uint32_t *nm);

/* src/eigensparsemarkerbuilder.cpp:40-78 (statistics with missing imputed by the mean). */
void orc_eigen_stats(const uint8_t *bedcol, uint32_t N, double stats[4]);

/* src/densebayesrkernel.cpp:10-15 */
double orc_dense_compute_num(const double *cx, const double *eps, uint32_t N, double beta_old);
/* src/densebayesrkernel.cpp:17-20 + src/BayesRBase.cpp:376 (eps += (beta_old-beta)*Cx) */
void orc_dense_eps_update(const double *cx, double *eps, uint32_t N, double beta_old, double beta);

/* src/sparsebayesrkernel.cpp:24-27 + src/raggedbayesrkernel.cpp:32-35 */
double orc_ragged_compute_num(const double stats[4], const int32_t *ones, uint32_t n1, const int32_t *twos,
                              uint32_t n2, const double *eps, uint32_t N, double beta_old, double eps_sum);
/* src/raggedbayesrkernel.cpp:10-30 + src/sparsebayesrkernel.cpp:29-34 + src/SparseBayesRRG.cpp:84-92.
 * Updates eps in place and *eps_sum += delta of the analytically tracked sum. */
void orc_ragged_eps_update(const double stats[4], const int32_t *ones, uint32_t n1, const int32_t *twos,
                           uint32_t n2, const int32_t *miss, uint32_t nm, double *eps, uint32_t N,
                           double beta_old, double beta, double *eps_sum);

/* src/eigenbayesrkernel.cpp:10-20 (Zg holds 1, 2, or the mean for missing) */
double orc_eigen_compute_num(const double stats[4], const uint8_t *cls, const double *eps, uint32_t N,
                             double beta_old, double eps_sum);
void orc_eigen_eps_update(const double stats[4], const uint8_t *cls, double *eps, uint32_t N, double beta_old,
                          double beta, double *eps_sum);

typedef struct {
    uint32_t N;          /* individuals */
    uint32_t K;          /* mixture components incl. the spike (cols(S)+1) */
    uint32_t nGroups;
    const double *cva;   /* nGroups x (K-1), row-major (Options::S) */
    const double *pi;    /* nGroups x K */
    const double *sigmaG;/* nGroups */
    double sigmaE;
} orc_model;

/* src/BayesRBase.cpp:306-366: the scalar part of processColumn.
 * Returns the chosen component, or -1 if the walk falls through (nothing assigned; beta_new=beta_old).
 * acum_out = P(component 0) as stored in m_acum (src/BayesRBase.cpp:341). */
int orc_sample_marker(const orc_model *m, int group, double num, double beta_old, double u, double z,
                      double *beta_new, double *acum_out);

/* One synchronous sweep = what AnalysisGraph::exec must do (src/sequential.cpp:10-27 driving
 * src/BayesRBase.cpp:287-392) on a dense layout.  bed: M columns of ceil(N/4) bytes.
 * rand_uz: M x 2 indexed by MARKER ID (src/BayesRBase.cpp:343-344).  groups may be NULL.
 * v (nGroups x K) and betasq (nGroups) are ACCUMULATED INTO (caller zeroes them, src/BayesRBase.cpp:200-202).
 * tokens = 0: synchronous.  tokens > 0: bounded-staleness async (src/BayesRBase.cpp:394-523 under
 * src/parallelgraph.cpp:106-117): markers are sampled in batches of `tokens` against the residual
 * as it stood at the start of the batch, then all their updates are applied; m_acum is not written
 * (reference quirk), v keeps correct per-group cells (documented deviation). */
int orc_sweep_dense(const uint8_t *bed, uint32_t M, const uint32_t *order, uint32_t n_order,
                    const double *rand_uz, const orc_model *m, const int32_t *groups, uint32_t tokens,
                    double *eps, double *beta, double *comp, double *acum, double *v, double *betasq);

/* Same for the ragged layout (src/SparseBayesRRG.cpp:76-92); *eps_sum is m_epsilonSum. */
int orc_sweep_ragged(const uint8_t *bed, uint32_t M, const uint32_t *order, uint32_t n_order,
                     const double *rand_uz, const orc_model *m, const int32_t *groups, uint32_t tokens,
                     double *eps, double *eps_sum, double *beta, double *comp, double *acum, double *v,
                     double *betasq);

/* Same for the eigen-sparse layout (src/eigenbayesrkernel.cpp, src/eigensparsemarkerbuilder.cpp). */
int orc_sweep_eigen(const uint8_t *bed, uint32_t M, const uint32_t *order, uint32_t n_order,
                    const double *rand_uz, const orc_model *m, const int32_t *groups, uint32_t tokens,
                    double *eps, double *eps_sum, double *beta, double *comp, double *acum, double *v,
                    double *betasq);

/* src/BayesRBase.cpp:85-89: y (float32, file order) -> standardised eps; returns sigmaE init. */
double orc_init_epsilon(const float *y, uint32_t N, double *eps);

/* src/DenseBayesRRmz.cpp:36-42 / src/SparseBayesRRG.cpp:106-113 given the N(0,1)-independent draw:
 * the caller supplies mu_new (drawn by the host RNG from N(sum/N, sigmaE/N)); these helpers return the
 * mean the draw must be centred on and apply the +old/-new shift. */
double orc_mu_mean_dense(const double *eps, uint32_t N, double mu_old);
void orc_mu_apply(double *eps, uint32_t N, double mu_old, double mu_new);

#ifdef __cplusplus
}
#endif
#endif
