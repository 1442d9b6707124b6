/*
 * lda_oracle.h — TEST INFRASTRUCTURE.  Plain-C restatement of the reference's hot path, used only as the
 * checker by tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py.
 * The product (trlda_b200/) never includes, links or loads anything in this directory.
 *
 * PARITY PINNED: every function here is checked in tests/test_oracle.py against
 *   (a) the golden psi / psi' / psi'' values of the reference's own tests (python/tests/utils_test.py:33-51),
 *   (b) oracle/_ref/libtrlda_ref.so = the unmodified reference core compiled here (oracle/Makefile), on
 *       seeded inputs (gamma, sstats, lambda, alpha, eta agree to ~1e-13 relative), and
 *   (c) fixtures under tests/golden/ generated from (b) by tests/golden/make_golden.py.
 */
#ifndef LDA_ORACLE_H
#define LDA_ORACLE_H

#include <stdint.h>
#include "../include/trlda_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_model {
	int kind;              /* TRLDA_KIND_* */
	int K, V;
	double* lambda;        /* K x V column-major                      (lda.h:134)            */
	double* alpha;         /* K                                       (lda.h:132)            */
	double eta;            /*                                         (lda.h:133)            */
	int num_documents;     /* OnlineLDA::mNumDocuments                (onlinelda.h:37)       */
	int update_counter;    /* OnlineLDA::mUpdateCounter               (onlinelda.h:38)       */
	double ada_rho, ada_tau, ada_sq_norm;   /*                        (onlinelda.h:41-43)    */
	double* ada_gradient;  /* K x V, allocated lazily                 (onlinelda.h:44)       */
	double* psi_gamma_diff;/* K, CumulativeLDA::mPsiGammaDiff         (cumulativelda.h:24)   */
	int cum_num_documents; /* CumulativeLDA::mNumDocuments            (cumulativelda.h:25)   */
} oracle_model;

double oracle_digamma(double x);                 /* digamma.cpp:116-178 */
double oracle_zeta(double x, double q);          /* zeta.cpp:67-134     */
double oracle_polygamma(int n, double x);        /* utils.cpp:107-111   */
void   oracle_sample_gamma(int m, int n, int k, double* out);  /* utils.cpp:224-231, divided by k */

oracle_model* oracle_create(int kind, int V, int K, int num_documents, const double* alpha, double eta);
void oracle_destroy(oracle_model* m);

/* lda.cpp:160-220.  iterations_out (B ints, may be NULL) receives the inner iteration count per document. */
int oracle_update_variables(const oracle_model* m, const trlda_docs* docs, const double* gamma0,
                            int max_iter, double threshold, double* gamma_out, double* sstats_out,
                            int* iterations_out);

/* onlinelda.cpp:53-180 / batchlda.cpp:43-209 / cumulativelda.cpp:49-153.  gamma0 (K x B) replaces every
 * internal fresh draw of gamma (NULL: draw from rand() like utils.cpp:224-231); lambda0 replaces the random
 * re-initialisation of cumulativelda.cpp:60 (NULL: draw).  gamma_out (K x B, may be NULL) receives the gamma of
 * the last E-step. */
/* LDA::updateVariablesGibbs (lda.cpp:224-293) with its index bug (:254) and data race (:284) removed; theta0 is K x B,
 * theta_out K x B, sstats_out K x V (either may be NULL); draws from rand() */
int oracle_update_variables_gibbs(const oracle_model* m, const trlda_docs* docs, const double* theta0,
                                  int num_samples, int burn_in, double* theta_out, double* sstats_out);

double oracle_update_parameters(oracle_model* m, const trlda_docs* docs, const trlda_params* p,
                                const double* gamma0, const double* lambda0, double* gamma_out);

/* The INTENDED variational bound (tests/onlineldavb.py:260-318; lda.cpp:297-360 minus the bug at :334).
 * per_doc_out may be NULL. */
double oracle_lower_bound(const oracle_model* m, const trlda_docs* docs, const double* gamma0,
                          const trlda_params* p, int num_documents, double* per_doc_out);

#ifdef __cplusplus
}
#endif
#endif
