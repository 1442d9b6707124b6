/*
 * lda_oracle.c — TEST INFRASTRUCTURE (see lda_oracle.h).  Plain-C restatement of the reference's variational
 * E-step, the three updateParameters variants and the special functions they use.  Written to keep the
 * reference's operation order wherever the order is visible in its source, so that it agrees with the
 * compiled reference (oracle/_ref) to rounding (~1e-13 relative after 10 trust-region iterations).
 *
 * Citations are relative to /root/reference/code/trlda/.
 */
#include "lda_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define ORACLE_PI    3.141592653589793238462643383279502884   /* include/utils.h:9  */
#define ORACLE_EULER 0.577215664901532860606512090082402431   /* include/utils.h:10 */

/* ---------------------------------------------------------------------------------------------------------
 * special functions
 * --------------------------------------------------------------------------------------------------------- */

/* asymptotic-series coefficients B_2k/(2k), src/digamma.cpp:42-52 */
static const double PSI_A[7] = {
	8.33333333333333333333E-2, -2.10927960927960927961E-2, 7.57575757575757575758E-3,
	-4.16666666666666666667E-3, 3.96825396825396825397E-3, -8.33333333333333333333E-3,
	8.33333333333333333333E-2};

/* src/digamma.cpp:116-178 (Cephes psi): reflection for x <= 0, exact harmonic numbers for integers <= 10,
 * upward recurrence to s >= 10, then the 7-term asymptotic series evaluated by Horner (polevl :95-110). */
double oracle_digamma(double x) {
	double p, q, nz = 0.0, s, w, y, z;
	int negative = 0;

	if(x <= 0.0) {
		negative = 1;
		q = x;
		p = floor(q);
		if(p == q)
			return INFINITY;
		nz = q - p;
		if(nz != 0.5) {
			if(nz > 0.5) {
				p += 1.0;
				nz = q - p;
			}
			nz = ORACLE_PI / tan(ORACLE_PI * nz);
		} else {
			nz = 0.0;
		}
		x = 1.0 - x;
	}

	if(x <= 10.0 && x == floor(x)) {
		int n = (int) x;
		y = 0.0;
		for(int i = 1; i < n; i++)
			y += 1.0 / (double) i;
		y -= ORACLE_EULER;
	} else {
		s = x;
		w = 0.0;
		while(s < 10.0) {
			w += 1.0 / s;
			s += 1.0;
		}
		if(s < 1.0e17) {
			z = 1.0 / (s * s);
			double ans = PSI_A[0];
			for(int i = 1; i <= 6; ++i)
				ans = ans * z + PSI_A[i];
			y = z * ans;
		} else {
			y = 0.0;
		}
		y = log(s) - (0.5 / s) - y - w;
	}

	if(negative)
		y -= nz;
	return y;
}

/* (2k)!/B_2k, src/zeta.cpp:47-60 */
static const double ZETA_A[12] = {
	12.0, -720.0, 30240.0, -1209600.0, 47900160.0, -1.8924375803183791606e9, 7.47242496e10,
	-2.950130727918164224e12, 1.1646782814350067249e14, -4.5979787224074726105e15,
	1.8152105401943546773e17, -7.1661652561756670113e18};

/* src/zeta.cpp:67-134 (Cephes Hurwitz zeta): direct sum until a > 9, then Euler–Maclaurin */
double oracle_zeta(double x, double q) {
	const double MACHEP = 1.11022302462515654042E-16;
	int i;
	double a, b, k, s, t, w;

	if(x == 1.0)
		return INFINITY;
	if(x < 1.0)
		return NAN;
	if(q <= 0.0) {
		if(q == floor(q))
			return INFINITY;
		if(x != floor(x))
			return NAN;
	}
	if(q > 1e8)
		return (1 / (x - 1) + 1 / (2 * q)) * pow(q, 1 - x);

	s = pow(q, -x);
	a = q;
	i = 0;
	b = 0.0;
	while((i < 9) || (a <= 9.0)) {
		i += 1;
		a += 1.0;
		b = pow(a, -x);
		s += b;
		if(fabs(b / s) < MACHEP)
			return s;
	}

	w = a;
	s += b * w / (x - 1.0);
	s -= 0.5 * b;
	a = 1.0;
	k = 0.0;
	for(i = 0; i < 12; i++) {
		a *= x + k;
		b /= w;
		t = a * b / ZETA_A[i];
		s = s + t;
		t = fabs(t / s);
		if(t < MACHEP)
			return s;
		k += 1.0;
		a *= x + k;
		b /= w;
		k += 1.0;
	}
	return s;
}

/* src/utils.cpp:107-111: (-1)^(n+1) n! zeta(n+1, x) */
double oracle_polygamma(int n, double x) {
	if(n < 1)
		return oracle_digamma(x);
	return pow(-1, n + 1) * tgamma(n + 1) * oracle_zeta(n + 1, x);
}

/* src/utils.cpp:224-231 with Eigen's Random (Eigen/src/Core/MathFunctions.h:441: x + (y-x)*rand()/RAND_MAX on
 * [-1, 1]): k passes over the m x n array in column-major order, each subtracting log|u|.  Divided by k here
 * because both call sites do (lda.cpp:71, :135). */
void oracle_sample_gamma(int m, int n, int k, double* out) {
	size_t size = (size_t) m * n;
	for(size_t e = 0; e < size; ++e)
		out[e] = 0.0;
	for(int i = 0; i < k; ++i)
		for(size_t e = 0; e < size; ++e) {
			double u = -1.0 + (1.0 - (-1.0)) * (double) rand() / (double) RAND_MAX;
			out[e] -= log(fabs(u));
		}
	for(size_t e = 0; e < size; ++e)
		out[e] /= (double) k;
}

/* ---------------------------------------------------------------------------------------------------------
 * model state
 * --------------------------------------------------------------------------------------------------------- */

/* constructors: lda.cpp:63-84, onlinelda.cpp:18-49, batchlda.cpp:22-39, cumulativelda.cpp:22-45 */
oracle_model* oracle_create(int kind, int V, int K, int num_documents, const double* alpha, double eta) {
	oracle_model* m = (oracle_model*) calloc(1, sizeof(oracle_model));
	m->kind = kind;
	m->K = K;
	m->V = V;
	m->eta = eta;
	m->alpha = (double*) malloc(sizeof(double) * K);
	memcpy(m->alpha, alpha, sizeof(double) * K);
	m->lambda = (double*) malloc(sizeof(double) * (size_t) K * V);
	oracle_sample_gamma(K, V, 100, m->lambda);
	m->num_documents = num_documents;
	m->update_counter = 0;
	m->ada_tau = 1000.;
	m->ada_rho = 1. / m->ada_tau;
	m->ada_sq_norm = 1.;
	m->ada_gradient = NULL;
	m->psi_gamma_diff = (double*) calloc(K, sizeof(double));
	m->cum_num_documents = 0;
	if(kind == TRLDA_KIND_CUMULATIVE)
		for(size_t e = 0; e < (size_t) K * V; ++e)
			m->lambda[e] = eta;
	return m;
}

void oracle_destroy(oracle_model* m) {
	if(!m) return;
	free(m->lambda);
	free(m->alpha);
	free(m->ada_gradient);
	free(m->psi_gamma_diff);
	free(m);
}

/* ---------------------------------------------------------------------------------------------------------
 * E-step, lda.cpp:160-220
 * --------------------------------------------------------------------------------------------------------- */

/* exp(psi(lambda) - psi(rowsum(lambda))), lda.cpp:172-173.  Row sums run over w in increasing order. */
static double* exp_elog_beta(const oracle_model* m) {
	const int K = m->K, V = m->V;
	double* psi_sum = (double*) calloc(K, sizeof(double));
	for(int w = 0; w < V; ++w)
		for(int k = 0; k < K; ++k)
			psi_sum[k] += m->lambda[(size_t) w * K + k];
	for(int k = 0; k < K; ++k)
		psi_sum[k] = oracle_digamma(psi_sum[k]);
	double* out = (double*) malloc(sizeof(double) * (size_t) K * V);
	#pragma omp parallel for
	for(int w = 0; w < V; ++w)
		for(int k = 0; k < K; ++k)
			out[(size_t) w * K + k] = exp(oracle_digamma(m->lambda[(size_t) w * K + k]) - psi_sum[k]);
	free(psi_sum);
	return out;
}

/* ---- LDA::updateVariablesGibbs (lda.cpp:224-293), restated WITHOUT its two defects --------------------------------
 *  (a) lda.cpp:254 reads theta.col(j) with j the token index; the document's own column theta.col(i) is meant;
 *  (b) lda.cpp:284 adds into sstats from all OpenMP threads unsynchronised; this restatement is serial.
 * Uniforms are rand() / (RAND_MAX + 1) (utils.cpp:203-206) and a topic is picked by walking the histogram in topic
 * order (sampleHistogram, utils.cpp:189-199).  The final theta ~ Dirichlet(counts) (lda.cpp:291) uses
 * std::gamma_distribution on an mt19937 in the reference, which plain C cannot replay: Marsaglia-Tsang on rand() here
 * — the law is the same, the stream is not.  Parity with the CUDA kernel is therefore distributional (the kernel has
 * its own counter-based generator anyway); tests/test_oracle.py and tests/test_gibbs_gpu.py hold the properties. */
static double gibbs_uniform01(void) {
	return (double) rand() / ((double) RAND_MAX + 1.0);
}

static int gibbs_histogram(const double* weight, int K) {
	double total = 0.0;
	for(int k = 0; k < K; ++k)
		total += weight[k];
	double r = gibbs_uniform01() * total;
	int last = 0;
	for(int k = 0; k < K; ++k) {
		if(weight[k] > 0.0)
			last = k;
		if(r < weight[k])
			return k;
		r -= weight[k];
	}
	return last;                                                  /* utils.cpp:198 throws here */
}

static double gibbs_gamma_draw(double shape) {
	const double a = shape < 1.0 ? shape + 1.0 : shape;
	const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
	for(;;) {
		const double u1 = (gibbs_uniform01() * RAND_MAX + 0.5) / ((double) RAND_MAX + 1.0);
		const double u2 = gibbs_uniform01();
		const double x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
		const double t = 1.0 + c * x;
		if(t <= 0.0)
			continue;
		const double v = t * t * t;
		const double u3 = (gibbs_uniform01() * RAND_MAX + 0.5) / ((double) RAND_MAX + 1.0);
		if(log(u3) < 0.5 * x * x + d - d * v + d * log(v)) {
			double g = d * v;
			if(shape < 1.0) {
				const double u4 = (gibbs_uniform01() * RAND_MAX + 0.5) / ((double) RAND_MAX + 1.0);
				g *= exp(log(u4) / shape);
			}
			return g;
		}
	}
}

int oracle_update_variables_gibbs(const oracle_model* m, const trlda_docs* docs, const double* theta0,
                                  int num_samples, int burn_in, double* theta_out, double* sstats_out)
{
	const int K = m->K, V = m->V;
	const int64_t B = docs->num_docs;
	const double unit = 1.0 / num_samples;                        /* :232 */
	double* beta = exp_elog_beta(m);                              /* :235-236 */
	double* counts = (double*) malloc(sizeof(double) * K);
	double* dist = (double*) malloc(sizeof(double) * K);
	if(sstats_out)
		memset(sstats_out, 0, sizeof(double) * (size_t) K * V);   /* :230 */

	for(int64_t d = 0; d < B; ++d) {
		const int64_t begin = docs->doc_ptr[d], end = docs->doc_ptr[d + 1];
		int64_t occurrences = 0;
		for(int64_t j = begin; j < end; ++j)
			occurrences += docs->counts[j];
		int* topics = (int*) malloc(sizeof(int) * (size_t) (occurrences ? occurrences : 1));
		const double* theta = theta0 + (size_t) d * K;            /* (a): column of the DOCUMENT */
		for(int k = 0; k < K; ++k)
			counts[k] = m->alpha[k];                              /* :244 */

		int64_t o = 0;
		for(int64_t j = begin; j < end; ++j) {                    /* :247-263 */
			const double* col = beta + (size_t) docs->word_ids[j] * K;
			for(int k = 0; k < K; ++k)
				dist[k] = col[k] * theta[k];
			for(int c = 0; c < docs->counts[j]; ++c, ++o) {
				topics[o] = gibbs_histogram(dist, K);
				counts[topics[o]] += 1.0;
			}
		}
		for(int s = 0; s < num_samples + burn_in; ++s) {          /* :265-288 */
			o = 0;
			for(int64_t j = begin; j < end; ++j) {
				const int w = docs->word_ids[j];
				const double* col = beta + (size_t) w * K;
				for(int c = 0; c < docs->counts[j]; ++c, ++o) {
					counts[topics[o]] -= 1.0;
					for(int k = 0; k < K; ++k)
						dist[k] = col[k] * counts[k];
					topics[o] = gibbs_histogram(dist, K);
					counts[topics[o]] += 1.0;
					if(s >= burn_in && sstats_out)
						sstats_out[(size_t) w * K + topics[o]] += unit;
				}
			}
		}
		if(theta_out) {                                           /* :291 */
			double sum = 0.0;
			double* out = theta_out + (size_t) d * K;
			for(int k = 0; k < K; ++k) {
				out[k] = gibbs_gamma_draw(counts[k]);
				sum += out[k];
			}
			for(int k = 0; k < K; ++k)
				out[k] /= sum;
		}
		free(topics);
	}
	free(dist);
	free(counts);
	free(beta);
	return 0;
}

int oracle_update_variables(const oracle_model* m, const trlda_docs* docs, const double* gamma0,
                            int max_iter, double threshold, double* gamma_out, double* sstats_out,
                            int* iterations_out)
{
	const int K = m->K, V = m->V;
	const int64_t B = docs->num_docs;
	const int64_t N = B ? docs->doc_ptr[B] : 0;

	double* beta = exp_elog_beta(m);                              /* expPsiLambda, :173 */
	double* gamma = (double*) malloc(sizeof(double) * (size_t) K * (B ? B : 1));
	double* etheta = (double*) malloc(sizeof(double) * (size_t) K * (B ? B : 1));
	double* weight = (double*) malloc(sizeof(double) * (size_t) (N ? N : 1));   /* c_j / phiNorm_j of the final state */
	memcpy(gamma, gamma0, sizeof(double) * (size_t) K * B);       /* :168 */

	#pragma omp parallel
	{
		double* last = (double*) malloc(sizeof(double) * K);
		#pragma omp for schedule(dynamic, 4)
		for(int64_t d = 0; d < B; ++d) {
			const int64_t begin = docs->doc_ptr[d];
			const int n = (int) (docs->doc_ptr[d + 1] - begin);
			const int32_t* ids = docs->word_ids + begin;
			const int32_t* cts = docs->counts + begin;
			double* g = gamma + (size_t) d * K;
			double* et = etheta + (size_t) d * K;
			double* phinorm = weight + begin;
			int it = 0;

			for(int k = 0; k < K; ++k)
				et[k] = exp(oracle_digamma(g[k]));                /* :174 — not normalised by psi(sum gamma) */

			for(int j = 0; j < n; ++j) {                          /* :183 */
				const double* col = beta + (size_t) ids[j] * K;
				double s = 0.0;
				for(int k = 0; k < K; ++k)
					s += et[k] * col[k];
				phinorm[j] = s + 1e-100;
			}

			for(it = 0; it < max_iter; ++it) {                    /* :185 */
				memcpy(last, g, sizeof(double) * K);
				for(int k = 0; k < K; ++k)
					g[k] = 0.0;
				for(int j = 0; j < n; ++j) {                      /* :190-193 */
					const double* col = beta + (size_t) ids[j] * K;
					const double wgt = cts[j] / phinorm[j];
					for(int k = 0; k < K; ++k)
						g[k] += wgt * col[k];
				}
				for(int k = 0; k < K; ++k) {                      /* :194-197 */
					g[k] *= et[k];
					g[k] += m->alpha[k];
				}
				for(int k = 0; k < K; ++k)
					et[k] = exp(oracle_digamma(g[k]));
				for(int j = 0; j < n; ++j) {                      /* :199 */
					const double* col = beta + (size_t) ids[j] * K;
					double s = 0.0;
					for(int k = 0; k < K; ++k)
						s += et[k] * col[k];
					phinorm[j] = s + 1e-100;
				}
				double change = 0.0;                              /* :202 */
				for(int k = 0; k < K; ++k)
					change += fabs(last[k] - g[k]);
				if(change / K < threshold) {
					++it;
					break;
				}
			}
			if(iterations_out)
				iterations_out[d] = it;
			for(int j = 0; j < n; ++j)
				phinorm[j] = cts[j] / phinorm[j];                 /* the factor of :212 */
		}
		free(last);
	}

	if(sstats_out) {
		memset(sstats_out, 0, sizeof(double) * (size_t) K * V);  /* :169 */
		for(int64_t d = 0; d < B; ++d)                            /* :207-213, document order */
			for(int64_t j = docs->doc_ptr[d]; j < docs->doc_ptr[d + 1]; ++j) {
				double* col = sstats_out + (size_t) docs->word_ids[j] * K;
				const double* et = etheta + (size_t) d * K;
				for(int k = 0; k < K; ++k)
					col[k] += weight[j] * et[k];
			}
		#pragma omp parallel for
		for(int w = 0; w < V; ++w)                                /* :217 */
			for(int k = 0; k < K; ++k)
				sstats_out[(size_t) w * K + k] *= beta[(size_t) w * K + k];
	}
	if(gamma_out)
		memcpy(gamma_out, gamma, sizeof(double) * (size_t) K * B);

	free(beta); free(gamma); free(etheta); free(weight);
	return 0;
}

/* ---------------------------------------------------------------------------------------------------------
 * helpers shared by the three updateParameters
 * --------------------------------------------------------------------------------------------------------- */

/* fresh-gamma E-step: lda.cpp:119-138 with the draw replaced by gamma0 when given */
static void fresh_estep(const oracle_model* m, const trlda_docs* docs, const trlda_params* p,
                        const double* gamma0, double* gamma, double* sstats)
{
	if(gamma0) {
		oracle_update_variables(m, docs, gamma0, p->max_iter_inference, p->threshold, gamma, sstats, NULL);
	} else {
		double* g0 = (double*) malloc(sizeof(double) * (size_t) m->K * docs->num_docs);
		oracle_sample_gamma(m->K, (int) docs->num_docs, 100, g0);
		oracle_update_variables(m, docs, g0, p->max_iter_inference, p->threshold, gamma, sstats, NULL);
		free(g0);
	}
}

/* sum_d [psi(gamma_dk) - psi(sum_k gamma_dk)], onlinelda.cpp:124-128 / batchlda.cpp:74-76 */
static void psi_gamma_diff(int K, int64_t B, const double* gamma, double* out) {
	for(int k = 0; k < K; ++k)
		out[k] = 0.0;
	for(int64_t d = 0; d < B; ++d) {
		const double* g = gamma + (size_t) d * K;
		double sum = 0.0;
		for(int k = 0; k < K; ++k)
			sum += g[k];
		const double psi_sum = oracle_digamma(sum);
		for(int k = 0; k < K; ++k)
			out[k] += oracle_digamma(g[k]) - psi_sum;
	}
}

static double sum_of(const double* x, int n) {
	double s = 0.0;
	for(int i = 0; i < n; ++i)
		s += x[i];
	return s;
}

/* Newton direction pieces shared by onlinelda.cpp:128-134, batchlda.cpp:90-96, cumulativelda.cpp:98-104:
 * g = diff - n (psi(alpha) - psi(sum alpha)); h = -n psi'(alpha); z = n psi'(sum alpha);
 * c = sum(g/h) / (1/z + sum(1/h)) */
static double alpha_newton(int K, const double* alpha, const double* diff, double n, double* g, double* h) {
	const double asum = sum_of(alpha, K);
	const double psi_asum = oracle_digamma(asum);
	double sgh = 0.0, sih = 0.0;
	for(int k = 0; k < K; ++k) {
		g[k] = diff[k] - n * (oracle_digamma(alpha[k]) - psi_asum);
		h[k] = -n * oracle_polygamma(1, alpha[k]);
		sgh += g[k] / h[k];
		sih += 1. / h[k];
	}
	const double z = n * oracle_polygamma(1, asum);
	return sgh / (1. / z + sih);
}

/* n (lnG(sum a) - sum lnG(a)) + sum diff (a - 1), batchlda.cpp:82-83 */
static double alpha_objective(int K, const double* a, const double* diff, double n) {
	double lg = 0.0, lin = 0.0;
	for(int k = 0; k < K; ++k) {
		lg += lgamma(a[k]);
		lin += diff[k] * (a[k] - 1.);
	}
	return n * (lgamma(sum_of(a, K)) - lg) + lin;
}

/* line-searched Newton ascent on alpha: batchlda.cpp:82-142, cumulativelda.cpp:91-150 */
static void alpha_line_search(oracle_model* m, const double* diff, double n, const trlda_params* p) {
	const int K = m->K;
	double* g = (double*) malloc(sizeof(double) * K);
	double* h = (double*) malloc(sizeof(double) * K);
	double* a = (double*) malloc(sizeof(double) * K);
	double L = alpha_objective(K, m->alpha, diff, n);
	double Lprime = L;

	for(int i = 0; i < p->max_iter_alpha; ++i) {
		const double c = alpha_newton(K, m->alpha, diff, n, g, h);
		double rho = .2;
		for(int j = 0; j < 20; ++j) {
			int small = 0;
			for(int k = 0; k < K; ++k) {
				a[k] = m->alpha[k] - rho * (g[k] - c) / h[k];
				if(a[k] < p->min_alpha)
					small = 1;
			}
			if(small) {
				rho /= 2.;
				continue;
			}
			Lprime = alpha_objective(K, a, diff, n);
			if(L <= Lprime) {
				memcpy(m->alpha, a, sizeof(double) * K);
				break;
			}
			rho /= 2.;
		}
		if(Lprime - L < p->emp_bayes_threshold)
			break;
		L = Lprime;
	}
	free(g); free(h); free(a);
}

/* sum psi(lambda) - V sum_k psi(sum_w lambda_kw), onlinelda.cpp:153 / batchlda.cpp:153 */
static double eta_constant(const oracle_model* m) {
	const int K = m->K, V = m->V;
	double total = 0.0;
	double* rows = (double*) calloc(K, sizeof(double));
	for(int w = 0; w < V; ++w)
		for(int k = 0; k < K; ++k) {
			total += oracle_digamma(m->lambda[(size_t) w * K + k]);
			rows[k] += m->lambda[(size_t) w * K + k];
		}
	double psi_rows = 0.0;
	for(int k = 0; k < K; ++k)
		psi_rows += oracle_digamma(rows[k]);
	free(rows);
	return total - V * psi_rows;
}

/* ---------------------------------------------------------------------------------------------------------
 * OnlineLDA::updateParameters, onlinelda.cpp:53-180
 * --------------------------------------------------------------------------------------------------------- */
static double online_update(oracle_model* m, const trlda_docs* docs, const trlda_params* p,
                            const double* gamma0, double* gamma_out)
{
	const int K = m->K, V = m->V;
	const int64_t B = docs->num_docs;
	const size_t KV = (size_t) K * V;
	if(B == 0)
		return 1.0;                                               /* :54-56, counter untouched */

	double rho = p->rho;                                          /* :59-66 */
	if(rho < 0.) {
		if(p->adaptive)
			rho = m->ada_rho;
		else
			rho = pow(p->tau + m->update_counter, -p->kappa);
	}

	double* lambda_prime = (double*) malloc(sizeof(double) * KV);
	double* lambda_hat = (double*) malloc(sizeof(double) * KV);
	double* sstats = (double*) malloc(sizeof(double) * KV);
	double* gamma = (double*) malloc(sizeof(double) * (size_t) K * B);
	double* gamma_next = (double*) malloc(sizeof(double) * (size_t) K * B);
	memcpy(lambda_prime, m->lambda, sizeof(double) * KV);         /* :68 */
	const double scale = (double) m->num_documents / B;           /* static_cast<double>(D) / documents.size() */

	if(p->update_lambda) {
		if(p->max_iter_tr > 0) {
			double* wordcounts = (double*) calloc(V, sizeof(double));   /* :79-82 */
			for(int64_t j = 0; j < docs->doc_ptr[B]; ++j)
				wordcounts[docs->word_ids[j]] += docs->counts[j];
			const double scale_k = (double) m->num_documents / B / K;
			for(int w = 0; w < V; ++w) {                          /* :85-86 */
				const double target = rho * (m->eta + scale_k * wordcounts[w]);
				for(int k = 0; k < K; ++k)
					m->lambda[(size_t) w * K + k] = (1. - rho) * lambda_prime[(size_t) w * K + k] + target;
			}
			free(wordcounts);

			for(int i = 0; i < p->max_iter_tr; ++i) {             /* :89-101 */
				if(i > 0 && p->init_gamma) {
					oracle_update_variables(m, docs, gamma, p->max_iter_inference, p->threshold, gamma_next, sstats, NULL);
					memcpy(gamma, gamma_next, sizeof(double) * (size_t) K * B);
				} else {
					fresh_estep(m, docs, p, gamma0, gamma, sstats);
				}
				for(size_t e = 0; e < KV; ++e) {
					lambda_hat[e] = m->eta + scale * sstats[e];
					m->lambda[e] = (1. - rho) * lambda_prime[e] + rho * lambda_hat[e];
				}
			}
		} else {                                                  /* :102-110 */
			fresh_estep(m, docs, p, gamma0, gamma, sstats);
			for(size_t e = 0; e < KV; ++e) {
				lambda_hat[e] = m->eta + scale * sstats[e];
				m->lambda[e] = (1. - rho) * lambda_prime[e] + rho * lambda_hat[e];
			}
		}
	}

	if(p->update_alpha) {                                         /* :116-142 */
		if(!p->update_lambda)
			fresh_estep(m, docs, p, gamma0, gamma, NULL);
		double* diff = (double*) malloc(sizeof(double) * K);
		double* g = (double*) malloc(sizeof(double) * K);
		double* h = (double*) malloc(sizeof(double) * K);
		psi_gamma_diff(K, B, gamma, diff);
		const double c = alpha_newton(K, m->alpha, diff, (double) B, g, h);
		for(int k = 0; k < K; ++k) {
			m->alpha[k] = m->alpha[k] - rho * (g[k] - c) / h[k];
			if(m->alpha[k] < p->min_alpha)
				m->alpha[k] = p->min_alpha;
		}
		free(diff); free(g); free(h);
	}

	if(p->update_eta) {                                           /* :147-162 */
		const int N = V;
		const double g = eta_constant(m) - K * N * (oracle_digamma(m->eta) - oracle_digamma(N * m->eta));
		const double h = K * N * (oracle_polygamma(1, N * m->eta) - oracle_polygamma(1, m->eta));
		m->eta = m->eta - rho * g / h;
		if(m->eta < p->min_eta)
			m->eta = p->min_eta;
	}

	if(p->update_lambda && p->adaptive) {                         /* :167-175 */
		if(!m->ada_gradient)
			m->ada_gradient = (double*) calloc(KV, sizeof(double));
		double sq = 0.0, gsq = 0.0;
		for(size_t e = 0; e < KV; ++e) {
			const double u = lambda_hat[e] - lambda_prime[e];
			sq += u * u;
			m->ada_gradient[e] = (1. - 1. / m->ada_tau) * m->ada_gradient[e] + 1. / m->ada_tau * u;
			gsq += m->ada_gradient[e] * m->ada_gradient[e];
		}
		m->ada_sq_norm = (1. - 1. / m->ada_tau) * m->ada_sq_norm + 1. / m->ada_tau * sq;
		m->ada_rho = gsq / m->ada_sq_norm;
		m->ada_tau = m->ada_tau * (1. - m->ada_rho) + 1.;
	}

	m->update_counter++;                                          /* :177 */
	if(gamma_out)
		memcpy(gamma_out, gamma, sizeof(double) * (size_t) K * B);
	free(lambda_prime); free(lambda_hat); free(sstats); free(gamma); free(gamma_next);
	return rho;
}

/* ---------------------------------------------------------------------------------------------------------
 * BatchLDA::updateParameters, batchlda.cpp:43-209
 * --------------------------------------------------------------------------------------------------------- */
static double batch_update(oracle_model* m, const trlda_docs* docs, const trlda_params* p,
                           const double* gamma0, double* gamma_out)
{
	const int K = m->K, V = m->V;
	const int64_t B = docs->num_docs;
	const size_t KV = (size_t) K * V;
	if(B == 0)
		return 1.;

	double* sstats = (double*) malloc(sizeof(double) * KV);
	double* gamma = (double*) malloc(sizeof(double) * (size_t) K * B);
	double* diff = (double*) malloc(sizeof(double) * K);

	for(int epoch = 0; epoch < p->max_epochs; ++epoch) {
		if(p->update_lambda) {                                    /* :54-61 */
			fresh_estep(m, docs, p, gamma0, gamma, sstats);
			for(size_t e = 0; e < KV; ++e)
				m->lambda[e] = m->eta + sstats[e];
		}

		if(p->update_alpha) {                                     /* :66-143 */
			if(!p->update_lambda)
				fresh_estep(m, docs, p, gamma0, gamma, NULL);
			psi_gamma_diff(K, B, gamma, diff);
			alpha_line_search(m, diff, (double) B, p);
		}

		if(p->update_eta) {                                       /* :147-205 */
			const int N = V;
			const double c = eta_constant(m);
			double L = (m->eta - 1) * c + K * lgamma(N * m->eta) - K * N * lgamma(m->eta);
			double Lprime = L;
			for(int i = 0; i < p->max_iter_eta; ++i) {
				const double g = c - K * N * (oracle_digamma(m->eta) - oracle_digamma(N * m->eta));
				const double h = K * N * (oracle_polygamma(1, N * m->eta) - oracle_polygamma(1, m->eta));
				double rho = .5;
				for(int j = 0; j < 20; ++j) {
					const double eta = m->eta - rho * g / h;
					if(eta < p->min_eta) {
						rho /= 2.;
						continue;
					}
					Lprime = (eta - 1) * c + K * lgamma(N * eta) - K * N * lgamma(eta);
					if(L <= Lprime) {
						m->eta = eta;
						break;
					}
					rho /= 2.;
				}
				if(Lprime - L < p->emp_bayes_threshold)
					break;
				L = Lprime;
			}
		}
	}

	if(gamma_out)
		memcpy(gamma_out, gamma, sizeof(double) * (size_t) K * B);
	free(sstats); free(gamma); free(diff);
	return 1.;
}

/* ---------------------------------------------------------------------------------------------------------
 * CumulativeLDA::updateParameters, cumulativelda.cpp:49-153
 * --------------------------------------------------------------------------------------------------------- */
static double cumulative_update(oracle_model* m, const trlda_docs* docs, const trlda_params* p,
                                const double* gamma0, const double* lambda0, double* gamma_out)
{
	const int K = m->K, V = m->V;
	const int64_t B = docs->num_docs;
	const size_t KV = (size_t) K * V;
	if(B == 0)
		return 1.;

	double* lambda_prime = (double*) malloc(sizeof(double) * KV);
	double* sstats = (double*) malloc(sizeof(double) * KV);
	double* gamma = (double*) malloc(sizeof(double) * (size_t) K * B);
	memcpy(lambda_prime, m->lambda, sizeof(double) * KV);         /* :57 */

	if(lambda0)                                                   /* :60 — happens even if !update_lambda */
		memcpy(m->lambda, lambda0, sizeof(double) * KV);
	else
		oracle_sample_gamma(K, V, 100, m->lambda);

	if(p->update_lambda)
		for(int epoch = 0; epoch < p->max_epochs; ++epoch) {      /* :62-71 */
			fresh_estep(m, docs, p, gamma0, gamma, sstats);
			for(size_t e = 0; e < KV; ++e)
				m->lambda[e] = lambda_prime[e] + sstats[e];
		}

	if(p->update_alpha) {                                         /* :76-150 */
		double* diff = (double*) malloc(sizeof(double) * K);
		fresh_estep(m, docs, p, gamma0, gamma, NULL);
		psi_gamma_diff(K, B, gamma, diff);
		for(int k = 0; k < K; ++k)
			m->psi_gamma_diff[k] += diff[k];
		m->cum_num_documents += (int) B;
		alpha_line_search(m, m->psi_gamma_diff, (double) m->cum_num_documents, p);
		free(diff);
	}

	if(gamma_out)
		memcpy(gamma_out, gamma, sizeof(double) * (size_t) K * B);
	free(lambda_prime); free(sstats); free(gamma);
	return 1.;
}

double oracle_update_parameters(oracle_model* m, const trlda_docs* docs, const trlda_params* p,
                                const double* gamma0, const double* lambda0, double* gamma_out)
{
	switch(m->kind) {
		case TRLDA_KIND_ONLINE: return online_update(m, docs, p, gamma0, gamma_out);
		case TRLDA_KIND_BATCH: return batch_update(m, docs, p, gamma0, gamma_out);
		default: return cumulative_update(m, docs, p, gamma0, lambda0, gamma_out);
	}
}

/* ---------------------------------------------------------------------------------------------------------
 * variational lower bound — the INTENDED formula
 *
 * lda.cpp:297-360 computes  E[log p(w|z,beta)] + E[log p(beta)] - E[log q(beta)]   (:317, :356-357)
 *                        + f (E[log p(z)] - E[log q(z)])                           (:325-347)
 *                        + f (E[log p(theta)] - E[log q(theta)])                   (:349-355)
 * with f = numDocuments / B, but reads psiLambda.row(wordID) at :334 where the word's column is meant.
 * With phi exactly the softmax of E[log theta] + E[log beta], the z- and w-terms collapse to
 * sum_j c_j log sum_k exp(E[log theta_k] + E[log beta_kw_j]), which is how Hoffman's reference VB writes it
 * (python/tests/onlineldavb.py:289-307); that is what is evaluated here.
 * --------------------------------------------------------------------------------------------------------- */
double oracle_lower_bound(const oracle_model* m, const trlda_docs* docs, const double* gamma0,
                          const trlda_params* p, int num_documents, double* per_doc_out)
{
	const int K = m->K, V = m->V;
	const int64_t B = docs->num_docs;
	const size_t KV = (size_t) K * V;
	const double factor = num_documents >= 0 ? num_documents / (double) B : 1;   /* :303-304 */

	double* gamma = (double*) malloc(sizeof(double) * (size_t) K * (B ? B : 1));
	fresh_estep(m, docs, p, gamma0, gamma, NULL);                 /* :307 */

	double* rows = (double*) calloc(K, sizeof(double));
	for(int w = 0; w < V; ++w)
		for(int k = 0; k < K; ++k)
			rows[k] += m->lambda[(size_t) w * K + k];
	double* psi_rows = (double*) malloc(sizeof(double) * K);
	for(int k = 0; k < K; ++k)
		psi_rows[k] = oracle_digamma(rows[k]);
	double* elogbeta = (double*) malloc(sizeof(double) * KV);
	#pragma omp parallel for
	for(int w = 0; w < V; ++w)
		for(int k = 0; k < K; ++k)
			elogbeta[(size_t) w * K + k] = oracle_digamma(m->lambda[(size_t) w * K + k]) - psi_rows[k];

	double lg_alpha = 0.0;
	for(int k = 0; k < K; ++k)
		lg_alpha += lgamma(m->alpha[k]);
	const double alpha_const = lgamma(sum_of(m->alpha, K)) - lg_alpha;   /* :355 */

	double docs_total = 0.0;
	double* elogtheta = (double*) malloc(sizeof(double) * K);
	double* tmp = (double*) malloc(sizeof(double) * K);
	for(int64_t d = 0; d < B; ++d) {
		const double* g = gamma + (size_t) d * K;
		const double gsum = sum_of(g, K);
		const double psi_gsum = oracle_digamma(gsum);
		double score = 0.0;
		for(int k = 0; k < K; ++k)
			elogtheta[k] = oracle_digamma(g[k]) - psi_gsum;       /* :341 */
		for(int64_t j = docs->doc_ptr[d]; j < docs->doc_ptr[d + 1]; ++j) {
			const double* col = elogbeta + (size_t) docs->word_ids[j] * K;
			double tmax = -INFINITY;
			for(int k = 0; k < K; ++k) {
				tmp[k] = elogtheta[k] + col[k];
				if(tmp[k] > tmax) tmax = tmp[k];
			}
			double s = 0.0;
			for(int k = 0; k < K; ++k)
				s += exp(tmp[k] - tmax);
			score += docs->counts[j] * (log(s) + tmax);           /* onlineldavb.py:289-296 */
		}
		for(int k = 0; k < K; ++k)                                /* :349-351 */
			score += (m->alpha[k] - g[k]) * elogtheta[k] + lgamma(g[k]);
		score -= lgamma(gsum);
		score += alpha_const;
		if(per_doc_out)
			per_doc_out[d] = score;
		docs_total += score;
	}

	double beta_terms = 0.0;                                      /* :317 without the sstats part, :356-357 */
	for(size_t e = 0; e < KV; ++e)
		beta_terms += (m->eta - m->lambda[e]) * elogbeta[e] + lgamma(m->lambda[e]);
	beta_terms -= (double) K * V * lgamma(m->eta);
	for(int k = 0; k < K; ++k)
		beta_terms += lgamma(V * m->eta) - lgamma(rows[k]);

	free(gamma); free(rows); free(psi_rows); free(elogbeta); free(elogtheta); free(tmp);
	return beta_terms + factor * docs_total;
}
