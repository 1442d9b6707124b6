/*
 * ref_shim.cpp — TEST INFRASTRUCTURE, not product code.
 *
 * An `extern "C"` surface over the UNMODIFIED reference core (code/trlda/src/*.cpp under /root/reference,
 * compiled where it lies by oracle/Makefile into oracle/_ref/libtrlda_ref.so) so that tests and the
 * `cpu_baseline` / `--impl reference` legs of bench.py can drive it with ctypes.  Nothing of the reference
 * is copied; this file only includes its public headers and calls its public methods.
 *
 * The one trick: `updateParameters` draws its initial gamma internally (lda.cpp:133-136).  To run the
 * reference and the CUDA path "on the same inputs and initial gamma", the subclasses below override the
 * virtual 2-argument `updateVariables` (lda.h:105-107) and forward to the 3-argument form with a
 * caller-supplied gamma0.  Zero edits to reference code, identical arithmetic.
 */
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "trlda/models"
#include "trlda/utils"
#include "../include/trlda_b200.h"

using namespace TRLDA;
using Eigen::ArrayXd;
using Eigen::ArrayXXd;
using Eigen::Map;
using std::pair;

namespace {

template <class Base>
struct Injected : public Base {
	template <class... Args>
	Injected(Args... args) : Base(args...), injected(0) {}

	const ArrayXXd* injected;

	// Replaces lambda (a protected member, lda.h:134) by a constant matrix of any width.  Used by ref_create_fast to
	// skip the constructor's 100 K V calls to rand() (lda.cpp:71 -> utils.cpp:224-231: minutes at K V = 1e8) when the
	// caller overwrites lambda anyway; the object is first constructed with numWords = 1.
	void resetLambda(int numTopics, int numWords, double value) {
		this->mLambda = ArrayXXd::Constant(numTopics, numWords, value);
	}

	using Base::updateVariables;
	virtual pair<ArrayXXd, ArrayXXd> updateVariables(
		const LDA::Documents& documents,
		const LDA::Parameters& parameters) const
	{
		if(injected)
			return Base::updateVariables(documents, *injected, parameters);
		return Base::updateVariables(documents, parameters);
	}
};

struct Handle {
	int kind;
	LDA* lda;
	const ArrayXXd** slot;
	const char* error;
	bool fast;     // built by ref_create_fast: OnlineLDA's private mAdaGradient has the wrong width, adaptive mode is off limits
};

LDA::Documents toDocuments(const trlda_docs* docs) {
	LDA::Documents documents(docs->num_docs);
	for(int64_t d = 0; d < docs->num_docs; ++d) {
		documents[d].reserve(docs->doc_ptr[d + 1] - docs->doc_ptr[d]);
		for(int64_t j = docs->doc_ptr[d]; j < docs->doc_ptr[d + 1]; ++j)
			documents[d].push_back(std::make_pair(docs->word_ids[j], docs->counts[j]));
	}
	return documents;
}

LDA::Parameters toParameters(const trlda_params* p) {
	return LDA::Parameters(
		p->inference_method == TRLDA_INFERENCE_GIBBS ? LDA::GIBBS : LDA::VI,
		p->threshold, p->max_iter_inference, p->max_iter_tr, p->tau, p->kappa, p->rho, p->adaptive != 0,
		p->num_samples, p->burn_in, p->init_gamma != 0, p->update_lambda != 0, p->update_alpha != 0,
		p->update_eta != 0, p->min_alpha, p->min_eta, p->max_epochs, p->max_iter_alpha, p->max_iter_eta,
		p->emp_bayes_threshold, p->verbosity);
}

}

extern "C" {

void* ref_create(int kind, int numWords, int numTopics, int numDocuments, const double* alpha, double eta) {
	Handle* h = new Handle();
	h->kind = kind;
	h->error = "";
	h->fast = false;
	ArrayXd a = Map<const ArrayXd>(alpha, numTopics);
	if(kind == TRLDA_KIND_ONLINE) {
		Injected<OnlineLDA>* m = new Injected<OnlineLDA>(numWords, numDocuments, a, eta);
		h->lda = m; h->slot = &m->injected;
	} else if(kind == TRLDA_KIND_BATCH) {
		Injected<BatchLDA>* m = new Injected<BatchLDA>(numWords, a, eta);
		h->lda = m; h->slot = &m->injected;
	} else {
		Injected<CumulativeLDA>* m = new Injected<CumulativeLDA>(numWords, a, eta);
		h->lda = m; h->slot = &m->injected;
	}
	return h;
}

/* Same object as ref_create, but lambda is set to eta everywhere instead of being drawn with rand(): for callers
 * that install their own lambda next (the benchmark), so that construction takes milliseconds instead of minutes. */
void* ref_create_fast(int kind, int numWords, int numTopics, int numDocuments, const double* alpha, double eta) {
	Handle* h = new Handle();
	h->kind = kind;
	h->error = "";
	h->fast = true;
	ArrayXd a = Map<const ArrayXd>(alpha, numTopics);
	if(kind == TRLDA_KIND_ONLINE) {
		Injected<OnlineLDA>* m = new Injected<OnlineLDA>(1, numDocuments, a, eta);
		m->resetLambda(numTopics, numWords, eta);
		h->lda = m; h->slot = &m->injected;
	} else if(kind == TRLDA_KIND_BATCH) {
		Injected<BatchLDA>* m = new Injected<BatchLDA>(1, a, eta);
		m->resetLambda(numTopics, numWords, eta);
		h->lda = m; h->slot = &m->injected;
	} else {
		Injected<CumulativeLDA>* m = new Injected<CumulativeLDA>(1, a, eta);
		m->resetLambda(numTopics, numWords, eta);
		h->lda = m; h->slot = &m->injected;
	}
	return h;
}

void ref_destroy(void* h_) {
	Handle* h = static_cast<Handle*>(h_);
	delete h->lda;
	delete h;
}

const char* ref_last_error(void* h_) { return static_cast<Handle*>(h_)->error; }

void ref_srand(unsigned seed) { srand(seed); }

int ref_num_topics(void* h) { return static_cast<Handle*>(h)->lda->numTopics(); }
int ref_num_words(void* h) { return static_cast<Handle*>(h)->lda->numWords(); }

void ref_get_lambda(void* h_, double* out) {
	LDA* lda = static_cast<Handle*>(h_)->lda;
	Map<ArrayXXd>(out, lda->numTopics(), lda->numWords()) = lda->lambda();
}

int ref_set_lambda(void* h_, const double* in, int rows, int cols) {
	Handle* h = static_cast<Handle*>(h_);
	try {
		h->lda->setLambda(Map<const ArrayXXd>(in, rows, cols));
	} catch(Exception& e) { h->error = e.message(); return 1; }
	return 0;
}

void ref_get_alpha(void* h_, double* out) {
	LDA* lda = static_cast<Handle*>(h_)->lda;
	Map<ArrayXd>(out, lda->numTopics()) = lda->alpha();
}

int ref_set_alpha(void* h_, const double* in, int n) {
	Handle* h = static_cast<Handle*>(h_);
	try {
		if(n == 1) h->lda->setAlpha(in[0]);
		else h->lda->setAlpha(ArrayXd(Map<const ArrayXd>(in, n)));
	} catch(Exception& e) { h->error = e.message(); return 1; }
	return 0;
}

double ref_get_eta(void* h_) { return static_cast<Handle*>(h_)->lda->eta(); }

int ref_set_eta(void* h_, double eta) {
	Handle* h = static_cast<Handle*>(h_);
	try { h->lda->setEta(eta); } catch(Exception& e) { h->error = e.message(); return 1; }
	return 0;
}

int ref_get_num_documents(void* h_) {
	Handle* h = static_cast<Handle*>(h_);
	return h->kind == TRLDA_KIND_ONLINE ? static_cast<OnlineLDA*>(h->lda)->numDocuments() : 0;
}

int ref_set_num_documents(void* h_, int n) {
	Handle* h = static_cast<Handle*>(h_);
	if(h->kind != TRLDA_KIND_ONLINE) return 1;
	try { static_cast<OnlineLDA*>(h->lda)->setNumDocuments(n); } catch(Exception& e) { h->error = e.message(); return 1; }
	return 0;
}

int ref_get_update_count(void* h_) {
	Handle* h = static_cast<Handle*>(h_);
	return h->kind == TRLDA_KIND_ONLINE ? static_cast<OnlineLDA*>(h->lda)->updateCount() : 0;
}

int ref_set_update_count(void* h_, int n) {
	Handle* h = static_cast<Handle*>(h_);
	if(h->kind != TRLDA_KIND_ONLINE) return 1;
	try { static_cast<OnlineLDA*>(h->lda)->setUpdateCount(n); } catch(Exception& e) { h->error = e.message(); return 1; }
	return 0;
}

/* LDA::updateVariables(documents, latents, parameters) — lda.cpp:142-156; latents == NULL -> lda.cpp:119-138 */
int ref_update_variables(void* h_, const trlda_docs* docs, const double* latents, int rows, long cols,
                         const trlda_params* params, double* gammaOut, double* sstatsOut)
{
	Handle* h = static_cast<Handle*>(h_);
	try {
		LDA::Documents documents = toDocuments(docs);
		LDA::Parameters parameters = toParameters(params);
		pair<ArrayXXd, ArrayXXd> results;
		if(latents)
			results = h->lda->updateVariables(documents, ArrayXXd(Map<const ArrayXXd>(latents, rows, cols)), parameters);
		else
			results = h->lda->updateVariables(documents, parameters);
		if(gammaOut)
			Map<ArrayXXd>(gammaOut, results.first.rows(), results.first.cols()) = results.first;
		if(sstatsOut)
			Map<ArrayXXd>(sstatsOut, results.second.rows(), results.second.cols()) = results.second;
	} catch(Exception& e) { h->error = e.message(); return 1; }
	return 0;
}

/* updateParameters with an optional injected initial gamma (see header comment) */
int ref_update_parameters(void* h_, const trlda_docs* docs, const trlda_params* params,
                          const double* gamma0, int rows, long cols, double* result)
{
	Handle* h = static_cast<Handle*>(h_);
	if(h->fast && params->adaptive) {
		h->error = "adaptive mode needs a model built by ref_create";
		return 1;
	}
	try {
		LDA::Documents documents = toDocuments(docs);
		LDA::Parameters parameters = toParameters(params);
		ArrayXXd g0;
		if(gamma0) {
			g0 = Map<const ArrayXXd>(gamma0, rows, cols);
			*h->slot = &g0;
		}
		double r = h->lda->updateParameters(documents, parameters);
		*h->slot = 0;
		if(result) *result = r;
	} catch(Exception& e) { *h->slot = 0; h->error = e.message(); return 1; }
	return 0;
}

/* LDA::lowerBound as written (including the mis-indexing at lda.cpp:334); only used to document that bug */
int ref_lower_bound(void* h_, const trlda_docs* docs, const trlda_params* params, int numDocuments,
                    const double* gamma0, int rows, long cols, double* result)
{
	Handle* h = static_cast<Handle*>(h_);
	try {
		LDA::Documents documents = toDocuments(docs);
		LDA::Parameters parameters = toParameters(params);
		ArrayXXd g0;
		if(gamma0) {
			g0 = Map<const ArrayXXd>(gamma0, rows, cols);
			*h->slot = &g0;
		}
		*result = h->lda->lowerBound(documents, parameters, numDocuments);
		*h->slot = 0;
	} catch(Exception& e) { *h->slot = 0; h->error = e.message(); return 1; }
	return 0;
}

double ref_digamma(double x) { return digamma(x); }
double ref_polygamma(int n, double x) { return polygamma(n, x); }
double ref_lngamma(double x) { return lngamma(x); }

/* sampleGamma(m, n, k) / k — the law of the initial lambda (lda.cpp:71) and gamma (lda.cpp:135) */
void ref_sample_gamma(int m, int n, int k, double* out) {
	Map<ArrayXXd>(out, m, n) = sampleGamma(m, n, k) / static_cast<double>(k);
}

}
