"""
The collapsed Gibbs E-step (`update_variables(..., inference_method='GIBBS')`, lda.cpp:224-293) on the device.

The reference's version cannot serve as an oracle: it initialises document i from theta.col(j) with j the TOKEN index
(lda.cpp:254, reads past the matrix once a document has more pairs than the minibatch has documents), races on sstats
(lda.cpp:284) and draws from rand(); its own test only checks that the call returns (onlinelda_test.py:99-109).
Parity is therefore pinned by properties every correct sampler must have: conservation of the counts, exactness when
the topics have disjoint supports, agreement of the sampled topic proportions with the variational posterior, and
reproducibility under `seed`.
"""
import numpy as np
import pytest

from common import random_docs

pytestmark = pytest.mark.gpu


def _word_totals(docs, V):
	total = np.zeros(V)
	np.add.at(total, docs.word_ids, docs.counts)
	return total


@pytest.mark.parametrize('precision', ['fp64', 'mixed'])
@pytest.mark.parametrize('num_samples', [1, 4])
def test_counts_are_conserved(precision, num_samples):
	from trlda_b200 import capi
	rng = np.random.default_rng(11)
	K, V, B = 24, 300, 40
	docs = capi.CSR.from_lists(random_docs(rng, B, V, 60) + [[]])          # one empty document
	model = capi.Model('online', V, K, 1000, .1, .2, precision=precision)
	capi.seed(5)
	theta, sstats = model.update_variables(docs, inference_method='GIBBS', num_samples=num_samples, burn_in=2)
	assert theta.shape == (K, B + 1) and sstats.shape == (K, V)
	assert np.all(theta > 0) and np.allclose(theta.sum(0), 1., rtol=0, atol=1e-12)
	assert np.all(sstats >= 0)
	# every token occurrence contributes num_samples times 1 / num_samples (lda.cpp:232, 278-286)
	assert np.allclose(sstats.sum(0), _word_totals(docs, V), rtol=0, atol=1e-9)
	# the empty document keeps its prior: theta ~ Dirichlet(alpha)
	model.close()


def test_disjoint_topics_are_recovered_exactly():
	"""topic k owns the words [10 k, 10 k + 10): expElogbeta is exactly zero elsewhere, so every assignment is forced"""
	from trlda_b200 import capi
	rng = np.random.default_rng(12)
	K, V, B = 4, 40, 30
	lam = np.full((K, V), 1e-16)
	for k in range(K):
		lam[k, 10 * k:10 * k + 10] = 50.
	model = capi.Model('online', V, K, 1000, .1, .2)
	model.lambdas = np.asfortranarray(lam)
	docs = capi.CSR.from_lists(random_docs(rng, B, V, 25))
	capi.seed(6)
	theta, sstats = model.update_variables(docs, inference_method='GIBBS', num_samples=3, burn_in=1)
	expected = np.zeros((K, V))
	totals = _word_totals(docs, V)
	for k in range(K):
		expected[k, 10 * k:10 * k + 10] = totals[10 * k:10 * k + 10]
	assert np.allclose(sstats, expected, rtol=0, atol=1e-9)
	# theta_d ~ Dirichlet(alpha + n_d) (lda.cpp:291) with n_d known here: standardised residuals of the K B draws
	conc = np.full((K, B), .1)
	for d in range(B):
		for j in range(docs.doc_ptr[d], docs.doc_ptr[d + 1]):
			conc[docs.word_ids[j] // 10, d] += docs.counts[j]
	total = conc.sum(0)
	mean = conc / total
	sigma = np.sqrt(mean * (1. - mean) / (total + 1.))
	stat = np.mean(((theta - mean) / sigma) ** 2)
	# ... against the same statistic of numpy's Dirichlet sampler at these concentrations (some are 0.1: far from normal)
	ref = np.array([np.mean(((np.stack([rng.dirichlet(conc[:, d]) for d in range(B)], 1) - mean) / sigma) ** 2) for _ in range(400)])
	assert np.quantile(ref, .002) < stat < np.quantile(ref, .998), (stat, np.quantile(ref, [.002, .5, .998]))
	model.close()


def test_agrees_with_variational_posterior():
	"""sstats averaged over many sweeps approach the variational sufficient statistics on a well separated model"""
	from trlda_b200 import capi
	from trlda_b200.synth import make_corpus
	K, V, B = 8, 400, 64
	rng = np.random.default_rng(13)
	lam = np.asfortranarray(rng.gamma(.05, 1., size=(V, K)).T * 200. + .01)    # sparse, peaked topics
	model = capi.Model('online', V, K, 1000, .1, .2)
	model.lambdas = lam
	docs = capi.CSR(*make_corpus(B, V, K, .1, .05, mean_length=120, seed=14))
	gamma, sstats_vi = model.update_variables(docs, max_iter=200, threshold=1e-6)
	capi.seed(7)
	theta, sstats_gibbs = model.update_variables(docs, inference_method='GIBBS', num_samples=64, burn_in=16)
	assert np.allclose(sstats_gibbs.sum(0), sstats_vi.sum(0), rtol=0, atol=1e-6)       # both conserve the counts
	assert np.corrcoef(sstats_gibbs.sum(1), sstats_vi.sum(1))[0, 1] > .98              # topic usage
	assert np.corrcoef(sstats_gibbs.ravel(), sstats_vi.ravel())[0, 1] > .95
	assert np.corrcoef(theta.ravel(), (gamma / gamma.sum(0)).ravel())[0, 1] > .8           # one Dirichlet draw per document
	model.close()


def test_seed_and_initial_theta():
	from trlda_b200 import capi
	rng = np.random.default_rng(15)
	K, V, B = 16, 200, 20
	docs = capi.CSR.from_lists(random_docs(rng, B, V, 40))
	model = capi.Model('online', V, K, 1000, .1, .2)
	out = []
	for s in (21, 21, 22):
		capi.seed(s)
		out.append(model.update_variables(docs, inference_method='GIBBS'))
	assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
	assert not np.array_equal(out[0][1], out[2][1])
	theta0 = rng.dirichlet(np.ones(K), size=B).T
	capi.seed(21)
	theta, sstats = model.update_variables(docs, latents=theta0, inference_method='GIBBS', num_samples=2, burn_in=0)
	assert np.allclose(sstats.sum(0), _word_totals(docs, V), rtol=0, atol=1e-9)
	with pytest.raises(RuntimeError, match='Initial theta has wrong dimensionality.'):          # lda.cpp:228
		model.update_variables(docs, latents=theta0[:, :-1], inference_method='GIBBS')
	with pytest.raises(RuntimeError):                                                          # parameter updates stay variational
		model.update_parameters(docs, inference_method=1)
	model.close()


def test_python_api_like_the_reference_test():
	# onlinelda_test.py:99-109
	import trlda.models
	W, K, D, N = 100, 20, 10, 100
	model = trlda.models.OnlineLDA(num_words=W, num_topics=K, num_documents=D)
	docs = model.sample(D, N)
	theta, sstats = model.update_variables(docs, inference_method='gibbs', num_samples=2, burn_in=2)
	assert theta.shape == (K, D) and sstats.shape == (K, W)
	assert abs(sstats.sum() - sum(c for doc in docs for _, c in doc)) < 1e-9
