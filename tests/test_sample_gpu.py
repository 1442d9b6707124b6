"""
LDA::sample on the device (trlda_sample; reference lda.cpp:88-115).  The reference never tests its sampler beyond using
it (onlinelda_test.py draws its corpora with it); like the reference's own sampler tests (utils_test.py:55-66: a
Kolmogorov-Smirnov test against numpy's generator) the laws are checked statistically, and the structure exactly.
"""
import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def capi():
	from trlda_b200 import capi
	capi.lib()
	return capi


def test_document_lengths_follow_poisson(capi):
	capi.seed(11)
	model = capi.Model('online', 500, 20, 1000, .1, .2)
	docs = model.sample(4000, 37.)
	n = np.diff(docs.doc_ptr)
	assert np.all(docs.counts == 1)                               # every word is emitted as (word, 1): lda.cpp:108
	assert abs(n.mean() - 37.) < .5 and abs(n.var() - 37.) < 3.
	# KS against numpy's Poisson generator, as utils_test.py:55-66 does for the Dirichlet sampler
	assert stats.ks_2samp(n, np.random.default_rng(0).poisson(37., size=4000))[1] > 1e-6
	assert docs.word_ids.min() >= 0 and docs.word_ids.max() < 500


def test_words_follow_the_topics(capi):
	"""lambda puts (almost) all mass of topic k on the words k, k + K, ...: with a tiny alpha a document uses one topic,
	so all its words are congruent modulo K; over the corpus every topic appears"""
	capi.seed(12)
	K, V = 8, 400
	lam = np.full((K, V), 1e-3)
	for k in range(K):
		lam[k, k::K] = 50.
	model = capi.Model('online', V, K, 1000, .001, .2)
	model.lambdas = np.asfortranarray(lam)
	docs = model.sample(600, 30.)
	ptr = docs.doc_ptr
	pure, used = 0, set()
	for d in range(600):
		w = docs.word_ids[ptr[d]:ptr[d + 1]]
		if w.size == 0:
			continue
		topics = np.bincount(w % K, minlength=K)
		pure += topics.max() >= .9 * w.size
		used.add(int(topics.argmax()))
	assert pure > 560 and used == set(range(K))
	# within a topic the words are uniform over its V / K words (all lambda equal): chi-square on topic 0's words
	w0 = docs.word_ids[docs.word_ids % K == 0] // K
	counts = np.bincount(w0, minlength=V // K)
	assert stats.chisquare(counts)[1] > 1e-6


def test_theta_follows_dirichlet(capi):
	"""every word identifies its topic (one word per topic): the topic proportions of long documents are Dirichlet(alpha)"""
	capi.seed(13)
	K = 5
	lam = np.full((K, K), 1e-6) + np.eye(K) * 1e3
	alpha = np.array([.5, 1., 2., 4., .25])
	model = capi.Model('online', K, K, 1000, alpha, .2)
	model.lambdas = np.asfortranarray(lam)
	docs = model.sample(1500, 400.)
	ptr = docs.doc_ptr
	theta = np.stack([np.bincount(docs.word_ids[ptr[d]:ptr[d + 1]], minlength=K) / max(ptr[d + 1] - ptr[d], 1) for d in range(1500)])
	# the same law drawn with numpy: theta ~ Dirichlet(alpha), then the document's words (a multinomial of its length)
	rng = np.random.default_rng(1)
	lengths = np.diff(ptr)
	want = np.stack([rng.multinomial(n, p) / max(n, 1) for n, p in zip(lengths, rng.dirichlet(alpha, size=1500))])
	for k in range(K):
		assert stats.ks_2samp(theta[:, k], want[:, k])[1] > 1e-6
	assert np.allclose(theta.mean(0), alpha / alpha.sum(), atol=.02)


def test_collapsed_form_and_reproducibility(capi):
	K, V = 16, 300
	model = capi.Model('online', V, K, 1000, .1, .2)
	capi.seed(14)
	a = model.sample(200, 80., collapse=True)
	capi.seed(14)
	model2 = capi.Model('online', V, K, 1000, .1, .2)
	model2.lambdas = model.lambdas
	# the stream counter advanced identically on both models after seeding: same seed, same lambda, same corpus
	ptr = a.doc_ptr
	for d in range(200):
		w = a.word_ids[ptr[d]:ptr[d + 1]]
		assert np.all(np.diff(w) > 0)                             # unique, sorted by id (load_documents.py:41-44)
	assert np.all(a.counts >= 1)
	tokens = np.add.reduceat(a.counts, ptr[:-1][np.diff(ptr) > 0])
	assert abs(tokens.mean() - 80.) < 2.
	# the collapsed corpus trains: one update_parameters step stays finite
	rho = model.update_parameters(a, max_iter_inference=20)
	assert np.isfinite(rho) and np.all(np.isfinite(model.lambdas))
	assert model.sample(0, 10.).num_docs == 0
