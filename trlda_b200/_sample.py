"""
Host-side samplers that sit beside the hot path in the reference (code/trlda/src/lda.cpp:88-115,
utils.cpp:235-287, 294-330).  They are utilities for tests and examples, written with numpy; nothing here is
timed or accelerated.
"""
import numpy as np


def sample_documents(lambdas, alpha, num_documents, length):
	"""LDA::sample: beta_k ~ Dirichlet(lambda_k), theta ~ Dirichlet(alpha), Poisson(length) words per document,
	every word emitted as (word_id, 1) — repeated ids are possible, exactly like the reference."""
	lambdas = np.asarray(lambdas, dtype=np.float64)
	alpha = np.asarray(alpha, dtype=np.float64).ravel()
	K, V = lambdas.shape
	beta = np.random.standard_gamma(np.maximum(lambdas, 1e-300))
	beta = np.maximum(beta, 1e-300)
	beta /= beta.sum(1, keepdims=True)
	cdf = np.cumsum(beta, axis=1)
	cdf /= cdf[:, -1:]
	documents = []
	for n in np.random.poisson(length, size=num_documents):
		theta = np.random.standard_gamma(np.maximum(alpha, 1e-300))
		theta = np.maximum(theta, 1e-300)
		theta /= theta.sum()
		topics = np.random.choice(K, size=n, p=theta)
		u = np.random.rand(n)
		words = [min(int(np.searchsorted(cdf[k], v, side='right')), V - 1) for k, v in zip(topics, u)]
		documents.append([(w, 1) for w in words])
	return documents


def random_select(k, n):
	"""k distinct indices out of range(n) (reference: utils.cpp randomSelect)."""
	if k > n:
		raise Exception('k must be smaller than n.')
	if k < 0 or n < 0:
		raise Exception('n and k must be non-negative.')
	return sorted(int(i) for i in np.random.permutation(n)[:k])


def sample_dirichlet(m, n=1, alpha=.1):
	"""m x n array whose columns are draws from a symmetric Dirichlet(alpha) (reference: utils.cpp:252-264)."""
	sample = np.maximum(np.random.standard_gamma(alpha, size=(m, n)), 1e-300)
	return sample / sample.sum(0, keepdims=True)
