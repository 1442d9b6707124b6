"""
Host-side helpers of trlda.utils that sit beside the hot path in the reference (utils.cpp:235-287, 294-330), written
with numpy.  LDA::sample itself runs on the device (csrc/sample.cu, trlda_sample).
"""
import numpy as np


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
