"""
Generates the golden fixtures in this directory from the UNMODIFIED reference core
(oracle/_ref/libtrlda_ref.so, built by oracle/Makefile from /root/reference).  Run in the container that has
/root/reference; the .npz files are committed so that the GPU box (which has no reference) can check both the
plain-C oracle and the CUDA path against real reference outputs.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.pyoracle import CSR, RefModel, ref_lib   # noqa: E402


def random_docs(rng, B, V, max_len, empty=(), duplicates=()):
	"""documents in the reference's list-of-(id, count) form; some empty, some with repeated word ids"""
	docs = []
	for d in range(B):
		if d in empty:
			docs.append([])
			continue
		n = 1 + rng.integers(max_len)
		ids = rng.permutation(V)[:n]
		doc = [(int(w), int(1 + rng.integers(9))) for w in ids]
		if d in duplicates:
			doc += [(doc[0][0], 2), (doc[-1][0], 1)]
		docs.append(doc)
	return docs


def gamma(rng, rows, cols):
	return np.asfortranarray(rng.gamma(100., 1. / 100., size=(cols, rows)).T)


def online_case(name, seed, V, K, B, D, alpha, eta, **kwargs):
	rng = np.random.default_rng(seed)
	docs = CSR.from_lists(random_docs(rng, B, V, min(V, 40), empty=(2,), duplicates=(4,)))
	lam0, g0 = gamma(rng, K, V), gamma(rng, K, B)
	model = RefModel('online', V, K, D, alpha, eta)
	model.lambdas = lam0
	model.update_count = 3
	e_gamma, e_sstats = model.update_variables(docs, g0, max_iter=kwargs.get('max_iter_inference', 20))
	rho = model.update_parameters(docs, gamma0=g0, **kwargs)
	np.savez_compressed(
		os.path.join(HERE, name), kind='online', V=V, K=K, B=B, D=D, alpha0=np.full(K, alpha), eta0=eta,
		doc_ptr=docs.doc_ptr, word_ids=docs.word_ids, counts=docs.counts, lambda0=lam0, gamma0=g0, update_count0=3,
		params=np.array(sorted(kwargs.items()), dtype=object), estep_gamma=e_gamma, estep_sstats=e_sstats,
		rho=rho, lambda1=model.lambdas, alpha1=model.alpha, eta1=model.eta, update_count1=model.update_count)


def batch_case(name, seed, V, K, B, alpha, eta, **kwargs):
	rng = np.random.default_rng(seed)
	docs = CSR.from_lists(random_docs(rng, B, V, min(V, 30), empty=(0,), duplicates=(1, 3)))
	lam0, g0 = gamma(rng, K, V), gamma(rng, K, B)
	model = RefModel('batch', V, K, 0, alpha, eta)
	model.lambdas = lam0
	result = model.update_parameters(docs, gamma0=g0, **kwargs)
	np.savez_compressed(
		os.path.join(HERE, name), kind='batch', V=V, K=K, B=B, D=0, alpha0=np.full(K, alpha), eta0=eta,
		doc_ptr=docs.doc_ptr, word_ids=docs.word_ids, counts=docs.counts, lambda0=lam0, gamma0=g0,
		params=np.array(sorted(kwargs.items()), dtype=object), rho=result, lambda1=model.lambdas,
		alpha1=model.alpha, eta1=model.eta)


def cumulative_case(name, seed, V, K, B, alpha, eta, **kwargs):
	rng = np.random.default_rng(seed)
	model = RefModel('cumulative', V, K, 0, alpha, eta)
	lam_start = model.lambdas                      # == eta everywhere (cumulativelda.cpp:30)
	out = dict(kind='cumulative', V=V, K=K, B=B, D=0, alpha0=np.full(K, alpha), eta0=eta, lambda0=lam_start,
		params=np.array(sorted(kwargs.items()), dtype=object))
	# two consecutive calls: the alpha statistics accumulate across calls (cumulativelda.cpp:84-85)
	for call in range(2):
		docs = CSR.from_lists(random_docs(rng, B, V, min(V, 25), duplicates=(2,)))
		g0 = gamma(rng, K, B)
		# the reference re-randomises lambda from rand() inside the call (cumulativelda.cpp:60); replay its draw
		lam_rand = model.sample_gamma(K, V, 100, seed=seed + call)
		result = model.update_parameters(docs, gamma0=g0, seed=seed + call, **kwargs)
		out.update({
			'doc_ptr_%d' % call: docs.doc_ptr, 'word_ids_%d' % call: docs.word_ids, 'counts_%d' % call: docs.counts,
			'gamma0_%d' % call: g0, 'lambda_rand_%d' % call: lam_rand, 'rho_%d' % call: result,
			'lambda1_%d' % call: model.lambdas, 'alpha1_%d' % call: model.alpha})
	np.savez_compressed(os.path.join(HERE, name), **out)


def special_case(name):
	lib = ref_lib()
	x = np.concatenate([
		np.logspace(-12, 3, 61), [.1, 1., 2., 5., 9.999, 10., 10.001, 120., 1e5, 1e17, 3e17], np.arange(1., 12.)])
	np.savez_compressed(
		os.path.join(HERE, name), x=x,
		digamma=np.array([lib.ref_digamma(v) for v in x]),
		trigamma=np.array([lib.ref_polygamma(1, v) for v in x]),
		tetragamma=np.array([lib.ref_polygamma(2, v) for v in x]),
		lngamma=np.array([lib.ref_lngamma(v) for v in x]))


if __name__ == '__main__':
	special_case('special.npz')
	online_case('online_tr.npz', 11, V=60, K=12, B=9, D=500, alpha=.1, eta=.2,
		max_iter_tr=3, max_iter_inference=20, kappa=.7, tau=10., update_alpha=1, update_eta=1)
	online_case('online_sgd.npz', 12, V=45, K=33, B=7, D=200, alpha=.3, eta=.4,
		max_iter_tr=0, max_iter_inference=50, rho=.25, update_alpha=1)
	online_case('online_adaptive.npz', 13, V=50, K=8, B=12, D=1000, alpha=.1, eta=.3,
		max_iter_tr=2, max_iter_inference=20, adaptive=1, init_gamma=0)
	batch_case('batch.npz', 21, V=40, K=7, B=15, alpha=.2, eta=.3,
		max_epochs=3, max_iter_inference=30, update_alpha=1, update_eta=1)
	cumulative_case('cumulative.npz', 31, V=30, K=5, B=10, alpha=.1, eta=.25,
		max_epochs=2, max_iter_inference=25, update_alpha=1, threshold=1e-4)
	print('wrote', sorted(f for f in os.listdir(HERE) if f.endswith('.npz')))
