"""
ctypes drivers for the two CPU checkers — TEST INFRASTRUCTURE, never imported by the product.

  RefModel     oracle/_ref/libtrlda_ref.so, the unmodified reference core + oracle/ref_shim.cpp
  PortModel    oracle/libtrlda_oracle.so, the plain-C restatement (oracle/lda_oracle.c)

Both expose the same methods so a test can be parametrised over them.  Matrices are numpy float64 in
Fortran (column-major) order, K x V for lambda / sstats and K x B for gamma, exactly what the reference
binding returns (python/src/pyutils.cpp:25).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, '_ref', 'libtrlda_ref.so')
PORT_SO = os.path.join(HERE, 'libtrlda_oracle.so')

KIND = {'online': 0, 'batch': 1, 'cumulative': 2}


class Params(C.Structure):
	"""POD mirror of LDA::Parameters (code/trlda/include/lda.h:32-78) == trlda_params in include/trlda_b200.h"""
	_fields_ = [
		('inference_method', C.c_int32),
		('threshold', C.c_double),
		('max_iter_inference', C.c_int32),
		('max_iter_tr', C.c_int32),
		('tau', C.c_double),
		('kappa', C.c_double),
		('rho', C.c_double),
		('adaptive', C.c_int32),
		('num_samples', C.c_int32),
		('burn_in', C.c_int32),
		('init_gamma', C.c_int32),
		('update_lambda', C.c_int32),
		('update_alpha', C.c_int32),
		('update_eta', C.c_int32),
		('min_alpha', C.c_double),
		('min_eta', C.c_double),
		('max_epochs', C.c_int32),
		('max_iter_alpha', C.c_int32),
		('max_iter_eta', C.c_int32),
		('emp_bayes_threshold', C.c_double),
		('verbosity', C.c_int32)]

	DEFAULTS = dict(
		inference_method=0, threshold=0.001, max_iter_inference=100, max_iter_tr=10, tau=100., kappa=.7,
		rho=-1., adaptive=0, num_samples=1, burn_in=2, init_gamma=1, update_lambda=1, update_alpha=0,
		update_eta=0, min_alpha=1e-6, min_eta=1e-6, max_epochs=100, max_iter_alpha=10, max_iter_eta=20,
		emp_bayes_threshold=1e-8, verbosity=0)

	def __init__(self, **kwargs):
		super().__init__()
		values = dict(self.DEFAULTS)
		for key, value in kwargs.items():
			if key not in values:
				raise TypeError('unknown parameter ' + key)
			values[key] = value
		for key, value in values.items():
			setattr(self, key, int(value) if isinstance(self.DEFAULTS[key], int) else float(value))


class Docs(C.Structure):
	"""CSR view of LDA::Documents == trlda_docs in include/trlda_b200.h"""
	_fields_ = [
		('num_docs', C.c_int64),
		('doc_ptr', C.POINTER(C.c_int64)),
		('word_ids', C.POINTER(C.c_int32)),
		('counts', C.POINTER(C.c_int32))]


class CSR(object):
	"""Owns the three arrays of a minibatch and hands out the C view."""

	def __init__(self, doc_ptr, word_ids, counts):
		self.doc_ptr = np.ascontiguousarray(doc_ptr, dtype=np.int64)
		self.word_ids = np.ascontiguousarray(word_ids, dtype=np.int32)
		self.counts = np.ascontiguousarray(counts, dtype=np.int32)
		assert self.doc_ptr.ndim == 1 and self.doc_ptr.size >= 1
		assert self.word_ids.size == self.counts.size == self.doc_ptr[-1]
		self.c = Docs(
			self.doc_ptr.size - 1,
			self.doc_ptr.ctypes.data_as(C.POINTER(C.c_int64)),
			self.word_ids.ctypes.data_as(C.POINTER(C.c_int32)),
			self.counts.ctypes.data_as(C.POINTER(C.c_int32)))

	@classmethod
	def from_lists(cls, docs):
		"""docs: list of lists of (word_id, count) — the reference's document format"""
		doc_ptr = np.zeros(len(docs) + 1, dtype=np.int64)
		for i, doc in enumerate(docs):
			doc_ptr[i + 1] = doc_ptr[i] + len(doc)
		word_ids = np.fromiter((w for doc in docs for w, _ in doc), dtype=np.int32, count=int(doc_ptr[-1]))
		counts = np.fromiter((c for doc in docs for _, c in doc), dtype=np.int32, count=int(doc_ptr[-1]))
		return cls(doc_ptr, word_ids, counts)

	def to_lists(self):
		return [
			[(int(w), int(c)) for w, c in zip(
				self.word_ids[self.doc_ptr[d]:self.doc_ptr[d + 1]], self.counts[self.doc_ptr[d]:self.doc_ptr[d + 1]])]
			for d in range(self.num_docs)]

	@property
	def num_docs(self):
		return self.doc_ptr.size - 1

	@property
	def num_pairs(self):
		return int(self.doc_ptr[-1])

	def slice(self, begin, end):
		"""documents [begin, end) as a new CSR"""
		lo, hi = self.doc_ptr[begin], self.doc_ptr[end]
		return CSR(self.doc_ptr[begin:end + 1] - lo, self.word_ids[lo:hi], self.counts[lo:hi])


def _dptr(a):
	return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _fortran(a):
	return np.asfortranarray(a, dtype=np.float64)


def build(ref=True, port=True):
	"""(Re)build the checkers with oracle/Makefile.  The reference target is a no-op without /root/reference."""
	targets = (['libtrlda_oracle.so'] if port else []) + (['ref'] if ref else [])
	subprocess.run(['make', '-s', '-C', HERE] + targets, check=True)


def have_ref():
	return os.path.exists(REF_SO)


def have_port():
	return os.path.exists(PORT_SO)


_libs = {}


def _load(path):
	if path not in _libs:
		if not os.path.exists(path):
			build()
		_libs[path] = C.CDLL(path)
	return _libs[path]


def ref_lib():
	lib = _load(REF_SO)
	lib.ref_create.restype = C.c_void_p
	lib.ref_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_double]
	lib.ref_create_fast.restype = C.c_void_p
	lib.ref_create_fast.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_double]
	lib.ref_destroy.argtypes = [C.c_void_p]
	lib.ref_last_error.restype = C.c_char_p
	lib.ref_last_error.argtypes = [C.c_void_p]
	lib.ref_get_lambda.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
	lib.ref_set_lambda.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_int]
	lib.ref_get_alpha.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
	lib.ref_set_alpha.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
	lib.ref_get_eta.restype = C.c_double
	lib.ref_get_eta.argtypes = [C.c_void_p]
	lib.ref_set_eta.argtypes = [C.c_void_p, C.c_double]
	for name in ('ref_get_num_documents', 'ref_get_update_count', 'ref_num_topics', 'ref_num_words'):
		getattr(lib, name).argtypes = [C.c_void_p]
	lib.ref_set_num_documents.argtypes = [C.c_void_p, C.c_int]
	lib.ref_set_update_count.argtypes = [C.c_void_p, C.c_int]
	lib.ref_update_variables.argtypes = [
		C.c_void_p, C.POINTER(Docs), C.POINTER(C.c_double), C.c_int, C.c_long, C.POINTER(Params),
		C.POINTER(C.c_double), C.POINTER(C.c_double)]
	lib.ref_update_parameters.argtypes = [
		C.c_void_p, C.POINTER(Docs), C.POINTER(Params), C.POINTER(C.c_double), C.c_int, C.c_long,
		C.POINTER(C.c_double)]
	lib.ref_lower_bound.argtypes = [
		C.c_void_p, C.POINTER(Docs), C.POINTER(Params), C.c_int, C.POINTER(C.c_double), C.c_int, C.c_long,
		C.POINTER(C.c_double)]
	lib.ref_digamma.restype = C.c_double
	lib.ref_digamma.argtypes = [C.c_double]
	lib.ref_polygamma.restype = C.c_double
	lib.ref_polygamma.argtypes = [C.c_int, C.c_double]
	lib.ref_lngamma.restype = C.c_double
	lib.ref_lngamma.argtypes = [C.c_double]
	lib.ref_sample_gamma.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
	lib.ref_srand.argtypes = [C.c_uint]
	return lib


class OracleModelStruct(C.Structure):
	_fields_ = [
		('kind', C.c_int), ('K', C.c_int), ('V', C.c_int),
		('lambda_', C.POINTER(C.c_double)), ('alpha', C.POINTER(C.c_double)), ('eta', C.c_double),
		('num_documents', C.c_int), ('update_counter', C.c_int),
		('ada_rho', C.c_double), ('ada_tau', C.c_double), ('ada_sq_norm', C.c_double),
		('ada_gradient', C.POINTER(C.c_double)), ('psi_gamma_diff', C.POINTER(C.c_double)),
		('cum_num_documents', C.c_int)]


def port_lib():
	lib = _load(PORT_SO)
	lib.oracle_digamma.restype = C.c_double
	lib.oracle_digamma.argtypes = [C.c_double]
	lib.oracle_zeta.restype = C.c_double
	lib.oracle_zeta.argtypes = [C.c_double, C.c_double]
	lib.oracle_polygamma.restype = C.c_double
	lib.oracle_polygamma.argtypes = [C.c_int, C.c_double]
	lib.oracle_sample_gamma.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
	lib.oracle_create.restype = C.POINTER(OracleModelStruct)
	lib.oracle_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_double]
	lib.oracle_destroy.argtypes = [C.POINTER(OracleModelStruct)]
	lib.oracle_update_variables.argtypes = [
		C.POINTER(OracleModelStruct), C.POINTER(Docs), C.POINTER(C.c_double), C.c_int, C.c_double,
		C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
	lib.oracle_update_variables_gibbs.argtypes = [
		C.POINTER(OracleModelStruct), C.POINTER(Docs), C.POINTER(C.c_double), C.c_int, C.c_int,
		C.POINTER(C.c_double), C.POINTER(C.c_double)]
	lib.oracle_update_parameters.restype = C.c_double
	lib.oracle_update_parameters.argtypes = [
		C.POINTER(OracleModelStruct), C.POINTER(Docs), C.POINTER(Params), C.POINTER(C.c_double),
		C.POINTER(C.c_double), C.POINTER(C.c_double)]
	lib.oracle_lower_bound.restype = C.c_double
	lib.oracle_lower_bound.argtypes = [
		C.POINTER(OracleModelStruct), C.POINTER(Docs), C.POINTER(C.c_double), C.POINTER(Params), C.c_int,
		C.POINTER(C.c_double)]
	return lib


def _alpha_vector(alpha, K):
	alpha = np.asarray(alpha, dtype=np.float64).ravel()
	if alpha.size == 1:
		alpha = np.full(K, float(alpha[0]))
	assert alpha.size == K
	return np.ascontiguousarray(alpha)


class RefModel(object):
	"""The compiled reference (TRLDA::OnlineLDA / BatchLDA / CumulativeLDA) behind oracle/ref_shim.cpp."""

	backend = 'reference'

	def __init__(self, kind, num_words, num_topics, num_documents=0, alpha=.1, eta=.3, seed=0, fast_init=False):
		"""fast_init: lambda starts as eta everywhere instead of the constructor's Gamma(100, 1/100) draw, which
		costs 100 K V calls to rand() (minutes at K V = 1e8); for callers that install their own lambda next."""
		self.lib = ref_lib()
		self.kind, self.V, self.K = kind, num_words, num_topics
		a = _alpha_vector(alpha, num_topics)
		self.lib.ref_srand(seed)
		create = self.lib.ref_create_fast if fast_init else self.lib.ref_create
		self.h = create(KIND[kind], num_words, num_topics, num_documents, _dptr(a), eta)

	def __del__(self):
		if getattr(self, 'h', None):
			self.lib.ref_destroy(self.h)
			self.h = None

	def _check(self, status):
		if status:
			raise RuntimeError(self.lib.ref_last_error(self.h).decode())

	@property
	def lambdas(self):
		out = np.empty((self.K, self.V), order='F')
		self.lib.ref_get_lambda(self.h, _dptr(out))
		return out

	@lambdas.setter
	def lambdas(self, value):
		value = _fortran(value)
		self._check(self.lib.ref_set_lambda(self.h, _dptr(value), value.shape[0], value.shape[1]))

	@property
	def alpha(self):
		out = np.empty(self.K)
		self.lib.ref_get_alpha(self.h, _dptr(out))
		return out

	@alpha.setter
	def alpha(self, value):
		value = np.ascontiguousarray(np.asarray(value, dtype=np.float64).ravel())
		self._check(self.lib.ref_set_alpha(self.h, _dptr(value), value.size))

	@property
	def eta(self):
		return self.lib.ref_get_eta(self.h)

	@eta.setter
	def eta(self, value):
		self._check(self.lib.ref_set_eta(self.h, value))

	@property
	def num_documents(self):
		return self.lib.ref_get_num_documents(self.h)

	@num_documents.setter
	def num_documents(self, value):
		self._check(self.lib.ref_set_num_documents(self.h, value))

	@property
	def update_count(self):
		return self.lib.ref_get_update_count(self.h)

	@update_count.setter
	def update_count(self, value):
		self._check(self.lib.ref_set_update_count(self.h, value))

	def update_variables(self, docs, latents=None, max_iter=100, threshold=.001, want_sstats=True):
		params = Params(max_iter_inference=max_iter, threshold=threshold)
		gamma = np.empty((self.K, docs.num_docs), order='F')
		sstats = np.empty((self.K, self.V), order='F') if want_sstats else None
		rows, cols = (0, 0)
		if latents is not None:
			latents = _fortran(latents)
			rows, cols = latents.shape
		self._check(self.lib.ref_update_variables(
			self.h, C.byref(docs.c), _dptr(latents), rows, cols, C.byref(params), _dptr(gamma), _dptr(sstats)))
		return gamma, sstats

	def update_parameters(self, docs, gamma0=None, lambda0=None, seed=None, **kwargs):
		"""gamma0 is injected through the virtual-override seam; lambda0 (Cumulative) cannot be injected into
		the reference, use `seed` + sample_gamma() to reproduce its internal draw instead."""
		assert lambda0 is None
		params = Params(**kwargs)
		result = C.c_double(0.)
		rows, cols = (0, 0)
		if gamma0 is not None:
			gamma0 = _fortran(gamma0)
			rows, cols = gamma0.shape
		if seed is not None:
			self.lib.ref_srand(seed)
		self._check(self.lib.ref_update_parameters(
			self.h, C.byref(docs.c), C.byref(params), _dptr(gamma0), rows, cols, C.byref(result)))
		return result.value

	def lower_bound_as_written(self, docs, gamma0, num_documents=-1, max_iter=100):
		"""LDA::lowerBound including the lda.cpp:334 bug — documentation only"""
		params = Params(max_iter_inference=max_iter)
		gamma0 = _fortran(gamma0)
		result = C.c_double(0.)
		self._check(self.lib.ref_lower_bound(
			self.h, C.byref(docs.c), C.byref(params), num_documents, _dptr(gamma0), gamma0.shape[0],
			gamma0.shape[1], C.byref(result)))
		return result.value

	def sample_gamma(self, m, n, k=100, seed=None):
		"""sampleGamma(m, n, k) / k from libc rand(), after an optional srand(seed)"""
		if seed is not None:
			self.lib.ref_srand(seed)
		out = np.empty((m, n), order='F')
		self.lib.ref_sample_gamma(m, n, k, _dptr(out))
		return out


class PortModel(object):
	"""The plain-C restatement, oracle/lda_oracle.c."""

	backend = 'port'

	def __init__(self, kind, num_words, num_topics, num_documents=0, alpha=.1, eta=.3, seed=0):
		self.lib = port_lib()
		self.kind, self.V, self.K = kind, num_words, num_topics
		a = _alpha_vector(alpha, num_topics)
		C.CDLL(None).srand(seed)
		self.m = self.lib.oracle_create(KIND[kind], num_words, num_topics, num_documents, _dptr(a), eta)

	def __del__(self):
		if getattr(self, 'm', None):
			self.lib.oracle_destroy(self.m)
			self.m = None

	def _lambda_view(self):
		return np.ctypeslib.as_array(self.m.contents.lambda_, shape=(self.V, self.K)).T

	@property
	def lambdas(self):
		return np.asfortranarray(self._lambda_view().copy())

	@lambdas.setter
	def lambdas(self, value):
		value = np.asarray(value, dtype=np.float64)
		if value.shape != (self.K, self.V):
			raise RuntimeError('Lambda has wrong dimensionality.')
		self._lambda_view()[...] = value

	@property
	def alpha(self):
		return np.ctypeslib.as_array(self.m.contents.alpha, shape=(self.K,)).copy()

	@alpha.setter
	def alpha(self, value):
		value = np.asarray(value, dtype=np.float64).ravel()
		if value.size == 1:
			value = np.full(self.K, value[0])
		if value.size != self.K:
			raise RuntimeError('Alpha has wrong dimensionality.')
		np.ctypeslib.as_array(self.m.contents.alpha, shape=(self.K,))[...] = value

	@property
	def eta(self):
		return self.m.contents.eta

	@eta.setter
	def eta(self, value):
		self.m.contents.eta = value

	@property
	def num_documents(self):
		return self.m.contents.num_documents

	@num_documents.setter
	def num_documents(self, value):
		self.m.contents.num_documents = value

	@property
	def update_count(self):
		return self.m.contents.update_counter

	@update_count.setter
	def update_count(self, value):
		self.m.contents.update_counter = value

	def update_variables(self, docs, latents=None, max_iter=100, threshold=.001, want_sstats=True,
			want_iterations=False):
		if latents is None:
			latents = self.sample_gamma(self.K, docs.num_docs)
		latents = _fortran(latents)
		if latents.shape != (self.K, docs.num_docs):
			raise RuntimeError('Initial gamma has wrong dimensionality.')
		gamma = np.empty((self.K, docs.num_docs), order='F')
		sstats = np.empty((self.K, self.V), order='F') if want_sstats else None
		iterations = np.zeros(docs.num_docs, dtype=np.int32)
		self.lib.oracle_update_variables(
			self.m, C.byref(docs.c), _dptr(latents), max_iter, threshold, _dptr(gamma), _dptr(sstats),
			iterations.ctypes.data_as(C.POINTER(C.c_int)))
		if want_iterations:
			return gamma, sstats, iterations
		return gamma, sstats

	def update_variables_gibbs(self, docs, theta0=None, num_samples=1, burn_in=2, seed=None):
		"""lda.cpp:224-293 without its index bug and race (see lda_oracle.c); theta0 defaults to Dirichlet(1) columns
		(lda.cpp:123-126) drawn with numpy"""
		if theta0 is None:
			theta0 = np.random.dirichlet(np.ones(self.K), size=docs.num_docs).T
		theta0 = _fortran(theta0)
		if theta0.shape != (self.K, docs.num_docs):
			raise RuntimeError('Initial theta has wrong dimensionality.')
		if seed is not None:
			C.CDLL(None).srand(seed)
		theta = np.empty((self.K, docs.num_docs), order='F')
		sstats = np.empty((self.K, self.V), order='F')
		self.lib.oracle_update_variables_gibbs(self.m, C.byref(docs.c), _dptr(theta0), num_samples, burn_in, _dptr(theta), _dptr(sstats))
		return theta, sstats

	def update_parameters(self, docs, gamma0=None, lambda0=None, seed=None, want_gamma=False, **kwargs):
		params = Params(**kwargs)
		if gamma0 is not None:
			gamma0 = _fortran(gamma0)
		if lambda0 is not None:
			lambda0 = _fortran(lambda0)
		if seed is not None:
			C.CDLL(None).srand(seed)
		gamma = np.empty((self.K, docs.num_docs), order='F') if want_gamma else None
		result = self.lib.oracle_update_parameters(
			self.m, C.byref(docs.c), C.byref(params), _dptr(gamma0), _dptr(lambda0), _dptr(gamma))
		if want_gamma:
			return result, gamma
		return result

	def lower_bound(self, docs, gamma0, num_documents=-1, max_iter=100, threshold=.001):
		params = Params(max_iter_inference=max_iter, threshold=threshold)
		gamma0 = _fortran(gamma0)
		per_doc = np.empty(docs.num_docs)
		if self.kind == 'online' and num_documents < 0:
			num_documents = self.num_documents            # OnlineLDA::lowerBound, onlinelda.cpp:184-191
		total = self.lib.oracle_lower_bound(
			self.m, C.byref(docs.c), _dptr(gamma0), C.byref(params), num_documents, _dptr(per_doc))
		return total, per_doc

	def sample_gamma(self, m, n, k=100, seed=None):
		if seed is not None:
			C.CDLL(None).srand(seed)
		out = np.empty((m, n), order='F')
		self.lib.oracle_sample_gamma(m, n, k, _dptr(out))
		return out
