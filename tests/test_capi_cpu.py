"""
CPU tests of the boundary (`-m "not gpu"`): the C-ABI shared library loads, exports every symbol that
include/trlda_b200.h declares, its POD structs match the ctypes mirrors, and it fails loudly without a GPU.
No compute entry point is exercised here.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def capi():
	from trlda_b200 import build, capi
	build.build_library()
	return capi


def declared_symbols():
	header = open(os.path.join(ROOT, 'include', 'trlda_b200.h')).read()
	return sorted(set(re.findall(r'^TRLDA_API [\w\s\*]+?\b(trlda_\w+)\(', header, flags=re.M)))


def test_library_exports_every_declared_symbol(capi):
	handle = C.CDLL(capi.LIB_PATH)
	names = declared_symbols()
	assert len(names) >= 35
	for name in names:
		assert hasattr(handle, name), name


def test_ctypes_prototypes_cover_the_header(capi):
	assert sorted(capi.PROTOTYPES) == declared_symbols()


def test_params_defaults_follow_reference(capi):
	# LDA::Parameters::Parameters defaults, reference code/trlda/include/lda.h:56-77
	p = capi.default_params()
	expect = dict(
		inference_method=0, threshold=0.001, max_iter_inference=100, max_iter_tr=10, tau=100., kappa=.7, rho=-1.,
		adaptive=0, num_samples=1, burn_in=2, init_gamma=1, update_lambda=1, update_alpha=0, update_eta=0,
		min_alpha=1e-6, min_eta=1e-6, max_epochs=100, max_iter_alpha=10, max_iter_eta=20,
		emp_bayes_threshold=1e-8, verbosity=0)
	for key, value in expect.items():
		assert getattr(p, key) == value, key
	with pytest.raises(TypeError):
		capi.default_params(no_such_parameter=1)


def test_kernel_kind_names(capi):
	names = [capi.lib().trlda_kernel_kind_name(i).decode() for i in range(capi.NUM_KERNEL_KINDS)]
	assert names[:4] == ['rowsum', 'beta_prep', 'estep', 'scatter_mstep']
	assert len(set(names)) == capi.NUM_KERNEL_KINDS


def test_host_polygamma_known_answers(capi):
	# the reference's only known-answer test on this path, python/tests/utils_test.py:33-51
	values = {
		(0, .1): -10.423754940411, (0, 1.): -0.5772156649015329, (0, 120.): 4.7833192891185,
		(1, .01): 10001.6212135283, (1, .1): 101.433299150792758817215450106, (1, .4): 7.275356590529597,
		(1, 11.): 0.09516633568168575, (2, 14.): -0.005479465690312488}
	for (n, x), y in values.items():
		assert abs(capi.polygamma(n, x) - y) < 5e-8 * max(1., abs(y)), (n, x)


def test_host_polygamma_matches_reference_fixture(capi):
	case = np.load(os.path.join(ROOT, 'tests', 'golden', 'special.npz'))
	for name, n in (('digamma', 0), ('trigamma', 1), ('tetragamma', 2)):
		got = np.array([capi.polygamma(n, float(v)) for v in case['x']])
		err = np.max(np.abs(got - case[name]) / np.abs(case[name]))
		assert err < 5e-14, (name, err)


def test_csr_from_lists_roundtrip(capi):
	docs = [[(3, 1), (7, 2)], [], [(1, 5)]]
	csr = capi.CSR.from_lists(docs)
	assert csr.num_docs == 3 and csr.num_pairs == 3
	assert csr.doc_ptr.tolist() == [0, 2, 2, 3]
	assert csr.word_ids.tolist() == [3, 7, 1] and csr.counts.tolist() == [1, 2, 5]
	part = csr.slice(1, 3)
	assert part.doc_ptr.tolist() == [0, 0, 1] and part.word_ids.tolist() == [1]
	with pytest.raises(ValueError):
		capi.CSR([0, 2], [1], [1])


def test_create_fails_loudly_without_gpu(capi):
	"""no CPU fallback: constructing a model without a B200 raises with the library's message"""
	import subprocess
	import sys
	code = (
		'import sys; sys.path.insert(0, %r)\n'
		'from trlda_b200 import capi\n'
		'try:\n'
		'    capi.Model("online", 10, 3, 5)\n'
		'    print("CREATED")\n'
		'except RuntimeError as e:\n'
		'    print("RAISED", e)\n' % ROOT)
	env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
	out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True).stdout
	assert 'RAISED' in out and 'no CPU fallback' in out


def test_argument_validation_messages(capi):
	"""validation that happens before any device work keeps the reference's exception texts"""
	handle = C.c_void_p()
	a = np.array([.1, -.2, .3])
	status = capi.lib().trlda_create(0, 10, 3, 5, a.ctypes.data_as(C.POINTER(C.c_double)), .3, 0, 0, C.byref(handle))
	assert status == capi.ERR_ARG
	assert capi.lib().trlda_last_error(None) == b'Alpha should not be negative.'      # lda.h:148
	a = np.array([.1, .2, .3])
	status = capi.lib().trlda_create(0, 10, 3, -5, a.ctypes.data_as(C.POINTER(C.c_double)), .3, 0, 0, C.byref(handle))
	assert status == capi.ERR_ARG
	assert capi.lib().trlda_last_error(None) == b'The number of documents should not be negative.'   # onlinelda.h:58
