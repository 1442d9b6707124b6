"""CPU tests of the Python-facing package: import paths of the reference, the document text format, host helpers."""
import numpy as np
import pytest


@pytest.fixture(scope='module', autouse=True)
def built():
	from trlda_b200 import build
	build.build_all()


def test_reference_import_paths():
	# code/trlda/python/{__init__,models/__init__,utils/__init__}.py
	import trlda
	from trlda.models import BatchLDA, CumulativeLDA, Distribution, LDA, OnlineLDA
	from trlda.utils import load_documents, polygamma, random_select, sample_dirichlet
	assert callable(trlda.seed)
	assert issubclass(OnlineLDA, LDA) and issubclass(BatchLDA, LDA) and issubclass(CumulativeLDA, LDA)
	assert issubclass(LDA, Distribution)
	with pytest.raises(NotImplementedError):
		LDA()
	with pytest.raises(NotImplementedError):
		Distribution()
	assert polygamma(1, .1) == pytest.approx(101.433299150792758817215450106, rel=1e-13)      # utils_test.py:39
	assert np.allclose(polygamma(1, np.asarray([.01, .1])), [10001.6212135283, 101.433299150792758], rtol=1e-10)


def test_constructor_without_gpu_raises_runtime_error():
	import os
	import subprocess
	import sys
	root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
	code = (
		'import sys; sys.path.insert(0, %r)\n'
		'from trlda.models import OnlineLDA\n'
		'try:\n'
		'    OnlineLDA(num_words=10, num_topics=3, num_documents=5)\n'
		'    print("CREATED")\n'
		'except RuntimeError as e:\n'
		'    print("RAISED", e)\n' % root)
	out = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, CUDA_VISIBLE_DEVICES=''),
		capture_output=True, text=True, timeout=120).stdout
	assert 'RAISED' in out and 'no CPU fallback' in out


def test_load_documents_format(tmp_path):
	# python/utils/load_documents.py:6-69: "6 5600:2 293:1 ..." — the first field is ignored
	from trlda.utils import load_documents, load_documents_csr
	path = tmp_path / 'docs.txt'
	lines = ['6 5600:2 293:1 5548:1 2577:1 3733:3 2677:2', '0', '2 1:1 7:4', '1 9:9', '3 4:1 5:1 6:2']
	path.write_text('\n'.join(lines) + '\n')
	docs = load_documents(str(path))
	assert docs[0] == [(5600, 2), (293, 1), (5548, 1), (2577, 1), (3733, 3), (2677, 2)]
	assert docs[1] == [] and docs[2] == [(1, 1), (7, 4)] and len(docs) == 5
	batches = list(load_documents(str(path), batch_size=2))
	assert [len(b) for b in batches] == [2, 2, 1]                 # final partial batch is yielded (line 63)
	assert batches[1] == [[(1, 1), (7, 4)], [(9, 9)]]
	np.random.seed(0)
	stochastic = list(load_documents(str(path), batch_size=2, stochastic=True))
	assert sum(len(b) for b in stochastic) == 5
	ptr, ids, cts = load_documents_csr(str(path))
	assert ptr.tolist() == [0, 6, 6, 8, 9, 12] and ids.dtype == np.int32
	assert ids[:6].tolist() == [5600, 293, 5548, 2577, 3733, 2677] and cts[-3:].tolist() == [1, 1, 2]
	csr_batches = list(load_documents_csr(str(path), batch_size=2))
	assert [len(b[0]) - 1 for b in csr_batches] == [2, 2, 1]


def test_host_samplers():
	from trlda.utils import random_select, sample_dirichlet
	idx = random_select(5, 20)
	assert len(set(idx)) == 5 and all(0 <= i < 20 for i in idx)
	with pytest.raises(Exception):
		random_select(10, 4)                                          # utils_test.py:29
	x = sample_dirichlet(10, 7, .5)
	assert x.shape == (10, 7) and np.allclose(x.sum(0), 1.)


def test_list_walk_of_the_binding():
	"""list[list[(word, count)]] -> CSR, the conversion every method of the binding applies to `docs` (replaces
	PyList_ToDocuments, ldainterface.cpp:152-190): the threaded fast path for plain tuples of small ints, the general
	loop for anything else, the same arrays either way."""
	import numpy as np
	from trlda_b200 import _trlda
	rng = np.random.default_rng(3)
	lengths = rng.integers(0, 120, size=2000)                    # > 65536 pairs: the threaded path; empty documents too
	docs = [[(int(w), int(c)) for w, c in zip(rng.integers(0, 50000, n), rng.integers(1, 9, n))] for n in lengths]
	ptr, ids, cts = _trlda._pack_documents(docs)
	assert ptr.dtype == np.int64 and ids.dtype == np.int32 and cts.dtype == np.int32
	assert np.array_equal(ptr, np.concatenate([[0], np.cumsum(lengths)]))
	assert np.array_equal(ids, np.array([w for d in docs for w, _ in d], dtype=np.int32))
	assert np.array_equal(cts, np.array([c for d in docs for _, c in d], dtype=np.int32))

	mixed = [list(d) for d in docs]
	mixed[7] = [(np.int64(w), np.int32(c)) for w, c in mixed[7]] or [(np.int64(3), np.int32(1))]    # not exact ints
	mixed[11] = mixed[11] + [(2 ** 31 - 1, 5)]                  # the largest value that fits
	got = _trlda._pack_documents(mixed)
	want = _trlda._pack_documents([[(int(w), int(c)) for w, c in d] for d in mixed])
	assert all(np.array_equal(a, b) for a, b in zip(got, want))

	small = _trlda._pack_documents(docs[:5])                     # below the threshold: the serial loop
	assert np.array_equal(small[1], ids[:small[1].size])

	with pytest.raises(OverflowError):
		_trlda._pack_documents(docs[:3] + [[(2 ** 31, 1)]])
	with pytest.raises(OverflowError):
		_trlda._pack_documents(docs + [[(2 ** 40, 1)]])           # a big int in a large batch: general loop, same error
	with pytest.raises(TypeError):
		_trlda._pack_documents([(1, 2)])                           # a document that is not a list
	with pytest.raises(TypeError):
		_trlda._pack_documents('docs')
