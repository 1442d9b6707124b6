"""
The native reader of the reference's document format (csrc/ingest.cu, trlda_reader_*) against the reference's own
generator semantics (python/utils/load_documents.py:6-69), restated by the pure-Python loader of this package.
Runs without a GPU: the reader falls back from pinned to pageable memory when no device exists.
"""
import os

import numpy as np
import pytest

from trlda_b200.utils.load_documents import _batches, _parse_pairs, load_documents, load_documents_csr


def write_corpus(path, docs):
	with open(path, 'w') as handle:
		for doc in docs:
			handle.write(' '.join([str(len(doc))] + ['%d:%d' % pair for pair in doc]) + '\n')


def random_corpus(rng, n, vocab=5000):
	docs = []
	for d in range(n):
		size = int(rng.integers(0, 40)) if d % 7 else 0       # empty documents are legal
		docs.append([(int(w), int(1 + rng.integers(9))) for w in rng.permutation(vocab)[:size]])
	return docs


@pytest.fixture(scope='module')
def built():
	from trlda_b200 import build
	build.build_library()


@pytest.mark.parametrize('n,batch', [(23, 5), (20, 5), (3, 10), (0, 4), (17, None)])
def test_native_reader_matches_reference_generator(built, tmp_path, n, batch):
	rng = np.random.default_rng(n)
	docs = random_corpus(rng, n)
	path = str(tmp_path / 'corpus.txt')
	write_corpus(path, docs)
	if batch:
		want = list(_batches(path, batch, False, _parse_pairs))       # the reference's generator, full batches + remainder
		got = list(load_documents_csr(path, batch_size=batch))
		# 20 documents in batches of 5: four full batches AND the empty remainder (load_documents.py:63)
		assert len(got) == len(want)
	else:
		want = [load_documents(path)]
		got = [load_documents_csr(path)]
	for (ptr, ids, cts), ref in zip(got, want):
		assert ptr.dtype == np.int64 and ids.dtype == np.int32 and cts.dtype == np.int32
		assert len(ptr) - 1 == len(ref)
		lists = [list(zip(ids[ptr[d]:ptr[d + 1]].tolist(), cts[ptr[d]:ptr[d + 1]].tolist())) for d in range(len(ref))]
		assert lists == ref


def test_native_reader_equals_python_path(built, tmp_path):
	rng = np.random.default_rng(1)
	path = str(tmp_path / 'corpus.txt')
	write_corpus(path, random_corpus(rng, 200))
	a = list(load_documents_csr(path, batch_size=64))
	b = list(load_documents_csr(path, batch_size=64, native=False))
	assert len(a) == len(b) == 4
	for x, y in zip(a, b):
		for u, v in zip(x, y):
			assert np.array_equal(u, v)


def test_reader_views_and_prefetch(built, tmp_path):
	from trlda_b200 import capi
	rng = np.random.default_rng(2)
	docs = random_corpus(rng, 64)
	path = str(tmp_path / 'corpus.txt')
	write_corpus(path, docs)
	reader = capi.Reader(path, batch_size=16, prefetch=3, copy=False)
	sizes = [batch.num_docs for batch in reader]
	assert sizes == [16, 16, 16, 16, 0]
	# tolerant of missing trailing newline and of tabs / carriage returns
	with open(path, 'w') as handle:
		handle.write('2 1:2\t7:1\r\n1 3:4')
	batch = next(capi.Reader(path))
	assert batch.doc_ptr.tolist() == [0, 2, 3] and batch.word_ids.tolist() == [1, 7, 3] and batch.counts.tolist() == [2, 1, 4]


def test_reader_errors(built, tmp_path):
	from trlda_b200 import capi
	with pytest.raises(IOError):
		capi.Reader(str(tmp_path / 'missing.txt'))
	path = str(tmp_path / 'bad.txt')
	with open(path, 'w') as handle:
		handle.write('1 12:3\n1 oops\n')
	with pytest.raises(ValueError, match='Malformed document line'):
		list(capi.Reader(path, batch_size=10))
