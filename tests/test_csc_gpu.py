"""The device-built token list of a gathered multi-GPU minibatch (csrc/csc.cu) against a stable sort in numpy."""
import numpy as np
import pytest

from trlda_b200 import capi

pytestmark = pytest.mark.gpu


def reference_lists(lengths, ids, v0, v1, V):
	ranks, max_docs = lengths.shape
	max_pairs = ids.shape[1]
	flat = ids.reshape(-1)
	doc_of = np.full(flat.size, -1, dtype=np.int64)
	for r in range(ranks):
		at = r * max_pairs
		for d in range(max_docs):
			n = int(lengths[r, d])
			doc_of[at:at + n] = r * max_docs + d
			at += n
	own = np.nonzero((flat >= v0) & (flat < v1))[0]
	order = own[np.argsort(flat[own], kind='stable')]        # by word, a word's tokens in ascending global order
	word_ptr = np.zeros(V + 1, dtype=np.int64)
	np.add.at(word_ptr, flat[own] + 1, 1)
	return np.cumsum(word_ptr), doc_of[order], order


@pytest.mark.parametrize('ranks,max_docs,V,v0,v1,mean_len,seed', [
	(1, 40, 50, 0, 50, 12, 0),          # one segment, every word owned
	(2, 300, 400, 100, 300, 30, 1),     # a word range in the middle
	(4, 257, 2000, 1500, 2000, 70, 2),  # last shard, chunks that end inside documents
	(8, 64, 30, 7, 19, 25, 3),          # few words: many equal words inside a group of 32 tokens
	(3, 5, 100, 0, 34, 0, 4),           # only empty documents
	(2, 100, 1000, 400, 400, 20, 5),    # empty word range
])
def test_global_token_list_matches_a_stable_sort(ranks, max_docs, V, v0, v1, mean_len, seed):
	rng = np.random.default_rng(seed)
	lengths = rng.poisson(mean_len, size=(ranks, max_docs)).astype(np.int32)
	lengths[:, -max(1, max_docs // 10):] = 0              # padding documents (and empty ones in between)
	lengths[rng.random(lengths.shape) < .05] = 0
	max_pairs = int(lengths.sum(1).max()) + 3
	ids = np.full((ranks, max_pairs), -1, dtype=np.int32)
	for r in range(ranks):
		n = int(lengths[r].sum())
		ids[r, :n] = rng.integers(0, V, size=n)             # repeated words inside a document are allowed
	model = capi.Model('online', V, 4, 1000, .1, .2, device=0, precision='mixed')
	word_ptr, tok_doc, tok_src = model.debug_global_csc(lengths, ids, ranks, v0, v1)
	model.close()
	want_ptr, want_doc, want_src = reference_lists(lengths, ids, v0, v1, V)
	assert np.array_equal(word_ptr, want_ptr)
	assert np.array_equal(tok_src, want_src)
	assert np.array_equal(tok_doc, want_doc)
