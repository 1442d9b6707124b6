"""
Document loaders for the text format of the reference (code/trlda/python/utils/load_documents.py:6-69):
one document per line, `N id:cnt id:cnt ...`; the first field is ignored (line 43 of the reference), every
other field is `word_id:count`.

`load_documents` keeps the reference's behaviour exactly (list of lists of (id, count) tuples; with `batch_size`
a generator of batches, fixed-size or Poisson-sized, the final partial batch always yielded).
`load_documents_csr` yields the same batches already packed as (doc_ptr, word_ids, counts) numpy arrays, which
the models accept directly and which skips the per-word Python objects.  With a fixed batch size it is served by the
native reader (csrc/ingest.cu): the file is memory-mapped and parsed by a background thread into pinned CSR batches a
few batches ahead of the training loop; the pure-Python path remains for `stochastic=True` (numpy's Poisson stream).
"""
import numpy as np
from numpy.random import poisson


def _batches(filepath, batch_size, stochastic, parse):
	documents = []
	current_batch_size = poisson(batch_size) if stochastic else batch_size
	with open(filepath) as handle:
		for lineno, line in enumerate(handle):
			documents.append(parse(line))
			if batch_size:
				while current_batch_size == 0:
					yield []
					current_batch_size = poisson(batch_size)
				if (lineno + 1) % current_batch_size == 0:
					yield documents
					documents = []
					if stochastic:
						current_batch_size = poisson(batch_size)
	yield documents


def _parse_pairs(line):
	document = []
	for word in line.split()[1:]:
		wid, wct = word.split(':')
		document.append((int(wid), int(wct)))
	return document


def load_documents(filepath, batch_size=None, stochastic=False):
	"""
	Load documents from a text file.  If C{batch_size} is given, behaves like a generator and returns one batch
	at a time.  Each document is a list of (word id, word count) tuples.

	@type  batch_size: C{int}
	@param batch_size: the number of documents to return at once

	@type  stochastic: C{bool}
	@param stochastic: if True, the batch size is drawn from a Poisson distribution

	@rtype: C{list}/C{generator}
	"""
	if batch_size:
		return _batches(filepath, batch_size, stochastic, _parse_pairs)
	return next(_batches(filepath, batch_size, stochastic, _parse_pairs))


def _to_csr(documents):
	doc_ptr = np.zeros(len(documents) + 1, dtype=np.int64)
	for i, (ids, _) in enumerate(documents):
		doc_ptr[i + 1] = doc_ptr[i] + len(ids)
	if documents:
		word_ids = np.concatenate([ids for ids, _ in documents]).astype(np.int32, copy=False)
		counts = np.concatenate([cts for _, cts in documents]).astype(np.int32, copy=False)
	else:
		word_ids = np.zeros(0, dtype=np.int32)
		counts = np.zeros(0, dtype=np.int32)
	return doc_ptr, word_ids, counts


def _parse_arrays(line):
	fields = line.split()[1:]
	if not fields:
		return np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32)
	flat = np.array(' '.join(fields).replace(':', ' ').split(), dtype=np.int64)
	return flat[0::2].astype(np.int32), flat[1::2].astype(np.int32)


def load_documents_csr(filepath, batch_size=None, stochastic=False, native=True, prefetch=2):
	"""Same batching as L{load_documents}, but every batch is a (doc_ptr, word_ids, counts) CSR triple."""
	if native and not stochastic:
		from .. import capi
		reader = capi.Reader(filepath, batch_size, prefetch=prefetch)
		if batch_size:
			return ((b.doc_ptr, b.word_ids, b.counts) for b in reader)
		batch = next(reader)
		reader.close()
		return batch.doc_ptr, batch.word_ids, batch.counts
	if batch_size:
		return (_to_csr(batch) for batch in _batches(filepath, batch_size, stochastic, _parse_arrays))
	return _to_csr(next(_batches(filepath, batch_size, stochastic, _parse_arrays)))
