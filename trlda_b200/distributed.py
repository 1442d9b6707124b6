"""
Host-side plumbing for the document-sharded multi-GPU path (one process per GPU).

The data path itself lives in the library: after `Model.comm_init`, every hot-path call treats its `docs` as this
rank's shard and sums the batch-wide quantities (sufficient statistics, their row sums, word counts, alpha
statistics, document counts) over ranks with NCCL on the model's stream.  What remains for the host is
 (1) cutting a minibatch into per-rank shards and (2) getting rank 0's NCCL unique id to everybody, for which any
`torch.distributed` backend will do (nccl on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(doc_ptr, world_size):
	"""Contiguous document ranges [begin, end) per rank, balanced by the number of (word, count) pairs rather than
	by documents, since the E-step cost of a document is proportional to its length."""
	doc_ptr = np.asarray(doc_ptr, dtype=np.int64)
	num_docs = doc_ptr.size - 1
	total = int(doc_ptr[-1])
	bounds = [0]
	for r in range(1, world_size):
		target = total * r / float(world_size)
		cut = int(np.searchsorted(doc_ptr, target, side='left'))
		bounds.append(min(max(cut, bounds[-1]), num_docs))
	bounds.append(num_docs)
	return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def shard_documents(doc_ptr, word_ids, counts, rank, world_size):
	"""This rank's shard of a CSR minibatch, as a new CSR triple."""
	begin, end = shard_bounds(doc_ptr, world_size)[rank]
	lo, hi = int(doc_ptr[begin]), int(doc_ptr[end])
	return (np.asarray(doc_ptr[begin:end + 1], dtype=np.int64) - lo,
		np.asarray(word_ids[lo:hi], dtype=np.int32), np.asarray(counts[lo:hi], dtype=np.int32))


def broadcast_unique_id(dist, make_id, device=None):
	"""Rank 0 creates the 128-byte NCCL unique id (make_id()), everybody receives it through torch.distributed."""
	import torch
	uid = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else 'cpu')
	if dist.get_rank() == 0:
		uid.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
	dist.broadcast(uid, 0)
	return uid.cpu().numpy().tobytes()


def init_comm(model, dist, device=None):
	"""Joins `model` (a trlda_b200.capi.Model) to the communicator of the current torch.distributed world."""
	from . import capi
	world = dist.get_world_size()
	if world == 1:
		return
	uid = broadcast_unique_id(dist, capi.comm_unique_id, device)
	model.comm_init(uid, dist.get_rank(), world)
