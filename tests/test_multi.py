"""Multi-process tests of the document-sharded path: world_size 2 over gloo on the CPU (host logic + the
decomposition itself, checked with the oracle) and, on a box with >= 2 GPUs, the CUDA path over NCCL."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
	with socket.socket() as s:
		s.bind(('127.0.0.1', 0))
		return s.getsockname()[1]


def launch(mode, world):
	cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
		'--master-addr', '127.0.0.1', '--master-port', str(free_port()),
		os.path.join(ROOT, 'tests', 'multi_worker.py'), '--mode', mode]
	proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
	lines = [l for l in proc.stdout.splitlines() if l.startswith('RESULT')]
	assert proc.returncode == 0 and lines, proc.stdout[-2000:] + proc.stderr[-4000:]
	return [float(x) for x in lines[0].split()[1:]]


def test_shard_bounds_cover_and_balance():
	from trlda_b200.distributed import shard_bounds, shard_documents
	rng = np.random.default_rng(0)
	lengths = rng.integers(0, 300, size=1000)
	ptr = np.concatenate([[0], np.cumsum(lengths)])
	for world in (1, 2, 3, 8):
		bounds = shard_bounds(ptr, world)
		assert bounds[0][0] == 0 and bounds[-1][1] == 1000
		assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
		pairs = [ptr[e] - ptr[b] for b, e in bounds]
		assert max(pairs) - min(pairs) <= 2 * lengths.max()
	ids = np.arange(ptr[-1], dtype=np.int32)
	p, i, c = shard_documents(ptr, ids, ids, 1, 2)
	b, e = shard_bounds(ptr, 2)[1]
	assert p[0] == 0 and p[-1] == i.size == ptr[e] - ptr[b] and i[0] == ptr[b]
	# degenerate: more ranks than documents
	assert shard_bounds(np.array([0, 5]), 4)[-1] == (1, 1) or sum(e - b for b, e in shard_bounds(np.array([0, 5]), 4)) == 1


def test_document_sharding_decomposes_over_gloo(oracle_built):
	"""world_size 2, gloo, CPU: per-shard E-steps + all_reduce of sstats == full-batch E-step"""
	err, err_gamma, docs, pairs, B, N = launch('cpu', 2)
	assert err < 1e-13 and err_gamma == 0.0
	assert (docs, pairs) == (B, N)


@pytest.mark.gpu
def test_two_gpus_match_one_gpu():
	"""2 ranks x half the minibatch over NCCL == 1 rank x the whole minibatch (sum order differs: ~1e-16)"""
	import torch
	if torch.cuda.device_count() < 2:
		pytest.skip('needs two GPUs')
	out = launch('gpu', 2)
	fp64, mixed = out[:5], out[5:]
	assert fp64[0] == 0. and max(fp64[1:]) < 1e-11
	assert mixed[0] == 0. and max(mixed[1:]) < 1e-4
