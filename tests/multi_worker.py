"""
Worker of the multi-process tests, launched with torch.distributed.run.

  --mode cpu   (gloo, no GPU)  every rank runs the CPU oracle's E-step on ITS shard of the minibatch and the ranks
                               sum the sufficient statistics with all_reduce: the decomposition the multi-GPU path
                               relies on (documents are independent given lambda; only sstats couple them).
  --mode gpu   (nccl)          the CUDA path with the library's own NCCL exchange: N ranks on N shards must give the
                               lambda / alpha / eta of one GPU on the whole minibatch.
Rank 0 prints one line starting with RESULT.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('--mode', default='cpu')
	args = ap.parse_args()
	import torch
	import torch.distributed as dist
	from trlda_b200.distributed import init_comm, shard_bounds, shard_documents
	from trlda_b200.synth import gamma_matrix, make_corpus

	rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
	K, V, B = 64, 600, 90
	ptr, ids, cts = make_corpus(B, V, K, .1, .2, mean_length=40, seed=5)
	lam0, g0 = gamma_matrix(K, V, 6), gamma_matrix(K, B, 7)
	begin, end = shard_bounds(ptr, world)[rank]
	shard = shard_documents(ptr, ids, cts, rank, world)
	kwargs = dict(max_iter_tr=3, max_iter_inference=20, update_alpha=1, update_eta=1)

	if args.mode == 'cpu':
		dist.init_process_group('gloo')
		from oracle import pyoracle
		model = pyoracle.PortModel('online', V, K, 5000, .1, .2)
		model.lambdas = lam0
		gamma, sstats = model.update_variables(pyoracle.CSR(*shard), g0[:, begin:end], max_iter=20)
		total = torch.from_numpy(np.ascontiguousarray(sstats))
		dist.all_reduce(total)
		counts = torch.tensor([end - begin, int(shard[0][-1])])
		dist.all_reduce(counts)
		if rank == 0:
			full_gamma, full_sstats = model.update_variables(pyoracle.CSR(ptr, ids, cts), g0, max_iter=20)
			err = float(np.max(np.abs(total.numpy() - full_sstats)) / np.max(np.abs(full_sstats)))
			err_gamma = float(np.max(np.abs(gamma - full_gamma[:, begin:end])))
			print('RESULT', err, err_gamma, int(counts[0]), int(counts[1]), B, int(ptr[-1]), flush=True)
	else:
		local_rank = int(os.environ.get('LOCAL_RANK', rank))
		torch.cuda.set_device(local_rank)
		dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
		from trlda_b200 import capi
		results = {}
		for precision in ('fp64', 'mixed'):
			model = capi.Model('online', V, K, 5000, .1, .2, device=local_rank, precision=precision)
			model.lambdas = lam0
			init_comm(model, dist, torch.device('cuda', local_rank))
			rho = model.update_parameters(capi.CSR(*shard), gamma0=g0[:, begin:end], **kwargs)
			elbo, _ = model.lower_bound(capi.CSR(*shard), g0[:, begin:end], max_iter=20)
			results[precision] = (rho, model.lambdas, model.alpha, model.eta, elbo)
			model.close()
		dist.barrier()
		if rank == 0:
			out = []
			for precision in ('fp64', 'mixed'):
				single = capi.Model('online', V, K, 5000, .1, .2, device=local_rank, precision=precision)
				single.lambdas = lam0
				rho = single.update_parameters(capi.CSR(ptr, ids, cts), gamma0=g0, **kwargs)
				elbo, _ = single.lower_bound(capi.CSR(ptr, ids, cts), g0, max_iter=20)
				r = results[precision]
				out += [abs(r[0] - rho), float(np.max(np.abs(r[1] - single.lambdas) / single.lambdas)),
					float(np.max(np.abs(r[2] - single.alpha) / single.alpha)), abs(r[3] - single.eta) / single.eta,
					abs(r[4] - elbo) / abs(elbo)]
			print('RESULT', *out, flush=True)
		dist.barrier()
	dist.destroy_process_group()


if __name__ == '__main__':
	main()
