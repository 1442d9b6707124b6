#!/usr/bin/env python
"""
bench.py — documents/s for one OnlineLDA `update_parameters` step (BASELINE.json's metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision mixed|fp64]

Workload (BASELINE.json configs[2], "cfg-3"): OnlineLDA K=1000 topics, V=100 000 words, minibatch of 8192
synthetic documents (~150 distinct words each, LDA generative process), max_iter_tr=10, max_iter_inference=20,
kappa=.7, tau=100.  One "step" = one update_parameters call on one minibatch.

  value   whole-job documents/s with the minibatch already resident in HBM (trlda_update_parameters_resident),
          timed on the device with CUDA events on the library's stream, max over ranks.
  e2e     the same step through the reference-facing C-ABI call with HOST buffers
          (trlda_update_parameters: pinned staging + H2D of the CSR minibatch every step, then a D2H read of the
          K row sums of the new lambda), timed by wall clock between barriers, max over ranks.

N > 1 (launched by torch.distributed.run, one rank per GPU): documents are sharded over ranks — weak scaling,
8192 documents PER GPU — and the sufficient statistics are summed across ranks once per trust-region
iteration (NCCL all-reduce on the library's stream).  torch is used for the process group, the barrier and the
event timers only.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libtrlda_ref.so = the unmodified
reference core compiled by oracle/Makefile; the plain-C port if that file is absent) on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
	# name: (V, K, D, B per GPU, alpha, eta, update_parameters kwargs)
	'cfg3': dict(V=100000, K=1000, D=1000000, B=8192, alpha=.1, eta=.2,
		params=dict(max_iter_tr=10, max_iter_inference=20, kappa=.7, tau=100.),
		desc='OnlineLDA K=1000 V=100k batch 8192 kappa=.7 tau=100 max_iter_tr=10 max_iter_inference=20'),
	'cfg1': dict(V=7000, K=100, D=1000000, B=200, alpha=.1, eta=.2,
		params=dict(max_iter_tr=10, max_iter_inference=20, kappa=.7, tau=100.),
		desc='OnlineLDA K=100 V=7000 batch 200 (README example)'),
	'cfg4': dict(V=50000, K=500, D=1000000, B=8192, alpha=.1, eta=.2,
		params=dict(max_iter_tr=10, max_iter_inference=20, kappa=.7, tau=100., update_alpha=1, update_eta=1),
		desc='OnlineLDA K=500 V=50k batch 8192 with update_alpha, update_eta'),
}
CFG_INDEX = {'cfg1': 1, 'cfg3': 3, 'cfg4': 4}
# DRAM bytes of one launch of the dominant kernel, from the committed `ncu --set full` capture (profiles/)
NCU_TRAFFIC = {('cfg3', 'mixed'): 8.04e9}


def parse_args():
	ap = argparse.ArgumentParser()
	ap.add_argument('--gpus', type=int, default=1)
	ap.add_argument('--steps', type=int, default=5)
	ap.add_argument('--warmup', type=int, default=3)
	ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
	ap.add_argument('--precision', default='mixed', choices=['mixed', 'fp64'])
	ap.add_argument('--workload', default='cfg3', choices=sorted(WORKLOADS))
	ap.add_argument('--batch', type=int, default=0, help='override documents per GPU (debugging only)')
	ap.add_argument('--no-cpu-baseline', action='store_true')
	ap.add_argument('--cpu-budget', type=float, default=25., help='seconds of CPU work for cpu_baseline')
	return ap.parse_args()


def load_peaks():
	path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
	if os.path.exists(path):
		with open(path) as handle:
			return float(json.load(handle)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
	return 6650., 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
	"""samples nvidia-smi clocks and throttle reasons of one GPU while the timed region runs"""
	QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
		'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
		'clocks_event_reasons.sw_power_cap')

	def __init__(self, index):
		self.index = index
		self.proc = None
		self.path = None

	def start(self):
		try:
			fd, self.path = tempfile.mkstemp(suffix='.csv')
			os.close(fd)
			self.proc = subprocess.Popen(
				['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits', '-lms', '100'],
				stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
		except OSError:
			self.proc = None

	def stop(self):
		if self.proc is None:
			return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
		self.proc.terminate()
		try:
			self.proc.wait(timeout=5)
		except subprocess.TimeoutExpired:
			self.proc.kill()
		sm, mx, power, reasons = [], [], [], set()
		names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
		with open(self.path) as handle:
			for line in handle:
				f = [x.strip() for x in line.split(',')]
				if len(f) < 7:
					continue
				try:
					sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
				except ValueError:
					continue
				for name, flag in zip(names, f[3:7]):
					if flag.lower().startswith('active'):
						reasons.add(name)
		os.unlink(self.path)
		if not sm:
			return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
		return {
			'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'power_w_max': float(max(power)),
			'samples': len(sm), 'reasons': sorted(reasons)}


def make_inputs(w, batch, rank, workload_name, num_batches=1):
	"""`num_batches` minibatches of `batch` documents each, drawn from ONE synthetic corpus (same topics), as one CSR
	triple; plus the initial lambda."""
	from trlda_b200.synth import gamma_matrix, make_corpus
	cfg = CFG_INDEX[workload_name]
	ptr, ids, cts = make_corpus(batch * num_batches, w['V'], w['K'], w['alpha'], w['eta'], seed=1000 + cfg + 7919 * rank)
	lam0 = gamma_matrix(w['K'], w['V'], 2000 + cfg)          # identical on every rank (replicated model)
	return (ptr, ids, cts), lam0


def split_batches(docs, batch, num_batches):
	ptr, ids, cts = docs
	out = []
	for i in range(num_batches):
		lo, hi = ptr[i * batch], ptr[(i + 1) * batch]
		out.append((ptr[i * batch:(i + 1) * batch + 1] - lo, ids[lo:hi], cts[lo:hi]))
	return out


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU implementation on the host cores
# ----------------------------------------------------------------------------------------------------------------------
_CPU_MODEL = {}


def cpu_model(w):
	"""The CPU implementation, constructed once per process."""
	key = (w['V'], w['K'])
	if key not in _CPU_MODEL:
		from oracle import pyoracle
		if pyoracle.have_ref():
			# fast_init: skip the constructor's rand() draw of lambda (minutes at cfg-3); lambda0 is installed before every timed call
			_CPU_MODEL[key] = (pyoracle, pyoracle.RefModel('online', w['V'], w['K'], w['D'], w['alpha'], w['eta'], fast_init=True), 'reference')
		else:
			if not pyoracle.have_port():
				pyoracle.build(ref=False)
			_CPU_MODEL[key] = (pyoracle, pyoracle.PortModel('online', w['V'], w['K'], w['D'], w['alpha'], w['eta']), 'port')
	return _CPU_MODEL[key]


def cpu_sample(w, docs, lam0, sizes, workload_name):
	"""Times updateParameters of the CPU implementation on a bounded sample and extrapolates to the full step.

	Three calls with an injected gamma0: (B1, T=1), (B2, T=1), (B1, T=2).  With t(B, T) = c0 + T (a + b B) they give
	the per-call cost c0 (rho, lambda' copy, phi=1/K warm start), the per-iteration fixed K x V cost a (psi / exp over
	lambda, M-step expressions) and the per-document cost b; the full step is c0 + T (a + b B) at the workload's
	B and T.  Returns (docs/s, seconds of the extrapolated step, seconds spent, kind, description)."""
	from trlda_b200.synth import gamma_matrix
	pyoracle, model, kind = cpu_model(w)
	ptr, ids, cts = docs
	params = dict(w['params'])
	T = params['max_iter_tr']
	b1, b2 = sizes
	runs = [(b1, 1), (b2, 1), (b1, 2)]
	start_all = time.perf_counter()

	def timed(n, iters):
		csr = pyoracle.CSR(ptr[:n + 1], ids[:ptr[n]], cts[:ptr[n]])
		g0 = gamma_matrix(w['K'], n, 3000 + CFG_INDEX[workload_name])
		model.lambdas = lam0
		model.update_count = 0
		params['max_iter_tr'] = iters
		t0 = time.perf_counter()
		model.update_parameters(csr, gamma0=g0, **params)
		return time.perf_counter() - t0

	# the first call of a process pays for the page faults of its K x V temporaries: one untimed call, and the small run is
	# timed before and after the others (the smaller of the two counts)
	if not _CPU_MODEL.get('warm'):
		timed(b1, 1)
		_CPU_MODEL['warm'] = True
	times = [timed(n, iters) for n, iters in runs]
	times[0] = min(times[0], timed(b1, 1))
	spent = time.perf_counter() - start_all
	b = max((times[1] - times[0]) / (b2 - b1), 1e-9)
	per_iter = max(times[2] - times[0], 1e-9)              # a + b * b1
	a = max(per_iter - b * b1, 0.)
	c0 = max(times[0] - per_iter, 0.)
	B = w['B']
	full = c0 + T * (a + b * B)
	text = ('%s on the host cores: updateParameters(injected gamma0) on the first documents of the workload, '
		'(B, T) = %s took %s s; model t = c0 + T (a + b B) with c0=%.2f s per call, a=%.2f s per TR iteration '
		'(fixed K*V cost), b=%.3f ms per document and iteration, extrapolated to B=%d, T=%d (the reference\'s rand() '
		'draw of gamma0 is excluded, which favours the CPU)') % (
		'unmodified reference core (oracle/_ref)' if kind == 'reference' else 'plain-C port (oracle/lda_oracle.c)',
		', '.join('(%d, %d)' % r for r in runs), '/'.join('%.2f' % t for t in times), c0, a, b * 1e3, B, T)
	return B / full, full, spent, kind, text


def reference_arm(args, w):
	rank = int(os.environ.get('RANK', '0'))
	if rank != 0:
		return
	docs, lam0 = make_inputs(w, 2048, 0, args.workload)
	cores = os.cpu_count() or 1
	os.environ.setdefault('OMP_NUM_THREADS', str(cores))
	sizes = [64, 2048]     # the per-document slope needs a sample whose cost stands out of the ~3.5 s of fixed K*V work
	# CPU code needs no warm-up beyond the first call; keep the whole run within a few minutes
	total_steps = args.steps + args.warmup
	values, fulls = [], []
	t_start = time.perf_counter()
	for step in range(total_steps):
		value, full, spent, kind, text = cpu_sample(w, docs, lam0, sizes, args.workload)
		if step >= args.warmup:
			values.append(value)
			fulls.append(full)
		if time.perf_counter() - t_start > 240. and len(values) >= 1:
			break
	value = float(np.mean(values))
	line = {
		'impl': 'reference', 'metric': 'docs/sec per update_parameters step', 'value': value, 'unit': 'docs/s',
		'n_gpus': args.gpus, 'steps': len(values), 'warmup': args.warmup, 'ms_per_step': float(np.mean(fulls)) * 1e3,
		'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
		'config': {'workload': w['desc'], 'global_batch': w['B'], 'parallelism': 'host cores (OpenMP)'},
		'cpu_baseline': {'value': value, 'unit': 'docs/s', 'cores': int(os.environ['OMP_NUM_THREADS']), 'kind': kind, 'sample': text},
		'e2e': {'value': value, 'unit': 'docs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
		'gpu_launches': 0}
	print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def main():
	args = parse_args()
	w = dict(WORKLOADS[args.workload])
	if args.batch:
		w['B'] = args.batch
	if args.impl == 'reference':
		reference_arm(args, w)
		return

	import torch
	import torch.distributed as dist
	from trlda_b200 import capi

	rank = int(os.environ.get('RANK', '0'))
	local_rank = int(os.environ.get('LOCAL_RANK', '0'))
	world = int(os.environ.get('WORLD_SIZE', '1'))
	if world != args.gpus and world > 1:
		raise SystemExit('--gpus %d does not match WORLD_SIZE %d' % (args.gpus, world))
	if not torch.cuda.is_available():
		raise SystemExit('bench.py needs a B200: trlda_b200 has no CPU fallback (use --impl reference for the CPU arm)')
	torch.cuda.set_device(local_rank)
	if world > 1:
		dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

	def barrier():
		if world > 1:
			dist.barrier()

	def max_over_ranks(x):
		if world == 1:
			return x
		t = torch.tensor([x], dtype=torch.float64, device='cuda')
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		return float(t.item())

	K, V, B = w['K'], w['V'], w['B']
	# Every step (warm-up and timed) gets a minibatch the model has never seen, as in real online training: the first
	# E-step of a step then starts from a fresh gamma against documents lambda has not been fitted to.
	num_batches = min(args.steps + args.warmup, 32)
	docs_all, lam0 = make_inputs(w, B, rank, args.workload, num_batches)
	batches_np = split_batches(docs_all, B, num_batches)
	batches = [capi.CSR(*b) for b in batches_np]
	docs_np = batches_np[0]
	N = int(np.mean([b.num_pairs for b in batches]))

	model = capi.Model('online', V, K, w['D'], w['alpha'], w['eta'], device=local_rank, precision=args.precision)
	model.lambdas = lam0
	if world > 1:
		from trlda_b200.distributed import init_comm
		init_comm(model, dist, torch.device('cuda', local_rank))
	capi.seed(1234 + rank)
	stream = torch.cuda.ExternalStream(model.stream, device=torch.device('cuda', local_rank))
	params = dict(w['params'])

	# ---- device-resident leg: `value` ----------------------------------------------------------------------------------
	for i, b in enumerate(batches):
		model.upload_docs_slot(b, i)
	step_index = 0
	for _ in range(args.warmup):
		model.select_docs(step_index % num_batches)
		model.update_parameters_resident(**params)
		step_index += 1
	model.set_profiling(True)
	model.reset_stats()
	sampler = ClockSampler(local_rank)
	barrier()
	torch.cuda.synchronize()
	sampler.start()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record(stream)
	for _ in range(args.steps):
		model.select_docs(step_index % num_batches)
		model.update_parameters_resident(**params)
		step_index += 1
	e1.record(stream)
	torch.cuda.synchronize()
	barrier()
	clocks = sampler.stop()
	device_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
	stats = model.stats()
	model.set_profiling(False)

	# ---- end-to-end leg: host CSR buffers -> C ABI -> D2H of the step's result -----------------------------------------
	# the model is reset so that the same minibatches are unseen again
	model.lambdas = lam0
	model.update_count = 0
	model.reset_stats()
	step_index = 0
	for _ in range(min(args.warmup, 2)):
		model.update_parameters(batches[step_index % num_batches], **params)
		model.row_sums()
		step_index += 1
	model.reset_stats()
	barrier()
	torch.cuda.synchronize()
	t0 = time.perf_counter()
	for _ in range(args.steps):
		model.update_parameters(batches[step_index % num_batches], **params)
		result = model.row_sums()
		step_index += 1
	torch.cuda.synchronize()
	barrier()
	e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
	e2e_stats = model.stats()
	assert np.all(np.isfinite(result))

	global_batch = B * world
	value = global_batch / (device_ms * 1e-3)
	e2e_value = global_batch / e2e_s

	# ---- roofline of the dominant kernel (per-document E-step) ---------------------------------------------------------
	s_bytes = 4 if args.precision == 'mixed' else 8
	peak, peak_source = load_peaks()
	# one E-step call = one launch per document-length bucket; aggregate over the timed region
	est_calls = args.steps * max(params.get('max_iter_tr', 10), 1)
	est_launches = max(stats['launches']['estep'], 1)
	est_ms = stats['ms']['estep'] / est_calls
	# per-document share n_d K s + 16 K + 8 n_d, summed over the documents of one E-step call
	est_bytes = N * K * s_bytes + 16 * B * K + 8 * N
	achieved = est_bytes / (est_ms * 1e-3) / 1e9 if est_ms > 0 else 0.
	kernel_ms = {k: v / args.steps for k, v in stats['ms'].items() if v > 0}
	roofline = {
		'kernel': 'k_estep_stream (per-document gamma/phi fixed point, one launch per E-step)', 'bound': 'hbm', 'achieved': achieved, 'peak': peak,
		'unit': 'GB/s', 'frac': achieved / peak, 'traffic': NCU_TRAFFIC.get((args.workload, args.precision)), 'peak_source': peak_source,
		'traffic_source': 'profiles/round1_final_estep_traffic.csv (dram__bytes_read.sum + dram__bytes_write.sum, mean over the 30 k_estep_stream launches of three steps; the same launches moved 33.5 GB each from L2 to the SMs: the re-sweeps of a document are served by L2)',
		'algorithmic_bytes_per_estep': est_bytes, 'avg_estep_ms': est_ms, 'launches_per_estep': est_launches / est_calls,
		'avg_inner_iterations_last_estep': (stats['estep_doc_iterations'] / max(stats['estep_docs'], 1)),
		'kernel_ms_per_step': kernel_ms}

	line = {
		'metric': 'docs/sec per update_parameters step', 'value': value, 'unit': 'docs/s', 'n_gpus': world,
		'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': device_ms, 'higher_is_better': True,
		'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 tile / f64 accumulate' if args.precision == 'mixed' else 'f64',
		'data': 'synthetic',
		'config': {
			'workload': w['desc'], 'global_batch': global_batch, 'docs_per_gpu': B, 'pairs_per_gpu': N,
			'minibatches': '%d distinct minibatches of one synthetic corpus, one per step: every step sees unseen documents' % num_batches,
			'precision': args.precision, 'parallelism': ('single GPU' if world == 1 else 'docs sharded over %d GPUs; per TR iteration one fused reduce-scatter + M-step + all-gather of beta over NVLink peer memory' % world),
			'l2': 'no flush needed: every step streams lambda/beta (%.1f GB working set >> 126 MB L2)' % (
				K * V * (16 + s_bytes) / 1e9)},
		'clocks': clocks,
		'e2e': {
			'value': e2e_value, 'unit': 'docs/s', 'ms_per_step': e2e_s * 1e3,
			'h2d_bytes_per_step': e2e_stats['h2d_bytes'] // args.steps,
			'd2h_bytes_per_step': e2e_stats['d2h_bytes'] // args.steps},
		'gpu_launches': stats['total_launches'],
		'roofline': roofline}

	if rank == 0 and world == 1 and not args.no_cpu_baseline:
		sizes = [64, min(2048, B)]     # the per-document slope needs a sample whose cost stands out of the ~3.5 s of fixed K*V work
		cpu_value, full, spent, kind, text = cpu_sample(w, docs_np, lam0, sizes, args.workload)
		line['cpu_baseline'] = {
			'value': cpu_value, 'unit': 'docs/s', 'cores': int(os.environ.get('OMP_NUM_THREADS', os.cpu_count() or 1)),
			'kind': kind, 'sample': text}
	elif rank == 0:
		line['cpu_baseline'] = None

	if rank == 0:
		print(json.dumps(line), flush=True)
	model.close()
	if world > 1:
		dist.destroy_process_group()


if __name__ == '__main__':
	main()
