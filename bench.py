#!/usr/bin/env python
"""
bench.py — documents/s for one OnlineLDA `update_parameters` step (BASELINE.json's metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision mixed|fp64]
                    [--scaling strong|weak] [--no-cpu-baseline] [--no-extras]

Workload (BASELINE.json configs[2], "cfg-3"): OnlineLDA K=1000 topics, V=100 000 words, minibatch of 8192
synthetic documents (~150 distinct words each, LDA generative process), max_iter_tr=10, max_iter_inference=20,
kappa=.7, tau=100.  One "step" = one update_parameters call on one minibatch the model has not seen.

  value   whole-job documents/s with the minibatch already resident in HBM (trlda_update_parameters_resident),
          timed on the device with CUDA events on the library's stream, max over ranks.
  e2e     the same step through the reference-facing C-ABI call with HOST buffers
          (trlda_update_parameters: pinned staging + H2D of the CSR minibatch every step, then a D2H read of the
          K row sums of the new lambda), timed by wall clock between barriers, max over ranks.

N > 1 (launched by torch.distributed.run, one rank per GPU).  The headline is STRONG scaling, BASELINE.json's
target: the same global minibatch of 8192 documents is sharded over the ranks (balanced by pairs) and the sufficient
statistics meet once per trust-region iteration.  `weak` (8192 documents per GPU, different documents of the same
corpus on every rank) is reported beside it.  Before anything is timed the N-rank path is checked against rank 0
alone on one minibatch (`multi_gpu_parity`); the run fails if they disagree.  torch is used for the process group,
the barrier and the event timers only.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libtrlda_ref.so = the unmodified
reference core compiled by oracle/Makefile; the plain-C port if that file is absent) on the host cores: one REAL
full step, with the three-point cost model as a cross-check.
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
	# name: V, K, D, B (global minibatch), alpha, eta, update_parameters kwargs
	'cfg3': dict(kind='online', V=100000, K=1000, D=1000000, B=8192, alpha=.1, eta=.2,
		params=dict(max_iter_tr=10, max_iter_inference=20, kappa=.7, tau=100.),
		desc='OnlineLDA K=1000 V=100k batch 8192 kappa=.7 tau=100 max_iter_tr=10 max_iter_inference=20'),
	'cfg1': dict(kind='online', V=7000, K=100, D=1000000, B=200, alpha=.1, eta=.2,
		params=dict(max_iter_tr=10, max_iter_inference=20, kappa=.7, tau=100.),
		desc='OnlineLDA K=100 V=7000 batch 200 (README example)'),
	'cfg2': dict(kind='batch', V=10000, K=100, D=0, B=100000, alpha=.1, eta=.2,
		params=dict(max_epochs=1, max_iter_inference=20),
		desc='BatchLDA K=100 V=10k, one epoch (full-corpus E-step + lambda = eta + sstats) over 100k docs, max_iter_inference=20'),
	'cfg4': dict(kind='online', V=50000, K=500, D=1000000, B=8192, alpha=.1, eta=.2,
		params=dict(max_iter_tr=10, max_iter_inference=20, kappa=.7, tau=100., update_alpha=1, update_eta=1),
		desc='OnlineLDA K=500 V=50k batch 8192 with update_alpha, update_eta'),
	'cfg5': dict(kind='cumulative', V=100000, K=200, D=0, B=4096, alpha=.1, eta=.2,
		params=dict(max_epochs=10, max_iter_inference=100, update_alpha=1),
		desc='CumulativeLDA K=200 V=100k, batches of 4096 docs, max_epochs=10, max_iter_inference=100'),
}
CFG_INDEX = {'cfg1': 1, 'cfg2': 2, 'cfg3': 3, 'cfg4': 4, 'cfg5': 5}

def parse_args():
	ap = argparse.ArgumentParser()
	ap.add_argument('--gpus', type=int, default=1)
	ap.add_argument('--steps', type=int, default=5)
	ap.add_argument('--warmup', type=int, default=3)
	ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
	ap.add_argument('--precision', default='mixed', choices=['mixed', 'fp64'])
	ap.add_argument('--workload', default='cfg3', choices=sorted(WORKLOADS))
	ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'], help='N > 1: which leg is the headline')
	ap.add_argument('--batch', type=int, default=0, help='override the global minibatch (debugging only)')
	ap.add_argument('--no-cpu-baseline', action='store_true')
	ap.add_argument('--no-extras', action='store_true', help='skip other_configs, e2e_python, the second scaling leg and the parity leg')
	ap.add_argument('--cpu-budget', type=float, default=25., help='seconds of CPU work for cpu_baseline')
	return ap.parse_args()


def load_peaks():
	path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
	if os.path.exists(path):
		with open(path) as handle:
			return float(json.load(handle)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
	return 6650., 'fallback (B200_PROFILING.md)'


def load_traffic(workload, precision):
	"""per-launch DRAM bytes of the dominant kernel from the committed ncu capture (profiles/round2_estep_traffic.json)"""
	path = os.path.join(ROOT, 'profiles', 'round2_estep_traffic.json')
	if os.path.exists(path):
		with open(path) as handle:
			table = json.load(handle)
		entry = table.get('%s/%s' % (workload, precision))
		if entry:
			return entry
	return None


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


def make_inputs(w, batch, workload_name, num_batches=1, doc_seed=None):
	"""`num_batches` minibatches of `batch` documents each, drawn from ONE synthetic corpus (same topics on every rank;
	`doc_seed` selects which documents of it), as one CSR triple; plus the initial lambda."""
	from trlda_b200.synth import gamma_matrix, make_corpus
	cfg = CFG_INDEX[workload_name]
	ptr, ids, cts = make_corpus(batch * num_batches, w['V'], w['K'], w['alpha'], w['eta'], seed=1000 + cfg, doc_seed=doc_seed)
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


def use_all_cores():
	"""torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm takes all host cores whoever launched it"""
	cores = os.cpu_count() or 1
	os.environ['OMP_NUM_THREADS'] = str(cores)
	return cores


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


def cpu_timed_call(w, docs, lam0, n, iters, workload_name):
	"""wall time of updateParameters (gamma0 injected) on the first n documents with max_iter_tr = iters"""
	from trlda_b200.synth import gamma_matrix
	pyoracle, model, kind = cpu_model(w)
	ptr, ids, cts = docs
	params = dict(w['params'])
	params['max_iter_tr'] = iters
	csr = pyoracle.CSR(ptr[:n + 1], ids[:ptr[n]], cts[:ptr[n]])
	g0 = gamma_matrix(w['K'], n, 3000 + CFG_INDEX[workload_name])
	model.lambdas = lam0
	model.update_count = 0
	t0 = time.perf_counter()
	model.update_parameters(csr, gamma0=g0, **params)
	return time.perf_counter() - t0


def cpu_sample(w, docs, lam0, sizes, workload_name):
	"""Bounded sample: times updateParameters on (B1, T=1), (B2, T=1), (B1, T=2) and fits t(B, T) = c0 + T (a + b B):
	c0 = per-call cost (rho, lambda' copy, phi = 1/K warm start), a = per-iteration fixed K x V cost (psi / exp over
	lambda, M-step expressions), b = per-document cost; extrapolated to the workload's B and T.
	Returns (docs/s, seconds of the extrapolated step, seconds spent, kind, description)."""
	kind = cpu_model(w)[2]
	T = w['params']['max_iter_tr']
	b1, b2 = sizes
	runs = [(b1, 1), (b2, 1), (b1, 2)]
	start_all = time.perf_counter()
	# the first call of a process pays for the page faults of its K x V temporaries: one untimed call, and the small run is
	# timed before and after the others (the smaller of the two counts)
	if not _CPU_MODEL.get('warm'):
		cpu_timed_call(w, docs, lam0, b1, 1, workload_name)
		_CPU_MODEL['warm'] = True
	times = [cpu_timed_call(w, docs, lam0, n, iters, workload_name) for n, iters in runs]
	times[0] = min(times[0], cpu_timed_call(w, docs, lam0, b1, 1, workload_name))
	spent = time.perf_counter() - start_all
	b = max((times[1] - times[0]) / (b2 - b1), 1e-9)
	per_iter = max(times[2] - times[0], 1e-9)              # a + b * b1
	a = max(per_iter - b * b1, 0.)
	c0 = max(times[0] - per_iter, 0.)
	B = w['B']
	full = c0 + T * (a + b * B)
	text = ('EXTRAPOLATED from a bounded sample: %s on the host cores, updateParameters(injected gamma0) on the first '
		'documents of the workload, (B, T) = %s took %s s; model t = c0 + T (a + b B) with c0=%.2f s per call, a=%.2f s '
		'per TR iteration (fixed K*V cost), b=%.3f ms per document and iteration, extrapolated to B=%d, T=%d (the '
		'reference\'s rand() draw of gamma0 is excluded, which favours the CPU; the CPU computes in fp64)') % (
		'unmodified reference core (oracle/_ref)' if kind == 'reference' else 'plain-C port (oracle/lda_oracle.c)',
		', '.join('(%d, %d)' % r for r in runs), '/'.join('%.2f' % t for t in times), c0, a, b * 1e3, B, T)
	return B / full, full, spent, kind, text


def cpu_one_thread_point(w, docs, lam0, b1, workload_name, all_cores_seconds):
	"""SURVEY section 8d asks for the CPU arm at one thread beside all cores (the reference serialises its accumulation
	with `omp critical`): the smallest sample point again with the OpenMP team cut to one thread."""
	import ctypes
	try:
		gomp = ctypes.CDLL('libgomp.so.1')
		gomp.omp_get_max_threads.restype = ctypes.c_int
		before = int(gomp.omp_get_max_threads())
		gomp.omp_set_num_threads(1)
	except (OSError, AttributeError):
		return None
	try:
		seconds = cpu_timed_call(w, docs, lam0, b1, 1, workload_name)
	finally:
		gomp.omp_set_num_threads(before)
	return {'sample': 'updateParameters(B=%d, max_iter_tr=1)' % b1, 'one_thread_seconds': seconds,
		'all_cores_seconds': all_cores_seconds, 'cores': before}


def reference_arm(args, w):
	"""One REAL update_parameters step of the reference's CPU implementation at the workload's full size (cfg-3: B=8192,
	T=10: about a minute), gamma0 injected; the three-point model of cpu_sample() is run afterwards as a cross-check.
	The driver's --steps / --warmup ask for more than a CPU step per minute allows: `steps` reports what was run."""
	rank = int(os.environ.get('RANK', '0'))
	if rank != 0:
		return
	cores = use_all_cores()
	B = w['B']
	docs, lam0 = make_inputs(w, B, args.workload)
	kind = cpu_model(w)[2]
	cpu_timed_call(w, docs, lam0, 64, 1, args.workload)        # untimed: page faults of the K x V temporaries
	_CPU_MODEL['warm'] = True
	full_s = cpu_timed_call(w, docs, lam0, B, w['params']['max_iter_tr'], args.workload)
	value = B / full_s
	model_value, model_full, spent, _, model_text = cpu_sample(w, docs, lam0, [64, 2048], args.workload)
	sample = ('ONE REAL full step: %s, updateParameters(B=%d, max_iter_tr=%d, max_iter_inference=%d, injected gamma0) on %d '
		'OpenMP threads took %.1f s (fp64; the reference\'s rand() draw of gamma0, ~30 s, is excluded).  Cross-check, '
		'three-point model: %.1f s (%.0f docs/s)') % (
		'unmodified reference core (oracle/_ref)' if kind == 'reference' else 'plain-C port (oracle/lda_oracle.c)',
		B, w['params']['max_iter_tr'], w['params']['max_iter_inference'], cores, full_s, model_full, model_value)
	line = {
		'impl': 'reference', 'metric': 'docs/sec per update_parameters step', 'value': value, 'unit': 'docs/s',
		'n_gpus': args.gpus, 'steps': 1, 'warmup': 1, 'requested_steps': args.steps, 'requested_warmup': args.warmup,
		'ms_per_step': full_s * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
		'data': 'synthetic', 'extrapolated': False,
		'config': {'workload': w['desc'], 'global_batch': B, 'parallelism': 'host cores (OpenMP), %d threads' % cores},
		'cpu_baseline': {'value': value, 'unit': 'docs/s', 'cores': cores, 'kind': kind, 'sample': sample,
			'extrapolated': False, 'cross_check_model_docs_per_s': model_value,
			'one_thread': cpu_one_thread_point(w, docs, lam0, 64, args.workload, cpu_timed_call(w, docs, lam0, 64, 1, args.workload))},
		'e2e': {'value': value, 'unit': 'docs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
		'gpu_launches': 0}
	print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
class Runner(object):
	"""process-group plumbing shared by the legs"""

	def __init__(self, args):
		import torch
		import torch.distributed as dist
		self.torch, self.dist = torch, dist
		self.rank = int(os.environ.get('RANK', '0'))
		self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
		self.world = int(os.environ.get('WORLD_SIZE', '1'))
		if self.world != args.gpus and self.world > 1:
			raise SystemExit('--gpus %d does not match WORLD_SIZE %d' % (args.gpus, self.world))
		if not torch.cuda.is_available():
			raise SystemExit('bench.py needs a B200: trlda_b200 has no CPU fallback (use --impl reference for the CPU arm)')
		torch.cuda.set_device(self.local_rank)
		self.device = torch.device('cuda', self.local_rank)
		if self.world > 1:
			dist.init_process_group('nccl', device_id=self.device)

	def barrier(self):
		if self.world > 1:
			self.dist.barrier()

	def max_over_ranks(self, x):
		if self.world == 1:
			return x
		t = self.torch.tensor([x], dtype=self.torch.float64, device='cuda')
		self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
		return float(t.item())

	def sum_over_ranks(self, x):
		if self.world == 1:
			return x
		t = self.torch.tensor([x], dtype=self.torch.float64, device='cuda')
		self.dist.all_reduce(t)
		return float(t.item())

	def model(self, w, precision, lam0):
		from trlda_b200 import capi
		m = capi.Model(w.get('kind', 'online'), w['V'], w['K'], w['D'], w['alpha'], w['eta'], device=self.local_rank, precision=precision)
		m.lambdas = lam0
		if self.world > 1:
			from trlda_b200.distributed import init_comm
			init_comm(m, self.dist, self.device)
		capi.seed(1234 + self.rank)
		return m


def timed_leg(run, w, precision, lam0, batches, steps, warmup, sample_clocks=False):
	"""device-resident leg (`value`) and host-buffer leg (`e2e`) over this rank's `batches`; returns a dict"""
	from trlda_b200 import capi
	torch = run.torch
	params = dict(w['params'])
	model = run.model(w, precision, lam0)
	stream = torch.cuda.ExternalStream(model.stream, device=run.device)
	num_batches = len(batches)
	csr = [capi.CSR(*b) for b in batches]

	# ---- device-resident leg ------------------------------------------------------------------------------------------
	for i, b in enumerate(csr):
		model.upload_docs_slot(b, i)
	step_index = 0
	for _ in range(warmup):
		model.select_docs(step_index % num_batches)
		model.update_parameters_resident(**params)
		step_index += 1
	model.set_profiling(True)
	model.reset_stats()
	sampler = ClockSampler(run.local_rank) if sample_clocks else None
	run.barrier()
	torch.cuda.synchronize()
	if sampler:
		sampler.start()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record(stream)
	for _ in range(steps):
		model.select_docs(step_index % num_batches)
		model.update_parameters_resident(**params)
		step_index += 1
	e1.record(stream)
	torch.cuda.synchronize()
	run.barrier()
	clocks = sampler.stop() if sampler else None
	device_ms = run.max_over_ranks(e0.elapsed_time(e1)) / steps
	stats = model.stats()
	model.set_profiling(False)

	# ---- end-to-end leg: host CSR buffers -> C ABI -> D2H of the step's result ----------------------------------------
	# the model is reset so that the same minibatches are unseen again
	model.lambdas = lam0
	model.update_count = 0
	model.reset_stats()
	step_index = 0
	for _ in range(min(warmup, 2)):
		model.update_parameters(csr[step_index % num_batches], **params)
		model.row_sums()
		step_index += 1
	model.reset_stats()
	run.barrier()
	torch.cuda.synchronize()
	t0 = time.perf_counter()
	for _ in range(steps):
		model.update_parameters(csr[step_index % num_batches], **params)
		result = model.row_sums()
		step_index += 1
	torch.cuda.synchronize()
	run.barrier()
	e2e_s = run.max_over_ranks(time.perf_counter() - t0) / steps
	e2e_stats = model.stats()
	assert np.all(np.isfinite(result))
	model.close()
	return dict(device_ms=device_ms, e2e_s=e2e_s, stats=stats, e2e_stats=e2e_stats, clocks=clocks,
		pairs=float(np.mean([b.num_pairs for b in csr])), docs=float(np.mean([b.num_docs for b in csr])))


def multi_gpu_parity(run):
	"""The N-rank path against rank 0 alone on the same minibatch, before anything is timed: K=1000, V=20 000, 512
	documents, T=3, I=20 with the empirical-Bayes updates, all three exchange implementations (all-gather of the factors +
	word-sharded M-step, the default; NVLink peer stores; NCCL all-reduce of the dense statistics), both precisions.  fp64 must agree to 1e-11 (sum order), mixed to 1e-4, in the metric of
	tests/common.py parity_err.  Raises SystemExit if not."""
	from trlda_b200 import capi
	from trlda_b200.distributed import init_comm, shard_bounds, shard_documents
	from trlda_b200.synth import gamma_matrix, make_corpus
	K, V, B, D = 1000, 20000, 512, 1000000
	ptr, ids, cts = make_corpus(B, V, K, .1, .2, seed=77)
	lam0, g0 = gamma_matrix(K, V, 78), gamma_matrix(K, B, 79)
	begin, end = shard_bounds(ptr, run.world)[run.rank]
	shard = shard_documents(ptr, ids, cts, run.rank, run.world)
	kwargs = dict(max_iter_tr=3, max_iter_inference=20, kappa=.7, tau=100., update_alpha=1, update_eta=1)
	out = {}
	ok = True
	saved = os.environ.get('TRLDA_MULTI_GPU')
	for precision, bound in (('fp64', 1e-11), ('mixed', 1e-4)):
		single = None
		if run.rank == 0:
			m1 = capi.Model('online', V, K, D, .1, .2, device=run.local_rank, precision=precision)
			m1.lambdas = lam0
			m1.update_parameters(capi.CSR(ptr, ids, cts), gamma0=g0, **kwargs)
			single = (m1.lambdas, m1.alpha, m1.eta)
			m1.close()
		for mode in ('gather', 'peer', 'allreduce'):
			os.environ['TRLDA_MULTI_GPU'] = mode
			m = capi.Model('online', V, K, D, .1, .2, device=run.local_rank, precision=precision)
			m.lambdas = lam0
			init_comm(m, run.dist, run.device)
			m.update_parameters(capi.CSR(*shard), gamma0=g0[:, begin:end], **kwargs)
			lam, alpha, eta = m.lambdas, m.alpha, m.eta
			m.close()
			if run.rank == 0:
				if precision == 'fp64':
					lam_err = float(np.max(np.abs(lam - single[0]) / np.abs(single[0])))
				else:
					lam_err = float(np.max(np.max(np.abs(lam - single[0]), axis=0) / np.max(np.abs(single[0]), axis=0)))
				alpha_err = float(np.max(np.abs(alpha - single[1])) / np.max(np.abs(single[1])))
				eta_err = abs(eta - single[2]) / abs(single[2])
				out['%s/%s' % (precision, mode)] = {'lambda_rel': lam_err, 'alpha_rel': alpha_err, 'eta_rel': eta_err, 'bound': bound}
				ok = ok and max(lam_err, alpha_err, eta_err) < bound
			run.barrier()
	if saved is None:
		os.environ.pop('TRLDA_MULTI_GPU', None)
	else:
		os.environ['TRLDA_MULTI_GPU'] = saved
	flag = run.torch.tensor([1 if ok else 0], device='cuda')
	run.dist.broadcast(flag, 0)
	out['shape'] = 'K=%d V=%d B=%d T=3 I=20 update_alpha update_eta, %d ranks vs 1' % (K, V, B, run.world)
	out['passed'] = bool(flag.item())
	return out


def quick_config(run, name, precision, steps=2, warmup=1):
	"""one of the other BASELINE.json configurations on rank 0's GPU: device-resident docs/s of its update_parameters"""
	from trlda_b200 import capi
	w = dict(WORKLOADS[name])
	batches_needed = steps + warmup
	if name == 'cfg2':
		w['B'] = 100000
		batches_needed = 1        # the whole corpus is the batch; an epoch does not depend on what the model has seen
	docs_all, lam0 = make_inputs(w, w['B'], name, batches_needed)
	batches = split_batches(docs_all, w['B'], batches_needed)
	model = capi.Model(w['kind'], w['V'], w['K'], w['D'], w['alpha'], w['eta'], device=run.local_rank, precision=precision)
	if w['kind'] != 'cumulative':
		model.lambdas = lam0
	stream = run.torch.cuda.ExternalStream(model.stream, device=run.device)
	for i, b in enumerate(batches):
		model.upload_docs_slot(capi.CSR(*b), i)
	index = 0
	for _ in range(warmup):
		model.select_docs(index % batches_needed)
		model.update_parameters_resident(**w['params'])
		index += 1
	model.set_profiling(True)
	model.reset_stats()
	e0, e1 = run.torch.cuda.Event(enable_timing=True), run.torch.cuda.Event(enable_timing=True)
	run.torch.cuda.synchronize()
	e0.record(stream)
	for _ in range(steps):
		model.select_docs(index % batches_needed)
		model.update_parameters_resident(**w['params'])
		index += 1
	e1.record(stream)
	run.torch.cuda.synchronize()
	ms = e0.elapsed_time(e1) / steps
	stats = model.stats()
	model.close()
	N = float(np.mean([b[0][-1] for b in batches]))
	s_bytes = 4 if precision == 'mixed' else 8
	peak, _ = load_peaks()
	calls = max(stats['estep_calls'], 1)
	est_ms = stats['ms']['estep'] / calls
	est_bytes = N * w['K'] * s_bytes + 16 * w['B'] * w['K'] + 8 * N
	return {
		'workload': w['desc'], 'precision': precision, 'docs_per_step': w['B'], 'steps': steps, 'ms_per_step': ms,
		'docs_per_s': w['B'] / (ms * 1e-3), 'estep_calls_per_step': calls / steps,
		'estep_roofline_frac': (est_bytes / (est_ms * 1e-3) / 1e9 / peak) if est_ms > 0 else None,
		'avg_sweeps_per_document_and_estep': stats['estep_sweeps'] / float(calls * w['B']) if stats['estep_sweeps'] else None,
		'kernel_ms_per_step': {k: v / steps for k, v in stats['ms'].items() if v > 0}}


def e2e_python(run, w, precision, lam0, batches_np, steps):
	"""the reference-facing Python call: trlda.models.OnlineLDA.update_parameters(list of lists of (word, count)),
	wall clock per step including the list walk of the binding; the list itself is built before the clock starts,
	as in the reference's examples (load_documents returns it)"""
	import trlda
	from trlda_b200.synth import to_lists
	params = dict(w['params'])

	def timed(inputs):
		# the same model state and the same minibatches for both input forms: the difference is the list walk alone
		model = trlda.models.OnlineLDA(num_words=w['V'], num_topics=w['K'], num_documents=w['D'], alpha=w['alpha'], eta=w['eta'],
			device=run.local_rank, precision=precision)
		model.lambdas = lam0
		model.update_parameters(inputs[0], **params)
		t0 = time.perf_counter()
		for i in range(steps):
			model.update_parameters(inputs[1 + i % (len(inputs) - 1)] if len(inputs) > 1 else inputs[0], **params)
		_ = model.eta
		return (time.perf_counter() - t0) / steps, getattr(model, 'precision', precision)

	s, mode = timed([to_lists(*b) for b in batches_np[:steps + 1]])
	s_csr, _ = timed([tuple(b) for b in batches_np[:steps + 1]])
	return {'value': w['B'] / s, 'unit': 'docs/s', 'ms_per_step': s * 1e3, 'steps': steps,
		'api': 'trlda.models.OnlineLDA.update_parameters(list of lists of (word_id, count))', 'precision': mode,
		'same_steps_with_csr_input': {'value': w['B'] / s_csr, 'ms_per_step': s_csr * 1e3,
			'api': 'update_parameters((doc_ptr, word_ids, counts)): the binding\'s numpy CSR form, no list walk'}}


def e2e_file(run, w, precision, lam0, batches_np, steps):
	"""training straight from the reference's text format: the native reader (memory-mapped file, background parser,
	pinned CSR batches `prefetch` ahead) feeding update_parameters; the Python loader of the reference's format is timed
	on the same file for comparison (parsing only)"""
	from trlda_b200 import capi
	from trlda_b200.utils.load_documents import load_documents
	use = batches_np[:steps + 1]
	fd, path = tempfile.mkstemp(suffix='.txt')
	os.close(fd)
	try:
		with open(path, 'w') as handle:
			for ptr, ids, cts in use:
				pairs = np.char.add(np.char.add(ids.astype(str), ':'), cts.astype(str))
				for d in range(len(ptr) - 1):
					handle.write(str(ptr[d + 1] - ptr[d]) + ' ' + ' '.join(pairs[ptr[d]:ptr[d + 1]]) + '\n')
		size = os.path.getsize(path)
		t0 = time.perf_counter()
		parsed = sum(b.num_docs for b in capi.Reader(path, batch_size=w['B'], prefetch=2, copy=False))
		native_s = time.perf_counter() - t0
		t0 = time.perf_counter()
		first = next(load_documents(path, batch_size=w['B']))
		python_s = (time.perf_counter() - t0) * len(use)
		model = capi.Model('online', w['V'], w['K'], w['D'], w['alpha'], w['eta'], device=run.local_rank, precision=precision)
		model.lambdas = lam0
		params = dict(w['params'])
		reader = capi.Reader(path, batch_size=w['B'], prefetch=2, copy=False)
		model.update_parameters(next(reader), **params)          # warm-up on the first batch
		run.torch.cuda.synchronize()
		t0 = time.perf_counter()
		docs = 0
		for batch in reader:
			if batch.num_docs:
				model.update_parameters(batch, **params)
				docs += batch.num_docs
		model.row_sums()
		s = time.perf_counter() - t0
		model.close()
		return {'value': docs / s, 'unit': 'docs/s', 'ms_per_step': s * 1e3 / max(docs / w['B'], 1), 'file_bytes': size,
			'what': 'text file -> trlda_reader (background parser, pinned CSR) -> trlda_update_parameters, wall clock',
			'native_parse_docs_per_s': parsed / native_s, 'python_load_documents_docs_per_s': len(first) * len(use) / python_s}
	finally:
		os.unlink(path)


def main():
	args = parse_args()
	w = dict(WORKLOADS[args.workload])
	if args.batch:
		w['B'] = args.batch
	if args.impl == 'reference':
		reference_arm(args, w)
		return

	run = Runner(args)
	from trlda_b200.distributed import shard_documents
	K, V, B = w['K'], w['V'], w['B']
	world = run.world
	s_bytes = 4 if args.precision == 'mixed' else 8
	peak, peak_source = load_peaks()

	parity = None
	if world > 1 and not args.no_extras:
		parity = multi_gpu_parity(run)
		if not parity['passed']:
			if run.rank == 0:
				print(json.dumps({'error': 'multi_gpu_parity failed', 'multi_gpu_parity': parity}), flush=True)
			raise SystemExit(3)

	# Every step (warm-up and timed) gets a minibatch the model has never seen, as in real online training: the first
	# E-step of a step then starts from a fresh gamma against documents lambda has not been fitted to.
	num_batches = min(args.steps + args.warmup, 32)
	legs = {}
	order = [args.scaling] + ([] if world == 1 or args.no_extras else ['weak' if args.scaling == 'strong' else 'strong'])
	docs_np = None
	for scaling in order:
		if scaling == 'strong' or world == 1:
			# ONE corpus, ONE sequence of global minibatches on every rank; a rank keeps its shard (balanced by pairs)
			docs_all, lam0 = make_inputs(w, B, args.workload, num_batches)
			global_batches = split_batches(docs_all, B, num_batches)
			batches = [shard_documents(*g, run.rank, world) for g in global_batches] if world > 1 else global_batches
			if docs_np is None:
				docs_np = global_batches
			global_batch = B
		else:
			# weak: 8192 documents per rank, different documents of the same corpus
			docs_all, lam0 = make_inputs(w, B, args.workload, num_batches, doc_seed=run.rank)
			batches = split_batches(docs_all, B, num_batches)
			global_batch = B * world
		leg = timed_leg(run, w, args.precision, lam0, batches, args.steps, args.warmup, sample_clocks=(scaling == order[0]))
		leg['global_batch'] = global_batch
		leg['value'] = global_batch / (leg['device_ms'] * 1e-3)
		leg['e2e_value'] = global_batch / leg['e2e_s']
		legs[scaling] = leg

	head = legs[order[0]]
	stats = head['stats']
	N_local = head['pairs']
	B_local = head['docs']

	# ---- roofline of the dominant kernel (per-document E-step) ---------------------------------------------------------
	est_calls = max(stats['estep_calls'], 1)
	est_launches = max(stats['launches']['estep'], 1)
	est_ms = stats['ms']['estep'] / est_calls
	# per-document share n_d K s + 16 K + 8 n_d, summed over the documents of one E-step call
	est_bytes = N_local * K * s_bytes + 16 * B_local * K + 8 * N_local
	achieved = est_bytes / (est_ms * 1e-3) / 1e9 if est_ms > 0 else 0.
	kernel_ms = {k: v / args.steps for k, v in stats['ms'].items() if v > 0}
	traffic = load_traffic(args.workload, args.precision)
	kernel_name = 'k_estep_tmem (per-document gamma/phi fixed point, tile resident in tensor memory)' if args.precision == 'mixed' else \
		'k_estep_stream (per-document gamma/phi fixed point, tile streamed from L2 once per inner iteration)'
	roofline = {
		'kernel': kernel_name, 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
		'traffic': traffic['bytes_per_estep'] if traffic else None,
		'traffic_source': traffic['source'] if traffic else 'no ncu capture committed for this workload / precision',
		'peak_source': peak_source,
		'algorithmic_bytes_per_estep': est_bytes, 'avg_estep_ms': est_ms, 'launches_per_estep': est_launches / float(est_calls),
		'avg_sweeps_per_document_and_estep': stats['estep_sweeps'] / float(est_calls * max(B_local, 1)),
		'avg_inner_iterations_last_estep': (stats['estep_doc_iterations'] / max(stats['estep_docs'], 1)),
		'whole_step': {'algorithmic_bytes': 16 * K * V + 8 * V + w['params'].get('max_iter_tr', 10) * (
			K * V * (8 + s_bytes) + N_local * K * s_bytes + 8 * N_local + 8 * (B_local + 1) + 16 * B_local * K + 8 * K * V + 16 * K * V + K * V * (8 + s_bytes)),
			'frac': None},
		'kernel_ms_per_step': kernel_ms}
	roofline['whole_step']['frac'] = roofline['whole_step']['algorithmic_bytes'] / (head['device_ms'] * 1e-3) / 1e9 / peak

	line = {
		'metric': 'docs/sec per update_parameters step', 'value': head['value'], 'unit': 'docs/s', 'n_gpus': world,
		'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': head['device_ms'], 'higher_is_better': True,
		'scaling': order[0] if world > 1 else 'weak', 'vs_baseline': None,
		'dtype': 'f32 tile / f64 accumulate' if args.precision == 'mixed' else 'f64', 'data': 'synthetic',
		'config': {
			'workload': w['desc'], 'global_batch': head['global_batch'], 'docs_per_gpu': B_local, 'pairs_per_gpu': N_local,
			'minibatches': '%d distinct minibatches of one synthetic corpus (same topics on every rank), one per step: every step sees unseen documents' % num_batches,
			'precision': args.precision,
			'parallelism': ('single GPU' if world == 1 else 'the global minibatch sharded over %d GPUs (balanced by pairs); per TR iteration an all-gather of etheta and the token weights, scatter + M-step + beta-prep on every rank\'s word range, an all-gather of beta (NCCL over NVLink)' % world),
			'l2': 'no flush needed: every step streams lambda/beta (%.1f GB working set >> 126 MB L2)' % (K * V * (16 + s_bytes) / 1e9)},
		'clocks': head['clocks'],
		'e2e': {
			'value': head['e2e_value'], 'unit': 'docs/s', 'ms_per_step': head['e2e_s'] * 1e3,
			'h2d_bytes_per_step': head['e2e_stats']['h2d_bytes'] // args.steps,
			'd2h_bytes_per_step': head['e2e_stats']['d2h_bytes'] // args.steps},
		'gpu_launches': stats['total_launches'],
		'roofline': roofline}
	if world > 1:
		other = [s for s in order if s != order[0]]
		if other:
			o = legs[other[0]]
			line[other[0]] = {
				'value': o['value'], 'unit': 'docs/s', 'ms_per_step': o['device_ms'], 'global_batch': o['global_batch'],
				'e2e': {'value': o['e2e_value'], 'ms_per_step': o['e2e_s'] * 1e3},
				'avg_sweeps_per_document_and_estep': o['stats']['estep_sweeps'] / float(max(o['stats']['estep_calls'], 1) * max(o['docs'], 1)),
				'kernel_ms_per_step': {k: v / args.steps for k, v in o['stats']['ms'].items() if v > 0}}
		line['multi_gpu_parity'] = parity

	if run.rank == 0 and world == 1 and not args.no_extras:
		extras = []
		for name, precision in (('cfg3', 'fp64' if args.precision == 'mixed' else 'mixed'), ('cfg1', 'mixed'), ('cfg2', 'mixed'), ('cfg4', 'mixed'), ('cfg5', 'mixed')):
			if name == args.workload and precision == args.precision:
				continue
			try:
				extras.append(quick_config(run, name, precision))
			except Exception as error:            # noqa: BLE001 — an extra must not take the headline down
				extras.append({'workload': WORKLOADS[name]['desc'], 'precision': precision, 'error': str(error)[:200]})
		line['other_configs'] = extras
		try:
			line['e2e_python'] = e2e_python(run, w, args.precision, lam0, docs_np, min(args.steps, 3))
		except Exception as error:                # noqa: BLE001
			line['e2e_python'] = {'error': str(error)[:200]}
		try:
			line['e2e_file'] = e2e_file(run, w, args.precision, lam0, docs_np, min(args.steps, 3))
		except Exception as error:                # noqa: BLE001
			line['e2e_file'] = {'error': str(error)[:200]}

	if run.rank == 0 and world == 1 and not args.no_cpu_baseline:
		cores = use_all_cores()
		sizes = [64, min(2048, B)]     # the per-document slope needs a sample whose cost stands out of the ~3.5 s of fixed K*V work
		cpu_value, full, spent, kind, text = cpu_sample(w, docs_np[0], lam0, sizes, args.workload)
		line['cpu_baseline'] = {
			'value': cpu_value, 'unit': 'docs/s', 'cores': cores, 'kind': kind, 'sample': text, 'extrapolated': True,
			'sample_seconds': spent}
	elif run.rank == 0:
		line['cpu_baseline'] = None

	if run.rank == 0:
		print(json.dumps(line), flush=True)
	if world > 1:
		run.dist.destroy_process_group()


if __name__ == '__main__':
	main()
