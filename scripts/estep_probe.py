"""E-step kernels on the cfg-3 shape: tensor-memory-resident cluster kernel against the streaming kernel.

One cold E-step (fresh gamma, 20 inner iterations), one warm E-step (restart from the converged gamma) and one with a
single inner iteration, each run with TRLDA_ESTEP_TMEM=1 and =0; prints the E-step kernel time and the largest relative
difference of gamma between the two.  The switch is read when a model is created, so every leg builds its own model."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from trlda_b200 import capi
from trlda_b200.synth import gamma_matrix, make_corpus

K = int(os.environ.get('K', 1000))
V = int(os.environ.get('V', 100000))
B = int(os.environ.get('B', 8192))
docs = capi.CSR(*make_corpus(B, V, K, .1, .2, seed=1003))
g0 = gamma_matrix(K, B, 3003)
lam0 = gamma_matrix(K, V, 2003)


MODES = {'stream': '0', 'tmem': '1'}


def run(mode, gamma, max_iter, reps=3):
	os.environ['TRLDA_ESTEP_TMEM'] = MODES[mode]
	m = capi.Model('online', V, K, 1000000, .1, .2, precision='mixed')
	m.lambdas = lam0
	m.update_variables(docs, gamma, max_iter=max_iter, want_sstats=False)     # warm-up (beta-prep, allocations)
	m.set_profiling(True)
	best = 1e9
	for _ in range(reps):
		m.reset_stats()
		g, _ = m.update_variables(docs, gamma, max_iter=max_iter, want_sstats=False)
		s = m.stats()
		best = min(best, s['ms']['estep'])
	its = s['estep_doc_iterations'] / B
	launches = s['launches']['estep']
	m.close()
	return g, best, its, launches


results = {}
for label, gamma, it in (('cold', g0, 20), ('warm', None, 20), ('one', g0, 1)):
	if gamma is None:
		gamma = results[('cold', 'stream')][0] if ('cold', 'stream') in results else g0
		# converge further so that the warm start stops after one or two iterations
		os.environ['TRLDA_ESTEP_TMEM'] = '0'
		m = capi.Model('online', V, K, 1000000, .1, .2, precision='mixed')
		m.lambdas = lam0
		gamma, _ = m.update_variables(docs, gamma, max_iter=100, want_sstats=False)
		m.close()
	for mode in os.environ.get('MODES', 'stream,tmem').split(','):
		try:
			results[(label, mode)] = run(mode, gamma, it)
		except Exception as e:            # noqa: BLE001
			print(label, mode, 'FAILED', e, flush=True)
			continue
		g, ms, its, launches = results[(label, mode)]
		print('%-5s %-8s E-step %.3f ms  (%.2f inner iterations per document, %d launches)' % (label, mode, ms, its, launches), flush=True)
		if mode != 'stream' and (label, 'stream') in results:
			a, b = g, results[(label, 'stream')][0]
			rel = np.abs(a - b).max(axis=0) / np.abs(b).max(axis=0)
			print('%-5s gamma %s vs stream: per-document inf-norm relative difference max %.3e, median %.3e, finite %s' % (
				label, mode, rel.max(), np.median(rel), bool(np.isfinite(a).all())), flush=True)
