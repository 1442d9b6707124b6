"""A small K=1000 workload (bulk-copy streaming E-step, fused scatter/M-step, Gibbs) for compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_probe.py
    compute-sanitizer --tool racecheck python scripts/sanitize_probe.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from trlda_b200 import capi
from trlda_b200.synth import make_corpus
K, V, B = 1000, 3000, 48
for precision in ('mixed', 'fp64'):
	docs = capi.CSR(*make_corpus(B, V, K, .1, .2, mean_length=60, seed=5))
	m = capi.Model('online', V, K, 10000, .1, .2, precision=precision)
	rho = m.update_parameters(docs, max_iter_tr=2, max_iter_inference=5, kappa=.7, tau=100., update_alpha=1, update_eta=1)
	g, s = m.update_variables(docs, max_iter=5)
	t, s2 = m.update_variables(docs, inference_method='GIBBS')
	print(precision, rho, float(np.abs(m.lambdas).sum()), float(g.sum()), float(s2.sum()), flush=True)
	m.close()
