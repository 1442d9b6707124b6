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

# enough documents, of all tile shapes, for the work counter of k_estep_tmem to hand documents out (2 x 74 teams are
# fixed) and for a few documents of more than 192 pairs (streaming kernel)
docs = capi.CSR(*make_corpus(420, V, K, .1, .2, mean_length=130, seed=6))
lengths = np.diff(docs.doc_ptr) if hasattr(docs, 'doc_ptr') else None
m = capi.Model('online', V, K, 10000, .1, .2, precision='mixed')
g, s = m.update_variables(docs, max_iter=4)
rho = m.update_parameters(docs, max_iter_tr=2, max_iter_inference=4, kappa=.7, tau=100.)
print('mixed, 420 documents', rho, float(g.sum()), float(np.abs(m.lambdas).sum()), m.stats()['estep_sweeps'],
	None if lengths is None else (int(lengths.min()), int(lengths.max())), flush=True)
m.close()
