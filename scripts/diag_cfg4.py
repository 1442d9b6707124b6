"""cfg-4 shape: where does the mixed-mode error of lambda sit, and do the inner-iteration counts of the two modes differ?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from trlda_b200 import capi
from trlda_b200.synth import gamma_matrix, make_corpus
K, V, B, D = 500, 50000, 256, 1000000
ptr, ids, cts = make_corpus(B, V, K, .1, .2, seed=1004)
lam0, g0 = gamma_matrix(K, V, 2004), gamma_matrix(K, B, 3004)
out = {}
for T in (1, 2, 3):
	kwargs = dict(max_iter_tr=T, max_iter_inference=20, kappa=.7, tau=100., update_alpha=1, update_eta=1)
	for precision in ('fp64', 'mixed'):
		m = capi.Model('online', V, K, D, .1, .2, precision=precision)
		m.lambdas = lam0
		m.reset_stats()
		m.update_parameters(capi.CSR(ptr, ids, cts), gamma0=g0, **kwargs)
		s = m.stats()
		out[precision] = (m.lambdas, s['estep_doc_iterations'])
		m.close()
	a, b = out['mixed'][0], out['fp64'][0]
	err = np.max(np.abs(a - b), axis=0) / np.max(np.abs(b), axis=0)
	top = np.argsort(err)[-5:][::-1]
	print('T=%d last E-step iterations fp64 %d mixed %d; per-word error: max %.3e, top words %s, errors %s, 99.9%% quantile %.2e, median %.2e' % (
		T, out['fp64'][1], out['mixed'][1], err.max(), top.tolist(), ['%.2e' % e for e in err[top]], np.quantile(err, .999), np.median(err)))
	docs_with = [int(np.searchsorted(ptr, np.nonzero(ids == w)[0][0], side='right') - 1) if np.any(ids == w) else -1 for w in top]
	print('   documents holding those words:', docs_with)
