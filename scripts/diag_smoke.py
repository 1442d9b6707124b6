"""Distribution of the mixed-mode error of the smoke() case: per word, and after each trust-region iteration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle
from trlda_b200 import capi
from trlda_b200.synth import gamma_matrix, make_corpus

if not pyoracle.have_port():
	pyoracle.build(ref=False)
K, V, B = 256, 2000, 96
ptr, ids, cts = make_corpus(B, V, K, .1, .2, mean_length=80, seed=7)
lam0, g0 = gamma_matrix(K, V, 8), gamma_matrix(K, B, 9)
for T in (1, 2, 3):
	kwargs = dict(max_iter_tr=T, max_iter_inference=20, kappa=.7, tau=100., update_alpha=1, update_eta=1)
	port = pyoracle.PortModel('online', V, K, 100000, .1, .2)
	port.lambdas = lam0
	port.update_parameters(pyoracle.CSR(ptr, ids, cts), gamma0=g0, **kwargs)
	for precision in ('fp64', 'mixed'):
		model = capi.Model('online', V, K, 100000, .1, .2, device=0, precision=precision)
		model.lambdas = lam0
		model.update_parameters(capi.CSR(ptr, ids, cts), gamma0=g0, **kwargs)
		lam, want = model.lambdas, port.lambdas
		err = np.max(np.abs(lam - want), axis=0) / np.max(np.abs(want), axis=0)
		print('T=%d %-5s per-word error: median %.2e  90%% %.2e  99%% %.2e  99.9%% %.2e  max %.2e  (words above 1e-5: %d)' % (
			T, precision, np.median(err), np.quantile(err, .9), np.quantile(err, .99), np.quantile(err, .999), err.max(), int((err > 1e-5).sum())))
		model.close()
