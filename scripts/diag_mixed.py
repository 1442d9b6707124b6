"""Mixed-mode error of every golden case and of the smoke / oracle shapes under the candidate metrics:
elementwise relative, column-wise infinity-norm relative (per document for gamma, per word for lambda), whole-array."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from common import (load_case, rel_err, rel_err_columns, rel_err_elementwise, run_batch_case, run_cumulative_case,
	run_online_case)
from trlda_b200 import capi


def report(tag, got, want):
	got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
	if got.ndim < 2:
		print('%-34s elementwise %.2e  vector inf-norm %.2e' % (tag, rel_err_elementwise(got, want), rel_err(got, want)))
	else:
		print('%-34s elementwise %.2e  column inf-norm %.2e  whole %.2e' % (
			tag, rel_err_elementwise(got, want), rel_err_columns(got, want), rel_err(got, want)))


for name in ('online_tr.npz', 'online_sgd.npz', 'online_adaptive.npz'):
	case = load_case(name)
	m = capi.Model('online', case['V'], case['K'], case['D'], case['alpha0'], case['eta0'], precision='mixed')
	out = run_online_case(m, capi.CSR, case)
	report(name + ' gamma', out['estep_gamma'], case['estep_gamma'])
	report(name + ' lambda', out['lambda1'], case['lambda1'])
	report(name + ' alpha', out['alpha1'], case['alpha1'])
	print('%-34s eta %.2e' % (name, abs(out['eta1'] - float(case['eta1'])) / float(case['eta1'])))
case = load_case('batch.npz')
m = capi.Model('batch', case['V'], case['K'], 0, case['alpha0'], case['eta0'], precision='mixed')
out = run_batch_case(m, capi.CSR, case)
report('batch lambda', out['lambda1'], case['lambda1'])
report('batch alpha', out['alpha1'], case['alpha1'])
print('%-34s eta %.2e' % ('batch', abs(out['eta1'] - float(case['eta1'])) / float(case['eta1'])))
case = load_case('cumulative.npz')
m = capi.Model('cumulative', case['V'], case['K'], 0, case['alpha0'], case['eta0'], precision='mixed')
out = run_cumulative_case(m, capi.CSR, case)
for call in range(2):
	report('cumulative lambda %d' % call, out['lambda1_%d' % call], case['lambda1_%d' % call])
	report('cumulative alpha %d' % call, out['alpha1_%d' % call], case['alpha1_%d' % call])

# the smoke shape and the two oracle shapes of test_online_update_parameters_vs_oracle
from oracle import pyoracle
from trlda_b200.synth import gamma_matrix, make_corpus
for K, V, B, T, ml, seeds in ((256, 2000, 96, 3, 80, (7, 8, 9)), (100, 7000, 200, 10, 150, (1001, 2001, 3001)), (1000, 2000, 64, 10, 150, (1001, 2001, 3001))):
	ptr, ids, cts = make_corpus(B, V, K, .1, .2, mean_length=ml, seed=seeds[0])
	lam0, g0 = gamma_matrix(K, V, seeds[1]), gamma_matrix(K, B, seeds[2])
	kwargs = dict(max_iter_tr=T, max_iter_inference=20, kappa=.7, tau=100., update_alpha=1, update_eta=1)
	port = pyoracle.PortModel('online', V, K, 100000, .1, .2)
	port.lambdas = lam0
	port.update_parameters(pyoracle.CSR(ptr, ids, cts), gamma0=g0, **kwargs)
	m = capi.Model('online', V, K, 100000, .1, .2, precision='mixed')
	m.lambdas = lam0
	m.update_parameters(capi.CSR(ptr, ids, cts), gamma0=g0, **kwargs)
	report('online K=%d V=%d B=%d lambda' % (K, V, B), m.lambdas, port.lambdas)
	report('online K=%d V=%d B=%d alpha' % (K, V, B), m.alpha, port.alpha)
	print('%-34s eta %.2e' % ('', abs(m.eta - port.eta) / port.eta))
