"""Phase timers of the fast E-step kernel on the cfg-3 shape: one cold E-step (20 inner iterations) and one warm
E-step (re-run from the converged gamma)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['TRLDA_ESTEP_TICKS'] = '1'
import numpy as np
from trlda_b200 import capi
from trlda_b200.synth import gamma_matrix, make_corpus
K, V, B = 1000, 100000, int(os.environ.get('B', 4096))
docs = capi.CSR(*make_corpus(B, V, K, .1, .2, seed=1003))
g0 = gamma_matrix(K, B, 3003)
for label, it in (('cold (20 iterations)', 20), ('warm (restart from converged gamma)', None)):
    m = capi.Model('online', V, K, 1000000, .1, .2, precision=os.environ.get('PREC', 'mixed'))
    if it is None:
        os.environ['TRLDA_ESTEP_TICKS'] = '0'
        m0 = capi.Model('online', V, K, 1000000, .1, .2, precision=os.environ.get('PREC', 'mixed'))
        m0.lambdas = m.lambdas
        g1, _ = m0.update_variables(docs, g0, max_iter=100, want_sstats=False)
        m0.close()
        os.environ['TRLDA_ESTEP_TICKS'] = '1'
        g, _ = m.update_variables(docs, g1, max_iter=20, want_sstats=False)
    else:
        g, _ = m.update_variables(docs, g0, max_iter=it, want_sstats=False)
    print(label, m.stats()['estep_doc_iterations'] / B, flush=True)
    sys.stderr.flush()
    m.close()
