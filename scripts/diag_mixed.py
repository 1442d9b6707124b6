"""Where does the mixed-precision error come from?  Compares mixed vs fp64 on the GPU (fp64 matches the oracle to
1e-11) for the smoke shape: gamma after one E-step at several iteration caps, then lambda after update_parameters."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from trlda_b200 import capi
from trlda_b200.synth import gamma_matrix, make_corpus

K, V, B = 256, 2000, 96
ptr, ids, cts = make_corpus(B, V, K, .1, .2, mean_length=80, seed=7)
docs = capi.CSR(ptr, ids, cts)
lam0, g0 = gamma_matrix(K, V, 8), gamma_matrix(K, B, 9)
for max_iter in (0, 1, 2, 5, 10, 20, 50):
    out = {}
    for prec in ('fp64', 'mixed'):
        m = capi.Model('online', V, K, 100000, .1, .2, precision=prec)
        m.lambdas = lam0
        g, s = m.update_variables(docs, g0, max_iter=max_iter)
        out[prec] = (g, s, m.stats()['estep_doc_iterations'])
    g64, s64, it64 = out['fp64']; g32, s32, it32 = out['mixed']
    col = np.max(np.abs(g32 - g64), axis=0) / np.max(np.abs(g64), axis=0)
    ew = np.abs(g32 - g64) / np.abs(g64)
    sw = np.abs(s32 - s64) / np.maximum(np.abs(s64), 1e-12 * np.abs(s64).max())
    print('max_iter %3d: gamma col-rel %.2e elementwise %.2e | sstats global %.2e elementwise %.2e | iters %d vs %d' % (
        max_iter, col.max(), ew.max(), np.abs(s32 - s64).max() / np.abs(s64).max(), sw.max(), it64, it32))
kw = dict(max_iter_tr=3, max_iter_inference=20, kappa=.7, tau=100.)
lams = {}
for prec in ('fp64', 'mixed'):
    m = capi.Model('online', V, K, 100000, .1, .2, precision=prec)
    m.lambdas = lam0
    m.update_parameters(docs, gamma0=g0, **kw)
    lams[prec] = m.lambdas
e = np.abs(lams['mixed'] - lams['fp64']) / np.abs(lams['fp64'])
i = np.unravel_index(np.argmax(e), e.shape)
print('lambda elementwise max %.2e at %s value %.4e (lambda0 %.3e); 99.99%% quantile %.2e; global %.2e' % (
    e.max(), i, lams['fp64'][i], lam0[i], np.quantile(e, .9999), np.abs(lams['mixed'] - lams['fp64']).max() / lams['fp64'].max()))
