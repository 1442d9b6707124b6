"""Runs the five BASELINE.json configurations (shortened where the full run is long) through the C ABI and prints
one timing line each.  cfg-3 is what bench.py measures; the others are functional checks with timings."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from trlda_b200 import capi
from trlda_b200.synth import gamma_matrix, make_corpus

precision = os.environ.get('PREC', 'mixed')
out = []

def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps

# cfg-1: README example
K, V, B = 100, 7000, 200
docs = capi.CSR(*make_corpus(B, V, K, .1, .2, seed=1001))
m = capi.Model('online', V, K, 1000000, .1, .2, precision=precision); m.lambdas = gamma_matrix(K, V, 2001)
t = timed(lambda: m.update_parameters(docs, max_iter_tr=10, max_iter_inference=20, kappa=.7, tau=100.))
out.append(dict(cfg=1, what='OnlineLDA K=100 V=7000 B=200 T=10 I=20', ms=t * 1e3, docs_per_s=B / t)); m.close()

# cfg-2: BatchLDA, one epoch over 100k documents
K, V, B = 100, 10000, 100000
docs = capi.CSR(*make_corpus(B, V, K, .1, .2, seed=1002))
m = capi.Model('batch', V, K, 0, .1, .2, precision=precision); m.lambdas = gamma_matrix(K, V, 2002)
for iters in (20, 100):
    t = timed(lambda: m.update_parameters(docs, max_epochs=1, max_iter_inference=iters), reps=2)
    out.append(dict(cfg=2, what='BatchLDA K=100 V=10k 100k docs, 1 epoch, max_iter_inference=%d' % iters, ms=t * 1e3, docs_per_s=B / t))
assert np.all(np.isfinite(m.lambdas)); m.close()

# cfg-4: empirical Bayes on
K, V, B = 500, 50000, 8192
docs = capi.CSR(*make_corpus(B, V, K, .1, .2, seed=1004))
m = capi.Model('online', V, K, 1000000, .1, .2, precision=precision); m.lambdas = gamma_matrix(K, V, 2004)
t = timed(lambda: m.update_parameters(docs, max_iter_tr=10, max_iter_inference=20, kappa=.7, tau=100., update_alpha=1, update_eta=1))
out.append(dict(cfg=4, what='OnlineLDA K=500 V=50k B=8192 update_alpha update_eta', ms=t * 1e3, docs_per_s=B / t, eta=m.eta, alpha_min=float(m.alpha.min()), alpha_max=float(m.alpha.max()))); m.close()

# cfg-5: CumulativeLDA streaming, 4 batches of 4096
K, V, B = 200, 100000, 4096
m = capi.Model('cumulative', V, K, 0, .1, .2, precision=precision)
batches = [capi.CSR(*make_corpus(B, V, K, .1, .2, seed=1005 + i)) for i in range(4)]
t0 = time.perf_counter()
for b in batches:
    m.update_parameters(b, max_epochs=10, max_iter_inference=100, update_alpha=1)
t = time.perf_counter() - t0
out.append(dict(cfg=5, what='CumulativeLDA K=200 V=100k, 4 batches of 4096, max_epochs=10, update_alpha', ms=t * 1e3 / len(batches), docs_per_s=B * len(batches) / t, alpha_mean=float(m.alpha.mean())))
assert np.all(np.isfinite(m.lambdas)); m.close()

for line in out:
    print(json.dumps(line))
