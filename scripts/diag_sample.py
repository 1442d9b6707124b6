import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from trlda_b200 import capi
capi.seed(5)
m = capi.Model('online', 4, 2, 1000, [.2, .01], .2)
m.lambdas = np.asfortranarray(np.array([[100, 100, 1e-16, 1e-16], [1e-16, 1e-16, 100, 100.]]))
d = m.sample(2000, 10.)
ptr = d.doc_ptr
frac1 = np.array([np.mean(d.word_ids[ptr[i]:ptr[i + 1]] >= 2) if ptr[i + 1] > ptr[i] else np.nan for i in range(2000)])
print('docs dominated by topic 1: %.3f (Dirichlet(.2,.01): ~%.3f); mixed docs %.3f; word histogram %s' % (
	np.nanmean(frac1 > .5), .01 / .21, np.nanmean((frac1 > 0) & (frac1 < 1)), np.bincount(d.word_ids, minlength=4)))
