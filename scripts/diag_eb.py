import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trlda, trlda.models as models
for seed in (20261017, 1, 2, 3, 4, 5, 6, 7):
	trlda.seed(seed)
	model = models.OnlineLDA(num_words=4, num_topics=2, num_documents=1000, alpha=[.2, .01], eta=.2)
	model.lambdas = [[100, 100, 1e-16, 1e-16], [1e-16, 1e-16, 100, 100]]
	documents = model.sample(100, 10)
	n1 = sum(1 for d in documents if d and np.mean([w >= 2 for w, _ in d]) > .5)
	model.alpha = [4., 4.]
	for _ in range(100):
		model.update_parameters(documents, rho=.1, max_iter_tr=0, update_lambda=False, update_alpha=True)
	a = model.alpha.ravel().copy()
	model = models.BatchLDA(num_words=4, num_topics=2, alpha=[.2, .05], eta=.2)
	model.lambdas = [[100, 100, 1e-16, 1e-16], [1e-16, 1e-16, 100, 100]]
	documents = model.sample(100, 10)
	model.alpha = [4., 4.]
	model.update_parameters(documents, max_epochs=10, update_lambda=False, update_alpha=True)
	print(seed, 'topic-1 docs', n1, 'online alpha', a, 'batch alpha', model.alpha.ravel(), flush=True)
