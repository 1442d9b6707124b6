"""
CPU tests of the checker itself (`-m "not gpu"`): the plain-C oracle is pinned against
  - the known-answer values of the reference's own test (code/trlda/python/tests/utils_test.py:33-51),
  - the golden fixtures in tests/golden/ (generated from the compiled reference by make_golden.py),
  - the compiled reference itself, when oracle/_ref/libtrlda_ref.so is present (this container).
"""
import numpy as np
import pytest

from common import (load_case, rel_err, rel_err_elementwise, run_batch_case, run_cumulative_case,
	run_online_case)

# utils_test.py:35-43
POLYGAMMA_GOLDEN = {
	(0, .1): -10.423754940411,
	(0, 1.): -0.5772156649015329,
	(0, 120.): 4.7833192891185,
	(1, .01): 10001.6212135283,
	(1, .1): 101.433299150792758817215450106,
	(1, .4): 7.275356590529597,
	(1, 11.): 0.09516633568168575,
	(2, 14.): -0.005479465690312488}


def test_polygamma_known_answers(oracle_built):
	lib = oracle_built.port_lib()
	for (n, x), y in POLYGAMMA_GOLDEN.items():
		assert abs(lib.oracle_polygamma(n, x) - y) < 5e-8 * max(1., abs(y))      # assertAlmostEqual: 7 places


def test_special_functions_match_reference_fixture(oracle_built):
	lib = oracle_built.port_lib()
	case = np.load(__import__('os').path.join(__import__('common').GOLDEN, 'special.npz'))
	for name, n in (('digamma', 0), ('trigamma', 1), ('tetragamma', 2)):
		got = np.array([lib.oracle_polygamma(n, float(v)) for v in case['x']])
		assert rel_err_elementwise(got, case[name], floor=0.) < 1e-14, name


@pytest.mark.parametrize('name', ['online_tr.npz', 'online_sgd.npz', 'online_adaptive.npz'])
def test_port_reproduces_online_golden(oracle_built, name):
	case = load_case(name)
	model = oracle_built.PortModel('online', case['V'], case['K'], case['D'], case['alpha0'], case['eta0'])
	out = run_online_case(model, oracle_built.CSR, case)
	assert rel_err_elementwise(out['estep_gamma'], case['estep_gamma']) < 1e-12
	assert rel_err(out['estep_sstats'], case['estep_sstats']) < 1e-12
	assert out['rho'] == pytest.approx(float(case['rho']), rel=1e-15)
	assert rel_err_elementwise(out['lambda1'], case['lambda1']) < 1e-11
	assert rel_err_elementwise(out['alpha1'], case['alpha1']) < 1e-11
	assert out['eta1'] == pytest.approx(float(case['eta1']), rel=1e-11)
	assert out['update_count1'] == int(case['update_count1'])


def test_port_reproduces_batch_golden(oracle_built):
	case = load_case('batch.npz')
	model = oracle_built.PortModel('batch', case['V'], case['K'], 0, case['alpha0'], case['eta0'])
	out = run_batch_case(model, oracle_built.CSR, case)
	assert out['rho'] == 1.
	assert rel_err_elementwise(out['lambda1'], case['lambda1']) < 1e-10
	assert rel_err_elementwise(out['alpha1'], case['alpha1']) < 1e-10
	assert out['eta1'] == pytest.approx(float(case['eta1']), rel=1e-10)


def test_port_reproduces_cumulative_golden(oracle_built):
	case = load_case('cumulative.npz')
	model = oracle_built.PortModel('cumulative', case['V'], case['K'], 0, case['alpha0'], case['eta0'])
	assert np.all(model.lambdas == case['eta0'])                                  # cumulativelda.cpp:30
	out = run_cumulative_case(model, oracle_built.CSR, case)
	for call in range(2):
		assert rel_err_elementwise(out['lambda1_%d' % call], case['lambda1_%d' % call]) < 1e-10
		assert rel_err_elementwise(out['alpha1_%d' % call], case['alpha1_%d' % call]) < 1e-10


def test_port_matches_compiled_reference_on_seeded_inputs(oracle_built):
	"""larger seeded case than the fixtures: K=100 V=7000 B=64 (cfg-1 shape), T=3"""
	if not oracle_built.have_ref():
		pytest.skip('compiled reference not present on this box (built only where /root/reference exists)')
	from trlda_b200.synth import gamma_matrix, make_corpus
	K, V, B = 100, 7000, 64
	docs = oracle_built.CSR(*make_corpus(B, V, K, .1, .2, seed=1001))
	lam0, g0 = gamma_matrix(K, V, 2001), gamma_matrix(K, B, 3001)
	results = []
	for cls in (oracle_built.RefModel, oracle_built.PortModel):
		model = cls('online', V, K, 1000000, .1, .2)
		model.lambdas = lam0
		gamma, sstats = model.update_variables(docs, g0, max_iter=20)
		rho = model.update_parameters(docs, gamma0=g0, max_iter_tr=3, max_iter_inference=20, update_alpha=1, update_eta=1)
		results.append((gamma, sstats, rho, model.lambdas, model.alpha, model.eta))
	ref, port = results
	assert rel_err_elementwise(port[0], ref[0]) < 1e-12
	assert rel_err(port[1], ref[1]) < 1e-12
	assert port[2] == ref[2]
	assert rel_err_elementwise(port[3], ref[3]) < 1e-11
	assert rel_err_elementwise(port[4], ref[4]) < 1e-11
	assert port[5] == pytest.approx(ref[5], rel=1e-11)


def test_port_gamma_sampler_replays_reference_rand_stream(oracle_built):
	if not oracle_built.have_ref():
		pytest.skip('compiled reference not present on this box')
	ref = oracle_built.RefModel('online', 5, 3, 10)
	port = oracle_built.PortModel('online', 5, 3, 10)
	a = ref.sample_gamma(7, 11, 100, seed=5)
	b = port.sample_gamma(7, 11, 100, seed=5)
	assert np.max(np.abs(a - b)) < 1e-15
	assert abs(a.mean() - 1.) < .05 and abs(a.std() - .1) < .03             # Gamma(100, 1/100)


def test_intended_lower_bound_against_independent_numpy(oracle_built):
	"""the oracle's ELBO follows Hoffman's approx_bound (python/tests/onlineldavb.py:260-318), restated here in
	numpy/scipy as an independent cross-check; lda.cpp:334 (bug) is not the oracle"""
	from scipy.special import gammaln, psi
	rng = np.random.default_rng(5)
	K, V, B, D = 9, 40, 14, 300
	from common import random_docs
	lists = random_docs(rng, B, V, 20, empty=(3,), duplicates=(5,))
	docs = oracle_built.CSR.from_lists(lists)
	model = oracle_built.PortModel('online', V, K, D, .15, .25)
	lam = model.lambdas
	g0 = np.asfortranarray(rng.gamma(100., .01, size=(B, K)).T)
	total, per_doc = model.lower_bound(docs, g0, max_iter=50)
	gamma, _ = model.update_variables(docs, g0, max_iter=50)

	alpha, eta = model.alpha, model.eta
	elogbeta = psi(lam) - psi(lam.sum(1))[:, None]
	score = 0.
	for d, doc in enumerate(lists):
		g = gamma[:, d]
		elogtheta = psi(g) - psi(g.sum())
		s = 0.
		for w, c in doc:
			t = elogtheta + elogbeta[:, w]
			s += c * (np.log(np.sum(np.exp(t - t.max()))) + t.max())
		s += np.sum((alpha - g) * elogtheta) + np.sum(gammaln(g)) - gammaln(g.sum())
		s += gammaln(alpha.sum()) - np.sum(gammaln(alpha))
		assert per_doc[d] == pytest.approx(s, rel=1e-10)
		score += s
	score *= D / float(B)
	score += np.sum((eta - lam) * elogbeta) + np.sum(gammaln(lam) - gammaln(eta))
	score += np.sum(gammaln(eta * V) - gammaln(lam.sum(1)))
	assert total == pytest.approx(score, rel=1e-10)


def test_estep_edge_cases_port_vs_reference(oracle_built):
	"""empty documents, duplicate word ids, max_iter = 0, single-topic model"""
	if not oracle_built.have_ref():
		pytest.skip('compiled reference not present on this box')
	from common import random_docs
	rng = np.random.default_rng(9)
	for K, V, B, max_iter in ((1, 10, 4, 5), (6, 25, 8, 0), (6, 25, 8, 100)):
		lists = random_docs(rng, B, V, 12, empty=(0, B - 1), duplicates=(1,))
		docs = oracle_built.CSR.from_lists(lists)
		lam0 = np.asfortranarray(rng.gamma(100., .01, size=(V, K)).T)
		g0 = np.asfortranarray(rng.gamma(100., .01, size=(B, K)).T)
		out = []
		for cls in (oracle_built.RefModel, oracle_built.PortModel):
			model = cls('online', V, K, 100, .2, .3)
			model.lambdas = lam0
			out.append(model.update_variables(docs, g0, max_iter=max_iter))
		assert rel_err_elementwise(out[1][0], out[0][0]) < 1e-13
		assert rel_err(out[1][1], out[0][1]) < 1e-13


def test_fast_constructed_reference_is_the_same_model(oracle_built):
	"""bench.py builds the reference with fast_init (no rand() draw of lambda in the constructor); once lambda is
	installed it must behave exactly like a normally constructed reference"""
	if not oracle_built.have_ref():
		pytest.skip('compiled reference not present on this box')
	from common import random_docs
	rng = np.random.default_rng(2)
	K, V, B = 7, 50, 11
	docs = oracle_built.CSR.from_lists(random_docs(rng, B, V, 20))
	lam0 = np.asfortranarray(rng.gamma(100., .01, size=(V, K)).T)
	g0 = np.asfortranarray(rng.gamma(100., .01, size=(B, K)).T)
	out = []
	for fast in (False, True):
		model = oracle_built.RefModel('online', V, K, 300, .1, .2, fast_init=fast)
		assert model.lambdas.shape == (K, V)
		if fast:
			assert np.all(model.lambdas == .2)
		model.lambdas = lam0
		rho = model.update_parameters(docs, gamma0=g0, max_iter_tr=3, max_iter_inference=20, update_alpha=1, update_eta=1)
		out.append((rho, model.lambdas, model.alpha, model.eta))
	assert out[0][0] == out[1][0] and out[0][3] == out[1][3]
	# the reference sums the per-thread statistics in arrival order (lda.cpp:211, omp critical): two runs of the
	# SAME object differ in the last bits, so this is a 1e-12 comparison, not a bitwise one
	assert np.allclose(out[0][1], out[1][1], rtol=1e-12, atol=0) and np.allclose(out[0][2], out[1][2], rtol=1e-12, atol=0)
	with pytest.raises(RuntimeError):
		oracle_built.RefModel('online', V, K, 300, .1, .2, fast_init=True).update_parameters(docs, gamma0=g0, adaptive=1)


def test_gibbs_restatement_properties(oracle_built):
	"""the corrected restatement of lda.cpp:224-293 (no reference output can pin it: the original indexes theta by token,
	lda.cpp:254, and races on sstats, :284): counts conserved, forced assignments exact, seeded runs reproducible, and
	many sweeps agree with the variational sufficient statistics — the same properties tests/test_gibbs_gpu.py asks of
	the CUDA kernel"""
	from common import random_docs
	rng = np.random.default_rng(31)
	K, V, B = 4, 40, 25
	lam = np.full((K, V), 1e-16)
	for k in range(K):
		lam[k, 10 * k:10 * k + 10] = 50.
	model = oracle_built.PortModel('online', V, K, 1000, .1, .2)
	model.lambdas = np.asfortranarray(lam)
	docs = oracle_built.CSR.from_lists(random_docs(rng, B, V, 25) + [[]])
	totals = np.zeros(V)
	np.add.at(totals, docs.word_ids, docs.counts)
	theta0 = rng.dirichlet(np.ones(K), size=B + 1).T
	theta, sstats = model.update_variables_gibbs(docs, theta0, num_samples=3, burn_in=1, seed=4)
	expected = np.zeros((K, V))
	for k in range(K):
		expected[k, 10 * k:10 * k + 10] = totals[10 * k:10 * k + 10]
	assert np.allclose(sstats, expected, rtol=0, atol=1e-9)
	assert np.all(theta > 0) and np.allclose(theta.sum(0), 1., rtol=0, atol=1e-12)
	again = model.update_variables_gibbs(docs, theta0, num_samples=3, burn_in=1, seed=4)
	assert np.array_equal(theta, again[0]) and np.array_equal(sstats, again[1])
	with pytest.raises(RuntimeError, match='Initial theta has wrong dimensionality.'):
		model.update_variables_gibbs(docs, theta0[:, :-1])

	# overlapping topics: averaged over many sweeps the assignments approach the variational statistics
	K, V, B = 6, 120, 30
	model = oracle_built.PortModel('online', V, K, 1000, .1, .2)
	model.lambdas = np.asfortranarray(rng.gamma(.05, 1., size=(V, K)).T * 200. + .01)
	docs = oracle_built.CSR.from_lists(random_docs(rng, B, V, 40))
	_, sstats_vi = model.update_variables(docs, rng.gamma(100., .01, size=(B, K)).T, max_iter=200, threshold=1e-6)
	_, sstats_gibbs = model.update_variables_gibbs(docs, num_samples=64, burn_in=16, seed=5)
	assert np.allclose(sstats_gibbs.sum(0), sstats_vi.sum(0), rtol=0, atol=1e-6)
	assert np.corrcoef(sstats_gibbs.ravel(), sstats_vi.ravel())[0, 1] > .9
