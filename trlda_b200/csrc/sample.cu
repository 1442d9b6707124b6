// sample.cu — LDA::sample on the device (reference: code/trlda/src/lda.cpp:88-115, utils.cpp:235-330): the generative
// process of the model, used by the reference's tests to draw corpora and by bench.py as a corpus generator.
//
//   beta_k  ~ Dirichlet(lambda_k)          lda.cpp:95-97     k_sample_beta: Gamma(lambda_kw) per entry, row-wise CDF
//   n_d     ~ Poisson(length)              lda.cpp:92        Knuth's product method, as utils.cpp:294-311
//   theta_d ~ Dirichlet(alpha)             lda.cpp:101       Gamma(alpha_k) per topic, CDF in shared memory
//   per token: k ~ theta_d, w ~ beta_k     lda.cpp:105-107   two binary searches (utils.cpp sampleHistogram)
//
// One warp per document.  The reference emits every token as (word, 1), repeated ids included; `collapse` instead
// sorts a document's tokens (bitonic sort in shared memory) and emits unique (word, count) pairs in word order, the
// form load_documents produces (python/utils/load_documents.py:41-44).  Random numbers are counter-based
// (rng.cuh): the corpus depends on (seed, lambda, alpha) only.
#include "kernels.cuh"
#include "rng.cuh"

namespace trlda {

namespace {

constexpr int kSampleThreads = 256;

// cdf[k * V + w] = sum_{v <= w} g_kv / sum_v g_kv with g_kv ~ Gamma(lambda_kv): one CTA per topic
__global__ void __launch_bounds__(kSampleThreads) k_sample_beta(const double* __restrict__ lambda, int K, int V, uint64_t seed, double* __restrict__ cdf) {
	__shared__ double warp_tot[kSampleThreads / 32];
	__shared__ double carry_s;
	const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	double* row = cdf + (int64_t) k * V;
	if(tid == 0)
		carry_s = 0.0;
	__syncthreads();
	for(int w0 = 0; w0 < V; w0 += kSampleThreads) {
		const int w = w0 + tid;
		double g = 0.0;
		if(w < V)
			g = fmax(philox_gamma(fmax(lambda[(int64_t) w * K + k], 1e-300), seed, (uint64_t) w * (uint64_t) K + k, 1u), 1e-300);
		// inclusive scan over the CTA
		double scan = g;
		#pragma unroll
		for(int o = 1; o < 32; o <<= 1) {
			const double other = __shfl_up_sync(0xffffffffu, scan, o);
			if(lane >= o)
				scan += other;
		}
		if(lane == 31)
			warp_tot[warp] = scan;
		__syncthreads();
		double before = carry_s;
		for(int q = 0; q < warp; ++q)
			before += warp_tot[q];
		if(w < V)
			row[w] = before + scan;
		__syncthreads();
		if(tid == kSampleThreads - 1)
			carry_s = before + scan;
		__syncthreads();
	}
	const double total = carry_s;
	for(int w = tid; w < V; w += kSampleThreads)
		row[w] /= total;
}

// first index whose cdf value exceeds u (utils.cpp sampleHistogram), clamped to n - 1
__device__ __forceinline__ int cdf_search(const double* __restrict__ cdf, int n, double u) {
	int lo = 0, hi = n - 1;
	while(lo < hi) {
		const int mid = (lo + hi) >> 1;
		if(cdf[mid] > u)
			hi = mid;
		else
			lo = mid + 1;
	}
	return lo;
}

// one warp per document; tokens[d * cap ..] receives the word ids (collapse: sorted unique ids, counts beside them)
template <bool COLLAPSE>
__global__ void __launch_bounds__(128) k_sample_docs(const double* __restrict__ cdf, const double* __restrict__ alpha, int K, int V,
                                                      int64_t B, double length, int cap, uint64_t seed, int32_t* __restrict__ tokens,
                                                      int32_t* __restrict__ counts, int32_t* __restrict__ lengths) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	double* theta = reinterpret_cast<double*>(smem_raw) + (size_t) warp * K;
	int32_t* sorted = reinterpret_cast<int32_t*>(reinterpret_cast<double*>(smem_raw) + (size_t) 4 * K) + (size_t) warp * cap;
	const int64_t d = (int64_t) blockIdx.x * 4 + warp;
	if(d >= B)
		return;
	// document length (utils.cpp:294-311)
	int n = 0;
	if(lane == 0) {
		const double threshold = exp(-length);
		double p = philox_uniform(seed, (uint64_t) d, 2u, 0u);
		while(p > threshold && n < cap) {
			++n;
			p *= philox_uniform(seed, (uint64_t) d, 2u, (uint32_t) n);
		}
	}
	n = __shfl_sync(0xffffffffu, n, 0);
	// theta ~ Dirichlet(alpha): CDF over the topics
	double run = 0.0;
	for(int k0 = 0; k0 < K; k0 += 32) {
		const int k = k0 + lane;
		double g = k < K ? fmax(philox_gamma(fmax(alpha[k], 1e-300), seed, (uint64_t) d * (uint64_t) K + k, 3u), 1e-300) : 0.0;
		double scan = g;
		#pragma unroll
		for(int o = 1; o < 32; o <<= 1) {
			const double other = __shfl_up_sync(0xffffffffu, scan, o);
			if(lane >= o)
				scan += other;
		}
		if(k < K)
			theta[k] = run + scan;
		run += __shfl_sync(0xffffffffu, scan, 31);
	}
	__syncwarp();
	const double total = theta[K - 1];
	// tokens
	int32_t* out = tokens + d * cap;
	for(int t = lane; t < n; t += 32) {
		const double u1 = philox_uniform(seed, (uint64_t) d, 4u, (uint32_t) t) * total;
		const int k = cdf_search(theta, K, u1);
		const double u2 = philox_uniform(seed, (uint64_t) d, 5u, (uint32_t) t);
		const int w = cdf_search(cdf + (int64_t) k * V, V, u2);
		if(COLLAPSE)
			sorted[t] = w;
		else
			out[t] = w;
	}
	if(!COLLAPSE) {
		if(lane == 0)
			lengths[d] = n;
		return;
	}
	// bitonic sort of the document's tokens (padded with INT32_MAX to a power of two)
	int m = 32;
	while(m < n)
		m <<= 1;
	for(int t = n + lane; t < m; t += 32)
		sorted[t] = INT32_MAX;
	__syncwarp();
	for(int size = 2; size <= m; size <<= 1)
		for(int stride = size >> 1; stride > 0; stride >>= 1) {
			for(int t = lane; t < m; t += 32) {
				const int partner = t ^ stride;
				if(partner > t) {
					const bool up = (t & size) == 0;
					const int32_t x = sorted[t], y = sorted[partner];
					if((x > y) == up) {
						sorted[t] = y;
						sorted[partner] = x;
					}
				}
			}
			__syncwarp();
		}
	// unique + count: a token starts a run if it differs from its predecessor
	int base = 0;
	for(int t0 = 0; t0 < n; t0 += 32) {
		const int t = t0 + lane;
		const bool start = t < n && (t == 0 || sorted[t] != sorted[t - 1]);
		const unsigned mask = __ballot_sync(0xffffffffu, start);
		if(start) {
			const int pos = base + __popc(mask & ((1u << lane) - 1u));
			int e = t + 1;
			while(e < n && sorted[e] == sorted[t])
				++e;
			out[pos] = sorted[t];
			counts[d * cap + pos] = e - t;
		}
		base += __popc(mask);
	}
	if(lane == 0)
		lengths[d] = base;
}

}  // namespace

// capacity per document: Poisson(length) stays below with overwhelming probability; longer draws are truncated
int sample_capacity(double length) {
	int cap = (int) (length + 12.0 * sqrt(length + 1.0) + 32.0);
	int pow2 = 32;
	while(pow2 < cap)
		pow2 <<= 1;
	return pow2;
}

size_t sample_smem_bytes(int K, int cap) { return (size_t) 4 * K * 8 + (size_t) 4 * cap * 4; }

void launch_sample_beta(const double* lambda, int K, int V, uint64_t seed, double* cdf, cudaStream_t s) {
	k_sample_beta<<<K, kSampleThreads, 0, s>>>(lambda, K, V, seed, cdf);
}

int launch_sample_docs(const double* cdf, const double* alpha, int K, int V, int64_t B, double length, int cap, bool collapse,
                       uint64_t seed, int32_t* tokens, int32_t* counts, int32_t* lengths, cudaStream_t s) {
	if(B == 0)
		return 0;
	const size_t smem = sample_smem_bytes(K, cap);
	const unsigned grid = (unsigned) ((B + 3) / 4);
	if(collapse) {
		if(cudaFuncSetAttribute(k_sample_docs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess)
			return -1;
		k_sample_docs<true><<<grid, 128, smem, s>>>(cdf, alpha, K, V, B, length, cap, seed, tokens, counts, lengths);
	} else {
		if(cudaFuncSetAttribute(k_sample_docs<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess)
			return -1;
		k_sample_docs<false><<<grid, 128, smem, s>>>(cdf, alpha, K, V, B, length, cap, seed, tokens, counts, lengths);
	}
	return 0;
}

}  // namespace trlda
