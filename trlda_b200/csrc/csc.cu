// csc.cu — the word-sorted token list of the GLOBAL minibatch of a multi-GPU step, restricted to this rank's words,
// built on the device.
//
// In the gather exchange (model.cu) every rank runs the scatter + M-step kernel on its word range
// [v0, v1) over the tokens of ALL ranks' documents; it needs, per word, the tokens in ascending global order (the order
// in which one GPU adds them: bitwise-equal results).  The ranks all-gather the word ids and document lengths; this
// file turns them into word_ptr / tok_doc / tok_src without a round trip through the host (the first version copied
// the ids back, sorted them with host threads and uploaded the lists: 3.5 ms per minibatch, the whole gap between the
// end-to-end and the device-resident step at 2-8 GPUs).
//
// Stable counting sort in three passes over the token stream cut into kCscChunks contiguous chunks:
//   count    hist[c][w]   = tokens of word w in chunk c                           (atomics: counts do not depend on order)
//   scan     hist[c][w]  -> tokens of w in chunks before c; totals -> word_ptr     (column scan, then one scan over words)
//   place    one warp per chunk walks it in order, 32 tokens at a time; equal words inside a group of 32 are ranked by
//            lane (match.any), so a word's tokens land in ascending token order whatever the timing.
#include "kernels.cuh"

#include <algorithm>

namespace trlda {

namespace {

constexpr int kCscChunks = 256;

// exclusive scan of one value per thread over a block of 1024 threads; returns the block total in `total`
__device__ __forceinline__ int block_exclusive_scan(int v, int& total, int* warp_sums) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int inc = v;
	#pragma unroll
	for(int o = 1; o < 32; o <<= 1) {
		const int up = __shfl_up_sync(0xffffffffu, inc, o);
		if(lane >= o)
			inc += up;
	}
	if(lane == 31)
		warp_sums[warp] = inc;
	__syncthreads();
	if(warp == 0) {
		int s = lane < (int) (blockDim.x >> 5) ? warp_sums[lane] : 0;
		#pragma unroll
		for(int o = 1; o < 32; o <<= 1) {
			const int up = __shfl_up_sync(0xffffffffu, s, o);
			if(lane >= o)
				s += up;
		}
		warp_sums[lane] = s;                // inclusive over warps
	}
	__syncthreads();
	const int before = warp > 0 ? warp_sums[warp - 1] : 0;
	total = warp_sums[(blockDim.x >> 5) - 1];
	__syncthreads();
	return before + inc - v;
}

// first token of every (padded) global document: rank r's documents are [r max_B, (r + 1) max_B), their tokens start at
// r max_N; one CTA per rank
__global__ void __launch_bounds__(1024) k_csc_doc_starts(const int32_t* __restrict__ len, int64_t max_B, int64_t max_N,
                                                         int32_t* __restrict__ gptr) {
	__shared__ int warp_sums[32];
	const int64_t d0 = (int64_t) blockIdx.x * max_B;
	int running = (int) ((int64_t) blockIdx.x * max_N);
	for(int64_t base = 0; base < max_B; base += blockDim.x) {
		const int64_t d = base + threadIdx.x;
		const int v = d < max_B ? len[d0 + d] : 0;
		int total;
		const int before = block_exclusive_scan(v, total, warp_sums);
		if(d < max_B)
			gptr[d0 + d] = running + before;
		running += total;
	}
}

// document of every token: one warp per document
__global__ void __launch_bounds__(256) k_csc_token_docs(const int32_t* __restrict__ len, const int32_t* __restrict__ gptr, int64_t B,
                                                        int32_t* __restrict__ tok_doc_all) {
	const int64_t d = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if(d >= B)
		return;
	const int begin = gptr[d], n = len[d];
	for(int t = threadIdx.x & 31; t < n; t += 32)
		tok_doc_all[begin + t] = (int32_t) d;
}

__global__ void __launch_bounds__(256) k_csc_count(const int32_t* __restrict__ ids, int64_t N, int64_t chunk, int v0, int v1,
                                                   int32_t* __restrict__ hist) {
	const int nv = v1 - v0;
	const int64_t begin = (int64_t) blockIdx.x * chunk, end = min(N, begin + chunk);
	int32_t* h = hist + (int64_t) blockIdx.x * nv;
	for(int64_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
		const int w = ids[i];
		if(w >= v0 && w < v1)
			atomicAdd(h + (w - v0), 1);
	}
}

// per word: tokens in the chunks before c (in place), and the word's total
__global__ void __launch_bounds__(256) k_csc_chunk_scan(int32_t* __restrict__ hist, int nv, int chunks, int32_t* __restrict__ totals) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if(j >= nv)
		return;
	int run = 0;
	for(int c = 0; c < chunks; ++c) {
		const int x = hist[(int64_t) c * nv + j];
		hist[(int64_t) c * nv + j] = run;
		run += x;
	}
	totals[j] = run;
}

// word_ptr over all V + 1 entries: 0 up to v0, the running count inside [v0, v1), the total from v1 on; one CTA
__global__ void __launch_bounds__(1024) k_csc_word_ptr(const int32_t* __restrict__ totals, int v0, int v1, int V,
                                                       int32_t* __restrict__ word_ptr) {
	__shared__ int warp_sums[32];
	const int nv = v1 - v0;
	int running = 0;
	for(int base = 0; base < nv; base += blockDim.x) {
		const int j = base + threadIdx.x;
		const int v = j < nv ? totals[j] : 0;
		int total;
		const int before = block_exclusive_scan(v, total, warp_sums);
		if(j < nv)
			word_ptr[v0 + j] = running + before;
		running += total;
	}
	for(int w = threadIdx.x; w < v0; w += blockDim.x)
		word_ptr[w] = 0;
	for(int w = v1 + threadIdx.x; w <= V; w += blockDim.x)
		word_ptr[w] = running;
}

// one warp per chunk, in token order
__global__ void __launch_bounds__(256) k_csc_place(const int32_t* __restrict__ ids, const int32_t* __restrict__ tok_doc_all, int64_t N,
                                                   int64_t chunk, int chunks, int v0, int v1, int32_t* hist,
                                                   const int32_t* __restrict__ word_ptr, int32_t* __restrict__ tok_doc,
                                                   int32_t* __restrict__ tok_src) {
	const int c = (int) ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	if(c >= chunks)
		return;
	const int lane = threadIdx.x & 31, nv = v1 - v0;
	const int64_t begin = (int64_t) c * chunk, end = min(N, begin + chunk);
	volatile int32_t* h = hist + (int64_t) c * nv;     // read and written by different lanes from one group to the next
	for(int64_t base = begin; base < end; base += 32) {
		const int64_t i = base + lane;
		const int w = i < end ? ids[i] : -1;
		const bool own = w >= v0 && w < v1;
		// lanes holding the same word of this rank; the others get a key of their own
		const unsigned peers = __match_any_sync(0xffffffffu, own ? w : -1 - lane);
		const int leader = __ffs(peers) - 1;
		const int rank_in_group = __popc(peers & ((1u << lane) - 1u));
		int start = 0;
		if(own && lane == leader) {
			start = h[w - v0];
			h[w - v0] = start + __popc(peers);
		}
		start = __shfl_sync(0xffffffffu, start, leader);
		if(own) {
			const int pos = word_ptr[w] + start + rank_in_group;
			tok_src[pos] = (int32_t) i;
			tok_doc[pos] = tok_doc_all[i];
		}
		__syncwarp();          // the leaders' updates of h are visible to the next group's leaders
	}
}

// lengths of this rank's documents, padded with empty ones: out[d] = doc_ptr[d + 1] - doc_ptr[d] (d < B), 0 (B <= d < pad_B)
__global__ void __launch_bounds__(256) k_csc_doc_lengths(const int64_t* __restrict__ doc_ptr, int64_t B, int64_t pad_B, int32_t* __restrict__ out) {
	const int64_t d = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if(d < pad_B)
		out[d] = d < B ? (int32_t) (doc_ptr[d + 1] - doc_ptr[d]) : 0;
}

}  // namespace

void launch_doc_lengths(const int64_t* doc_ptr, int64_t B, int64_t pad_B, int32_t* out, cudaStream_t s) {
	if(pad_B > 0)
		k_csc_doc_lengths<<<(unsigned) ((pad_B + 255) / 256), 256, 0, s>>>(doc_ptr, B, pad_B, out);
}

size_t global_csc_scratch_ints(int64_t B, int64_t N, int v0, int v1) {
	return (size_t) (B + 1) + (size_t) std::max<int64_t>(N, 1) + (size_t) kCscChunks * (size_t) std::max(v1 - v0, 1) + (size_t) std::max(v1 - v0, 1);
}

// len[B], ids[N] (-1: padding): the gathered minibatch, R rank segments of max_B documents / max_N tokens each.
// scratch: global_csc_scratch_ints(...) int32.  Outputs: word_ptr[V + 1], tok_doc / tok_src (capacity N).
void launch_global_csc(const int32_t* len, const int32_t* ids, int R, int64_t max_B, int64_t max_N, int v0, int v1, int V,
                       int32_t* scratch, int32_t* word_ptr, int32_t* tok_doc, int32_t* tok_src, cudaStream_t s) {
	const int64_t B = (int64_t) R * max_B, N = (int64_t) R * max_N;
	const int nv = std::max(v1 - v0, 0);
	int32_t* gptr = scratch;
	int32_t* tok_doc_all = gptr + (B + 1);
	int32_t* hist = tok_doc_all + std::max<int64_t>(N, 1);
	int32_t* totals = hist + (size_t) kCscChunks * (size_t) std::max(nv, 1);
	if(B > 0 && max_B > 0) {
		k_csc_doc_starts<<<R, 1024, 0, s>>>(len, max_B, max_N, gptr);
		k_csc_token_docs<<<(unsigned) ((B * 32 + 255) / 256), 256, 0, s>>>(len, gptr, B, tok_doc_all);
	}
	if(nv == 0 || N == 0) {
		cudaMemsetAsync(word_ptr, 0, sizeof(int32_t) * ((size_t) V + 1), s);
		return;
	}
	const int64_t chunk = (N + kCscChunks - 1) / kCscChunks;
	cudaMemsetAsync(hist, 0, sizeof(int32_t) * (size_t) kCscChunks * (size_t) nv, s);
	k_csc_count<<<kCscChunks, 256, 0, s>>>(ids, N, chunk, v0, v1, hist);
	k_csc_chunk_scan<<<(nv + 255) / 256, 256, 0, s>>>(hist, nv, kCscChunks, totals);
	k_csc_word_ptr<<<1, 1024, 0, s>>>(totals, v0, v1, V, word_ptr);
	k_csc_place<<<(kCscChunks * 32 + 255) / 256, 256, 0, s>>>(ids, tok_doc_all, N, chunk, kCscChunks, v0, v1, hist, word_ptr, tok_doc, tok_src);
}

}  // namespace trlda
