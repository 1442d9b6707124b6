// estep_stream.cu — the per-document E-step (lda.cpp:174-204 of the reference) for WARM-STARTED documents.
//
// In trust-region iterations >= 1 the fixed point restarts from the previous iteration's gamma
// (onlinelda.cpp:91-93) and almost every document converges after one or two inner iterations.  For such
// documents keeping the K x n_d tile resident (estep_fast.cu) buys nothing: the cost is the gather plus a few
// latency-bound phases.  This kernel instead STREAMS the tile: one CTA owns a whole document (all K rows), every
// warp pulls its own columns of expElogbeta through a private shared-memory ring with `cp.async` (no block-wide
// barrier on the data path), and — because a warp sees a complete column — the two passes of the reference's
// inner iteration fuse into ONE sweep:
//
//     phi_j = etheta . col_j  (+1e-100)      lda.cpp:183,199      warp-shuffle reduction over the K rows
//     W_j   = c_j / phi_j                    lda.cpp:192
//     acc  += W_j col_j                      lda.cpp:189-193      lane-local: every lane keeps its rows in registers
//
// After a sweep the per-warp partial sums meet in shared memory (fixed order: deterministic), gamma and
// exp(psi(gamma)) are updated (lda.cpp:194-197) and the convergence test of lda.cpp:202 is CTA-local: no
// cluster, no DSMEM.  An inner iteration costs one sweep over the document's columns, which the 126 MB L2
// serves from the second sweep on; I inner iterations cost I+1 sweeps (the last one produces the token weights
// and the document's share of the row sums of the sufficient statistics).
#include "kernels.cuh"
#include "special.cuh"

#include <algorithm>
#include <cstdlib>

namespace trlda {

template <typename T> struct SVec;
template <> struct SVec<float> { using type = float4; static constexpr int N = 4; };
template <> struct SVec<double> { using type = double2; static constexpr int N = 2; };

__device__ __forceinline__ void svec_get(const float4& v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void svec_get(const double2& v, double (&o)[2]) { o[0] = v.x; o[1] = v.y; }

constexpr int STREAM_DEPTH = 2;   // columns in flight per warp

struct StreamSmem {
	size_t ring, red, eth, gam, wsum, wid, cnt, total;
};

__host__ __device__ inline StreamSmem stream_smem_layout(int NW, int kp, int n_cap, int elem) {
	StreamSmem L;
	size_t o = 0;
	auto take = [&o](size_t bytes) { size_t at = o; o += (bytes + 15) & ~size_t(15); return at; };
	L.ring = take((size_t) NW * STREAM_DEPTH * kp * elem);
	L.red = take((size_t) NW * kp * elem);
	L.eth = take((size_t) kp * elem);
	L.gam = take((size_t) kp * 8);
	L.wsum = take((size_t) NW * 8);
	L.wid = take((size_t) n_cap * 4);
	L.cnt = take((size_t) n_cap * 4);
	L.total = o;
	return L;
}

static bool g_stream_shuffle = false; // TRLDA_STREAM_SHUFFLE=1: strided document order (measured: no gain over longest-first)
static int g_stream_warps_f32 = 8;    // TRLDA_STREAM_WARPS: 8 (two CTAs per SM, default) or 16 (one)
static inline int stream_warps(int elem) { return elem == 4 ? g_stream_warps_f32 : 8; }
static inline int stream_nvec(int K, int elem) {
	const int per_sweep = 32 * (16 / elem);
	int nvec = 1;
	while(nvec * per_sweep < K)
		nvec *= 2;
	return nvec;
}

// applicable?  (aligned columns, K small enough for the per-lane register tile, word ids fit in shared memory)
bool stream_estep_applicable(int K, int n_max, int elem, int smem_optin) {
	if((K * elem) % 16 != 0)
		return false;
	const int nvec = stream_nvec(K, elem);
	if(nvec > (elem == 4 ? 8 : 16))
		return false;
	const int kp = nvec * 32 * (16 / elem);
	const int n_cap = std::max(32, (n_max + 31) / 32 * 32);
	return stream_smem_layout(stream_warps(elem), kp, n_cap, elem).total <= (size_t) smem_optin - 1024;
}

template <typename T, int NW, int NVEC>
__global__ void __launch_bounds__(NW * 32, (NW == 4 ? 3 : (NW == 8 && sizeof(T) == 4 ? 2 : 1)))
k_estep_stream(EStepArgs a, DeviceDocs docs, const int32_t* __restrict__ order, int64_t doc_offset, int n_cap,
               unsigned stride, unsigned count) {
	extern __shared__ __align__(128) unsigned char smem[];
	using V = typename SVec<T>::type;
	constexpr int VN = SVec<T>::N;
	constexpr int KP = NVEC * 32 * VN;                       // padded number of topic rows
	constexpr int NT = NW * 32;
	const StreamSmem L = stream_smem_layout(NW, KP, n_cap, (int) sizeof(T));
	T* ring = reinterpret_cast<T*>(smem + L.ring);
	T* red = reinterpret_cast<T*>(smem + L.red);
	T* eth = reinterpret_cast<T*>(smem + L.eth);
	double* gam = reinterpret_cast<double*>(smem + L.gam);
	double* wsum = reinterpret_cast<double*>(smem + L.wsum);
	int* wid = reinterpret_cast<int*>(smem + L.wid);
	int* cnt = reinterpret_cast<int*>(smem + L.cnt);

	// CTA i takes the (i * stride mod count)-th longest document: every wave of CTAs then holds a mix of lengths, the
	// CTAs drift out of phase, and the HBM-bound first sweep of one document overlaps the L2-bound re-sweeps and
	// psi phases of its neighbours instead of all CTAs hitting HBM (then L2) in lockstep
	const int64_t slot = doc_offset + (int64_t) (((uint64_t) blockIdx.x * stride) % count);
	const int64_t d = order ? order[slot] : slot;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int K = a.K;
	const int64_t begin = docs.doc_ptr[d];
	const int n = (int) (docs.doc_ptr[d + 1] - begin);
	const T* __restrict__ beta = static_cast<const T*>(a.beta);

	for(int j = tid; j < n; j += NT) {
		wid[j] = docs.word_ids[begin + j];
		cnt[j] = docs.counts[begin + j];
	}
	for(int r = tid; r < KP; r += NT) {
		double g = 0.0, e = 0.0;
		if(r < K) {
			g = a.gamma[d * K + r];
			e = exp_digamma_for<T>(g, 0.0);                                   // lda.cpp:174
		}
		gam[r] = g;
		eth[r] = (T) e;
	}
	__syncthreads();

	// this lane's rows: vectors v = 0..NVEC-1 cover rows (v * 32 + lane) * VN .. + VN
	T* my_ring = ring + (size_t) warp * STREAM_DEPTH * KP;
	const uint32_t ring_addr = (uint32_t) __cvta_generic_to_shared(my_ring);
	auto issue = [&](int j, int stage) {                           // gather column j into ring slot `stage`
		const T* src = beta + (int64_t) wid[j] * K;
		#pragma unroll
		for(int v = 0; v < NVEC; ++v) {
			const int row = (v * 32 + lane) * VN;
			if(row < K)
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
					::"r"(ring_addr + (uint32_t) ((stage * KP + row) * sizeof(T))), "l"(src + row) : "memory");
		}
	};

	int it = 0;
	bool converged = false;
	bool primed = false;   // the ring already holds the first columns of the coming sweep
	int stage = 0;
	while(true) {
		const bool final_sweep = converged || it >= a.max_iter;
		// ---- one sweep over the document's columns: this warp takes columns warp, warp + NW, ... -----------------------
		T e[NVEC][VN], acc[NVEC][VN];
		#pragma unroll
		for(int v = 0; v < NVEC; ++v) {
			svec_get(*reinterpret_cast<const V*>(eth + (v * 32 + lane) * VN), e[v]);
			#pragma unroll
			for(int q = 0; q < VN; ++q)
				acc[v][q] = T(0);
		}
		// M = number of columns of this warp.  If the ring was not primed by the previous sweep, fill it now.
		const int M = warp < n ? (n - warp + NW - 1) / NW : 0;
		if(!primed) {
			#pragma unroll
			for(int s = 0; s < STREAM_DEPTH; ++s) {
				if(s < M)
					issue(warp + s * NW, s);
				asm volatile("cp.async.commit_group;" ::: "memory");
			}
			stage = 0;
		}
		// a non-final sweep is always followed by another one over the same columns: its first STREAM_DEPTH columns
		// are requested while the tail of this sweep is consumed, so the ring never drains at a sweep boundary and
		// the (L2) latency of the next sweep hides behind the reduction and the psi evaluations
		const bool prime_next = !final_sweep && M >= STREAM_DEPTH;
		for(int mcol = 0; mcol < M; ++mcol) {
			const int j = warp + mcol * NW;
			asm volatile("cp.async.wait_group %0;" ::"n"(STREAM_DEPTH - 1) : "memory");
			__syncwarp();
			const T* col = my_ring + (size_t) stage * KP;
			T c[NVEC][VN];
			double part = 0.0;
			#pragma unroll
			for(int v = 0; v < NVEC; ++v) {
				const int row = (v * 32 + lane) * VN;
				if(row < K) {
					svec_get(*reinterpret_cast<const V*>(col + row), c[v]);
				} else {
					#pragma unroll
					for(int q = 0; q < VN; ++q)
						c[v][q] = T(0);
				}
				// products of one 16-byte chunk are summed in T, chunks are added in float64
				T chunk = e[v][0] * c[v][0];
				#pragma unroll
				for(int q = 1; q < VN; ++q)
					chunk = fma(e[v][q], c[v][q], chunk);
				part += (double) chunk;
			}
			const double phi = warp_sum(part) + 1e-100;                    // lda.cpp:183,199
			const double w = (double) cnt[j] / phi;                        // lda.cpp:192
			const T wt = (T) w;
			#pragma unroll
			for(int v = 0; v < NVEC; ++v)
				#pragma unroll
				for(int q = 0; q < VN; ++q)
					acc[v][q] = fma(wt, c[v][q], acc[v][q]);
			if(final_sweep && lane == 0)
				a.weight[begin + j] = (double) wt;
			__syncwarp();
			const int mn = mcol + STREAM_DEPTH;
			if(mn < M)
				issue(warp + mn * NW, stage);
			else if(prime_next)
				issue(warp + (mn - M) * NW, stage);
			asm volatile("cp.async.commit_group;" ::: "memory");
			stage = stage + 1 == STREAM_DEPTH ? 0 : stage + 1;
		}
		primed = prime_next;
		if(!primed)
			asm volatile("cp.async.wait_group 0;" ::: "memory");
		// ---- the per-warp partial sums meet in shared memory, fixed order ------------------------------------------------
		#pragma unroll
		for(int v = 0; v < NVEC; ++v) {
			V out;
			T* o = reinterpret_cast<T*>(&out);
			#pragma unroll
			for(int q = 0; q < VN; ++q)
				o[q] = acc[v][q];
			*reinterpret_cast<V*>(red + (size_t) warp * KP + (v * 32 + lane) * VN) = out;
		}
		__syncthreads();
		double delta_local = 0.0;
		for(int r = tid; r < K; r += NT) {
			double total = 0.0;
			#pragma unroll 4
			for(int q = 0; q < NW; ++q)
				total += (double) red[(size_t) q * KP + r];
			const double eo = (double) eth[r];
			if(final_sweep) {
				a.doc_stat[d * K + r] = total * eo;
				a.gamma[d * K + r] = gam[r];
				a.etheta[d * K + r] = eo;
				if(a.etheta32)
					a.etheta32[d * K + r] = (float) eo;
			} else {                                                   // lda.cpp:186-197
				const double g_old = gam[r];
				double g_new = total * eo;
				g_new += a.alpha[r];
				delta_local += fabs(g_old - g_new);
				gam[r] = g_new;
				eth[r] = (T) exp_digamma_for<T>(g_new, 0.0);
			}
		}
		if(final_sweep)
			break;
		delta_local = warp_sum(delta_local);
		if(lane == 0)
			wsum[warp] = delta_local;
		__syncthreads();
		double delta = 0.0;
		for(int q = 0; q < NW; ++q)
			delta += wsum[q];
		++it;
		converged = delta / K < a.threshold;                           // lda.cpp:202
		__syncthreads();
	}
	if(tid == 0 && a.iterations)
		a.iterations[d] = it;
}

template <typename T, int NW, int NVEC>
static void launch_stream_t(const EStepArgs& args, const DeviceDocs& docs, const int32_t* order, int64_t offset,
                            int64_t count, int n_cap, size_t smem, cudaStream_t s) {
	cudaFuncSetAttribute(k_estep_stream<T, NW, NVEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	// a stride coprime to the document count visits every document exactly once
	static const unsigned primes[4] = {7919u, 7907u, 7901u, 7883u};
	unsigned stride = 1;
	if(g_stream_shuffle && count > 1)
		for(unsigned p : primes)
			if(count % p != 0) {
				stride = p;
				break;
			}
	k_estep_stream<T, NW, NVEC><<<(unsigned) count, NW * 32, smem, s>>>(args, docs, order, offset, n_cap, stride, (unsigned) count);
}

void launch_estep_stream(const EStepArgs& args, const DeviceDocs& docs, const int32_t* order, int64_t offset,
                         int64_t count, int n_max, int elem_size, cudaStream_t s) {
	if(count == 0)
		return;
	if(const char* e = getenv("TRLDA_STREAM_SHUFFLE"))
		g_stream_shuffle = atoi(e) != 0;
	if(const char* e = getenv("TRLDA_STREAM_WARPS"))
		g_stream_warps_f32 = atoi(e) == 16 ? 16 : (atoi(e) == 4 ? 4 : 8);
	const int nvec = stream_nvec(args.K, elem_size);
	const int n_cap = std::max(32, (n_max + 31) / 32 * 32);
	const int kp = nvec * 32 * (16 / elem_size);
	const size_t smem = stream_smem_layout(stream_warps(elem_size), kp, n_cap, elem_size).total;
	if(elem_size == 4 && stream_warps(4) == 16) {
		switch(nvec) {
			case 1: launch_stream_t<float, 16, 1>(args, docs, order, offset, count, n_cap, smem, s); break;
			case 2: launch_stream_t<float, 16, 2>(args, docs, order, offset, count, n_cap, smem, s); break;
			case 4: launch_stream_t<float, 16, 4>(args, docs, order, offset, count, n_cap, smem, s); break;
			default: launch_stream_t<float, 16, 8>(args, docs, order, offset, count, n_cap, smem, s); break;
		}
	} else if(elem_size == 4 && stream_warps(4) == 4) {
		switch(nvec) {
			case 1: launch_stream_t<float, 4, 1>(args, docs, order, offset, count, n_cap, smem, s); break;
			case 2: launch_stream_t<float, 4, 2>(args, docs, order, offset, count, n_cap, smem, s); break;
			case 4: launch_stream_t<float, 4, 4>(args, docs, order, offset, count, n_cap, smem, s); break;
			default: launch_stream_t<float, 4, 8>(args, docs, order, offset, count, n_cap, smem, s); break;
		}
	} else if(elem_size == 4) {
		switch(nvec) {
			case 1: launch_stream_t<float, 8, 1>(args, docs, order, offset, count, n_cap, smem, s); break;
			case 2: launch_stream_t<float, 8, 2>(args, docs, order, offset, count, n_cap, smem, s); break;
			case 4: launch_stream_t<float, 8, 4>(args, docs, order, offset, count, n_cap, smem, s); break;
			default: launch_stream_t<float, 8, 8>(args, docs, order, offset, count, n_cap, smem, s); break;
		}
	} else {
		switch(nvec) {
			case 1: launch_stream_t<double, 8, 1>(args, docs, order, offset, count, n_cap, smem, s); break;
			case 2: launch_stream_t<double, 8, 2>(args, docs, order, offset, count, n_cap, smem, s); break;
			case 4: launch_stream_t<double, 8, 4>(args, docs, order, offset, count, n_cap, smem, s); break;
			case 8: launch_stream_t<double, 8, 8>(args, docs, order, offset, count, n_cap, smem, s); break;
			default: launch_stream_t<double, 8, 16>(args, docs, order, offset, count, n_cap, smem, s); break;
		}
	}
}

}  // namespace trlda
