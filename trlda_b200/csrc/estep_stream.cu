// estep_stream.cu — the per-document E-step (lda.cpp:174-204 of the reference), streaming design (the default path).
//
// The K x n_d tile of expElogbeta columns of a document is NOT kept on chip (150 columns x 4 KB at cfg-3: no SM holds
// it); it is streamed once per inner iteration.  Every warp pulls its own columns (K contiguous values each) through a
// private shared-memory ring — one `cp.async.bulk` per column, issued by one lane, completing on the slot's mbarrier
// (16-byte `cp.async` for columns shorter than 2 KB); there is no block-wide barrier on the data path — and, because a
// warp sees a complete column, the two passes of the reference's inner iteration fuse into ONE sweep:
//
//     phi_j = etheta . col_j  (+1e-100)      lda.cpp:183,199      warp-shuffle reduction over the K rows
//     W_j   = c_j / phi_j                    lda.cpp:192
//     acc  += W_j col_j                      lda.cpp:189-193      lane-local: every lane keeps its rows in registers
//
// After a sweep the per-warp partial sums meet in shared memory (fixed order: deterministic), gamma and
// exp(psi(gamma)) are updated (lda.cpp:194-197) and the convergence test of lda.cpp:202 is evaluated.  I inner
// iterations cost I+1 sweeps (the last one produces the token weights and the document's share of the row sums of
// the sufficient statistics); the first sweep comes from HBM, the re-sweeps from L2 (148 documents x 600 KB in flight).
//
// Launch shape: one CTA of 16 warps per document (fp32 tile; 8 warps for the fp64 tile and for short columns), handed
// out by the hardware scheduler.  Kept behind environment switches because they were measured, and lost:
// a CLUSTER of C CTAs per document that split its COLUMNS and exchange one K-vector of partial sums per sweep by DSMEM
// bulk copies (every CTA then adds the C partial vectors in rank order — identical bits everywhere — and updates gamma
// redundantly); a persistent grid with fewer documents in flight; deeper rings with fewer warps.
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


// how many per-warp partial vectors meet in shared memory at once: above 64 KB the warps fold pairwise first
__host__ __device__ constexpr int stream_reduce_width(int NW, int kp, int elem) {
	return (NW * kp * elem > 65536 && NW % 2 == 0) ? NW / 2 : NW;
}

struct StreamSmem {
	size_t ring, red, eth, gam, wsum, xstage, xall, wid, cnt, bar, rbar, total;
};

__host__ __device__ inline StreamSmem stream_smem_layout(int NW, int C, int depth, int kp, int n_cap, int elem) {
	StreamSmem L;
	size_t o = 0;
	auto take = [&o](size_t bytes) { size_t at = o; o += (bytes + 15) & ~size_t(15); return at; };
	L.ring = take((size_t) NW * depth * kp * elem);
	L.red = take((size_t) stream_reduce_width(NW, kp, elem) * kp * elem);
	L.eth = take((size_t) kp * elem);
	L.gam = take((size_t) kp * 8);
	L.wsum = take((size_t) NW * 8);
	L.xstage = take(C > 1 ? (size_t) 2 * kp * elem : 0);          // [2] outgoing partial vector
	L.xall = take(C > 1 ? (size_t) 2 * C * kp * elem : 0);        // [2][C] incoming partial vectors
	L.wid = take((size_t) n_cap * 4);
	L.cnt = take((size_t) n_cap * 4);
	L.bar = take(32);
	L.rbar = take((size_t) NW * depth * 8);                // one mbarrier per ring slot (bulk-copy variant)
	L.total = o;
	return L;
}

static inline int stream_nvec(int K, int elem) {
	const int per_sweep = 32 * (16 / elem);
	int nvec = 1;
	while(nvec * per_sweep < K)
		nvec *= 2;
	return nvec;
}

// launch shape: cluster size, warps per CTA, ring slots per warp.  Measured at cfg-3 (fresh minibatch per step):
// clusters lose to their per-sweep exchange (2 x 8 warps: 102 ms/step against 74); with the bulk-copy ring the E-step
// takes 56.9 ms/step with 16 warps x 2 slots, 59.5 with 12 x 3, 60.3 with 12 x 2, 75.1 with 8 x 4: warps, not slots.
// TRLDA_STREAM_CLUSTER / TRLDA_STREAM_WARPS / TRLDA_STREAM_DEPTH override (only compiled combinations are accepted).
struct StreamShape { int C, NW, D; };
static const StreamShape kFloatShapes[] = {{1, 16, 2}, {1, 12, 3}, {1, 12, 2}, {1, 8, 4}, {1, 8, 2}, {1, 4, 2}, {2, 8, 2}};
static const StreamShape kDoubleShapes[] = {{1, 8, 2}, {1, 4, 2}, {1, 4, 4}, {2, 4, 2}};

static StreamShape stream_shape(int elem, int K) {
	const StreamShape* shapes = elem == 4 ? kFloatShapes : kDoubleShapes;
	const int count = elem == 4 ? (int) (sizeof(kFloatShapes) / sizeof(StreamShape)) : (int) (sizeof(kDoubleShapes) / sizeof(StreamShape));
	const char* ec = getenv("TRLDA_STREAM_CLUSTER");
	const char* ew = getenv("TRLDA_STREAM_WARPS");
	const char* ed = getenv("TRLDA_STREAM_DEPTH");
	const int C = ec ? atoi(ec) : 0, NW = ew ? atoi(ew) : 0, D = ed ? atoi(ed) : 0;
	for(int i = 0; i < count; ++i)
		if((!C || shapes[i].C == C) && (!NW || shapes[i].NW == NW) && (!D || shapes[i].D == D))
			return (C || NW || D) ? shapes[i] : (elem == 4 && K * elem < 2048 ? StreamShape{1, 8, 2} : shapes[i]);
	return shapes[0];
}

// applicable?  (aligned columns, K small enough for the per-lane register tile, everything fits in shared memory)
bool stream_estep_applicable(int K, int n_max, int elem, int smem_optin) {
	if((K * elem) % 16 != 0)
		return false;
	const int nvec = stream_nvec(K, elem);
	if(nvec > (elem == 4 ? 8 : 16))
		return false;
	const int kp = nvec * 32 * (16 / elem);
	const int n_cap = std::max(32, (n_max + 31) / 32 * 32);
	const StreamShape sh = stream_shape(elem, K);
	return stream_smem_layout(sh.NW, sh.C, sh.D, kp, n_cap, elem).total <= (size_t) smem_optin - 1024;
}

__device__ __forceinline__ uint32_t s_smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t s_map_to_rank(uint32_t smem_addr, int rank) {
	uint32_t remote;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_addr), "r"(rank));
	return remote;
}

__device__ __forceinline__ void s_mbar_wait(uint32_t bar, uint32_t parity) {
	uint32_t done = 0;
	while(!done)
		asm volatile(
			"{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

template <typename T, int NW, int NVEC, int C, bool BULK, int STREAM_DEPTH>
__global__ void __launch_bounds__(NW * 32, (C == 1 && NW == 4 ? 3 : (C == 1 && NW == 8 && sizeof(T) == 4 ? 2 : 1)))
k_estep_stream(EStepArgs a, DeviceDocs docs, const int32_t* __restrict__ order, int64_t doc_offset, int64_t count, int n_cap) {
	extern __shared__ __align__(128) unsigned char smem[];
	using V = typename SVec<T>::type;
	constexpr int VN = SVec<T>::N;
	constexpr int KP = NVEC * 32 * VN;                       // padded number of topic rows
	constexpr int NT = NW * 32;
	constexpr int TW = NW * C;                               // warps working on one document
	constexpr int RW = stream_reduce_width(NW, KP, (int) sizeof(T));
	const StreamSmem L = stream_smem_layout(NW, C, STREAM_DEPTH, KP, n_cap, (int) sizeof(T));
	T* ring = reinterpret_cast<T*>(smem + L.ring);
	T* red = reinterpret_cast<T*>(smem + L.red);
	T* eth = reinterpret_cast<T*>(smem + L.eth);
	double* gam = reinterpret_cast<double*>(smem + L.gam);
	double* wsum = reinterpret_cast<double*>(smem + L.wsum);
	T* xstage = reinterpret_cast<T*>(smem + L.xstage);
	T* xall = reinterpret_cast<T*>(smem + L.xall);
	int* wid = reinterpret_cast<int*>(smem + L.wid);
	int* cnt = reinterpret_cast<int*>(smem + L.cnt);
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);
	uint64_t* rbar = reinterpret_cast<uint64_t*>(smem + L.rbar);

	int rank = 0;
	if(C > 1)
		asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int gw = rank * NW + warp;                          // this warp's index among the document's warps
	const int K = a.K;
	const T* __restrict__ beta = static_cast<const T*>(a.beta);

	if(C > 1) {
		if(tid == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_smem_u32(bar)));
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_smem_u32(bar + 1)));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
		// start-up cluster barrier (exchange barriers initialised everywhere): arrive now, wait before the first push
		asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	}

	// this lane's rows: vectors v = 0..NVEC-1 cover rows (v * 32 + lane) * VN .. + VN
	T* my_ring = ring + (size_t) warp * STREAM_DEPTH * KP;
	const uint32_t ring_addr = s_smem_u32(my_ring);
	const uint32_t rbar_addr = s_smem_u32(rbar + (size_t) warp * STREAM_DEPTH);
	// whole rounds of 32 vectors are moved, rows K..KP-1 of a column are the head of the next column (or the zeroed slack
	// behind the matrix): finite values that meet etheta = 0 in the dot product and rows nobody reads in the sums
	const uint32_t lane_bytes = (uint32_t) lane * 16u;
	const uint32_t col_bytes = (uint32_t) K * (uint32_t) sizeof(T);
	uint32_t phases = 0;                                           // bulk variant: parity of the next completion per slot
	auto issue = [&](int j, int stage) {                           // gather column j into ring slot `stage`
		if(BULK) {
			// one bulk copy of the K values per column, issued by one lane, written by the copy engine: the load/store
			// pipe only sees the LDS that read the column back (LDGSTS costs it 8 cycles per 512 bytes, LDS.128 4)
			if(lane == 0) {
				const T* src = beta + (int64_t) wid[j] * K;
				const uint32_t dst = ring_addr + (uint32_t) (stage * KP * (int) sizeof(T));
				const uint32_t mb = rbar_addr + 8u * (uint32_t) stage;
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(col_bytes) : "memory");
				asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					::"r"(dst), "l"(src), "r"(col_bytes), "r"(mb) : "memory");
			}
		} else {
			const char* src = reinterpret_cast<const char*>(beta + (int64_t) wid[j] * K) + lane_bytes;
			const uint32_t dst = ring_addr + (uint32_t) (stage * KP * (int) sizeof(T)) + lane_bytes;
			#pragma unroll
			for(int v = 0; v < NVEC; ++v)
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + v * 512u), "l"(src + v * 512) : "memory");
		}
	};
	auto commit = [&]() {
		if(!BULK)
			asm volatile("cp.async.commit_group;" ::: "memory");
	};
	auto wait_slot = [&](int stage) {
		if(BULK) {
			s_mbar_wait(rbar_addr + 8u * (uint32_t) stage, (phases >> stage) & 1u);
			phases ^= 1u << stage;
		} else {
			asm volatile("cp.async.wait_group %0;" ::"n"(STREAM_DEPTH - 1) : "memory");
		}
	};
	if(BULK) {
		// the copies move K values, the lanes read KP: the tail of every slot is zeroed once
		for(int i = tid; i < NW * STREAM_DEPTH * (KP - K); i += NT)
			ring[(size_t) (i / (KP - K)) * KP + K + i % (KP - K)] = T(0);
		if(lane == 0) {
			for(int st = 0; st < STREAM_DEPTH; ++st)
				asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rbar_addr + 8u * (uint32_t) st));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		__syncthreads();
	}

	const uint32_t xstage_addr = s_smem_u32(xstage), xall_addr = s_smem_u32(xall), bar_addr = s_smem_u32(bar);
	constexpr uint32_t XBYTES = KP * sizeof(T);

	int sweep = 0;         // counts the sweeps (= cluster exchanges) of all documents of this CTA
	// debug phase timers (TRLDA_ESTEP_TICKS=1), thread 0's view: setup, stream, fold, exchange, update, test
	const bool timing = a.ticks != nullptr && tid == 0;
	long long tk[6] = {0, 0, 0, 0, 0, 0}, t_mark = 0;
	long long docs_done = 0;
	#define TRLDA_TICK(i) if(timing) { const long long now = clock64(); tk[i] += now - t_mark; t_mark = now; }
	// persistent grid: the documents in flight (and with them the footprint of their tiles in L2) are set by the grid
	for(int64_t item = blockIdx.x / C; item < count; item += gridDim.x / C) {
	const int64_t slot = doc_offset + item;
	const int64_t d = order ? order[slot] : slot;
	const int64_t begin = docs.doc_ptr[d];
	const int n = (int) (docs.doc_ptr[d + 1] - begin);
	__syncthreads();       // the previous document's last reads of the shared vectors
	if(timing)
		t_mark = clock64();
	for(int j = tid; j < n; j += NT) {
		wid[j] = docs.word_ids[begin + j];
		cnt[j] = docs.counts[begin + j];
	}
	for(int r0 = tid; r0 < KP; r0 += 2 * NT) {                      // two rows per trip: their psi/exp chains overlap
		double g[2];
		T e[2];
		#pragma unroll
		for(int u = 0; u < 2; ++u) {
			const int r = r0 + u * NT;
			g[u] = r < K ? a.gamma[d * K + r] : 1.0;
		}
		#pragma unroll
		for(int u = 0; u < 2; ++u)
			e[u] = (T) exp_digamma_for<T>(g[u], 0.0);             // lda.cpp:174
		#pragma unroll
		for(int u = 0; u < 2; ++u) {
			const int r = r0 + u * NT;
			if(r < KP) {
				gam[r] = r < K ? g[u] : 0.0;
				eth[r] = r < K ? e[u] : T(0);
			}
		}
	}
	__syncthreads();

	TRLDA_TICK(0)
	int it = 0;
	bool converged = false;
	bool primed = false;   // the ring already holds the first columns of the coming sweep
	int stage = 0;
	while(true) {
		const bool final_sweep = converged || it >= a.max_iter;
		// ---- one sweep over the document's columns: this warp takes columns gw, gw + TW, ... ---------------------------
		T e[NVEC][VN], acc[NVEC][VN];
		#pragma unroll
		for(int v = 0; v < NVEC; ++v) {
			svec_get(*reinterpret_cast<const V*>(eth + (v * 32 + lane) * VN), e[v]);
			#pragma unroll
			for(int q = 0; q < VN; ++q)
				acc[v][q] = T(0);
		}
		// M = number of columns of this warp.  If the ring was not primed by the previous sweep, fill it now.
		const int M = gw < n ? (n - gw + TW - 1) / TW : 0;
		if(!primed) {
			#pragma unroll
			for(int s = 0; s < STREAM_DEPTH; ++s) {
				if(s < M)
					issue(gw + s * TW, s);
				commit();
			}
			stage = 0;
		}
		// a non-final sweep is always followed by another one over the same columns: its first STREAM_DEPTH columns
		// are requested while the tail of this sweep is consumed, so the ring never drains at a sweep boundary and
		// the (L2) latency of the next sweep hides behind the reduction and the psi evaluations
		const bool prime_next = !final_sweep && M >= STREAM_DEPTH;
		for(int mcol = 0; mcol < M; ++mcol) {
			const int j = gw + mcol * TW;
			wait_slot(stage);
			__syncwarp();
			const T* col = my_ring + (size_t) stage * KP;
			T c[NVEC][VN];
			#pragma unroll
			for(int v = 0; v < NVEC; ++v)
				svec_get(*reinterpret_cast<const V*>(col + (v * 32 + lane) * VN), c[v]);
			T wt;
			if(sizeof(T) == 4) {
				// mixed mode: four independent float chains per lane, float butterfly over the warp, float division
				float part[4] = {0.f, 0.f, 0.f, 0.f};
				#pragma unroll
				for(int v = 0; v < NVEC; ++v)
					#pragma unroll
					for(int q = 0; q < VN; ++q)
						part[q & 3] = fmaf((float) e[v][q], (float) c[v][q], part[q & 3]);
				float phi = (part[0] + part[1]) + (part[2] + part[3]);
				#pragma unroll
				for(int o = 16; o > 0; o >>= 1)
					phi += __shfl_xor_sync(0xffffffffu, phi, o);
				// lda.cpp:183,192,199; the +1e-100 only matters at 0, where the weight must stay finite in float32
				wt = (T) fminf(__fdividef((float) cnt[j], fmaxf(phi, 1e-37f)), 1e30f);
			} else {
				double part = 0.0;
				#pragma unroll
				for(int v = 0; v < NVEC; ++v) {
					T chunk = e[v][0] * c[v][0];
					#pragma unroll
					for(int q = 1; q < VN; ++q)
						chunk = fma(e[v][q], c[v][q], chunk);
					part += (double) chunk;
				}
				const double phi = warp_sum(part) + 1e-100;                    // lda.cpp:183,199
				wt = (T) ((double) cnt[j] / phi);                              // lda.cpp:192
			}
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
				issue(gw + mn * TW, stage);
			else if(prime_next)
				issue(gw + (mn - M) * TW, stage);
			commit();
			stage = stage + 1 == STREAM_DEPTH ? 0 : stage + 1;
		}
		primed = prime_next;
		if(!primed && !BULK)
			asm volatile("cp.async.wait_group 0;" ::: "memory");
		// ---- the per-warp partial sums meet in shared memory, fixed order ------------------------------------------------
		TRLDA_TICK(1)
		auto store_acc = [&](int slot) {
			#pragma unroll
			for(int v = 0; v < NVEC; ++v) {
				V out;
				T* o = reinterpret_cast<T*>(&out);
				#pragma unroll
				for(int q = 0; q < VN; ++q)
					o[q] = acc[v][q];
				*reinterpret_cast<V*>(red + (size_t) slot * KP + (v * 32 + lane) * VN) = out;
			}
		};
		if(RW < NW) {                                                  // upper half of the warps folds into the lower half
			if(warp >= RW)
				store_acc(warp - RW);
			__syncthreads();
			if(warp < RW) {
				#pragma unroll
				for(int v = 0; v < NVEC; ++v) {
					T other[VN];
					svec_get(*reinterpret_cast<const V*>(red + (size_t) warp * KP + (v * 32 + lane) * VN), other);
					#pragma unroll
					for(int q = 0; q < VN; ++q)
						acc[v][q] += other[q];
				}
			}
			__syncthreads();
			if(warp < RW)
				store_acc(warp);
		} else {
			store_acc(warp);
		}
		__syncthreads();

		TRLDA_TICK(2)
		const int buf = sweep & 1;
		if(C > 1) {
			// ---- cluster exchange: this CTA's partial K-vector goes to every CTA of the cluster (itself included) -------
			if(tid == 0)
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
					::"r"(bar_addr + 8 * buf), "r"((uint32_t) C * XBYTES) : "memory");
			T* out = xstage + (size_t) buf * KP;
			for(int r = tid; r < KP; r += NT) {
				double partial = 0.0;
				#pragma unroll 4
				for(int q = 0; q < RW; ++q)
					partial += (double) red[(size_t) q * KP + r];
				out[r] = (T) partial;
			}
			__syncthreads();
			if(sweep == 0)
				asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
			if(tid < C) {
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				const uint32_t dst = s_map_to_rank(xall_addr + (uint32_t) ((buf * C + rank) * XBYTES), tid);
				const uint32_t dst_bar = s_map_to_rank(bar_addr + 8 * buf, tid);
				asm volatile(
					"cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					::"r"(dst), "r"(xstage_addr + (uint32_t) buf * XBYTES), "r"(XBYTES), "r"(dst_bar) : "memory");
			}
			s_mbar_wait(bar_addr + 8 * buf, (uint32_t) ((sweep >> 1) & 1));
		}

		TRLDA_TICK(3)
		double delta_local = 0.0;
		// two rows per thread and trip, so that their chains (sum of the partials, psi, exp) overlap
		for(int r0 = tid; r0 < K; r0 += 2 * NT) {
			double total[2], eo[2];
			int row[2];
			bool live[2];
			#pragma unroll
			for(int u = 0; u < 2; ++u) {
				live[u] = r0 + u * NT < K;
				row[u] = live[u] ? r0 + u * NT : r0;
				const int r = row[u];
				total[u] = 0.0;
				if(C > 1) {
					#pragma unroll
					for(int src = 0; src < C; ++src)
						total[u] += (double) xall[(size_t) (buf * C + src) * KP + r];   // rank order: identical bits in every CTA
				} else {
					// the partials are T: they are added in T (mixed mode: one float32 rounding per partial, like the
					// accumulators themselves) and converted once — F2F.F64.F32 per partial was a cost of its own
					T part = red[r];
					#pragma unroll
					for(int q = 1; q < RW; ++q)
						part += red[(size_t) q * KP + r];
					total[u] = (double) part;
				}
				eo[u] = (double) eth[r];
			}
			if(final_sweep) {
				#pragma unroll
				for(int u = 0; u < 2; ++u) {
					const int r = row[u];
					if(live[u] && (C == 1 || r % C == rank)) {             // the CTAs share the output rows
						a.doc_stat[d * K + r] = total[u] * eo[u];
						a.gamma[d * K + r] = gam[r];
						a.etheta[d * K + r] = eo[u];
						if(a.etheta32)
							a.etheta32[d * K + r] = (float) eo[u];
					}
				}
			} else {                                                   // lda.cpp:186-197
				double g_new[2];
				T e_new[2];
				#pragma unroll
				for(int u = 0; u < 2; ++u) {
					g_new[u] = total[u] * eo[u];
					g_new[u] += a.alpha[row[u]];
				}
				#pragma unroll
				for(int u = 0; u < 2; ++u)
					e_new[u] = (T) exp_digamma_for<T>(g_new[u], 0.0);     // fp64 also in mixed mode: a float32 evaluation
					                                                      // (5e-7) pushed lambda past the 1e-4 bound of the mode
				#pragma unroll
				for(int u = 0; u < 2; ++u) {
					if(live[u]) {
						const int r = row[u];
						delta_local += fabs(gam[r] - g_new[u]);
						gam[r] = g_new[u];
						eth[r] = e_new[u];
					}
				}
			}
		}
		++sweep;
		TRLDA_TICK(4)
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
		// no second barrier: wsum and red are next written behind the barriers of the coming sweep's fold
		TRLDA_TICK(5)
	}
	++docs_done;
	if(rank == 0 && tid == 0 && a.iterations)
		a.iterations[d] = it;
	if(rank == 0 && tid == 0 && a.sweeps)
		atomicAdd(a.sweeps, (unsigned long long) (it + 1));
	}
	if(timing && rank == 0) {
		for(int i = 0; i < 6; ++i)
			atomicAdd(a.ticks + i, (unsigned long long) tk[i]);
		atomicAdd(a.ticks + 14, (unsigned long long) sweep);
		atomicAdd(a.ticks + 15, (unsigned long long) docs_done);
	}
	#undef TRLDA_TICK
	if(C > 1) {
		// nobody reads this CTA's staging buffers any more once every CTA has passed its last exchange
		asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
		asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	}
}

// documents in flight: one CTA (cluster) per document, handed out by the hardware scheduler as SMs free up (a static
// persistent grid measured 7 % slower: the documents of a CTA differ in sweeps).  TRLDA_STREAM_GRID = n runs a
// persistent grid of n documents in flight instead (experiments on the L2 footprint of the tiles).
static int64_t stream_grid_docs(int64_t count) {
	if(const char* e = getenv("TRLDA_STREAM_GRID"))
		if(atoi(e) > 0)
			return std::min<int64_t>(count, atoi(e));
	return count;
}

template <typename T, int NW, int NVEC, int C, bool BULK, int DEPTH>
static void launch_stream_b(const EStepArgs& args, const DeviceDocs& docs, const int32_t* order, int64_t offset,
                            int64_t count, int n_cap, size_t smem, cudaStream_t s) {
	auto kernel = k_estep_stream<T, NW, NVEC, C, BULK, DEPTH>;
	static size_t configured = 0;      // per instantiation: the attribute is set once (and again only if the tile grows)
	if(smem > configured) {
		cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		configured = smem;
	}
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned) (stream_grid_docs(count) * C));
	cfg.blockDim = dim3(NW * 32);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = C;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = C > 1 ? 1 : 0;
	cudaLaunchKernelEx(&cfg, kernel, args, docs, order, offset, count, n_cap);
}

// columns travel by cp.async.bulk (one copy per column, written by the copy engine); TRLDA_STREAM_BULK = 0 goes back to
// 16-byte cp.async (LDGSTS), which costs the load/store pipe 8 cycles per 512 bytes: 63.8 against 56.9 ms/step of E-step
// Short columns (K = 100 or 200: 400-1600 bytes) stay with cp.async, the per-copy cost of the bulk path does not pay:
// cfg-5 (K = 200) 117 ms per batch with bulk copies and one 16-warp CTA per SM against 33 ms with two 8-warp CTAs.
static bool stream_bulk(int column_bytes) {
	static const int mode = [] { const char* e = getenv("TRLDA_STREAM_BULK"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
	return mode < 0 ? column_bytes >= 2048 : mode == 1;
}

template <typename T, int NW, int NVEC, int C, int DEPTH>
static void launch_stream_t(const EStepArgs& args, const DeviceDocs& docs, const int32_t* order, int64_t offset,
                            int64_t count, int n_cap, size_t smem, cudaStream_t s) {
	if(C == 1 && stream_bulk(args.K * (int) sizeof(T)))
		launch_stream_b<T, NW, NVEC, 1, true, DEPTH>(args, docs, order, offset, count, n_cap, smem, s);
	else
		launch_stream_b<T, NW, NVEC, C, false, DEPTH>(args, docs, order, offset, count, n_cap, smem, s);
}

template <typename T, int NW, int C, int DEPTH>
static void launch_stream_v(int nvec, const EStepArgs& args, const DeviceDocs& docs, const int32_t* order, int64_t offset,
                            int64_t count, int n_cap, size_t smem, cudaStream_t s) {
	constexpr int MAXV = sizeof(T) == 4 ? 8 : 16;
	switch(nvec) {
		case 1: launch_stream_t<T, NW, 1, C, DEPTH>(args, docs, order, offset, count, n_cap, smem, s); break;
		case 2: launch_stream_t<T, NW, 2, C, DEPTH>(args, docs, order, offset, count, n_cap, smem, s); break;
		case 4: launch_stream_t<T, NW, 4, C, DEPTH>(args, docs, order, offset, count, n_cap, smem, s); break;
		case 8: launch_stream_t<T, NW, 8, C, DEPTH>(args, docs, order, offset, count, n_cap, smem, s); break;
		default: launch_stream_t<T, NW, MAXV, C, DEPTH>(args, docs, order, offset, count, n_cap, smem, s); break;
	}
}

void launch_estep_stream(const EStepArgs& args, const DeviceDocs& docs, const int32_t* order, int64_t offset,
                         int64_t count, int n_max, int elem_size, bool cold, cudaStream_t s) {
	(void) cold;
	if(count == 0)
		return;
	const StreamShape sh = stream_shape(elem_size, args.K);
	const int nvec = stream_nvec(args.K, elem_size);
	const int n_cap = std::max(32, (n_max + 31) / 32 * 32);
	const int kp = nvec * 32 * (16 / elem_size);
	const size_t smem = stream_smem_layout(sh.NW, sh.C, sh.D, kp, n_cap, elem_size).total;
#define TRLDA_STREAM_CASE(TYPE, CLUSTER, WARPS, DEPTH) \
	if(sh.C == CLUSTER && sh.NW == WARPS && sh.D == DEPTH) { \
		launch_stream_v<TYPE, WARPS, CLUSTER, DEPTH>(nvec, args, docs, order, offset, count, n_cap, smem, s); return; }
	if(elem_size == 4) {                                  // keep in step with kFloatShapes / kDoubleShapes
		TRLDA_STREAM_CASE(float, 1, 12, 3)
		TRLDA_STREAM_CASE(float, 1, 16, 2)
		TRLDA_STREAM_CASE(float, 1, 12, 2)
		TRLDA_STREAM_CASE(float, 1, 8, 4)
		TRLDA_STREAM_CASE(float, 1, 8, 2)
		TRLDA_STREAM_CASE(float, 1, 4, 2)
		TRLDA_STREAM_CASE(float, 2, 8, 2)
	} else {
		TRLDA_STREAM_CASE(double, 1, 8, 2)
		TRLDA_STREAM_CASE(double, 1, 4, 2)
		TRLDA_STREAM_CASE(double, 1, 4, 4)
		TRLDA_STREAM_CASE(double, 2, 4, 2)
	}
#undef TRLDA_STREAM_CASE
}

}  // namespace trlda
