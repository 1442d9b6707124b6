// estep_fast.cu — the tuned per-document E-step kernel (lda.cpp:174-204 of the reference) for documents whose
// K x n_d tile of expElogbeta columns fits in the shared memory of one thread-block cluster.
//
// Design (B200-first):
//   * one cluster of C CTAs per document, tile split by TOPIC ROWS: CTA r keeps rows [r*kc, r*kc+kc) of every
//     column of the document resident in shared memory for the whole fixed point — HBM sees each column once;
//   * the gather is asynchronous: one 16-byte `cp.async` (LDGSTS) per lane moves a 512-byte column slice per warp
//     instruction straight into shared memory, all slices in flight at once, while the threads evaluate
//     exp(psi(gamma0)) (TRLDA_ESTEP_TMA=1 selects one `cp.async.bulk` per slice on the TMA engine instead);
//   * tile columns are padded by 16 bytes so that BOTH passes are conflict-free 128-bit shared-memory loads:
//       pass 1  acc_k  = sum_j W_j tile[j][k]      lanes along rows (4 floats / 2 doubles per lane), columns split
//                                                  over thread groups and combined through shared memory
//       pass 2  phi_j  = sum_k etheta_k tile[j][k]  one THREAD per (column, row half) walking down its rows — no
//                                                  cross-lane reduction at all (a quarter warp touches 8 x 16 B =
//                                                  all 32 banks); the two row halves meet in shared memory
//   * NO cluster barrier inside the fixed point: every CTA stages its partial phi-norms (and its share of
//     sum|delta gamma|) in its own shared memory and sends the row to every CTA of the cluster with ONE
//     DSMEM bulk copy per destination (`cp.async.bulk.shared::cluster.shared::cta`), which counts itself on
//     the RECEIVER's mbarrier (double-buffered); a CTA continues as soon as the C rows addressed to it have
//     landed, then sums them in rank order — identical bits everywhere, so the convergence decision is
//     cluster-uniform and the result is deterministic.  The one cluster barrier at start-up (exchange
//     barriers initialised everywhere) is split into arrive / wait around the gather.
#include "kernels.cuh"
#include "special.cuh"

#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace trlda {

constexpr int FAST_THREADS = 384;   // 12 warps; two co-resident CTAs per SM still fit 85 registers per thread
constexpr int FAST_WARPS = FAST_THREADS / 32;

struct FastSmem {
	size_t tile, pall, pstage, W, gam, eth, red, wsum, wid, cnt, bar, total;
	int ld;   // column stride of the tile in elements (kc + 16 bytes of padding)
};

__host__ __device__ inline FastSmem fast_smem_layout(int C, int kc, int n_cap, int elem) {
	FastSmem L;
	size_t o = 0;
	auto take = [&o](size_t bytes) { size_t at = o; o += (bytes + 15) & ~size_t(15); return at; };
	const int vec = 16 / elem;
	L.ld = kc + vec;
	L.tile = take((size_t) n_cap * L.ld * elem);
	L.pall = take((size_t) 2 * C * ((size_t) n_cap * elem + 16));     // [2][C] rows: n_cap partials + the delta part
	L.pstage = take((size_t) 2 * ((size_t) n_cap * elem + 16));        // [2] outgoing row
	L.W = take((size_t) n_cap * elem);
	L.gam = take((size_t) kc * 8);
	L.eth = take((size_t) kc * elem);
	L.red = take((size_t) (FAST_THREADS / (kc / vec)) * kc * 8 + 256);
	L.wsum = take(FAST_WARPS * 8);
	L.wid = take((size_t) n_cap * 4);
	L.cnt = take((size_t) n_cap * 4);
	L.bar = take(32);
	L.total = o;
	return L;
}

// is the fast kernel applicable, and with which cluster size?  Returns cluster = 0 if not.
EStepPlan plan_estep_fast(int K, int n_max, int elem, int smem_optin, int force_cluster) {
	EStepPlan plan;
	plan.cluster = 0;
	if((K * elem) % 16 != 0)
		return plan;                                      // column slices must be 16-byte aligned for the async copies
	const int vec = 16 / elem;
	const int n_cap = std::max(4, (n_max + 3) / 4 * 4);
	const size_t budget_full = (size_t) smem_optin - 1024;
	const size_t budget_half = ((size_t) smem_optin + 1024) / 2 - 2048;
	const int candidates[4] = {1, 2, 4, 8};
	for(int pass = 0; pass < 2; ++pass)
		for(int ci = 0; ci < 4; ++ci) {
			const int C = candidates[ci];
			if(force_cluster > 0 && C != force_cluster)
				continue;
			const int kc = (((K + C - 1) / C) + 31) / 32 * 32;
			if(C > 1 && (C - 1) * kc >= K)
				continue;
			if(kc > FAST_THREADS)
				continue;                                 // one gamma row per thread
			if(kc / vec > FAST_THREADS)
				continue;
			const FastSmem L = fast_smem_layout(C, kc, n_cap, elem);
			if(L.total <= (pass == 0 ? budget_half : budget_full)) {
				plan.cluster = C;
				plan.kc = kc;
				plan.n_cap = n_cap;
				plan.n_fit = n_cap;
				plan.smem = L.total;
				plan.fast = true;
				return plan;
			}
		}
	return plan;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, int rank) {
	uint32_t remote;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_addr), "r"(rank));
	return remote;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	uint32_t done = 0;
	while(!done)
		asm volatile(
			"{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };

__device__ __forceinline__ void vec_get(const float4& v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void vec_get(const double2& v, double (&o)[2]) { o[0] = v.x; o[1] = v.y; }

template <typename T, bool CLUSTERED, bool USE_TMA>
__global__ void __launch_bounds__(FAST_THREADS, 2)
k_estep_fast(EStepArgs a, DeviceDocs docs, const int32_t* __restrict__ order, int64_t doc_offset, int C, int kc, int n_cap) {
	extern __shared__ __align__(128) unsigned char smem[];
	using V = typename Vec<T>::type;
	constexpr int VN = Vec<T>::N;
	const FastSmem L = fast_smem_layout(C, kc, n_cap, (int) sizeof(T));
	const int ld = L.ld;
	T* tile = reinterpret_cast<T*>(smem + L.tile);
	unsigned char* pall = smem + L.pall;
	unsigned char* pstage = smem + L.pstage;
	T* W = reinterpret_cast<T*>(smem + L.W);
	double* gam = reinterpret_cast<double*>(smem + L.gam);
	T* eth = reinterpret_cast<T*>(smem + L.eth);
	double* red = reinterpret_cast<double*>(smem + L.red);
	int* wid = reinterpret_cast<int*>(smem + L.wid);
	int* cnt = reinterpret_cast<int*>(smem + L.cnt);
	double* wsum = reinterpret_cast<double*>(smem + L.wsum);
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);   // [0] TMA gather, [1..2] exchange buffers

	cg::cluster_group cluster = cg::this_cluster();
	const int rank = CLUSTERED ? (int) cluster.block_rank() : 0;
	const int64_t slot = doc_offset + blockIdx.x / C;
	const int64_t d = order ? order[slot] : slot;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int K = a.K;
	const int k0 = rank * kc;
	const int kn = max(0, min(kc, K - k0));
	const int64_t begin = docs.doc_ptr[d];
	const int n = (int) (docs.doc_ptr[d + 1] - begin);
	const T* __restrict__ beta = static_cast<const T*>(a.beta);

	// this thread's gamma0 row: issue the load now so that its latency hides behind the document staging
	const double g0_early = tid < kn ? a.gamma[d * K + k0 + tid] : 0.0;

	// phase timers: CTA rank 0, thread 0 only, when a.ticks is given
	long long t_prev = 0;
	const bool timing = a.ticks != nullptr && rank == 0 && tid == 0;
	if(timing)
		t_prev = clock64();
	auto tick = [&](int phase) {
		if(timing) {
			const long long now = clock64();
			atomicAdd(a.ticks + phase, (unsigned long long) (now - t_prev));
			t_prev = now;
		}
	};

	// ---- barriers first, so that the start-up cluster barrier overlaps everything below -----------------------------
	if(tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 2)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if(CLUSTERED)
		asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");

	// ---- stage word ids / counts, start the gather ----------------------------------------------------------------------
	for(int j = tid; j < n; j += FAST_THREADS) {
		wid[j] = docs.word_ids[begin + j];
		cnt[j] = docs.counts[begin + j];
	}
	__syncthreads();
	if(USE_TMA) {
		if(warp == 0 && kn > 0 && n > 0) {
			const uint32_t col_bytes = (uint32_t) kn * sizeof(T);
			if(lane == 0)
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(col_bytes * (uint32_t) n) : "memory");
			__syncwarp();
			for(int j = lane; j < n; j += 32) {
				const T* src = beta + (int64_t) wid[j] * K + k0;
				asm volatile(
					"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					::"r"(smem_u32(tile + (size_t) j * ld)), "l"(src), "r"(col_bytes), "r"(smem_u32(bar)) : "memory");
			}
		}
	} else {
		// one warp instruction copies 32 x 16 B of a column slice; every slice of the document is in flight at once
		for(int j = warp; j < n; j += FAST_WARPS) {
			const T* src = beta + (int64_t) wid[j] * K + k0;
			const uint32_t dst = smem_u32(tile + (size_t) j * ld);
			for(int r = lane * VN; r < kn; r += 32 * VN)
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t) (r * sizeof(T))), "l"(src + r) : "memory");
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	}
	// rows beyond this CTA's share (last rank only) contribute zeros
	if(kn < kc)
		for(int j = warp; j < n; j += FAST_WARPS)
			for(int r = kn + lane; r < kc; r += 32)
				tile[(size_t) j * ld + r] = T(0);
	// gamma0 slice and exp(psi(gamma0)) while the copies are in flight (lda.cpp:174)
	if(tid < kc) {
		const double e = tid < kn ? exp_digamma_for<T>(g0_early, 0.0) : 0.0;
		gam[tid] = g0_early;
		eth[tid] = (T) e;
	}
	tick(0);   // staging, gather issue, exp(psi(gamma0))
	if(USE_TMA) {
		if(kn > 0 && n > 0)
			mbar_wait(smem_u32(bar), 0);
	} else {
		asm volatile("cp.async.wait_group 0;" ::: "memory");
	}
	__syncthreads();
	tick(1);   // wait for the gather

	// ---- pass 2: one thread per column, walking down the rows with 128-bit loads -----------------------------------------
	const uint32_t pall_addr = smem_u32(pall), pstage_addr = smem_u32(pstage), bar_addr = smem_u32(bar);
	const uint32_t row_bytes = (uint32_t) (n_cap * sizeof(T) + 16);   // n_cap partials, then the delta part
	// exchange e uses buffer e & 1 and phase (e >> 1) & 1 of that buffer's barrier
	auto pass2_push = [&](int e, double delta_part) {
		const int buf = e & 1;
		if(tid == 0)   // this CTA expects one row from each of the C CTAs on its own barrier
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
				::"r"(bar_addr + 8 * (1 + buf)), "r"((uint32_t) C * row_bytes) : "memory");
		T* stage = reinterpret_cast<T*>(pstage + (size_t) buf * row_bytes);
		// threads [0, NH) take the upper half of the rows of column tid, threads [NH, 2 NH) the lower half
		constexpr int NH = FAST_THREADS / 2;
		const int half = tid / NH, jt = tid % NH;
		const int r_begin = half ? kc / 2 : 0, r_end = half ? kc : kc / 2;
		for(int j0 = 0; j0 < n; j0 += NH) {
			const int j = j0 + jt;
			T part = T(0);
			if(j < n) {
				const T* col = tile + (size_t) j * ld;
				T acc0 = T(0), acc1 = T(0);
				#pragma unroll 4
				for(int r = r_begin; r < r_end; r += 2 * VN) {
					T e0[VN], v0[VN], e1[VN], v1[VN];
					vec_get(*reinterpret_cast<const V*>(eth + r), e0);
					vec_get(*reinterpret_cast<const V*>(col + r), v0);
					vec_get(*reinterpret_cast<const V*>(eth + r + VN), e1);
					vec_get(*reinterpret_cast<const V*>(col + r + VN), v1);
					#pragma unroll
					for(int q = 0; q < VN; ++q) {
						acc0 = fma(e0[q], v0[q], acc0);
						acc1 = fma(e1[q], v1[q], acc1);
					}
				}
				part = acc0 + acc1;
				if(half == 0)
					stage[j] = part;
			}
			__syncthreads();
			if(half == 1 && j < n)
				stage[j] += part;
		}
		if(tid == 0)
			*reinterpret_cast<double*>(pstage + (size_t) buf * row_bytes + (size_t) n_cap * sizeof(T)) = delta_part;
		__syncthreads();
		if(e == 0 && CLUSTERED)   // start-up barrier: every CTA has initialised its exchange barriers
			asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
		if(tid < C) {
			// make the staged row visible to the async proxy, then one DSMEM bulk copy to CTA `tid`
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			const uint32_t dst = map_to_rank(pall_addr + (uint32_t) ((buf * C + rank) * row_bytes), tid);
			const uint32_t dst_bar = map_to_rank(bar_addr + 8 * (1 + buf), tid);
			asm volatile(
				"cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				::"r"(dst), "r"(pstage_addr + (uint32_t) buf * row_bytes), "r"(row_bytes), "r"(dst_bar) : "memory");
		}
	};

	// once the C rows have landed: phiNorm_j = sum of the partials in rank order (+1e-100),
	// W_j = c_j / phiNorm_j (lda.cpp:183,192,199).  Returns the cluster-wide sum of |delta gamma|.
	auto finish_exchange = [&](int e) -> double {
		const int buf = e & 1;
		mbar_wait(bar_addr + 8 * (1 + buf), (uint32_t) ((e >> 1) & 1));
		const unsigned char* rows = pall + (size_t) buf * C * row_bytes;
		double delta = 0.0;
		for(int src = 0; src < C; ++src)
			delta += *reinterpret_cast<const double*>(rows + (size_t) src * row_bytes + (size_t) n_cap * sizeof(T));
		for(int j = tid; j < n; j += FAST_THREADS) {
			double phi = 0.0;
			for(int src = 0; src < C; ++src)
				phi += (double) reinterpret_cast<const T*>(rows + (size_t) src * row_bytes)[j];
			phi += 1e-100;
			// a weight beyond the float32 range (phi underflowed) would turn 0 * inf into NaN in the float32 products
			W[j] = (T) (sizeof(T) == 4 ? fmin((double) cnt[j] / phi, 1e30) : (double) cnt[j] / phi);
		}
		__syncthreads();
		return delta;
	};

	// ---- pass 1: acc_k = sum_j W_j tile[j][k]; VN rows per thread, columns split over CS thread groups ---------------
	const int RG = kc / VN;                                  // row groups
	const int CS = FAST_THREADS / RG;                        // column splits
	const int rg = tid % RG, cs = tid / RG;
	auto pass1 = [&]() {
		if(cs < CS) {
			T acc[VN];
			#pragma unroll
			for(int q = 0; q < VN; ++q)
				acc[q] = T(0);
			const T* col = tile + rg * VN;
			#pragma unroll 4
			for(int j = cs; j < n; j += CS) {
				const T w = W[j];
				T v[VN];
				vec_get(*reinterpret_cast<const V*>(col + (size_t) j * ld), v);
				#pragma unroll
				for(int q = 0; q < VN; ++q)
					acc[q] = fma(w, v[q], acc[q]);
			}
			#pragma unroll
			for(int q = 0; q < VN; ++q)
				red[cs * kc + rg * VN + q] = (double) acc[q];
		}
		__syncthreads();
	};
	auto pass1_total = [&](int row) -> double {
		double acc = 0.0;
		for(int q = 0; q < CS; ++q)
			acc += red[q * kc + row];
		return acc;
	};

	pass2_push(0, 0.0);
	tick(2);   // initial pass 2 + push (+ start-up cluster barrier wait)
	finish_exchange(0);
	tick(3);   // initial exchange wait + W

	int it = 0;
	while(it < a.max_iter) {
		pass1();
		tick(4);   // pass 1
		double delta_local = 0.0;
		if(tid < kn) {                                        // lda.cpp:186-197
			const double acc = pass1_total(tid);
			const double g_old = gam[tid];
			double g_new = acc * (double) eth[tid];
			g_new += a.alpha[k0 + tid];
			delta_local = fabs(g_old - g_new);
			gam[tid] = g_new;
			eth[tid] = (T) exp_digamma_for<T>(g_new, 0.0);
		}
		// block sum of |delta gamma| (the barrier also orders the eth writes before pass 2)
		delta_local = warp_sum(delta_local);
		if(lane == 0)
			wsum[warp] = delta_local;
		__syncthreads();
		double delta_part = 0.0;
		for(int q = 0; q < FAST_WARPS; ++q)
			delta_part += wsum[q];
		tick(5);   // gamma update, psi, delta reduction
		++it;
		pass2_push(it, delta_part);
		tick(6);   // pass 2 + push
		const double delta = finish_exchange(it);
		tick(7);   // exchange wait + W
		if(delta / K < a.threshold)                           // lda.cpp:202
			break;
	}

	// every CTA that passes this point has received all rows addressed to it, i.e. nobody reads this CTA's
	// staging buffers any more once the matching wait at the end returns (DSMEM must outlive its readers)
	if(CLUSTERED)
		asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");

	// ---- results -------------------------------------------------------------------------------------------------------
	if(tid < kn) {
		a.gamma[d * K + k0 + tid] = gam[tid];
		a.etheta[d * K + k0 + tid] = (double) eth[tid];
		if(a.etheta32)
			a.etheta32[d * K + k0 + tid] = (float) eth[tid];
	}
	for(int j = tid * C + rank; j < n; j += FAST_THREADS * C)
		a.weight[begin + j] = (double) W[j];
	if(a.doc_stat) {
		pass1();
		if(tid < kn)
			a.doc_stat[d * K + k0 + tid] = pass1_total(tid) * (double) eth[tid];
	}
	if(rank == 0 && tid == 0 && a.iterations)
		a.iterations[d] = it;
	if(CLUSTERED)
		asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	tick(8);   // results + doc_stat pass
	if(timing) {
		atomicAdd(a.ticks + 14, (unsigned long long) it);
		atomicAdd(a.ticks + 15, 1ull);
	}
}

static bool g_use_tma = false;

template <typename T, bool CLUSTERED, bool USE_TMA>
static void configure_one(int smem_optin) {
	cudaFuncSetAttribute(k_estep_fast<T, CLUSTERED, USE_TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
}

void configure_estep_fast(int smem_optin) {
	if(const char* e = getenv("TRLDA_ESTEP_TMA"))
		g_use_tma = atoi(e) != 0;
	configure_one<double, true, true>(smem_optin);
	configure_one<double, false, true>(smem_optin);
	configure_one<float, true, true>(smem_optin);
	configure_one<float, false, true>(smem_optin);
	configure_one<double, true, false>(smem_optin);
	configure_one<double, false, false>(smem_optin);
	configure_one<float, true, false>(smem_optin);
	configure_one<float, false, false>(smem_optin);
}

template <typename T, bool CLUSTERED, bool USE_TMA>
static void launch_fast_t(const EStepPlan& plan, const EStepArgs& args, const DeviceDocs& docs, const int32_t* order,
                          int64_t offset, int64_t count, cudaStream_t s) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned) (count * plan.cluster));
	cfg.blockDim = dim3(FAST_THREADS);
	cfg.dynamicSmemBytes = plan.smem;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = plan.cluster;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;   // always a cluster launch (size 1 included): mapa / DSMEM bulk copies require one
	cudaLaunchKernelEx(&cfg, k_estep_fast<T, CLUSTERED, USE_TMA>, args, docs, order, offset, plan.cluster, plan.kc, plan.n_cap);
}

template <typename T>
static void launch_fast_p(const EStepPlan& plan, const EStepArgs& args, const DeviceDocs& docs, const int32_t* order,
                          int64_t offset, int64_t count, cudaStream_t s) {
	if(plan.cluster > 1) {
		if(g_use_tma) launch_fast_t<T, true, true>(plan, args, docs, order, offset, count, s);
		else launch_fast_t<T, true, false>(plan, args, docs, order, offset, count, s);
	} else {
		if(g_use_tma) launch_fast_t<T, false, true>(plan, args, docs, order, offset, count, s);
		else launch_fast_t<T, false, false>(plan, args, docs, order, offset, count, s);
	}
}

void launch_estep_fast(const EStepPlan& plan, const EStepArgs& args, const DeviceDocs& docs, const int32_t* order,
                       int64_t offset, int64_t count, int elem_size, cudaStream_t s) {
	if(count == 0)
		return;
	if(elem_size == 8)
		launch_fast_p<double>(plan, args, docs, order, offset, count, s);
	else
		launch_fast_p<float>(plan, args, docs, order, offset, count, s);
}

}  // namespace trlda
