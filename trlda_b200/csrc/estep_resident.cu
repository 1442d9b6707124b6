// estep_resident.cu — the per-document E-step (lda.cpp:174-204 of the reference) with the document's tile of
// expElogbeta columns RESIDENT IN REGISTERS for all inner iterations (mixed mode: float32 tile).
//
// Why registers: a cfg-3 document's tile is 150 columns x 4 KB = 600 KB.  No SM's shared memory holds it, the
// streaming kernel (estep_stream.cu) re-reads it from L2 once per inner iteration and is bound by the L2 (9.4 TB/s:
// a 20-iteration E-step moves 103 GB).  The register files of a thread-block cluster do hold it: 8 SMs x 256 KB.
//
// Decomposition.  A document is worked on by a TEAM: one group of WG warps in each of the C CTAs of a cluster.  The
// K topic rows are cut into C slabs of ROWS = 32 WG rows, one per CTA; inside a group warp w owns 32 rows of the slab
// and ALL columns of the document.  Inside a warp lane = (lr, lc), lr = lane / 8, lc = lane % 8: the lane keeps
// 8 rows (those of row group lr) x NJ/8 columns (those congruent to lc mod 8) = 8 x 4 NU floats in registers.
//
//   pass A   phi_j  = sum_k etheta_k D[k, j]       lda.cpp:183,199   in-thread over 8 rows, butterfly over lr (4 lanes)
//   pass B   acc_k  = sum_j (c_j / phi_j) D[k, j]  lda.cpp:189-193   in-thread over 4 NU columns, butterfly over lc (8 lanes)
//
// Both butterflies are "transposed" (reduce-scatter): after pass B lane l holds the complete sum of row 32 w + l —
// every lane then updates ONE topic (gamma, exp(psi(gamma)), lda.cpp:194-197) with full use of the fp64 pipe — and
// after pass A lane l holds the warp's partial phi of columns l + 32 u.  The registers are laid out with a
// lane-dependent XOR permutation (register i of a lane holds row 8 lr + (i ^ lc); register (q, u) holds column
// lc + 8 (q ^ lr) + 32 u), which makes "which half do I keep, which do I send" the same for every lane: the
// butterflies are plain shfl.bfly + add, no selects.  The inner products run on packed FFMA2 (fma.rn.f32x2).
//
// Per inner iteration there is ONE exchange: the group's partial phi (NJ floats, summed over its warps in shared
// memory) and its share of sum_k |delta gamma_k| go to every CTA of the cluster by st.async (DSMEM stores that
// complete on the receiver's mbarrier); every CTA adds the C partials in rank order — identical bits, identical
// convergence decision (lda.cpp:202) everywhere.  No cluster barrier inside the loop.
//
// A CTA hosts G = 8 / WG independent groups working on different documents, so that one team's exchange and psi
// latency hides behind the other's inner products.  Teams are persistent; the tile of a team's NEXT document is
// brought into shared memory by cp.async.bulk (one copy per column slab) while the current one iterates, so the HBM
// gather overlaps the arithmetic; shared memory is only a landing zone, it is read once per document.
#include "kernels.cuh"
#include "special.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <tuple>

namespace trlda {

namespace {

using u64 = unsigned long long;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
	u64 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
	u64 d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
	u64 d;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
	u64 d;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ u64 shfl_xor2(u64 v, int mask) {
	float lo, hi;
	unpack2(v, lo, hi);
	lo = __shfl_xor_sync(0xffffffffu, lo, mask);
	hi = __shfl_xor_sync(0xffffffffu, hi, mask);
	return pack2(lo, hi);
}

__device__ __forceinline__ uint32_t r_smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t r_map_to_rank(uint32_t smem_addr, int rank) {
	uint32_t remote;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_addr), "r"(rank));
	return remote;
}

__device__ __forceinline__ void r_mbar_wait(uint32_t bar, uint32_t parity) {
	uint32_t done = 0;
	while(!done)
		asm volatile(
			"{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

template <int GT>
__device__ __forceinline__ void group_barrier(int g) {
	if(GT == 256)
		__syncthreads();
	else
		asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(GT) : "memory");
}

}  // namespace

// shared memory of one group (bytes); shared by host and device
struct ResidentSmem {
	size_t stage, red, xbuf, dl, bars, ticks, group_total;
};

__host__ __device__ constexpr ResidentSmem resident_smem_layout(int C, int NU, int WG) {
	const int NJ = 32 * NU, ROWS = 32 * WG;
	ResidentSmem L{};
	size_t o = 0;
	L.stage = o; o += (size_t) NJ * ROWS * 4;                       // landing zone of the next document's tile [NJ][ROWS]
	L.red = o; o += (size_t) (C == 1 ? 2 : 1) * WG * NJ * 4;       // per-warp partial phi [parity][WG][NJ]
	L.xbuf = o; o += C > 1 ? (size_t) 2 * C * (NJ + 4) * 4 : 0;    // incoming partials [parity][C][NJ + 4]
	L.dl = o; o += (size_t) 2 * WG * 8;                            // per-warp |delta gamma| sums [parity][WG]
	L.bars = o; o += 32;                                           // xbar[2], tbar
	L.ticks = o; o += 16 * 8;                                      // debug phase timers (TRLDA_ESTEP_TICKS=1)
	L.group_total = (o + 127) & ~size_t(127);
	return L;
}

template <int C, int NU, int WG>
__global__ void __launch_bounds__(256, 1)
k_estep_resident(EStepArgs a, DeviceDocs docs, const int32_t* __restrict__ order, int64_t doc_offset, int64_t count) {
	constexpr int G = 8 / WG;                  // groups (documents in flight) per CTA
	constexpr int GT = 32 * WG;                // threads per group
	constexpr int ROWS = 32 * WG;              // topic rows per CTA
	constexpr int NJ = 32 * NU;                // column capacity
	constexpr int XS = NJ + 4;                 // floats per rank slot of the exchange buffer (the delta sits at [NJ])
	constexpr ResidentSmem L = resident_smem_layout(C, NU, WG);
	extern __shared__ __align__(128) unsigned char smem[];

	const int tid = threadIdx.x;
	const int g = tid / GT, tg = tid % GT, wg = tg >> 5, lane = tid & 31, lr = lane >> 3, lc = lane & 7;
	unsigned char* base = smem + (size_t) g * L.group_total;
	float* stage = reinterpret_cast<float*>(base + L.stage);
	float* red = reinterpret_cast<float*>(base + L.red);
	float* xbuf = reinterpret_cast<float*>(base + L.xbuf);
	double* dl = reinterpret_cast<double*>(base + L.dl);
	uint64_t* bars = reinterpret_cast<uint64_t*>(base + L.bars);
	const uint32_t xbar_addr = r_smem_u32(bars), tbar_addr = r_smem_u32(bars + 2);
	const uint32_t xbuf_addr = r_smem_u32(xbuf), stage_addr = r_smem_u32(stage);
	// debug phase timers, one thread's view, kept in shared memory (the register budget is spent on the tile)
	long long* tk = reinterpret_cast<long long*>(base + L.ticks);
	const bool timing = a.ticks != nullptr && tg == 0;
	#define TRLDA_RTICK(i) if(timing) { const long long now = clock64(); tk[i] += now - tk[15]; tk[15] = now; }

	int rank = 0;
	unsigned cluster_id = blockIdx.x;
	if(C > 1) {
		asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
		asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cluster_id));
	}
	const int64_t n_teams = (int64_t) (gridDim.x / C) * G;
	const int64_t team = (int64_t) cluster_id * G + g;

	const int K = a.K;
	const float* __restrict__ beta = static_cast<const float*>(a.beta);
	const int row0 = rank * ROWS;                                  // first topic row of this CTA's slab
	const int rows_valid = max(0, min(ROWS, K - row0));
	const uint32_t col_bytes = (uint32_t) rows_valid * 4u;
	const int k_mine = row0 + 32 * wg + lane;                      // the topic this lane updates
	const bool live = k_mine < K;
	const double alpha_mine = live ? a.alpha[k_mine] : 0.0;

	// the landing zone starts zeroed: rows beyond K and columns beyond a document's length are never written, and
	// whatever they hold meets etheta = 0 (rows) or a zero count (columns)
	for(int i = tg; i < NJ * ROWS / 4; i += GT)
		reinterpret_cast<float4*>(stage)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
	if(timing)
		for(int i = 0; i < 16; ++i)
			tk[i] = 0;
	if(tg == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xbar_addr));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xbar_addr + 8));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tbar_addr));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	__syncthreads();
	if(C > 1) {
		asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
		asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	}

	// gather of a document's tile slab into the landing zone: one bulk copy per column, completing on tbar
	auto prefetch = [&](int64_t item) {
		const int64_t d = order ? order[doc_offset + item] : doc_offset + item;
		const int64_t begin = docs.doc_ptr[d];
		const int n = (int) (docs.doc_ptr[d + 1] - begin);
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		if(tg == 0)
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tbar_addr), "r"((uint32_t) n * col_bytes) : "memory");
		// warp wg takes the columns 32 wg + 32 WG m + l: its lanes read 32 word ids at once, lane 0 issues the copies
		if(col_bytes)
			for(int j0 = 32 * wg; j0 < n; j0 += GT) {
				const int mine = j0 + lane < n ? docs.word_ids[begin + j0 + lane] : -1;
				const int cols = min(32, n - j0);
				for(int l = 0; l < cols; ++l) {
					const int w = __shfl_sync(0xffffffffu, mine, l);
					if(lane == 0) {
						const float* src = beta + (int64_t) w * K + row0;
						asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
							::"r"(stage_addr + (uint32_t) (j0 + l) * (ROWS * 4u)), "l"(src), "r"(col_bytes), "r"(tbar_addr) : "memory");
					}
				}
			}
	};

	uint32_t seq = 0;          // exchanges done by this team (identical in all its CTAs)
	uint32_t tile_phase = 0;
	if(team < count)
		prefetch(team);

	for(int64_t item = team; item < count; item += n_teams) {
		const int64_t d = order ? order[doc_offset + item] : doc_offset + item;
		const int64_t begin = docs.doc_ptr[d];
		const int n = (int) (docs.doc_ptr[d + 1] - begin);
		float cntf[NU];
		#pragma unroll
		for(int u = 0; u < NU; ++u) {
			const int j = lane + 32 * u;
			cntf[u] = j < n ? (float) docs.counts[begin + j] : 0.f;
		}
		double gam = live ? a.gamma[d * K + k_mine] : 1.0;
		if(timing)
			tk[15] = clock64();

		// ---- tile: landing zone -> registers ---------------------------------------------------------------------
		r_mbar_wait(tbar_addr, tile_phase);
		tile_phase ^= 1u;
		u64 D2[8][2][NU];       // [row register i][column pair t = q / 2][u]: (q = 2t, q = 2t + 1)
		{
			const float* srow = stage + 32 * wg + 8 * lr;
			#pragma unroll
			for(int u = 0; u < NU; ++u)
				#pragma unroll
				for(int t = 0; t < 2; ++t) {
					const float* c0 = srow + (size_t) (lc + 8 * ((2 * t) ^ lr) + 32 * u) * ROWS;
					const float* c1 = srow + (size_t) (lc + 8 * ((2 * t + 1) ^ lr) + 32 * u) * ROWS;
					#pragma unroll
					for(int i = 0; i < 8; ++i)
						D2[i][t][u] = pack2(c0[i ^ lc], c1[i ^ lc]);
				}
		}
		group_barrier<GT>(g);                       // everybody has read the landing zone: it may take the next tile
		if(item + n_teams < count)
			prefetch(item + n_teams);

		double e = live ? exp_digamma_lean(gam) : 0.0;       // lda.cpp:174
		float ef = (float) e;
		TRLDA_RTICK(0)
		double delta_local = 0.0;
		float W[NU];
		double delta_total = 0.0;

		// pass A + exchange: phi of the current etheta -> token weights W of this lane's columns (lane + 32 u)
		auto pass_a = [&]() {
			u64 phi2[2][NU];
			#pragma unroll
			for(int i = 0; i < 8; ++i) {
				const float ei = __shfl_xor_sync(0xffffffffu, ef, i);      // etheta of row 8 lr + (i ^ lc)
				const u64 e2 = pack2(ei, ei);
				#pragma unroll
				for(int t = 0; t < 2; ++t)
					#pragma unroll
					for(int u = 0; u < NU; ++u)
						phi2[t][u] = i == 0 ? fmul2(D2[i][t][u], e2) : ffma2(D2[i][t][u], e2, phi2[t][u]);
			}
			// butterfly over lr (lane bits 3, 4): registers q = 2, 3 go to lane ^ 16, then register 1 to lane ^ 8
			float phiP[NU];
			#pragma unroll
			for(int u = 0; u < NU; ++u) {
				const u64 s = fadd2(phi2[0][u], shfl_xor2(phi2[1][u], 16));
				float lo, hi;
				unpack2(s, lo, hi);
				phiP[u] = lo + __shfl_xor_sync(0xffffffffu, hi, 8);
			}
			const uint32_t par = seq & 1u;
			double dsum = delta_local;
			#pragma unroll
			for(int o = 16; o > 0; o >>= 1)
				dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
			TRLDA_RTICK(3)
			if(C == 1) {
				float* r = red + (size_t) par * WG * NJ;
				#pragma unroll
				for(int u = 0; u < NU; ++u)
					r[wg * NJ + lane + 32 * u] = phiP[u];
				if(lane == 0)
					dl[par * WG + wg] = dsum;
				group_barrier<GT>(g);
				delta_total = 0.0;
				#pragma unroll
				for(int w = 0; w < WG; ++w)
					delta_total += dl[par * WG + w];
				#pragma unroll
				for(int u = 0; u < NU; ++u) {
					float phi = r[lane + 32 * u];
					#pragma unroll
					for(int w = 1; w < WG; ++w)
						phi += r[w * NJ + lane + 32 * u];
					W[u] = fminf(__fdividef(cntf[u], fmaxf(phi, 1e-37f)), 1e30f);      // lda.cpp:183,192,199
				}
			} else {
				#pragma unroll
				for(int u = 0; u < NU; ++u)
					red[wg * NJ + lane + 32 * u] = phiP[u];
				if(lane == 0)
					dl[wg] = dsum;
				group_barrier<GT>(g);
				const uint32_t slot = xbuf_addr + (uint32_t) ((par * C + rank) * XS) * 4u;
				const uint32_t mbar = xbar_addr + 8u * par;
				if(tg == 0) {
					double dt = 0.0;
					#pragma unroll
					for(int w = 0; w < WG; ++w)
						dt += dl[w];
					asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "n"(C * (NJ * 4 + 8)) : "memory");
					#pragma unroll
					for(int r = 0; r < C; ++r)
						asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
							::"r"(r_map_to_rank(slot + NJ * 4u, r)), "l"(__double_as_longlong(dt)), "r"(r_map_to_rank(mbar, r)) : "memory");
				}
				if(tg < NJ / 4) {
					float4 s = *reinterpret_cast<const float4*>(red + 4 * tg);
					#pragma unroll
					for(int w = 1; w < WG; ++w) {
						const float4 o = *reinterpret_cast<const float4*>(red + w * NJ + 4 * tg);
						s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
					}
					#pragma unroll
					for(int r = 0; r < C; ++r)
						asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
							::"r"(r_map_to_rank(slot + 16u * (uint32_t) tg, r)), "r"(__float_as_uint(s.x)), "r"(__float_as_uint(s.y)),
							  "r"(__float_as_uint(s.z)), "r"(__float_as_uint(s.w)), "r"(r_map_to_rank(mbar, r)) : "memory");
				}
				TRLDA_RTICK(4)
				r_mbar_wait(mbar, (seq >> 1) & 1u);
				TRLDA_RTICK(5)
				const float* xb = xbuf + (size_t) par * C * XS;
				delta_total = 0.0;
				#pragma unroll
				for(int r = 0; r < C; ++r)
					delta_total += *reinterpret_cast<const double*>(xb + r * XS + NJ);
				#pragma unroll
				for(int u = 0; u < NU; ++u) {
					float phi = xb[lane + 32 * u];
					#pragma unroll
					for(int r = 1; r < C; ++r)
						phi += xb[r * XS + lane + 32 * u];                           // rank order: identical bits in every CTA
					W[u] = fminf(__fdividef(cntf[u], fmaxf(phi, 1e-37f)), 1e30f);      // lda.cpp:183,192,199
				}
			}
			++seq;
			TRLDA_RTICK(6)
		};

		// pass B: acc of this lane's topic = sum_j W_j D[k, j]
		auto pass_b = [&]() -> float {
			u64 part2[8];
			#pragma unroll
			for(int u = 0; u < NU; ++u) {
				#pragma unroll
				for(int t = 0; t < 2; ++t) {
					// W of column lc + 8 (q ^ lr) + 32 u lives in lane ^ (8 q)
					const float w0 = t == 0 ? W[u] : __shfl_xor_sync(0xffffffffu, W[u], 16);
					const float w1 = __shfl_xor_sync(0xffffffffu, W[u], 16 * t + 8);
					const u64 w2 = pack2(w0, w1);
					#pragma unroll
					for(int i = 0; i < 8; ++i)
						part2[i] = (u == 0 && t == 0) ? fmul2(D2[i][t][u], w2) : ffma2(D2[i][t][u], w2, part2[i]);
				}
			}
			float v[8];
			#pragma unroll
			for(int i = 0; i < 8; ++i) {
				float lo, hi;
				unpack2(part2[i], lo, hi);
				v[i] = lo + hi;
			}
			// butterfly over lc (lane bits 0-2): registers 4..7 go to lane ^ 4, then 2, 3 to lane ^ 2, then 1 to lane ^ 1
			#pragma unroll
			for(int i = 0; i < 4; ++i)
				v[i] += __shfl_xor_sync(0xffffffffu, v[i ^ 4], 4);
			#pragma unroll
			for(int i = 0; i < 2; ++i)
				v[i] += __shfl_xor_sync(0xffffffffu, v[i ^ 2], 2);
			const float total = v[0] + __shfl_xor_sync(0xffffffffu, v[1], 1);
			TRLDA_RTICK(1)
			return total;
		};

		pass_a();                                                   // lda.cpp:183
		int it = 0;
		while(it < a.max_iter) {                                    // lda.cpp:185-204
			const float acc = pass_b();
			const double g_new = fma(e, (double) acc, alpha_mine);  // lda.cpp:189-195
			const double e_new = live ? exp_digamma_lean(g_new) : 0.0;   // lda.cpp:197
			delta_local = live ? fabs(gam - g_new) : 0.0;
			gam = g_new;
			e = e_new;
			ef = (float) e;
			++it;
			TRLDA_RTICK(2)
			pass_a();                                               // lda.cpp:199
			if(delta_total / K < a.threshold)                       // lda.cpp:202
				break;
		}
		delta_local = 0.0;
		const float acc = pass_b();                                 // row sums of this document's sufficient statistics
		if(live) {
			a.gamma[d * K + k_mine] = gam;
			a.etheta[d * K + k_mine] = e;
			if(a.etheta32)
				a.etheta32[d * K + k_mine] = ef;
			a.doc_stat[d * K + k_mine] = (double) acc * e;
		}
		if(rank == 0 && wg == 0) {
			#pragma unroll
			for(int u = 0; u < NU; ++u)
				if(lane + 32 * u < n)
					a.weight[begin + lane + 32 * u] = (double) W[u];
			if(lane == 0 && a.iterations)
				a.iterations[d] = it;
		}
		TRLDA_RTICK(7)
		if(timing) {
			tk[13] += 1;
			tk[14] += it + 1;
		}
	}
	if(timing && rank == 0 && g == 0) {
		for(int i = 0; i < 8; ++i)
			atomicAdd(a.ticks + i, (unsigned long long) tk[i]);
		atomicAdd(a.ticks + 14, (unsigned long long) tk[14]);
		atomicAdd(a.ticks + 15, (unsigned long long) tk[13]);
	}
	#undef TRLDA_RTICK
	if(C > 1) {
		// no CTA may leave while a peer can still store into its exchange buffers
		asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
		asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	}
}

// ---- host side -------------------------------------------------------------------------------------------------

namespace {

struct ResidentKernel {
	const void* fn;
	int C, NU, WG;
	size_t smem;
	int max_clusters;     // co-resident clusters (persistent grid), 0 = not yet known
};

template <int C, int NU, int WG>
ResidentKernel* resident_kernel() {
	static ResidentKernel k = [] {
		ResidentKernel r{};
		r.fn = reinterpret_cast<const void*>(&k_estep_resident<C, NU, WG>);
		r.C = C; r.NU = NU; r.WG = WG;
		r.smem = resident_smem_layout(C, NU, WG).group_total * (8 / WG);
		return r;
	}();
	return &k;
}

template <int NU, int WG>
ResidentKernel* resident_pick_c(int C) {
	switch(C) {
		case 1: return resident_kernel<1, NU, WG>();
		case 2: return resident_kernel<2, NU, WG>();
		case 4: return resident_kernel<4, NU, WG>();
		case 8: if(WG == 4) return resident_kernel<8, NU, 4>();
	}
	return nullptr;
}

template <int WG>
ResidentKernel* resident_pick_nu(int C, int NU) {
	switch(NU) {
		case 2: return resident_pick_c<2, WG>(C);
		case 4: return resident_pick_c<4, WG>(C);
		case 5: return resident_pick_c<5, WG>(C);
		case 6: return resident_pick_c<6, WG>(C);
	}
	return nullptr;
}

int resident_wg(int K) {
	static const int forced = [] { const char* e = getenv("TRLDA_RESIDENT_WG"); return e ? atoi(e) : 0; }();
	if(forced == 4 || forced == 8)
		return forced;
	(void) K;
	return 4;
}

int resident_cluster(int K, int WG) {
	const int rows = 32 * WG;
	int C = 1;
	while(C * rows < K)
		C *= 2;
	return C;
}

}  // namespace

// longest document (pairs) the register tile covers
int resident_estep_max_len() { return 192; }

// applicable?  (mixed mode only: the float32 tile of a document fits the cluster's register files)
bool resident_estep_applicable(int K, int elem_size) {
	if(elem_size != 4 || K % 4 != 0 || K < 1)
		return false;
	const int WG = resident_wg(K);
	const int C = resident_cluster(K, WG);
	return C <= (WG == 4 ? 8 : 4);
}

// runs documents order[offset .. offset + count), all of at most n_max <= resident_estep_max_len() pairs
int launch_estep_resident(const EStepArgs& args, const DeviceDocs& docs, const int32_t* order, int64_t offset,
                          int64_t count, int n_max, cudaStream_t s) {
	if(count == 0)
		return 0;
	const int WG = resident_wg(args.K);
	const int C = resident_cluster(args.K, WG);
	const int NU = n_max <= 64 ? 2 : (n_max <= 128 ? 4 : (n_max <= 160 ? 5 : 6));
	ResidentKernel* k = WG == 4 ? resident_pick_nu<4>(C, NU) : resident_pick_nu<8>(C, NU);
	if(!k)
		return -1;
	cudaLaunchConfig_t cfg = {};
	cfg.blockDim = dim3(256);
	cfg.dynamicSmemBytes = k->smem;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = C;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = C > 1 ? 1 : 0;
	if(k->max_clusters == 0) {
		if(cudaFuncSetAttribute(k->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) k->smem) != cudaSuccess)
			return -1;
		int n = 0;
		if(C > 1) {
			cfg.gridDim = dim3(C);
			if(cudaOccupancyMaxActiveClusters(&n, k->fn, &cfg) != cudaSuccess || n < 1) {
				cudaGetLastError();
				return -1;
			}
		} else {
			int dev = 0, sms = 0, per_sm = 0;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
			if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k->fn, 256, k->smem) != cudaSuccess || per_sm < 1) {
				cudaGetLastError();
				return -1;
			}
			n = sms * per_sm;
		}
		if(const char* e = getenv("TRLDA_RESIDENT_CLUSTERS"))
			if(atoi(e) > 0)
				n = std::min(n, atoi(e));
		k->max_clusters = n;
		if(getenv("TRLDA_RESIDENT_VERBOSE"))
			fprintf(stderr, "[trlda] k_estep_resident<C=%d, NU=%d, WG=%d>: %zu B shared memory, %d co-resident clusters\n",
			        C, NU, WG, k->smem, n);
	}
	const int G = 8 / WG;
	const int64_t teams_needed = (count + G - 1) / G;
	const int clusters = (int) std::min<int64_t>(k->max_clusters, teams_needed);
	cfg.gridDim = dim3((unsigned) (clusters * C));
	void* params[] = {(void*) &args, (void*) &docs, (void*) &order, (void*) &offset, (void*) &count};
	return cudaLaunchKernelExC(&cfg, k->fn, params) == cudaSuccess ? 0 : -1;
}

}  // namespace trlda
