// estep_tmem.cu — the per-document E-step (lda.cpp:174-204 of the reference), mixed mode, with the document's tile of
// expElogbeta columns resident in TENSOR MEMORY (+ registers) for all inner iterations.
//
// Why TMEM.  A cfg-3 document's tile is 150 columns x 4 KB = 600 KB; the inner loop sweeps it twice per iteration, up
// to 20 times.  Streamed from L2 (estep_stream.cu) the E-step is bound by the L2 (9.4 TB/s, 11 ms for a 20-iteration
// E-step).  Blackwell's tensor memory — 256 KB per SM, 128 lanes x 512 32-bit columns, meant for tcgen05.mma
// accumulators — is also a per-thread scratchpad: a warp reads 32 columns of its own 32 lanes with one tcgen05.ld
// (SASS LDTM.x32), measured here at >= 400 B/cycle/SM with an FFMA per value (scripts/tmem_probe.cu), three times the
// shared-memory bandwidth.  TMEM (256 values per thread at 8 warps) plus 64-128 registers per thread hold 320-384 KB
// of tile per SM: a cluster of TWO SMs keeps a whole document on chip, all 148 SMs are used, and the per-iteration
// exchange has a single peer.
//
// Decomposition.  A document is worked on by a TEAM: one group of WG warps in each of the C CTAs of a cluster; the K
// topic rows are cut into C slabs of ROWS = 64 WG rows (K = 1000: C = 2, WG = 8).  Inside a group warp w owns 64 rows
// and ALL columns; lane = (lr, lc) = (lane / 8, lane % 8) keeps 16 rows x 4 NU columns, column blocks u < 4 in its
// private TMEM columns, u >= 4 in registers.
//
//   pass A   phi_j  = sum_k etheta_k D[k, j]       lda.cpp:183,199   in-thread over 16 rows, butterfly over lr (4 lanes)
//   pass B   acc_k  = sum_j (c_j / phi_j) D[k, j]  lda.cpp:189-193   in-thread over 4 NU columns, butterfly over lc (8 lanes)
//
// Both butterflies are transposed reductions (reduce-scatter): after pass B lane l owns rows 64 w + 2 l, + 1 and
// updates those TWO topics (gamma, exp(psi(gamma)), lda.cpp:194-197: two independent fp64 chains per lane); after pass
// A lane l holds the warp's partial phi of columns l + 32 u.  A lane-dependent XOR permutation of the layout (slot i
// of a lane holds row 16 lr + (i ^ 2 lc); slot (q, u) column lc + 8 (q ^ lr) + 32 u) makes the butterflies plain
// shfl.bfly + add without selects; the inner products are packed FFMA2 (fma.rn.f32x2) whose 64-bit tile operand is
// the ROW PAIR (2 m, 2 m + 1) of a column, adjacent in registers and TMEM: pass A multiplies it by (etheta_2m,
// etheta_2m+1), pass B by the duplicated weight (W_j, W_j) - no register moves to form operands in either pass.
//
// Per inner iteration ONE exchange: the group's partial phi (summed over its warps in shared memory) and its share of
// sum |delta gamma| go to every CTA of the cluster by st.async (DSMEM stores completing on the receiver's mbarrier);
// all CTAs add the partials in rank order — identical bits, identical convergence decisions (lda.cpp:202) — then one
// thread per column forms the token weight W_j = c_j / phi_j and leaves it, duplicated, in shared memory in the four
// lane permutations, from where every lane fetches the weights of a column block with two 128-bit loads in pass B.
//
// The tile arrives from HBM in chunks of 32 columns through a three-slot landing ring in shared memory (cp.async.bulk per
// column slab, completing on the slot's mbarrier) and is moved to TMEM by tcgen05.st; the first three chunks of a team's
// NEXT document are requested while the current one still iterates.
//
// ONE persistent launch serves all documents of up to 192 pairs: the body is instantiated per tile shape (NU = 2, 4,
// 5, 6 column blocks) and picked per document; documents are handed out by a work counter, longest first, two
// documents ahead of their use (the draw travels with the exchange message, see "Schedule" in the kernel).
#include "kernels.cuh"
#include "special.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace trlda {

namespace {

using u64 = unsigned long long;

__device__ __forceinline__ u64 t_pack2(float lo, float hi) {
	u64 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ u64 t_pack2u(uint32_t lo, uint32_t hi) {
	u64 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
	return r;
}
__device__ __forceinline__ void t_unpack2(u64 v, float& lo, float& hi) {
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 t_ffma2(u64 a, u64 b, u64 c) {
	u64 d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}
__device__ __forceinline__ u64 t_fmul2(u64 a, u64 b) {
	u64 d;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ u64 t_fadd2(u64 a, u64 b) {
	u64 d;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ u64 t_shfl_xor2(u64 v, int mask) {
	float lo, hi;
	t_unpack2(v, lo, hi);
	lo = __shfl_xor_sync(0xffffffffu, lo, mask);
	hi = __shfl_xor_sync(0xffffffffu, hi, mask);
	return t_pack2(lo, hi);
}

__device__ __forceinline__ uint32_t t_smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t t_map_to_rank(uint32_t smem_addr, int rank) {
	uint32_t remote;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_addr), "r"(rank));
	return remote;
}

__device__ __forceinline__ void t_mbar_wait(uint32_t bar, uint32_t parity) {
	uint32_t done = 0;
	while(!done)
		asm volatile(
			"{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

template <int GT>
__device__ __forceinline__ void t_group_barrier(int g) {
	if(GT == 512)
		__syncthreads();
	else
		asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(GT) : "memory");
}

// 32 consecutive TMEM columns of this thread's lane <-> 32 registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
		  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
		  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
		  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
		: "r"(taddr) : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
	asm volatile(
		"tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
		:: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
		   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
		   "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
		   "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
		   "r"(taddr) : "memory");
}


constexpr int kTmemBlocks = 4;     // column blocks (64 values each) of a thread's tile kept in TMEM: 8 warps x 256 columns
constexpr int kTmemMaxBlocks = 6;  // widest tile shape: 6 blocks = documents of up to 192 pairs

}  // namespace

// shared memory of one group (bytes)
struct TmemSmem {
	size_t ring, red, xbuf, wperm, dl, bars, ticks, group_total;
};

__host__ __device__ constexpr TmemSmem tmem_smem_layout(int C, int NU, int WG) {
	const int NJ = 32 * NU, ROWS = 64 * WG;
	TmemSmem L{};
	size_t o = 0;
	const int NS = 3;
	L.ring = o; o += (size_t) NS * 32 * ROWS * 4;                  // landing ring: NS slots x [32 columns][ROWS]
	L.red = o; o += (size_t) WG * NJ * 4;                          // per-warp partial phi [WG][NJ]
	L.xbuf = o; o += C > 1 ? (size_t) 2 * C * (NJ + 4) * 4 : 0;    // incoming partials [parity][C][NJ + 4]
	L.wperm = o; o += (size_t) 256 * NU * 4;                       // token weights, duplicated, [NU][2][lane][2 x 2]
	L.dl = o; o += (size_t) WG * 4 + 16;                           // per-warp |delta gamma| sums
	o = (o + 15) & ~size_t(15);
	L.bars = o; o += 40;                                           // xbar[2], full[3]
	L.ticks = o; o += 16 * 8;                                      // debug phase timers (TRLDA_ESTEP_TICKS=1)
	L.group_total = (o + 127) & ~size_t(127);
	return L;
}

template <int C, int WG>
__global__ void __launch_bounds__(256, 1)
k_estep_tmem(EStepArgs a, DeviceDocs docs, const int32_t* __restrict__ order, int64_t doc_offset, int64_t count) {
	constexpr int G = 8 / WG;                  // groups (documents in flight) per CTA
	constexpr int GT = 32 * WG;                // threads per group
	constexpr int ROWS = 64 * WG;              // topic rows per CTA
	constexpr int NS = 3;                      // landing slots (chunks of 32 columns in flight)
	constexpr TmemSmem L = tmem_smem_layout(C, kTmemMaxBlocks, WG);   // buffers sized for the widest tile shape
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ uint32_t tmem_base_slot;

	const int tid = threadIdx.x, warp = tid >> 5;
	const int g = tid / GT, tg = tid % GT, wg = tg >> 5, lane = tid & 31, lr = lane >> 3, lc = lane & 7;
	unsigned char* base = smem + (size_t) g * L.group_total;
	float* ring = reinterpret_cast<float*>(base + L.ring);
	float* red = reinterpret_cast<float*>(base + L.red);
	float* xbuf = reinterpret_cast<float*>(base + L.xbuf);
	float* wperm = reinterpret_cast<float*>(base + L.wperm);
	float* dl = reinterpret_cast<float*>(base + L.dl);
	uint64_t* bars = reinterpret_cast<uint64_t*>(base + L.bars);
	const uint32_t xbar_addr = t_smem_u32(bars), full_addr = t_smem_u32(bars + 2);
	const uint32_t xbuf_addr = t_smem_u32(xbuf), ring_addr = t_smem_u32(ring);
	long long* tk = reinterpret_cast<long long*>(base + L.ticks);
	const bool timing = a.ticks != nullptr && tg == 0;
	#define TRLDA_TTICK(i) if(timing) { const long long now = clock64(); tk[i] += now - tk[15]; tk[15] = now; }

	int rank = 0;
	unsigned cluster_id = blockIdx.x;
	if(C > 1) {
		asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
		asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cluster_id));
	}
	const int n_teams = (int) (gridDim.x / C) * G;
	const int team = (int) cluster_id * G + g;

	const int K = a.K;
	const float* __restrict__ beta = static_cast<const float*>(a.beta);
	const int row0 = rank * ROWS;                                  // first topic row of this CTA's slab
	const int rows_valid = max(0, min(ROWS, K - row0));
	const uint32_t col_bytes = (uint32_t) rows_valid * 4u;
	const int k_mine = row0 + 64 * wg + 2 * lane;                  // the two topics this lane updates: k_mine, k_mine + 1
	const bool live = k_mine < K;                                  // K is a multiple of 4: both or none
	double alpha_mine[2] = {0.0, 0.0};
	if(live) {
		alpha_mine[0] = a.alpha[k_mine];
		alpha_mine[1] = a.alpha[k_mine + 1];
	}

	// ---- set-up: TMEM, barriers, zeroed landing ring -----------------------------------------------------------------
	if(warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(t_smem_u32(&tmem_base_slot)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	// rows beyond K of a slab are never written by the copies: they stay zero (and meet etheta = 0 anyway)
	for(int i = tg; i < NS * 32 * ROWS / 4; i += GT)
		reinterpret_cast<float4*>(ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
	if(timing)
		for(int i = 0; i < 16; ++i)
			tk[i] = 0;
	if(tg == 0) {
		for(int i = 0; i < 5; ++i)
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xbar_addr + 8u * i));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	// this thread's private TMEM columns: lanes 32 (warp % 4) .., columns 256 (warp / 4) ..
	const uint32_t taddr = tmem_base_slot + ((uint32_t) (32 * (warp & 3)) << 16) + (uint32_t) ((warp >> 2) * 256);
	if(C > 1) {
		asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
		asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	}

	// document of a work item: index, first pair, number of pairs
	auto doc_of = [&](int item, int& d, int& begin, int& n) {
		d = order ? order[doc_offset + item] : (int) (doc_offset + item);
		const int64_t b = docs.doc_ptr[d];
		begin = (int) b;
		n = (int) (docs.doc_ptr[d + 1] - b);
	};
	// chunk u of a document (columns 32 u ..) into ring slot `slot`: `ids` = word id of column 32 u + lane (every warp of
	// the group holds them); warp wg issues the columns wg, wg + WG, ..., one lane each copy
	auto issue_chunk = [&](int ids, int nv, int slot) {
		if(tg == 0)
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_addr + 8u * slot), "r"((uint32_t) nv * col_bytes) : "memory");
		if(lane == 0)
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		if(col_bytes)
			for(int l = wg; l < nv; l += WG) {
				const int w = __shfl_sync(0xffffffffu, ids, l);
				if(lane == 0) {
					const float* src = beta + (int64_t) w * K + row0;
					asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
						::"r"(ring_addr + (uint32_t) (slot * 32 + l) * (ROWS * 4u)), "l"(src), "r"(col_bytes), "r"(full_addr + 8u * slot) : "memory");
				}
			}
	};
	auto chunk_cols = [](int n, int u) { return max(0, min(32, n - 32 * u)); };

	uint32_t seq = 0;          // exchanges done by this team (identical in all its CTAs)
	uint32_t cc = 0;           // chunks consumed by this group
	int d = 0, begin = 0, n = 0, d_next = 0, begin_next = 0, n_next = 0;
	int ids_next[NS];          // word ids (column 32 u + lane) of the first NS chunks of the next document
	#pragma unroll
	for(int u = 0; u < NS; ++u)
		ids_next[u] = 0;
	// Schedule.  A team's first two documents are fixed (items team, team + n_teams); every further one is drawn from
	// the launch's work counter, longest documents first, TWO documents ahead of its use: the team's leader (rank 0)
	// draws it when a document starts and hands (document, first pair, pairs) to all threads of the team, in every CTA,
	// with the per-iteration exchange - so that the document after the current one is always known early enough for
	// its first chunks to be prefetched.
	bool has_next = team < count;
	int after_d = -1, after_begin = 0, after_n = 0;       // the document after the next one (-1: none)
	if((int64_t) team + n_teams < count)
		doc_of(team + n_teams, after_d, after_begin, after_n);
	const bool leader = rank == 0 && tg == GT - 1;
	if(team < count) {
		doc_of(team, d_next, begin_next, n_next);
		#pragma unroll
		for(int u = 0; u < NS; ++u) {
			ids_next[u] = 32 * u + lane < n_next ? docs.word_ids[begin_next + 32 * u + lane] : 0;
			issue_chunk(ids_next[u], chunk_cols(n_next, u), u);
		}
	}

	// One document, tile shape NU (32 NU columns): the body is instantiated per shape and picked per document, so that
	// ONE persistent launch serves all document lengths (no tail per shape).  The landing ring always sees at least NS
	// chunks per document (empty ones complete at once), which keeps the look-ahead within the next document.
	auto process = [&](auto shape) {
		constexpr int NU = decltype(shape)::value;
		constexpr int RC = NU < NS ? NS : NU;      // chunks of this document in the landing ring
		constexpr int NJ = 32 * NU;                // column capacity
		constexpr int XS = NJ + 4;                 // floats per rank slot of the exchange buffer (the delta sits at [NJ])
		constexpr int UT = NU < kTmemBlocks ? NU : kTmemBlocks;   // column blocks in TMEM
		constexpr int UR = NU - UT;                // column blocks in registers
		constexpr int NCW = (NJ + GT - 1) / GT;    // columns per weight thread
		d = d_next; begin = begin_next; n = n_next;
		const bool more = after_d >= 0;
		has_next = more;
		if(more) {
			d_next = after_d; begin_next = after_begin; n_next = after_n;
		}
		int drawn = 0;             // leader: the item drawn for the document after the next one
		if(leader)
			drawn = 2 * n_teams + atomicAdd(a.work, 1);
		int ids[RC > NS ? RC - NS : 1];    // word ids of this document's chunks NS ..
		#pragma unroll
		for(int u = NS; u < RC; ++u)
			ids[u - NS] = 32 * u + lane < n ? docs.word_ids[begin + 32 * u + lane] : 0;
		if(more) {
			#pragma unroll
			for(int u = 0; u < NS; ++u)
				ids_next[u] = 32 * u + lane < n_next ? docs.word_ids[begin_next + 32 * u + lane] : 0;
		}
		float cntw[NCW], Wmine[NCW];       // weight threads: count and weight of columns tg + c GT
		#pragma unroll
		for(int c = 0; c < NCW; ++c) {
			const int j = tg + c * GT;
			cntw[c] = j < n ? (float) docs.counts[begin + j] : 0.f;
			Wmine[c] = 0.f;
		}
		double gam[2] = {1.0, 1.0};
		if(live) {
			const double2 g2 = *reinterpret_cast<const double2*>(a.gamma + (int64_t) d * K + k_mine);
			gam[0] = g2.x;
			gam[1] = g2.y;
		}
		if(timing)
			tk[15] = clock64();

		// ---- tile: landing ring -> TMEM / registers, chunk by chunk --------------------------------------------------
		// value slot c = 8 (i / 2) + 2 q + i % 2 of column block u: row 16 lr + (i ^ 2 lc), column lc + 8 (q ^ lr) + 32 u;
		// rows 2 m, 2 m + 1 of a column sit in adjacent registers: the operand pairs of the packed FFMA2 of both passes
		uint32_t Dreg[UR > 0 ? UR : 1][64];
		#pragma unroll
		for(int u = 0; u < RC; ++u) {
			const uint32_t slot = cc % NS;
			t_mbar_wait(full_addr + 8u * slot, (cc / NS) & 1u);
			++cc;
			const float* srow = ring + (size_t) slot * 32 * ROWS + 64 * wg + 16 * lr;
			#pragma unroll
			for(int h = 0; h < (u < NU ? 2 : 0); ++h) {
				uint32_t v[32];    // rows i = 8 h .. 8 h + 7
				#pragma unroll
				for(int q = 0; q < 4; ++q) {
					const float* col = srow + (size_t) (lc + 8 * (q ^ lr)) * ROWS;
					#pragma unroll
					for(int i2 = 0; i2 < 4; ++i2) {
						const float2 pr = *reinterpret_cast<const float2*>(col + ((8 * h + 2 * i2) ^ (2 * lc)));
						v[8 * i2 + 2 * q] = __float_as_uint(pr.x);
						v[8 * i2 + 2 * q + 1] = __float_as_uint(pr.y);
					}
				}
				if(u < UT)
					tmem_st32(taddr + 64u * u + 32u * h, v);
				else {
					#pragma unroll
					for(int c = 0; c < 32; ++c)
						Dreg[u < UT ? 0 : u - UT][32 * h + c] = v[c];
				}
			}
			t_group_barrier<GT>(g);                 // everybody has read the slot: it may take another chunk
			if(u + NS < RC)
				issue_chunk(ids[u + NS < RC ? u : 0], chunk_cols(n, u + NS), (int) slot);
			else if(more)
				issue_chunk(ids_next[u + NS - RC < NS ? u + NS - RC : 0], chunk_cols(n_next, u + NS - RC), (int) slot);
		}
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
		int lead_d = -1, lead_begin = 0, lead_n = 0;
		if(leader && drawn < count)
			doc_of(drawn, lead_d, lead_begin, lead_n);

		double e[2];
		float ef[2];
		#pragma unroll
		for(int x = 0; x < 2; ++x) {
			e[x] = live ? exp_digamma_lean(gam[x]) : 0.0;           // lda.cpp:174
			ef[x] = (float) e[x];
		}
		float delta_local = 0.f;
		float delta_total = 0.f;
		TRLDA_TTICK(0)

		// half a column block of the tile: 32 values, rows i = 8 h .. 8 h + 7
		auto block_half = [&](int u, int h, uint32_t (&r)[32]) {
			if(u < UT)
				tmem_ld32(taddr + 64u * u + 32u * h, r);
			else {
				#pragma unroll
				for(int c = 0; c < 32; ++c)
					r[c] = Dreg[u < UT ? 0 : u - UT][32 * h + c];
			}
		};

		// pass A + exchange: phi of the current etheta -> token weights of this lane's columns
		auto pass_a = [&]() {
			u64 es2[8];            // etheta of the row pairs 16 lr + ((2 m, 2 m + 1) ^ 2 lc)
			#pragma unroll
			for(int m = 0; m < 8; ++m)
				es2[m] = t_pack2(__shfl_xor_sync(0xffffffffu, ef[0], m), __shfl_xor_sync(0xffffffffu, ef[1], m));
			float phiP[NU];
			#pragma unroll
			for(int u = 0; u < NU; ++u) {
				u64 phi2[4];       // per column q: (even rows, odd rows)
				#pragma unroll
				for(int h = 0; h < 2; ++h) {
					uint32_t r[32];
					block_half(u, h, r);
					#pragma unroll
					for(int i2 = 0; i2 < 4; ++i2)
						#pragma unroll
						for(int q = 0; q < 4; ++q) {
							const u64 dd = t_pack2u(r[8 * i2 + 2 * q], r[8 * i2 + 2 * q + 1]);
							phi2[q] = (h == 0 && i2 == 0) ? t_fmul2(dd, es2[0]) : t_ffma2(dd, es2[4 * h + i2], phi2[q]);
						}
				}
				// butterfly over lr (lane bits 3, 4): slots q = 2, 3 go to lane ^ 16, then slot 1 to lane ^ 8
				float s[4];
				#pragma unroll
				for(int q = 0; q < 4; ++q) {
					float lo, hi;
					t_unpack2(phi2[q], lo, hi);
					s[q] = lo + hi;
				}
				s[0] += __shfl_xor_sync(0xffffffffu, s[2], 16);
				s[1] += __shfl_xor_sync(0xffffffffu, s[3], 16);
				phiP[u] = s[0] + __shfl_xor_sync(0xffffffffu, s[1], 8);
			}
			float dsum = delta_local;
			#pragma unroll
			for(int o = 16; o > 0; o >>= 1)
				dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
			TRLDA_TTICK(3)
			#pragma unroll
			for(int u = 0; u < NU; ++u)
				red[wg * NJ + lane + 32 * u] = phiP[u];
			if(lane == 0)
				dl[wg] = dsum;
			if(C == 1 && leader) {
				dl[WG] = __int_as_float(lead_d);
				dl[WG + 1] = __int_as_float(lead_begin);
				dl[WG + 2] = __int_as_float(lead_n);
			}
			t_group_barrier<GT>(g);
			const uint32_t par = seq & 1u;
			const float* xb = xbuf + (size_t) par * C * XS;
			if(C > 1) {
				const uint32_t slot = xbuf_addr + (uint32_t) ((par * C + rank) * XS) * 4u;
				const uint32_t mbar = xbar_addr + 8u * par;
				if(tg == GT - 1) {
					float dt = dl[0];
					#pragma unroll
					for(int w = 1; w < WG; ++w)
						dt += dl[w];
					asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "n"(C * (NJ * 4 + 16)) : "memory");
					// {delta, the leader's draw}: the last three words count from rank 0 only
					#pragma unroll
					for(int r = 0; r < C; ++r)
						asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
							::"r"(t_map_to_rank(slot + NJ * 4u, r)), "r"(__float_as_uint(dt)), "r"(lead_d), "r"(lead_begin), "r"(lead_n),
							  "r"(t_map_to_rank(mbar, r)) : "memory");
				}
				for(int j4 = tg; j4 < NJ / 4; j4 += GT) {
					float4 s = *reinterpret_cast<const float4*>(red + 4 * j4);
					#pragma unroll
					for(int w = 1; w < WG; ++w) {
						const float4 o = *reinterpret_cast<const float4*>(red + w * NJ + 4 * j4);
						s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
					}
					#pragma unroll
					for(int r = 0; r < C; ++r)
						asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
							::"r"(t_map_to_rank(slot + 16u * (uint32_t) j4, r)), "r"(__float_as_uint(s.x)), "r"(__float_as_uint(s.y)),
							  "r"(__float_as_uint(s.z)), "r"(__float_as_uint(s.w)), "r"(t_map_to_rank(mbar, r)) : "memory");
				}
				TRLDA_TTICK(4)
				t_mbar_wait(mbar, (seq >> 1) & 1u);
				TRLDA_TTICK(5)
				delta_total = xb[NJ];
				#pragma unroll
				for(int r = 1; r < C; ++r)
					delta_total += xb[r * XS + NJ];
				after_d = __float_as_int(xb[NJ + 1]);
				after_begin = __float_as_int(xb[NJ + 2]);
				after_n = __float_as_int(xb[NJ + 3]);
			} else {
				delta_total = dl[0];
				#pragma unroll
				for(int w = 1; w < WG; ++w)
					delta_total += dl[w];
				after_d = __float_as_int(dl[WG]);
				after_begin = __float_as_int(dl[WG + 1]);
				after_n = __float_as_int(dl[WG + 2]);
			}
			// one thread per column: phi (partials added in rank / warp order: identical bits in every CTA), weight,
			// and the weight's place in the four lane permutations: lane (p, j % 8) finds it, DUPLICATED (the FFMA2 operand
			// of pass B), at [u][q / 2][lane][q % 2] with q = (j / 8 % 4) ^ p
			#pragma unroll
			for(int c = 0; c < NCW; ++c) {
				const int j = tg + c * GT;
				if(NJ % GT == 0 || j < NJ) {
					float phi;
					if(C > 1) {
						phi = xb[j];
						#pragma unroll
						for(int r = 1; r < C; ++r)
							phi += xb[r * XS + j];
					} else {
						phi = red[j];
						#pragma unroll
						for(int w = 1; w < WG; ++w)
							phi += red[w * NJ + j];
					}
					const float wj = fminf(__fdividef(cntw[c], fmaxf(phi, 1e-37f)), 1e30f);   // lda.cpp:183,192,199
					Wmine[c] = wj;
					const int jc = j & 7, jq = (j >> 3) & 3, ju = j >> 5;
					#pragma unroll
					for(int p = 0; p < 4; ++p) {
						const int q = jq ^ p;
						*reinterpret_cast<float2*>(wperm + ((ju * 2 + (q >> 1)) * 32 + p * 8 + jc) * 4 + 2 * (q & 1)) = make_float2(wj, wj);
					}
				}
			}
			t_group_barrier<GT>(g);
			++seq;
			TRLDA_TTICK(6)
		};

		// pass B: acc of this lane's two topics = sum_j W_j D[k, j]
		auto pass_b = [&](float (&acc)[2]) {
			u64 part2[8];          // row pairs (2 m, 2 m + 1)
			#pragma unroll
			for(int u = 0; u < NU; ++u) {
				const float4 wa = *reinterpret_cast<const float4*>(wperm + ((u * 2) * 32 + lane) * 4);
				const float4 wb = *reinterpret_cast<const float4*>(wperm + ((u * 2 + 1) * 32 + lane) * 4);
				const u64 wd[4] = {t_pack2(wa.x, wa.y), t_pack2(wa.z, wa.w), t_pack2(wb.x, wb.y), t_pack2(wb.z, wb.w)};
				#pragma unroll
				for(int h = 0; h < 2; ++h) {
					uint32_t r[32];
					block_half(u, h, r);
					#pragma unroll
					for(int q = 0; q < 4; ++q)
						#pragma unroll
						for(int i2 = 0; i2 < 4; ++i2) {
							const u64 dd = t_pack2u(r[8 * i2 + 2 * q], r[8 * i2 + 2 * q + 1]);
							part2[4 * h + i2] = (u == 0 && q == 0) ? t_fmul2(dd, wd[0]) : t_ffma2(dd, wd[q], part2[4 * h + i2]);
						}
				}
			}
			// butterfly over lc (lane bits 0-2): row slots 8..15 go to lane ^ 4, then 4..7 to lane ^ 2, then 2, 3 to lane ^ 1
			#pragma unroll
			for(int m = 0; m < 4; ++m)
				part2[m] = t_fadd2(part2[m], t_shfl_xor2(part2[m ^ 4], 4));
			#pragma unroll
			for(int m = 0; m < 2; ++m)
				part2[m] = t_fadd2(part2[m], t_shfl_xor2(part2[m ^ 2], 2));
			part2[0] = t_fadd2(part2[0], t_shfl_xor2(part2[1], 1));
			t_unpack2(part2[0], acc[0], acc[1]);
			TRLDA_TTICK(1)
		};

		pass_a();                                                   // lda.cpp:183
		int it = 0;
		float acc[2];
		while(it < a.max_iter) {                                    // lda.cpp:185-204
			pass_b(acc);
			delta_local = 0.f;
			#pragma unroll
			for(int x = 0; x < 2; ++x) {
				const double g_new = fma(e[x], (double) acc[x], alpha_mine[x]);   // lda.cpp:189-195
				const double e_new = live ? exp_digamma_lean(g_new) : 0.0;        // lda.cpp:197
				delta_local += live ? (float) fabs(gam[x] - g_new) : 0.f;
				gam[x] = g_new;
				e[x] = e_new;
				ef[x] = (float) e_new;
			}
			++it;
			TRLDA_TTICK(2)
			pass_a();                                               // lda.cpp:199
			if(delta_total / (float) K < (float) a.threshold)       // lda.cpp:202
				break;
		}
		delta_local = 0.f;
		pass_b(acc);                                                // row sums of this document's sufficient statistics
		if(live) {
			const int64_t o = (int64_t) d * K + k_mine;
			*reinterpret_cast<double2*>(a.gamma + o) = make_double2(gam[0], gam[1]);
			*reinterpret_cast<double2*>(a.etheta + o) = make_double2(e[0], e[1]);
			if(a.etheta32)
				*reinterpret_cast<float2*>(a.etheta32 + o) = make_float2(ef[0], ef[1]);
			*reinterpret_cast<double2*>(a.doc_stat + o) = make_double2((double) acc[0] * e[0], (double) acc[1] * e[1]);
		}
		if(rank == 0) {
			#pragma unroll
			for(int c = 0; c < NCW; ++c)
				if(tg + c * GT < n)
					a.weight[begin + tg + c * GT] = (double) Wmine[c];
			if(tg == 0 && a.iterations)
				a.iterations[d] = it;
			if(tg == 0 && a.sweeps)
				atomicAdd(a.sweeps, (unsigned long long) (it + 1));
		}
		TRLDA_TTICK(7)
		if(timing) {
			tk[13] += 1;
			tk[14] += it + 1;
		}
	};
	while(has_next) {
		const int len = n_next;                    // the document about to be processed
		if(len <= 64)
			process(std::integral_constant<int, 2>{});
		else if(len <= 128)
			process(std::integral_constant<int, 4>{});
		else if(len <= 160)
			process(std::integral_constant<int, 5>{});
		else
			process(std::integral_constant<int, 6>{});
	}
	if(timing && rank == 0 && g == 0) {
		for(int i = 0; i < 8; ++i)
			atomicAdd(a.ticks + i, (unsigned long long) tk[i]);
		atomicAdd(a.ticks + 14, (unsigned long long) tk[14]);
		atomicAdd(a.ticks + 15, (unsigned long long) tk[13]);
	}
	if(timing && rank == 0) {
		// load balance of the static schedule: busiest team against the mean
		long long busy = 0;
		for(int i = 0; i < 8; ++i)
			busy += tk[i];
		atomicMax(a.ticks + 8, (unsigned long long) busy);
		atomicAdd(a.ticks + 9, (unsigned long long) busy);
		atomicAdd(a.ticks + 10, 1ull);
	}
	#undef TRLDA_TTICK
	if(C > 1) {
		// no CTA may leave while a peer can still store into its exchange buffers
		asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
		asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if(warp == 0)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_slot) : "memory");
}

// ---- host side -------------------------------------------------------------------------------------------------

namespace {

struct TmemKernel {
	const void* fn;
	int C, WG;
	size_t smem;
	int max_clusters;     // co-resident clusters (persistent grid), 0 = not yet known
};

template <int C, int WG>
TmemKernel* tmem_kernel() {
	static TmemKernel k = [] {
		TmemKernel r{};
		r.fn = reinterpret_cast<const void*>(&k_estep_tmem<C, WG>);
		r.C = C; r.WG = WG;
		r.smem = tmem_smem_layout(C, kTmemMaxBlocks, WG).group_total * (8 / WG);
		return r;
	}();
	return &k;
}

// cluster size and group width for K topics: ROWS = 64 WG rows per CTA, C CTAs per document
TmemKernel* tmem_select(int K) {
	// TRLDA_TMEM_WG = 4 / 2: narrower groups, more CTAs per document (experiments: two / four documents per SM)
	static const int forced = [] { const char* e = getenv("TRLDA_TMEM_WG"); return e ? atoi(e) : 0; }();
	if(forced == 4 && K > 512 && K <= 1024)
		return tmem_kernel<4, 4>();
	if(K <= 128)
		return tmem_kernel<1, 2>();
	if(K <= 256)
		return tmem_kernel<1, 4>();
	if(K <= 512)
		return tmem_kernel<1, 8>();
	if(K <= 1024)
		return tmem_kernel<2, 8>();
	if(K <= 2048)
		return tmem_kernel<4, 8>();
	return nullptr;
}

}  // namespace

// longest document (pairs) the TMEM tile covers
int tmem_estep_max_len() { return 192; }

bool tmem_estep_applicable(int K, int elem_size) {
	if(elem_size != 4 || K % 4 != 0 || K < 1)
		return false;
	return tmem_select(K) != nullptr;
}

// runs documents order[offset .. offset + count), all of at most n_max <= tmem_estep_max_len() pairs
int launch_estep_tmem(const EStepArgs& args, const DeviceDocs& docs, const int32_t* order, int64_t offset,
                      int64_t count, int n_max, cudaStream_t s) {
	if(count == 0)
		return 0;
	TmemKernel* k = n_max <= tmem_estep_max_len() ? tmem_select(args.K) : nullptr;
	if(!k)
		return -1;
	const int C = k->C, G = 8 / k->WG;
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
			int dev = 0;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);     // one CTA per SM: it owns the SM's TMEM
		}
		if(const char* e = getenv("TRLDA_TMEM_CLUSTERS"))
			if(atoi(e) > 0)
				n = std::min(n, atoi(e));
		k->max_clusters = n;
		if(getenv("TRLDA_RESIDENT_VERBOSE"))
			fprintf(stderr, "[trlda] k_estep_tmem<C=%d, WG=%d>: %zu B shared memory, %d co-resident clusters\n", C, k->WG, k->smem, n);
	}
	if(!args.work || cudaMemsetAsync(args.work, 0, sizeof(int), s) != cudaSuccess)     // the launch's work counter
		return -1;
	const int64_t teams_needed = (count + G - 1) / G;
	const int clusters = (int) std::min<int64_t>(k->max_clusters, teams_needed);
	cfg.gridDim = dim3((unsigned) (clusters * C));
	void* params[] = {(void*) &args, (void*) &docs, (void*) &order, (void*) &offset, (void*) &count};
	return cudaLaunchKernelExC(&cfg, k->fn, params) == cudaSuccess ? 0 : -1;
}


}  // namespace trlda
