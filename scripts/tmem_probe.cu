// TMEM as per-thread scratch: tcgen05.st / tcgen05.ld round trip and the read bandwidth seen by plain CUDA threads.
// Each warp owns TMEM lanes 32 (w % 4) .. +31; warps w and w + 4 (and w + 8, w + 12) use disjoint column ranges.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
		  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
		  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
		  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
		: "r"(taddr) : "memory");
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

// NW warps; every warp owns 512 * 4 / NW columns of its lane quarter (NW = 4: all 512; 8: 256; 16: 128)
template <int NW, int MODE>
__global__ void __launch_bounds__(NW * 32, 1) k_tmem(int iters, unsigned long long* cycles, uint32_t* check, float* sink) {
	__shared__ uint32_t tmem_base;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if(warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t) __cvta_generic_to_shared(&tmem_base)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	constexpr int COLS = 512 * 4 / NW;                        // columns owned by this warp
	const uint32_t mine = tmem_base + ((uint32_t) (32 * (warp & 3)) << 16) + (uint32_t) ((warp >> 2) * COLS);
	// fill: value = f(thread, column)
	for(int c = 0; c < COLS; c += 32) {
		uint32_t v[32];
		#pragma unroll
		for(int i = 0; i < 32; ++i)
			v[i] = (threadIdx.x << 16) | (uint32_t) (c + i);
		tmem_st32(mine + c, v);
	}
	asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	// verify
	uint32_t bad = 0;
	for(int c = 0; c < COLS; c += 32) {
		uint32_t v[32];
		tmem_ld32(mine + c, v);
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
		#pragma unroll
		for(int i = 0; i < 32; ++i)
			bad += v[i] != ((threadIdx.x << 16) | (uint32_t) (c + i));
	}
	if(bad)
		atomicAdd(check, bad);
	__syncthreads();
	// bandwidth: read the whole range `iters` times, two loads in flight, one FFMA per value (like an inner product)
	float acc[4] = {0.f, 0.f, 0.f, 0.f};
	const long long t0 = clock64();
	for(int it = 0; it < iters; ++it) {
		#pragma unroll 1
		for(int c = 0; c < COLS; c += 64) {
			uint32_t v0[32], v1[32];
			tmem_ld32(mine + c, v0);
			tmem_ld32(mine + c + 32, v1);
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			if(MODE == 0) {
				#pragma unroll
				for(int i = 0; i < 32; ++i) {
					acc[i & 3] = fmaf(__uint_as_float(v0[i]), 1.0001f, acc[i & 3]);
					acc[(i + 2) & 3] = fmaf(__uint_as_float(v1[i]), 0.9999f, acc[(i + 2) & 3]);
				}
			} else if(MODE == 1) {
				#pragma unroll
				for(int i = 0; i < 32; i += 2) {
					unsigned long long a0 = ((unsigned long long) v0[i + 1] << 32) | v0[i], a1 = ((unsigned long long) v1[i + 1] << 32) | v1[i];
					unsigned long long c0 = ((unsigned long long) __float_as_uint(acc[1]) << 32) | __float_as_uint(acc[0]);
					unsigned long long c1 = ((unsigned long long) __float_as_uint(acc[3]) << 32) | __float_as_uint(acc[2]);
					const unsigned long long w = 0x3f8003473f800347ull;
					asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c0) : "l"(a0), "l"(w));
					asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c1) : "l"(a1), "l"(w));
					acc[0] = __uint_as_float((unsigned) c0); acc[1] = __uint_as_float((unsigned) (c0 >> 32));
					acc[2] = __uint_as_float((unsigned) c1); acc[3] = __uint_as_float((unsigned) (c1 >> 32));
				}
			} else {
				unsigned x = 0;
				#pragma unroll
				for(int i = 0; i < 32; i += 4)
					x ^= (v0[i] ^ v0[i + 1]) ^ (v0[i + 2] ^ v0[i + 3]) ^ (v1[i] ^ v1[i + 1]) ^ (v1[i + 2] ^ v1[i + 3]);
				acc[0] += __uint_as_float(x & 0x3fffffff);
			}
		}
	}
	const long long t1 = clock64();
	sink[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
	if(threadIdx.x == 0 && blockIdx.x == 0)
		cycles[0] = (unsigned long long) (t1 - t0);
	__syncthreads();
	if(warp == 0)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <int NW, int MODE>
void run(unsigned long long* cycles, uint32_t* check, float* sink) {
	const int iters = 200;
	cudaMemset(check, 0, 4);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_tmem<NW, MODE><<<148, NW * 32>>>(2, cycles, check, sink);
	cudaEventRecord(e0);
	k_tmem<NW, MODE><<<148, NW * 32>>>(iters, cycles, check, sink);
	cudaEventRecord(e1);
	cudaError_t err = cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	unsigned long long cyc; uint32_t bad;
	cudaMemcpy(&cyc, cycles, 8, cudaMemcpyDeviceToHost);
	cudaMemcpy(&bad, check, 4, cudaMemcpyDeviceToHost);
	const double bytes = 256.0 * 1024 * iters;      // per SM
	printf("mode %d (0: FFMA per value, 1: FFMA2 per pair, 2: XOR only) warps=%2d: %s, mismatches %u, %.1f B/cycle/SM (clock64), whole kernel %.3f ms -> %.1f TB/s chip\n", MODE, NW,
	       cudaGetErrorString(err), bad, bytes / (double) cyc, ms, bytes * 148 / (ms * 1e-3) / 1e12);
}

int main() {
	unsigned long long* cycles; uint32_t* check; float* sink;
	cudaMalloc(&cycles, 8); cudaMalloc(&check, 4); cudaMalloc(&sink, 148 * 512 * 4);
	run<4, 0>(cycles, check, sink);
	run<8, 0>(cycles, check, sink);
	run<16, 0>(cycles, check, sink);
	run<4, 1>(cycles, check, sink);
	run<8, 1>(cycles, check, sink);
	run<16, 1>(cycles, check, sink);
	run<4, 2>(cycles, check, sink);
	run<8, 2>(cycles, check, sink);
	run<16, 2>(cycles, check, sink);
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
