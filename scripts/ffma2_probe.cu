// FFMA vs FFMA2 (fma.rn.f32x2) issue rate on one SM-filling grid: 8 warps per SM, long independent chains.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
	float a[16]; u64 p[16];
	for(int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 1e-3f + i; p[i] = ((u64) __float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
	u64 s2 = ((u64) __float_as_uint(s) << 32) | __float_as_uint(s);
	for(int it = 0; it < iters; ++it) {
		#pragma unroll
		for(int i = 0; i < 16; ++i) {
			if(MODE == 0) a[i] = fmaf(a[i], s, 1e-3f);
			else p[i] = ffma2(p[i], s2, s2);
		}
	}
	float r = 0; for(int i = 0; i < 16; ++i) r += MODE == 0 ? a[i] : __uint_as_float((unsigned) p[i]) + __uint_as_float((unsigned) (p[i] >> 32));
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
	float* out; cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for(int warps = 4; warps <= 32; warps *= 2)
		for(int mode = 0; mode < 2; ++mode) {
			const int iters = 20000;
			for(int rep = 0; rep < 2; ++rep) {
				cudaEventRecord(e0);
				if(mode == 0) k<0><<<148, warps * 32>>>(out, iters, 0.999f); else k<1><<<148, warps * 32>>>(out, iters, 0.999f);
				cudaEventRecord(e1); cudaEventSynchronize(e1);
			}
			float ms; cudaEventElapsedTime(&ms, e0, e1);
			const double inst = 148.0 * warps * iters * 16;
			printf("%s warps/SM=%2d: %.3f ms, %.1f warp-instr/cycle/SM at 1.9 GHz, %.1f Tflop/s\n", mode ? "FFMA2" : "FFMA ", warps, ms,
			       inst / 148 / (ms * 1e-3 * 1.9e9), inst * 32 * 2 * (mode ? 2 : 1) / (ms * 1e-3) / 1e12);
		}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
