// l2_probe.cu — how much of a cyclic per-CTA working set does the B200 L2 keep, and how fast does it serve it?
//
// Every CTA re-reads its own private region `sweeps` times (the access pattern of the streaming E-step: a document's
// tile of expElogbeta columns is swept once per inner iteration).  Printed: the footprint of all CTAs together and
// the bandwidth achieved.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/l2_probe scripts/l2_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int U>
__global__ void __launch_bounds__(256) k_probe(const float4* __restrict__ base, size_t region_vec, int sweeps, float* out) {
	const float4* p = base + (size_t) blockIdx.x * region_vec;
	float acc = 0.f;
	for(int s = 0; s < sweeps; ++s) {
		for(size_t i = threadIdx.x; i < region_vec; i += 256 * U) {
			float4 v[U];
			#pragma unroll
			for(int u = 0; u < U; ++u)
				if(i + u * 256 < region_vec)
					asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
						: "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p + i + u * 256));
				else
					v[u] = make_float4(0, 0, 0, 0);
			#pragma unroll
			for(int u = 0; u < U; ++u)
				acc += v[u].x + v[u].y + v[u].z + v[u].w;
		}
	}
	if(acc == 123.456f)
		out[0] = acc;
}

int main() {
	const size_t total = (size_t) 1 << 30;
	float4* buf;
	float* out;
	cudaMalloc(&buf, total);
	cudaMalloc(&out, 4);
	cudaMemset(buf, 0, total);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const int ctas_list[] = {148, 296};
	const int kb_list[] = {75, 150, 300, 450, 600, 900, 1200};
	for(int ctas : ctas_list)
		for(int kb : kb_list) {
			const size_t region_vec = (size_t) kb * 1024 / 16;
			if((size_t) ctas * region_vec * 16 > total)
				continue;
			const int sweeps = 20;
			k_probe<8><<<ctas, 256>>>(buf, region_vec, 2, out);
			cudaEventRecord(e0);
			k_probe<8><<<ctas, 256>>>(buf, region_vec, sweeps, out);
			cudaEventRecord(e1);
			cudaEventSynchronize(e1);
			float ms;
			cudaEventElapsedTime(&ms, e0, e1);
			const double bytes = (double) ctas * region_vec * 16 * sweeps;
			printf("ctas %3d  region %5d KB  footprint %6.1f MB  %7.1f GB/s  (%.3f ms)\n", ctas, kb,
				ctas * region_vec * 16 / 1e6, bytes / ms / 1e6, ms);
		}
	cudaError_t err = cudaDeviceSynchronize();
	printf("status %s\n", cudaGetErrorString(err));
	return 0;
}
