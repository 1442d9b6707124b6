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

// the streaming E-step's transport without its arithmetic: every warp pulls 4000-byte columns of its CTA's region through
// a private ring of DEPTH slots with cp.async.bulk, waits on the slot's mbarrier, reads the column back with LDS.128
template <int DEPTH>
__global__ void __launch_bounds__(512) k_bulk(const float* __restrict__ base, int cols_per_cta, int sweeps, float* out) {
	extern __shared__ __align__(128) unsigned char smem[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
	float* ring = reinterpret_cast<float*>(smem) + (size_t) warp * DEPTH * 1024;
	unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t) nw * DEPTH * 4096) + warp * DEPTH;
	const unsigned ring_addr = (unsigned) __cvta_generic_to_shared(ring), bar_addr = (unsigned) __cvta_generic_to_shared(bars);
	if(lane == 0)
		for(int s = 0; s < DEPTH; ++s)
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr + 8 * s));
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	__syncthreads();
	const float* region = base + (size_t) blockIdx.x * cols_per_cta * 1000;
	const int M = (cols_per_cta - warp + nw - 1) / nw;
	auto issue = [&](int m, int stage) {
		if(lane == 0) {
			const float* src = region + (size_t) (warp + m * nw) * 1000;
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr + 8 * stage), "r"(4000) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				::"r"(ring_addr + stage * 4096), "l"(src), "r"(4000), "r"(bar_addr + 8 * stage) : "memory");
		}
	};
	unsigned phases = 0;
	float acc = 0.f;
	const int total = M * sweeps;
	for(int s = 0; s < DEPTH && s < total; ++s)
		issue(s % M, s);
	int stage = 0;
	for(int t = 0; t < total; ++t) {
		unsigned done = 0;
		while(!done)
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
				: "=r"(done) : "r"(bar_addr + 8 * stage), "r"((phases >> stage) & 1u) : "memory");
		phases ^= 1u << stage;
		const float4* col = reinterpret_cast<const float4*>(ring + stage * 1024) + lane;
		#pragma unroll
		for(int v = 0; v < 7; ++v) {
			const float4 x = col[v * 32];
			acc += x.x + x.y + x.z + x.w;
		}
		__syncwarp();
		if(t + DEPTH < total)
			issue((t + DEPTH) % M, stage);
		stage = stage + 1 == DEPTH ? 0 : stage + 1;
	}
	if(acc == 123.456f)
		out[0] = acc;
}

template <int DEPTH>
static void run_bulk(const float* buf, float* out, int warps, int ctas) {
	const int cols = 150, sweeps = 20;
	const size_t smem = (size_t) warps * DEPTH * 4096 + warps * DEPTH * 8;
	cudaFuncSetAttribute(k_bulk<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	k_bulk<DEPTH><<<ctas, warps * 32, smem>>>(buf, cols, 2, out);
	cudaEventRecord(e0);
	k_bulk<DEPTH><<<ctas, warps * 32, smem>>>(buf, cols, sweeps, out);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double bytes = (double) ctas * cols * 4000.0 * sweeps;
	printf("bulk ring: ctas %3d  warps %2d  depth %d  footprint %6.1f MB  %7.1f GB/s  (%.3f ms)  %s\n", ctas, warps, DEPTH,
		ctas * cols * 4000.0 / 1e6, bytes / ms / 1e6, ms, cudaGetErrorString(cudaGetLastError()));
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
	for(int ctas : {148, 74}) {
		run_bulk<2>(reinterpret_cast<const float*>(buf), out, 16, ctas);
		run_bulk<3>(reinterpret_cast<const float*>(buf), out, 16, ctas);
		run_bulk<2>(reinterpret_cast<const float*>(buf), out, 8, ctas);
		run_bulk<4>(reinterpret_cast<const float*>(buf), out, 8, ctas);
		run_bulk<6>(reinterpret_cast<const float*>(buf), out, 8, ctas);
	}
	cudaError_t err = cudaDeviceSynchronize();
	printf("status %s\n", cudaGetErrorString(err));
	return 0;
}
