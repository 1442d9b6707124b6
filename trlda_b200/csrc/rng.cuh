// rng.cuh — counter-based random numbers for the device samplers (initial gamma / lambda, Gibbs E-step, LDA::sample).
// Philox4x32-10 keyed by the seed; the counter names WHAT is drawn (element, stream, attempt), so a value never depends
// on the launch shape or on the order in which threads run.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace trlda {

__device__ __forceinline__ void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
	#pragma unroll
	for(int r = 0; r < 10; ++r) {
		const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
		const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
		const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
		c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
}

// uniform in [0, 1) with 53 bits, named by (a, b, stream, draw)
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t a, uint32_t stream, uint32_t draw) {
	uint32_t c[4] = {(uint32_t) a, (uint32_t) (a >> 32), stream, draw};
	philox4x32(c, (uint32_t) seed, (uint32_t) (seed >> 32));
	return ((double) c[0] * 4294967296.0 + (double) c[1]) * (1.0 / 18446744073709551616.0);
}

// Gamma(shape, 1), shape > 0, named by (a, stream): Marsaglia-Tsang rejection with Box-Muller normals for
// shape >= 1, boosted by U^(1/shape) below 1
__device__ inline double philox_gamma(double shape, uint64_t seed, uint64_t a, uint32_t stream) {
	const double s = shape < 1.0 ? shape + 1.0 : shape;
	const double dd = s - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
	double result = s;
	for(uint32_t attempt = 0; attempt < 64; ++attempt) {
		uint32_t c[4] = {(uint32_t) a, (uint32_t) (a >> 32), stream, attempt};
		philox4x32(c, (uint32_t) seed ^ 0x2545F491u, (uint32_t) (seed >> 32) ^ 0x4F6CDD1Du);
		const double u1 = ((double) c[0] + 0.5) * (1.0 / 4294967296.0);
		const double u2 = ((double) c[1] + 0.5) * (1.0 / 4294967296.0);
		const double u3 = ((double) c[2] + 0.5) * (1.0 / 4294967296.0);
		const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
		const double t = 1.0 + cc * x;
		if(t <= 0.0)
			continue;
		const double v = t * t * t;
		if(log(u3) < 0.5 * x * x + dd - dd * v + dd * log(v)) {
			result = dd * v;
			if(shape < 1.0) {
				const double u4 = ((double) c[3] + 0.5) * (1.0 / 4294967296.0);
				result *= exp(log(u4) / shape);
			}
			break;
		}
	}
	return result;
}

}  // namespace trlda
