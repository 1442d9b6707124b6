// special.cuh — in-register special functions for the sm_100a kernels.
//
// The reference evaluates psi with the scalar Cephes routine (code/trlda/src/digamma.cpp:116-178): upward
// recurrence w = sum 1/(x+i) until s >= 10, then log(s) - 1/(2s) - sum_k B_2k/(2k s^2k) with the seven
// coefficients of digamma.cpp:42-52, and psi' as the Hurwitz zeta(2, x) (zeta.cpp:67-134 via utils.cpp:107).
// The device versions keep the same mathematical decomposition (same shift point, same series) so that the
// results agree to a few ulp, but are restructured for a GPU: the recurrence is accumulated as a rational
// num/den (two FMAs per step, ONE division in total) instead of one division per step.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace trlda {

// B_2k / (2k), k = 7..1 in Horner order (same constants as digamma.cpp:42-52)
__host__ __device__ __forceinline__ double psi_series(double z) {
	double p = 8.33333333333333333333E-2;
	p = fma(p, z, -2.10927960927960927961E-2);
	p = fma(p, z, 7.57575757575757575758E-3);
	p = fma(p, z, -4.16666666666666666667E-3);
	p = fma(p, z, 3.96825396825396825397E-3);
	p = fma(p, z, -8.33333333333333333333E-3);
	p = fma(p, z, 8.33333333333333333333E-2);
	return z * p;
}

// psi(x) for x > 0.
__host__ __device__ __forceinline__ double digamma_pos(double x) {
	double num = 0.0, den = 1.0, s = x;
	// sum_{i<n} 1/(x+i) = num/den; at most 10 steps (all terms positive: no cancellation)
	#pragma unroll 1
	while(s < 10.0) {
		num = fma(num, s, den);
		den *= s;
		s += 1.0;
	}
	const double r = 1.0 / s;
	const double y = psi_series(r * r);
	return log(s) - 0.5 * r - y - num / den;
}

// psi(x) for any x, following the reflection branch of digamma.cpp:123-144 for x <= 0 (never taken on the
// hot path, where all arguments are positive).
__host__ __device__ inline double digamma_reflect(double x) {
	const double kPi = 3.141592653589793238462643383279502884;
	double p = floor(x);
	if(p == x)
		return HUGE_VAL;
	double nz = x - p;
	if(nz != 0.5) {
		if(nz > 0.5) {
			p += 1.0;
			nz = x - p;
		}
		nz = kPi / tan(kPi * nz);
	} else {
		nz = 0.0;
	}
	return digamma_pos(1.0 - x) - nz;
}

__host__ __device__ __forceinline__ double digamma(double x) {
	if(x <= 0.0)
		return digamma_reflect(x);
	return digamma_pos(x);
}

// exp(psi(x) - c) for x > 0 WITHOUT a logarithm: with s = x + n >= 10 after the recurrence,
//   psi(x) = log(s) - 1/(2s) - series(1/s^2) - sum_{i<n} 1/(x+i)   =>   exp(psi(x) - c) = s * exp(-(t + y + c)),
// where t = 1/(2s) + num/den is formed with ONE reciprocal q = 1/(s den): 1/s = den q, t = (num s + den/2) q.
// This is the quantity the E-step needs (c = 0: lda.cpp:174,197) and the beta-prep needs (c = psi(row sum):
// lda.cpp:173); it costs one exp and one division instead of log + exp + two divisions.
__device__ __forceinline__ double exp_digamma_shifted(double x, double c) {
	if(x <= 0.0)
		return exp(digamma_reflect(x) - c);
	double num = 0.0, den = 1.0, s = x;
	#pragma unroll 1
	while(s < 10.0) {
		num = fma(num, s, den);
		den *= s;
		s += 1.0;
	}
	const double q = 1.0 / (s * den);
	const double r = den * q;
	const double t = fma(num, s, 0.5 * den) * q;
	return s * exp(-(t + psi_series(r * r) + c));
}

__device__ __forceinline__ double exp_digamma(double x) {
	return exp_digamma_shifted(x, 0.0);
}

// Same quantity for the MIXED-precision mode, where the result is rounded to float32 anyway: the recurrence shifts
// to s >= 6 only (the 7-term series is then still good to ~2e-13 absolute in psi) and the one reciprocal is a float32
// MUFU seed refined by two fp64 Newton steps (relative error ~1e-16) instead of a full fp64 division.
__device__ __forceinline__ double exp_digamma_shifted_mixed(double x, double c) {
	if(x <= 0.0)
		return exp(digamma_reflect(x) - c);
	// recurrence to s = x + 6 in closed form for x < 6 (see exp_digamma_scaled_f32), none for x >= 6
	double num = 0.0, den = 1.0, s = x;
	if(x < 6.0) {
		den = fma(fma(fma(fma(fma(x + 15.0, x, 85.0), x, 225.0), x, 274.0), x, 120.0), x, 0.0);
		num = fma(fma(fma(fma(fma(6.0, x, 75.0), x, 340.0), x, 675.0), x, 548.0), x, 120.0);
		s = x + 6.0;
	}
	const double y = s * den;
	double q;
	if(y > 1e-30 && y < 1e30) {
		q = (double) __frcp_rn((float) y);
		q = fma(q, fma(-y, q, 1.0), q);
		q = fma(q, fma(-y, q, 1.0), q);
	} else {
		q = 1.0 / y;
	}
	const double r = den * q;
	const double t = fma(num, s, 0.5 * den) * q;
	return s * exp(-(t + psi_series(r * r) + c));
}

// expElogbeta element for the MIXED mode, evaluated in float32: beta = exp(psi(lambda)) * ek with ek = exp(-psi(row sum))
// prepared once per topic.  Same decomposition (recurrence to s >= 6, one reciprocal, three series terms — the
// fourth is 2.5e-9 at s = 6); relative error ~5e-7, consistent with the float32 storage and the float32 inner
// products of this mode.  Arguments outside the float32 range of the recurrence take the fp64 path.
__device__ __forceinline__ float exp_digamma_scaled_f32(double lambda, float ek) {
	const float x = (float) lambda;
	if(!(x > 1e-30f) || x > 1e30f)
		return (float) (exp_digamma_shifted_mixed(lambda, 0.0) * (double) ek);
	// Branch-free recurrence: for x < 64 always shift by six, sum_{i<6} 1/(x+i) = D'(x)/D(x) with
	// D(x) = x (x+1) ... (x+5) = x^6 + 15 x^5 + 85 x^4 + 225 x^3 + 274 x^2 + 120 x (all coefficients positive:
	// Horner without cancellation); for x >= 64 no shift is needed (and D would overflow float32 much later).
	const bool big = x >= 64.0f;
	float den = fmaf(fmaf(fmaf(fmaf(fmaf(x + 15.0f, x, 85.0f), x, 225.0f), x, 274.0f), x, 120.0f), x, 0.0f);
	float num = fmaf(fmaf(fmaf(fmaf(fmaf(6.0f, x, 75.0f), x, 340.0f), x, 675.0f), x, 548.0f), x, 120.0f);
	const float s = big ? x : x + 6.0f;
	den = big ? 1.0f : den;
	num = big ? 0.0f : num;
	const float q = __frcp_rn(s * den);
	const float r = den * q;
	const float t = fmaf(num, s, 0.5f * den) * q;
	const float z = r * r;
	const float y = z * (8.33333333333e-2f - z * (8.33333333333e-3f - z * 3.96825396825e-3f));
	return s * expf(-(t + y)) * ek;
}

// exp(psi(x)) for the inner loop of the MIXED-mode E-step kernels that keep the tile on chip (estep_tmem.cu,
// estep_resident.cu), where the update of gamma is a third of all issued instructions: branch-free, ~75 instructions,
// relative error <= 2e-11 for x in [1e-6, 1e5] (checked against a long-double evaluation; the float32 tile and inner
// products of the mode carry 1e-7).  Same decomposition as above — shift by six below 64 (none above), one reciprocal
// (MUFU seed + one fp64 Newton step), five series terms (the sixth is 1e-11 at s = 6), and an inline exp: range
// reduction by the 1.5 * 2^52 trick, degree-9 Taylor polynomial on |r| <= ln2 / 2, exponent patched into the high
// word.  The constants sit in constant memory so that they reach the DFMAs as c[bank][offset] operands.
static __constant__ double kLeanPsi[32] = {
	120.0, 274.0, 225.0, 85.0, 15.0,                                    // 0..4   D(x) / x = 120 + 274 x + 225 x^2 + 85 x^3 + 15 x^4 + x^5
	548.0, 675.0, 340.0, 75.0, 6.0,                                     // 5..9   D'(x) = 120 + 548 x + ... + 6 x^5
	1.0 / 12, -1.0 / 120, 1.0 / 252, -1.0 / 240, 1.0 / 132,             // 10..14 B_2k / (2k), k = 1..5
	1.4426950408889634074, 6755399441055744.0,                          // 15, 16 log2(e), 1.5 * 2^52
	-6.93147180369123816490e-01, -1.90821492927058770002e-10,           // 17, 18 -ln2 (high, low)
	1.0 / 6, 0.5, 1.0 / 120, 1.0 / 24, 1.0 / 5040, 1.0 / 720, 1.0 / 362880, 1.0 / 40320,   // 19..26 1/k! in Estrin pairs
	64.0, 0.0, 0.0, 0.0, 0.0};                                          // 27
// Polynomials in Estrin form (pairs, then powers of x^2): the dependent chain is ~28 fp64 operations instead of ~45
// with Horner's rule — the evaluation is latency-bound, the kernels run two warps per scheduler.
__device__ __forceinline__ double exp_digamma_lean(double x) {
	const double* kT = kLeanPsi;
	const bool big = x >= kT[27];
	const double x2 = x * x, x4 = x2 * x2;
	double den = x * fma(x4, x + kT[4], fma(x2, fma(kT[3], x, kT[2]), fma(kT[1], x, kT[0])));
	double num = fma(x4, fma(kT[9], x, kT[8]), fma(x2, fma(kT[7], x, kT[6]), fma(kT[5], x, kT[0])));
	const double s = big ? x : x + kT[9];
	den = big ? 1.0 : den;
	num = big ? 0.0 : num;
	const double y = s * den;
	float qf;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(qf) : "f"(fmaxf((float) y, 1e-37f)));
	double q = (double) qf;
	q = fma(q, fma(-y, q, 1.0), q);
	const double r = den * q;                                // 1 / s
	const double t = fma(num, s, 0.5 * den) * q;             // sum_{i<6} 1/(x+i) + 1/(2s)
	const double z = r * r, z2 = z * z;
	const double ser = fma(z2, fma(z2, kT[14], fma(z, kT[13], kT[12])), fma(z, kT[11], kT[10]));
	const double u = fma(ser, z, t);                         // exp(psi(x)) = s exp(-u), u > 0
	double nd = fma(u, kT[15], kT[16]);
	const int n = __double2loint(nd);                        // rint(u log2 e)
	nd -= kT[16];
	double rr = fma(nd, kT[17], u);
	rr = fma(nd, kT[18], rr);                                // u - n ln2
	const double m = -rr, m2 = m * m, m4 = m2 * m2, m8 = m4 * m4;
	const double p01 = 1.0 + m, p23 = fma(m, kT[19], kT[20]), p45 = fma(m, kT[21], kT[22]), p67 = fma(m, kT[23], kT[24]),
	             p89 = fma(m, kT[25], kT[26]);
	const double p = fma(m8, p89, fma(m4, fma(m2, p67, p45), fma(m2, p23, p01)));     // exp(-(u - n ln2))
	const double res = s * p;
	const int hi = __double2hiint(res) - (n << 20);          // * 2^-n
	return n > 1000 ? 0.0 : __hiloint2double(hi, __double2loint(res));
}

// picks the evaluation matching the element type of the expElogbeta working copy
template <typename T> __device__ __forceinline__ double exp_digamma_for(double x, double c);
template <> __device__ __forceinline__ double exp_digamma_for<double>(double x, double c) { return exp_digamma_shifted(x, c); }
template <> __device__ __forceinline__ double exp_digamma_for<float>(double x, double c) { return exp_digamma_shifted_mixed(x, c); }


// psi'(x) = zeta(2, x) for x > 0: shift to s >= 10 by the recurrence psi'(x) = psi'(x+1) + 1/x^2, then the
// asymptotic series 1/s + 1/(2 s^2) + sum B_2k / s^(2k+1).  Only K+1 values per alpha update are needed
// (onlinelda.cpp:132-133), so clarity beats speed here.
__host__ __device__ __forceinline__ double trigamma(double x) {
	double w = 0.0, s = x;
	while(s < 10.0) {
		w += 1.0 / (s * s);
		s += 1.0;
	}
	const double r = 1.0 / s, z = r * r;
	// B2 = 1/6, B4 = -1/30, B6 = 1/42, B8 = -1/30, B10 = 5/66, B12 = -691/2730, B14 = 7/6
	double p = 7.0 / 6.0;
	p = fma(p, z, -691.0 / 2730.0);
	p = fma(p, z, 5.0 / 66.0);
	p = fma(p, z, -1.0 / 30.0);
	p = fma(p, z, 1.0 / 42.0);
	p = fma(p, z, -1.0 / 30.0);
	p = fma(p, z, 1.0 / 6.0);
	return w + r + 0.5 * z + p * z * r;
}

// ---- warp / block reductions ---------------------------------------------------------------------------------

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
	#pragma unroll
	for(int o = 16; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

}  // namespace trlda
