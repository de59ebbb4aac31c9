// djb_dmath.cuh -- lean double exp / log for the analytic BRDF kernels (djb::sgd evaluates 9 logarithms and up to 15
// exponentials per query: dj_brdf.h:3415-3432).
//
// CUDA's double libm carries its polynomial coefficients as immediates: two uniform moves per coefficient and use, so one
// polynomial term costs three issue slots (SASS of the round-1 kernels: more UMOV + IMAD than DFMA).  Here the coefficients live in
// the constant bank, where a double operand of a fused multiply-add costs no instruction.  Accuracy: ~1 ulp of double (checked
// against libm on the host, tests/cpp/dmath_check.cpp) -- the callers round to float, so even 100 ulp would change only 2e-7 of
// their results; arguments outside the plain range (zero, negative, subnormal, infinite, NaN, |x| >= 700 for exp) take libm.
//   exp: x = k ln2 + r, |r| <= ln2 / 2, exp(r) by its Taylor series to degree 13 (remainder < 4e-18), scaled by 2^k
//   log: x = 2^e m, m in [sqrt(1/2), sqrt(2)); with f = m - 1, s = f / (2 + f): log m = 2 s + s^3 (2/3 + 2/5 s^2 + ...), the classic
//        seven-coefficient form (W. Kahan / fdlibm e_log.c, whose published coefficients these are), one division by Newton steps
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define DJB_DM __device__ __forceinline__
#define DJB_DM_CONST static __constant__
#else
#define DJB_DM static inline
#define DJB_DM_CONST static const
#endif

namespace djb200 {

enum {
	DMC_LOG2E, DMC_LN2_HI, DMC_LN2_LO, DMC_SHIFT,
	DMC_E13, DMC_E12, DMC_E11, DMC_E10, DMC_E9, DMC_E8, DMC_E7, DMC_E6, DMC_E5, DMC_E4, DMC_E3, DMC_E2,
	DMC_LG1, DMC_LG2, DMC_LG3, DMC_LG4, DMC_LG5, DMC_LG6, DMC_LG7, DMC_SQRT2,
	DMC_COUNT
};
DJB_DM_CONST double g_dm_const[DMC_COUNT] = {
	1.4426950408889634074, 6.93147180369123816490e-01, 1.90821492927058770002e-10, 6755399441055744.0,
	1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
	1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5,
	6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
	1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01, 1.4142135623730951455,
};
#define DMC(k) g_dm_const[k]

DJB_DM double dm_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
	return __fma_rn(a, b, c);
#else
	return std::fma(a, b, c);
#endif
}
DJB_DM int64_t dm_bits(double x)
{
#if defined(__CUDA_ARCH__)
	return __double_as_longlong(x);
#else
	int64_t b;
	memcpy(&b, &x, 8);
	return b;
#endif
}
DJB_DM double dm_from_bits(int64_t b)
{
#if defined(__CUDA_ARCH__)
	return __longlong_as_double(b);
#else
	double x;
	memcpy(&x, &b, 8);
	return x;
#endif
}

DJB_DM double exp_d(double x)
{
	if (!(fabs(x) < 700.0)) return exp(x); // overflow / underflow / NaN: the library's answer
	const double kd = dm_fma(x, DMC(DMC_LOG2E), DMC(DMC_SHIFT)) - DMC(DMC_SHIFT); // rint(x log2 e)
	double r = dm_fma(-kd, DMC(DMC_LN2_HI), x);
	r = dm_fma(-kd, DMC(DMC_LN2_LO), r);
	double p = DMC(DMC_E13);
	p = dm_fma(p, r, DMC(DMC_E12));
	p = dm_fma(p, r, DMC(DMC_E11));
	p = dm_fma(p, r, DMC(DMC_E10));
	p = dm_fma(p, r, DMC(DMC_E9));
	p = dm_fma(p, r, DMC(DMC_E8));
	p = dm_fma(p, r, DMC(DMC_E7));
	p = dm_fma(p, r, DMC(DMC_E6));
	p = dm_fma(p, r, DMC(DMC_E5));
	p = dm_fma(p, r, DMC(DMC_E4));
	p = dm_fma(p, r, DMC(DMC_E3));
	p = dm_fma(p, r, DMC(DMC_E2));
	const double er = dm_fma(dm_fma(p, r, 1.0), r, 1.0); // 1 + r (1 + r p)
	return er * dm_from_bits(((int64_t)kd + 1023) << 52); // |k| <= 1010: 2^k is a normal number
}

// x positive, normal, finite
DJB_DM bool log_d_ok(double x) { return x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308; }
DJB_DM double log_d(double x)
{
	if (!log_d_ok(x)) return log(x);
	int64_t b = dm_bits(x);
	int e = (int)(b >> 52) - 1023;
	double m = dm_from_bits((b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL); // [1, 2)
	if (m > DMC(DMC_SQRT2)) { m *= 0.5; e += 1; }
	const double f = m - 1.0, d = 2.0 + f; // d in [1.7, 3.42)
	// s = f / d: reciprocal by three Newton steps from a float seed, then one residual correction
	double y = (double)(1.0f / (float)d);
	y = dm_fma(dm_fma(-d, y, 1.0), y, y);
	y = dm_fma(dm_fma(-d, y, 1.0), y, y);
	double s = f * y;
	s = dm_fma(dm_fma(-d, s, f), y, s);
	const double z = s * s, w = z * z;
	const double t1 = w * dm_fma(w, dm_fma(w, DMC(DMC_LG6), DMC(DMC_LG4)), DMC(DMC_LG2));
	const double t2 = z * dm_fma(w, dm_fma(w, dm_fma(w, DMC(DMC_LG7), DMC(DMC_LG5)), DMC(DMC_LG3)), DMC(DMC_LG1));
	const double R = t2 + t1, hfsq = 0.5 * f * f, dk = (double)e;
	return dm_fma(dk, DMC(DMC_LN2_HI), -((hfsq - dm_fma(s, hfsq + R, dk * DMC(DMC_LN2_LO))) - f));
}

} // namespace djb200
