// djb_glibcf.h -- the single-precision libm calls of the reference's sampling path, evaluated the way the platform's
// libm evaluates them.
//
// The reference calls logf (dj_brdf.h:695, inside djb::erfinv), std::pow(float, float) (:1917) and std::exp(float) (:1935):
// with libstdc++ these are glibc's logf / powf / expf.  glibc is a third-party dependency that is not under /root/reference;
// the pinned version is the image's GLIBC 2.39 (Ubuntu 2.39-0ubuntu8.5), whose float functions are the published
// table-driven algorithms of the ARM optimized-routines project (glibc sysdeps/ieee754/flt-32/e_{logf,expf,powf}.c): a 16-entry
// (1/c, log c) table, a 32-entry 2^(i/32) table and short polynomials, carried in double and rounded once to float.  They
// return the correctly rounded float for all but ~1e-3 of arguments, so "evaluate in double and round" (round 1 of this
// project) leaves 3e-4..8e-4 of sampled directions different from the reference.  This file restates the algorithms
// operation for operation -- including the fused multiply-adds of the x86-64 "fma" multiarch variant glibc selects on
// every AVX2 + FMA host (read from the disassembly of libm-2.39.a: e_logf-fma.o, e_expf-fma.o, e_powf-fma.o) -- with
// the tables libm.a carries in __logf_data, __exp2f_data and __powf_log2_data.  Same operations in the same order in IEEE
// double: the device results ARE glibc's results, and one logf costs 7 double operations instead of a 60-instruction
// float-float evaluation.
//
// The same source is compiled for the device (nvcc, -fmad=false, explicit __fma_rn) and for the host (tests/cpp/
// glibcf_check.cpp, g++ -ffp-contract=off, std::fma), where it is compared with libm itself: exhaustively over every
// float of the domains the sampling path uses (see that file), which pins the constants below.
//
// Only the main branch of each function is restated; `ok` tells the caller when an argument is outside it (zero,
// subnormal, negative, infinite, NaN, overflowing products), and the caller takes its literal path then.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GLF_FN __device__ __forceinline__
#define GLF_MEM __device__ __forceinline__
#define GLF_FMA(a, b, c) __fma_rn((a), (b), (c))
#define GLF_AS_U64(d) ((uint64_t)__double_as_longlong(d))
#define GLF_AS_F64(u) __longlong_as_double((long long)(u))
#define GLF_AS_U32(f) ((uint32_t)__float_as_int(f))
#define GLF_AS_F32(u) __int_as_float((int)(u))
#else
#include <cmath>
#include <cstring>
#define GLF_FN static inline
#define GLF_MEM inline
#define GLF_FMA(a, b, c) std::fma((a), (b), (c))
static inline uint64_t glf_as_u64(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double glf_as_f64(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint32_t glf_as_u32(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float glf_as_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define GLF_AS_U64(d) glf_as_u64(d)
#define GLF_AS_F64(u) glf_as_f64(u)
#define GLF_AS_U32(f) glf_as_u32(f)
#define GLF_AS_F32(u) glf_as_f32(u)
#endif

namespace djb200 {

// the three tables, in one block so that a kernel can stage them in shared memory with one loop:
//   [0, 32)   __logf_data.tab:       (1/c, log c)   for the 16 sub-intervals of [0x1.66p-1, 0x1.66p0)
//   [32, 64)  __powf_log2_data.tab:  (1/c, log2 c)  same sub-intervals
//   [64, 96)  __exp2f_data.tab:      bits of 2^(i/32) minus (i << 47), i = 0..31
//   [96, 99)  the three "hot" constants of GlfHot below (logf's A2, expf's Shift and C1), [99] padding
constexpr int GLF_TABLE_WORDS = 100;
constexpr int GLF_OFF_LOG = 0, GLF_OFF_LOG2 = 32, GLF_OFF_EXP2 = 64, GLF_OFF_HOT = 96;

#define GLF_TABLE_INIT                                                                                                  \
	{                                                                                                                   \
		0x3ff661ec79f8f3beull, 0xbfd57bf7808caadeull, 0x3ff571ed4aaf883dull, 0xbfd2bef0a7c06ddbull,                     \
		0x3ff49539f0f010b0ull, 0xbfd01eae7f513a67ull, 0x3ff3c995b0b80385ull, 0xbfcb31d8a68224e9ull,                     \
		0x3ff30d190c8864a5ull, 0xbfc6574f0ac07758ull, 0x3ff25e227b0b8ea0ull, 0xbfc1aa2bc79c8100ull,                     \
		0x3ff1bb4a4a1a343full, 0xbfba4e76ce8c0e5eull, 0x3ff12358f08ae5baull, 0xbfb1973c5a611cccull,                     \
		0x3ff0953f419900a7ull, 0xbfa252f438e10c1eull, 0x3ff0000000000000ull, 0x0000000000000000ull,                     \
		0x3fee608cfd9a47acull, 0x3faaa5aa5df25984ull, 0x3feca4b31f026aa0ull, 0x3fbc5e53aa362eb4ull,                     \
		0x3feb2036576afce6ull, 0x3fc526e57720db08ull, 0x3fe9c2d163a1aa2dull, 0x3fcbc2860d224770ull,                     \
		0x3fe886e6037841edull, 0x3fd1058bc8a07ee1ull, 0x3fe767dcf5534862ull, 0x3fd4043057b6ee09ull,                     \
		/* powf log2 */                                                                                                 \
		0x3ff661ec79f8f3beull, 0xbfdefec65b963019ull, 0x3ff571ed4aaf883dull, 0xbfdb0b6832d4fca4ull,                     \
		0x3ff49539f0f010b0ull, 0xbfd7418b0a1fb77bull, 0x3ff3c995b0b80385ull, 0xbfd39de91a6dcf7bull,                     \
		0x3ff30d190c8864a5ull, 0xbfd01d9bf3f2b631ull, 0x3ff25e227b0b8ea0ull, 0xbfc97c1d1b3b7af0ull,                     \
		0x3ff1bb4a4a1a343full, 0xbfc2f9e393af3c9full, 0x3ff12358f08ae5baull, 0xbfb960cbbf788d5cull,                     \
		0x3ff0953f419900a7ull, 0xbfaa6f9db6475fceull, 0x3ff0000000000000ull, 0x0000000000000000ull,                     \
		0x3fee608cfd9a47acull, 0x3fb338ca9f24f53dull, 0x3feca4b31f026aa0ull, 0x3fc476a9543891baull,                     \
		0x3feb2036576afce6ull, 0x3fce840b4ac4e4d2ull, 0x3fe9c2d163a1aa2dull, 0x3fd40645f0c6651cull,                     \
		0x3fe886e6037841edull, 0x3fd88e9c2c1b9ff8ull, 0x3fe767dcf5534862ull, 0x3fdce0a44eb17bccull,                     \
		/* exp2f */                                                                                                     \
		0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,                     \
		0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,                     \
		0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,                     \
		0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,                     \
		0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,                     \
		0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,                     \
		0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,                     \
		0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,                     \
		/* hot constants: -0x1.ffffef20a4123p-2, 0x1.8p+52, 0x1.ebfce50fac4f3p-13 */                                    \
		0xbfdffffef20a4123ull, 0x4338000000000000ull, 0x3f2ebfce50fac4f3ull, 0ull                                       \
	}

#if defined(__CUDACC__)
__device__ const uint64_t g_glf_table[GLF_TABLE_WORDS] = GLF_TABLE_INIT;
#else
static const uint64_t g_glf_table[GLF_TABLE_WORDS] = GLF_TABLE_INIT;
#endif

// the polynomial coefficients and scaling constants (__logf_data.ln2 / .poly, __exp2f_data.shift / .invln2_scaled / .poly_scaled /
// .shift_scaled / .poly, __powf_log2_data.poly).  On the device they live in the constant bank, where a double operand of an
// FMA costs no instruction (as immediates each one is two moves per use).
enum {
	GLC_LN2, GLC_LOG_A0, GLC_LOG_A1, GLC_INVLN2N, GLC_EXP_C0, GLC_EXP_C2, // single-constant FMA operands of logf / expf
	GLC_LOG_A2, GLC_SHIFT, GLC_EXP_C1,                                      // the GlfHot three
	GLC_POW_A0, GLC_POW_A1, GLC_POW_A2, GLC_POW_A3, GLC_POW_A4,
	GLC_SHIFT47, GLC_EXP2_C0, GLC_EXP2_C1, GLC_EXP2_C2,
	GLC_COUNT
};
#define GLF_CONST_INIT                                                                                                  \
	{                                                                                                                   \
		0x1.62e42fefa39efp-1, -0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, 0x1.71547652b82fep+5, 0x1.c6af84b912394p-20, \
		0x1.62e42ff0c52d6p-6, -0x1.ffffef20a4123p-2, 0x1.8p+52, 0x1.ebfce50fac4f3p-13,                                  \
		0x1.27616c9496e0bp-2, -0x1.71969a075c67ap-2, 0x1.ec70a6ca7baddp-2, -0x1.7154748bef6c8p-1, 0x1.71547652ab82bp0,  \
		0x1.8p+47, 0x1.c6af84b912394p-5, 0x1.ebfce50fac4f3p-3, 0x1.62e42ff0c52d6p-1                                     \
	}
#if defined(__CUDACC__)
__constant__ double g_glf_const[GLC_COUNT] = GLF_CONST_INIT;
#else
static const double g_glf_const[GLC_COUNT] = GLF_CONST_INIT;
#endif
#define GLC(k) g_glf_const[k]

// table access: a plain pointer (host; device global memory), or a shared-memory window address (device kernels that
// stage the block: a 32-bit shared address costs nothing to form, a generic pointer to shared memory does)
struct GlfTablePtr {
	const uint64_t *p;
	GLF_MEM void pair(int w, double &a, double &b) const { a = GLF_AS_F64(p[w]); b = GLF_AS_F64(p[w + 1]); }
	GLF_MEM uint64_t word(int w) const { return p[w]; }
};
#if defined(__CUDACC__)
struct GlfTableShared {
	uint32_t s; // __cvta_generic_to_shared of the staged block (16-byte aligned)
	GLF_MEM void pair(int w, double &a, double &b) const
	{
		asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(s + 8u * (uint32_t)w));
	}
	GLF_MEM uint64_t word(int w) const
	{
		uint64_t v;
		asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(s + 8u * (uint32_t)w));
		return v;
	}
};
#endif

// Three of the constants are the second constant operand of a fused multiply-add (A1 r + A2, InvLn2N x + Shift, C0 r + C1).
// A device FMA takes one operand from the constant bank / a uniform register for free; the other has to sit in a register.
// A kernel that calls these functions in a loop reads the three once from its staged table block (glf_hot): the assembler
// re-materialises constant-bank loads at every use, but not loads from shared memory.
struct GlfHot { double log_a2, exp_shift, exp_c1; };
template <class TB>
GLF_FN GlfHot glf_hot(TB T)
{
	GlfHot h;
	h.log_a2 = GLF_AS_F64(T.word(GLF_OFF_HOT));
	h.exp_shift = GLF_AS_F64(T.word(GLF_OFF_HOT + 1));
	h.exp_c1 = GLF_AS_F64(T.word(GLF_OFF_HOT + 2));
	return h;
}
GLF_FN GlfHot glf_hot() // from the constant bank: one-off calls
{
	GlfHot h;
	h.log_a2 = GLC(GLC_LOG_A2);
	h.exp_shift = GLC(GLC_SHIFT);
	h.exp_c1 = GLC(GLC_EXP_C1);
	return h;
}

// glibc logf, main branch: x positive, normal, finite (glibc e_logf.c; __logf_fma).  T = the table block above.
GLF_FN bool glf_logf_ok(float x) { return GLF_AS_U32(x) - 0x00800000u < 0x7f800000u - 0x00800000u; }
template <class TB>
GLF_FN float glf_logf(TB T, const GlfHot &H, float x)
{
	const uint32_t ix = GLF_AS_U32(x);
	// x = 2^k z with z in [0x1.66p-1, 0x1.66p0), c the centre of z's sub-interval: log x = k ln2 + log c + log1p(z / c - 1)
	const uint32_t tmp = ix - 0x3f330000u;
	const int i = (int)((tmp >> 19) & 15u);
	const int k = (int32_t)tmp >> 23;
	const uint32_t iz = ix - (tmp & 0xff800000u);
	double invc, logc;
	T.pair(GLF_OFF_LOG + 2 * i, invc, logc);
	const double z = (double)GLF_AS_F32(iz);
	const double r = GLF_FMA(z, invc, -1.0);
	const double y0 = GLF_FMA((double)k, GLC(GLC_LN2), logc);
	const double r2 = r * r;
	double y = GLF_FMA(GLC(GLC_LOG_A1), r, H.log_a2);
	y = GLF_FMA(GLC(GLC_LOG_A0), r2, y);
	y = GLF_FMA(y, r2, y0 + r);
	// glibc returns +0 for x == 1 before the evaluation; the evaluation gives the same (k = 0, i = 9: 1/c = 1, log c = 0, r = 0)
	return (float)y;
}

// glibc expf, main branch: |x| < 88 (glibc e_expf.c; __expf_fma)
GLF_FN bool glf_expf_ok(float x) { return ((GLF_AS_U32(x) >> 20) & 0x7ffu) <= 0x42au; }
template <class TB>
GLF_FN float glf_expf(TB T, const GlfHot &H, float x)
{
	const double xd = (double)x;
	// x = (k + r) ln2 / 32 with |r| <= 1/2: exp x = 2^(k / 32) 2^(r / 32); the shift constant leaves k in the low bits
	const double zs = GLF_FMA(GLC(GLC_INVLN2N), xd, H.exp_shift);
	const uint64_t ki = GLF_AS_U64(zs);
	const double kd = zs - H.exp_shift;
	const double r = GLF_FMA(GLC(GLC_INVLN2N), xd, -kd);
	const uint64_t t = T.word(GLF_OFF_EXP2 + (int)(ki & 31u)) + (ki << 47);
	const double s = GLF_AS_F64(t);
	const double z = GLF_FMA(GLC(GLC_EXP_C0), r, H.exp_c1);
	const double r2 = r * r;
	double y = GLF_FMA(GLC(GLC_EXP_C2), r, 1.0);
	y = GLF_FMA(z, r2, y);
	y = y * s;
	return (float)y;
}

// glibc powf, main branch: x positive, normal, finite; y finite, non-zero; |y log2 x| < 126 (glibc e_powf.c; __powf_fma)
GLF_FN bool glf_powf_ok(float x, float y)
{
	const uint32_t iy = GLF_AS_U32(y);
	return GLF_AS_U32(x) - 0x00800000u < 0x7f800000u - 0x00800000u && 2u * iy - 1u < 2u * 0x7f800000u - 1u;
}
template <class TB>
GLF_FN float glf_powf(TB T, float x, float y, bool &ok)
{
	const uint32_t ix = GLF_AS_U32(x);
	const uint32_t tmp = ix - 0x3f330000u;
	const int i = (int)((tmp >> 19) & 15u);
	const uint32_t top = tmp & 0xff800000u;
	const uint32_t iz = ix - top;
	const int k = (int32_t)top >> 23;
	double invc, logc;
	T.pair(GLF_OFF_LOG2 + 2 * i, invc, logc);
	const double z = (double)GLF_AS_F32(iz);
	const double r = GLF_FMA(z, invc, -1.0);
	const double y0 = logc + (double)k;
	const double r2 = r * r;
	const double yy = GLF_FMA(GLC(GLC_POW_A0), r, GLC(GLC_POW_A1));
	const double p = GLF_FMA(GLC(GLC_POW_A2), r, GLC(GLC_POW_A3));
	const double r4 = r2 * r2;
	double q = GLF_FMA(GLC(GLC_POW_A4), r, y0);
	q = GLF_FMA(p, r2, q);
	const double logx = GLF_FMA(yy, r4, q);
	const double ylogx = (double)y * logx;
	// |y log2 x| >= 126: glibc's overflow / underflow handling, not restated
	ok = ((GLF_AS_U64(ylogx) >> 47) & 0xffffu) < (0x405f800000000000ull >> 47);
	double kd = ylogx + GLC(GLC_SHIFT47); // 0x1.8p52 / 32
	const uint64_t ki = GLF_AS_U64(kd);
	kd = kd - GLC(GLC_SHIFT47);
	const double rr = ylogx - kd;
	const uint64_t t = T.word(GLF_OFF_EXP2 + (int)(ki & 31u)) + (ki << 47);
	const double s = GLF_AS_F64(t);
	const double zz = GLF_FMA(GLC(GLC_EXP2_C0), rr, GLC(GLC_EXP2_C1));
	const double rr2 = rr * rr;
	double e = GLF_FMA(GLC(GLC_EXP2_C2), rr, 1.0);
	e = GLF_FMA(zz, rr2, e);
	e = e * s;
	return (float)e;
}

} // namespace djb200
