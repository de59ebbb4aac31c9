// djb_lean.cuh -- the microfacet eval / evalp / pdf hot path in lean FP32.
//
// Same values as the mirrored-rounding functions in djb_device.cuh, different instruction sequences:
//   * every division and square root whose operands are floats rounds once whether it is done in double and
//     rounded back (the reference) or done in float -- so they are done in float, with Markstein-style FMA
//     sequences instead of the library's guarded IEEE routines:
//         rcp:  y = MUFU.RCP(b) (1 ulp);  y <- fma(y, fma(-b, y, 1), y)                      correctly rounded 1/b
//         div:  q = a y;  q <- fma(fma(-b, q, a), y, q)                                       correctly rounded a/b
//         sqrt: g = x MUFU.RSQ(x), h = rsq/2, one coupled Newton step, g <- fma(fma(-g, g, x), h, g)
//     (correct rounding holds for operands and results in the normal range, which is where these kernels use
//     them; callers keep the guarded routine where a zero, infinity or denormal can reach the operation);
//   * divisors that depend only on the material (ax, ax ay sqrt(1 - rho^2)) or only on the pair (h.z, h.z^4,
//     4 o.z, 4 i.h, i.z) have their correctly rounded reciprocal computed once, outside the material loop;
//   * the FMAs above are explicit intrinsics: the file is still compiled with -fmad=false, so none of the
//     reference's separate multiply / add pairs is contracted.
#pragma once
#include "djb_device.cuh"

namespace djb200 {

DJB_DEV float mufu_rcp(float x)
{
	float y;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}
DJB_DEV float mufu_rsq(float x)
{
	float y;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}

// correctly rounded 1/b for normal b with a normal reciprocal (Markstein: one FMA step from a 1-ulp estimate)
DJB_DEV float rcp_lean(float b)
{
	float y = mufu_rcp(b);
	return __fmaf_rn(y, __fmaf_rn(-b, y, 1.0f), y);
}
// a / b given y = RN(1 / b)
DJB_DEV float div_by(float a, float b, float y)
{
	float q = a * y;
	return __fmaf_rn(__fmaf_rn(-b, q, a), y, q);
}
DJB_DEV float div_lean(float a, float b) { return div_by(a, b, rcp_lean(b)); }
// a / b for a >= 0 that may be tiny (Beckmann tails): the FMA residual of div_by must not underflow, so small
// numerators are lifted by 2^64 (exact), divided, and lowered again (exact unless the quotient is subnormal,
// in which case the guarded IEEE division does the single rounding)
DJB_DEV float div_by_small(float a, float b, float y)
{
	if (a > 1e-18f) return div_by(a, b, y);
	float q = div_by(a * 0x1p64f, b, y) * 0x1p-64f;
	return fabsf(q) > 0x1p-120f ? q : __fdiv_rn(a, b);
}
// correctly rounded sqrt(x) for normal x > 0
DJB_DEV float sqrt_lean(float x)
{
	float y = mufu_rsq(x);
	float g = x * y, h = 0.5f * y;
	float r = __fmaf_rn(-g, h, 0.5f);
	g = __fmaf_rn(g, r, g);
	h = __fmaf_rn(h, r, h);
	return __fmaf_rn(__fmaf_rn(-g, g, x), h, g);
}

// ---- Beckmann's standard-space functions (sigma_std_radial, p22_radial; dj_brdf.h:1866-1879, djb::erf :667-688) ----
// The reference evaluates exp() in double and rounds the surrounding expression to float.  Here exp(x), x <= 0, is
// produced as an unevaluated float-float sum 2^m (Ph + Pl) with relative error < 2^-44 (table-driven: x = (64 m + j)
// ln2/64 + r, |r| < 0.0055, exp(r) from a short series whose first terms are carried in float-float), and each
// surrounding expression is carried in float-float up to its final rounding.  The rounded floats agree with the
// double evaluation except at double-rounding ties (< 1e-6 of inputs, 1 ulp): tests compare the two at scale.
// 2^(j/64) as float-float, j = 0..63 (host long-double exp2, split)
__device__ const float2 g_exp2_64[64] = {
	{0x1.0000000000000p+0f, 0x0.0p+0f},
	{0x1.02c9a40000000p+0f, -0x1.887fa00000000p-28f},
	{0x1.059b0e0000000p+0f, -0x1.9d4f520000000p-25f},
	{0x1.0874520000000p+0f, -0x1.e2990e0000000p-26f},
	{0x1.0b55860000000p+0f, 0x1.9f31220000000p-25f},
	{0x1.0e3ec40000000p+0f, -0x1.a585cc0000000p-25f},
	{0x1.11301e0000000p+0f, -0x1.fdb4960000000p-25f},
	{0x1.1429aa0000000p+0f, 0x1.d525bc0000000p-25f},
	{0x1.172b840000000p+0f, -0x1.c157420000000p-27f},
	{0x1.1a35be0000000p+0f, 0x1.6df96e0000000p-25f},
	{0x1.1d48740000000p+0f, -0x1.d2e8ca0000000p-25f},
	{0x1.2063b80000000p+0f, 0x1.0c519a0000000p-25f},
	{0x1.2387a60000000p+0f, 0x1.ceac480000000p-25f},
	{0x1.26b4560000000p+0f, 0x1.789f380000000p-26f},
	{0x1.29e9e00000000p+0f, -0x1.5c04240000000p-25f},
	{0x1.2d285a0000000p+0f, 0x1.b900c20000000p-26f},
	{0x1.306fe00000000p+0f, 0x1.4636e20000000p-25f},
	{0x1.33c08c0000000p+0f, -0x1.b37d200000000p-25f},
	{0x1.371a740000000p+0f, -0x1.18aac60000000p-25f},
	{0x1.3a7db40000000p+0f, -0x1.634c020000000p-25f},
	{0x1.3dea640000000p+0f, 0x1.8246840000000p-25f},
	{0x1.4160a20000000p+0f, 0x1.f72e2a0000000p-28f},
	{0x1.44e0860000000p+0f, 0x1.8624b40000000p-30f},
	{0x1.486a2c0000000p+0f, -0x1.47d8660000000p-25f},
	{0x1.4bfdae0000000p+0f, -0x1.593abc0000000p-25f},
	{0x1.4f9b280000000p+0f, -0x1.2c5a6c0000000p-25f},
	{0x1.5342b60000000p+0f, -0x1.2c56100000000p-25f},
	{0x1.56f4740000000p+0f, -0x1.295b040000000p-25f},
	{0x1.5ab07e0000000p+0f, -0x1.5bd5ec0000000p-27f},
	{0x1.5e76f20000000p+0f, -0x1.4a5bd60000000p-25f},
	{0x1.6247ec0000000p+0f, -0x1.f8b5500000000p-25f},
	{0x1.6623880000000p+0f, 0x1.2a91120000000p-27f},
	{0x1.6a09e60000000p+0f, 0x1.9fcef40000000p-26f},
	{0x1.6dfb240000000p+0f, -0x1.cd72e80000000p-27f},
	{0x1.71f75e0000000p+0f, 0x1.1d8bee0000000p-25f},
	{0x1.75feb60000000p+0f, -0x1.37b3060000000p-25f},
	{0x1.7a11480000000p+0f, -0x1.829fd00000000p-25f},
	{0x1.7e2f340000000p+0f, -0x1.2616340000000p-25f},
	{0x1.82589a0000000p+0f, -0x1.accc7c0000000p-26f},
	{0x1.868d9a0000000p+0f, -0x1.2edb440000000p-26f},
	{0x1.8ace540000000p+0f, 0x1.15506e0000000p-27f},
	{0x1.8f1aea0000000p+0f, -0x1.baa2320000000p-26f},
	{0x1.93737c0000000p+0f, -0x1.e647440000000p-25f},
	{0x1.97d82a0000000p+0f, -0x1.0d8d840000000p-31f},
	{0x1.9c49180000000p+0f, 0x1.51f8480000000p-27f},
	{0x1.a0c6680000000p+0f, -0x1.2886a60000000p-26f},
	{0x1.a5503c0000000p+0f, -0x1.b83b540000000p-25f},
	{0x1.a9e6b60000000p+0f, -0x1.50c0480000000p-25f},
	{0x1.ae89fa0000000p+0f, -0x1.a94b140000000p-26f},
	{0x1.b33a2c0000000p+0f, -0x1.ec3a820000000p-26f},
	{0x1.b7f7700000000p+0f, -0x1.a094380000000p-25f},
	{0x1.bcc1ea0000000p+0f, -0x1.f687c60000000p-25f},
	{0x1.c199be0000000p+0f, -0x1.3d56b20000000p-27f},
	{0x1.c67f120000000p+0f, 0x1.cafa2a0000000p-25f},
	{0x1.cb720e0000000p+0f, -0x1.8837cc0000000p-27f},
	{0x1.d072d40000000p+0f, 0x1.40f1300000000p-25f},
	{0x1.d5818e0000000p+0f, -0x1.822dbc0000000p-27f},
	{0x1.da9e600000000p+0f, 0x1.ed99420000000p-27f},
	{0x1.dfc9740000000p+0f, -0x1.908c940000000p-25f},
	{0x1.e502ee0000000p+0f, 0x1.e2cffe0000000p-26f},
	{0x1.ea4afa0000000p+0f, 0x1.52486c0000000p-27f},
	{0x1.efa1be0000000p+0f, 0x1.cc2b440000000p-25f},
	{0x1.f507660000000p+0f, -0x1.246eb00000000p-26f},
	{0x1.fa7c180000000p+0f, 0x1.9e90d80000000p-28f}
};

struct ExpFF { float ph, pl; int m; }; // exp(x) = 2^m (ph + pl), ph in [1, 2.01)

// x in [-104, 0]
DJB_DEV ExpFF exp_ff(const float2 *__restrict__ T, float x)
{
	const float INV = 0x1.715476p+6f;                                           // 64 / ln 2
	const float C1 = 0x1.63p-7f, C2 = -0x1.bdp-19f, C3 = -0x1.05c61p-35f;       // ln2 / 64 = C1 + C2 + C3, kf C1 and kf C2 exact
	const int k = __float2int_rn(x * INV);
	const float kf = (float)k;
	const float r2 = __fmaf_rn(-kf, C2, __fmaf_rn(-kf, C1, x));                  // both exact
	const float th = kf * C3, tl = __fmaf_rn(kf, C3, -th);
	const float rh = r2 - th;
	const float rl = ((r2 - rh) - th) - tl;                                     // r = rh + rl
	// exp(r) = 1 + r + r^2/2 + r^3 (1/6 + r/24 + r^2/120): the first three terms in float-float
	const float sh = rh * rh, sl = __fmaf_rn(rh, rh, -sh);
	const float p3 = __fmaf_rn(rh, __fmaf_rn(rh, 1.0f / 120.0f, 1.0f / 24.0f), 1.0f / 6.0f);
	const float w = (sh * rh) * p3;
	const float hs = 0.5f * sh;
	const float Ah = rh + hs, Al = (rh - Ah) + hs;
	const float B = (((Al + rl) + 0.5f * sl) + rh * rl) + w;
	const float Eh = 1.0f + Ah, El = ((1.0f - Eh) + Ah) + B;
	const int j = k & 63; // two's complement: k = 64 m + j with 0 <= j < 64
	const float2 t = T[j];
	ExpFF e;
	e.m = (k - j) >> 6;
	e.ph = t.x * Eh;
	e.pl = (__fmaf_rn(t.x, Eh, -e.ph) + t.x * El) + t.y * Eh;
	return e;
}

DJB_DEV float pow2i(int m) { return __int_as_float((m + 127) << 23); } // m in [-126, 127]

// beckmann::p22_radial, dj_brdf.h:1866-1869: float(exp(-r2) / M_PI)
DJB_DEV float beck_p22_lean(const float2 *__restrict__ T, float r2)
{
	if (r2 > 103.5f) return 0.0f; // exp(-103.5) / pi < 2^-150: rounds to zero
	if (!(r2 >= 0.0f)) return (float)(exp((double)(-r2)) / DJB_PI); // NaN (or a negative argument): literal path
	const ExpFF e = exp_ff(T, -r2);
	const float IPH = 0x1.45f306p-2f, IPL = 0x1.b93910p-27f; // 1 / M_PI as float-float
	const float qh = e.ph * IPH;
	const float ql = (__fmaf_rn(e.ph, IPH, -qh) + e.ph * IPL) + e.pl * IPH;
	if (e.m >= -124) return (qh + ql) * pow2i(e.m); // normal result: one rounding, exact scaling
	// Result below 2^-125: the float grid there is the integer grid in units of 2^-149, so the one rounding is
	// done by hand -- value in grid units as head + tail (TwoSum), round half to even, the tail breaks ties --
	// and the integer is the bit pattern of the (sub)normal float.
	const float s = pow2i(e.m + 149); // m >= -150: s >= 1/2
	const float t = qh * s, u = ql * s; // exact
	const float vh = t + u, bb = vh - t;
	const float vl = (t - (vh - bb)) + (u - bb);
	const float n0 = rintf(vh), d = vh - n0;
	const int adj = (d == 0.5f && vl > 0.0f) ? 1 : ((d == -0.5f && vl < 0.0f) ? -1 : 0);
	return __int_as_float((int)n0 + adj);
}

// beckmann::sigma_std_radial, dj_brdf.h:1871-1879
DJB_DEV float beck_sigma_std_lean(const float2 *__restrict__ T, float c)
{
	if (c == 1.0f) return 1.0f;
	// s = float(sqrt(1.0 - c c)): 1 - cc as float-float, then one corrected square root
	const float cc = c * c;
	const float vh = 1.0f - cc, vl = (1.0f - vh) - cc;
	if (!(vh > 1e-28f)) return sigma_std_radial<NDF_BECKMANN>(c); // |c| == 1 to rounding, NaN: literal path
	const float y = mufu_rsq(vh);
	float g = vh * y, hh = 0.5f * y;
	const float rr = __fmaf_rn(-g, hh, 0.5f);
	g = __fmaf_rn(g, rr, g);
	hh = __fmaf_rn(hh, rr, hh);
	const float s = __fmaf_rn(__fmaf_rn(-g, g, vh) + vl, hh, g);
	const float nu = div_lean(c, s);
	const float x = -nu * nu;
	// for nu > 4.06 the exponential is below every rounding threshold of the expression: erf rounds to 1 and
	// s exp(-nu^2) / sqrt(pi) is less than half an ulp of 2 c, so the value is c itself
	if (x < -16.5f && c > 0.0f) return c;
	if (x < -100.0f) return sigma_std_radial<NDF_BECKMANN>(c);
	const ExpFF e = exp_ff(T, x);
	// tmp = float(exp * inv_sqrt_pi), inv_sqrt_pi = float(1 / sqrt(float(M_PI)))
	const float ISP = 0x1.20dd74p-1f;
	const float th = e.ph * ISP, tl = __fmaf_rn(e.ph, ISP, -th) + e.pl * ISP;
	// erf(nu), A&S 7.1.26 as the reference evaluates it: t = float(1.0 / (1.0 + p |nu|)), Horner in float,
	// y = float(1.0 - (poly t) exp)
	const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f;
	const float px = 0.3275911f * fabsf(nu);
	const float uh = 1.0f + px;
	const float ul = px <= 1.0f ? px - (uh - 1.0f) : 1.0f - (uh - px);
	const float y0 = rcp_lean(uh);
	const float t = __fmaf_rn(y0, __fmaf_rn(-uh, y0, 1.0f) - ul * y0, y0);
	const float poly = ((((a5 * t + a4) * t) + a3) * t + a2) * t + a1;
	const float pt = poly * t;
	float tmp, yerf;
	if (e.m >= -120) {
		const float sc = pow2i(e.m);
		tmp = (th + tl) * sc;
		const float mh = pt * e.ph, ml = __fmaf_rn(pt, e.ph, -mh) + pt * e.pl; // (poly t) exp = sc (mh + ml)
		const float Mh = mh * sc, Ml = ml * sc;
		const float dh = 1.0f - Mh, dl = (1.0f - dh) - Mh;
		yerf = dh + (dl - Ml);
	} else { // exp < 2^-119: invisible next to 1, and s tmp is invisible next to anything it is added to
		tmp = 0.0f;
		yerf = 1.0f;
		if (!(c > 0.0f)) return sigma_std_radial<NDF_BECKMANN>(c); // c (1 + erf) == 0: the tiny term is the result
	}
	const float erfv = nu < 0.0f ? -yerf : yerf;
	// float((c (1.0 + erf) + s tmp) / 2.0): c (1 + erf) exact in float-float, s tmp a float product
	const float wh = 1.0f + erfv, wl = erfv - (wh - 1.0f);
	const float ph = c * wh, pl = __fmaf_rn(c, wh, -ph) + c * wl;
	const float q = s * tmp;
	const float ah = ph + q;
	const float bb = ah - ph;
	const float al = (ph - (ah - bb)) + (q - bb); // TwoSum
	return 0.5f * (ah + (al + pl));
}


struct FF { float h, l; }; // unevaluated sum h + l

// ---- the sampling path's logf / powf / expf (dj_brdf.h:695, 1917, 1935) ------------------------------------------
// glibc's own algorithms in double (djb_glibcf.h): the results are glibc's results, bit for bit, at 7-12 double
// operations each.  GT = the table block of djb_glibcf.h (staged in shared memory by the kernels).  Arguments outside
// the restated main branch (zero, subnormal, negative, NaN, huge) take the literal double path of djb_device.cuh.
struct GlfCtx { GlfTableShared T; GlfHot H; }; // the staged tables + the register-held constants, set up once per kernel
DJB_DEV float logf_lean(const GlfCtx &GT, float x)
{
	if (!glf_logf_ok(x)) return logf_literal(x);
	return glf_logf(GT.T, GT.H, x);
}
DJB_DEV float powf_lean(const GlfCtx &GT, float x, float y)
{
	if (!glf_powf_ok(x, y)) return powf_literal(x, y);
	bool ok;
	const float r = glf_powf(GT.T, x, y, ok);
	return ok ? r : powf_literal(x, y);
}
DJB_DEV float expf_lean(const GlfCtx &GT, float x)
{
	if (!glf_expf_ok(x)) return expf_literal(x);
	return glf_expf(GT.T, GT.H, x);
}

// ---- pieces of the visible-normal sampling path (dj_brdf.h:1669-1709, 1818-1846, 1897-1957, 2089-2146) --------
// float(sqrt(1.0 - double(c * c))): 1 - cc as float-float, one corrected square root
DJB_DEV float sqrt_1m_sq(float c)
{
	const float cc = c * c;
	const float vh = 1.0f - cc, vl = (1.0f - vh) - cc;
	if (!(vh > 1e-28f)) return (float)sqrt(1.0 - (double)cc); // |c| >= 1 to rounding, NaN: literal path
	const float y = mufu_rsq(vh);
	float g = vh * y, hh = 0.5f * y;
	const float rr = __fmaf_rn(-g, hh, 0.5f);
	g = __fmaf_rn(g, rr, g);
	hh = __fmaf_rn(hh, rr, hh);
	return __fmaf_rn(__fmaf_rn(-g, g, vh) + vl, hh, g);
}
// float((nh + nl) / (dh + dl)): quotient of two float-float values, one rounding
DJB_DEV float div_ff(float nh, float nl, float dh, float dl)
{
	const float y = rcp_lean(dh);
	const float q = nh * y;
	const float r = __fmaf_rn(-dh, q, nh) + (nl - dl * q);
	return __fmaf_rn(r, y, q);
}
DJB_DEV FF two_sum(float a, float b) // a + b == h + l exactly
{
	FF s;
	s.h = a + b;
	const float bb = s.h - a;
	s.l = (a - (s.h - bb)) + (b - bb);
	return s;
}

// ggx::qf2_radial, dj_brdf.h:2089-2119
DJB_DEV float ggx_qf2_lean(float u, float ck, float sk)
{
	const float T45 = 0x1.6a09eep-1f; // the smallest float above the reference's double constant 0.707107
	// st = float(u * (1.0 + ck) - 1.0): the product and the difference are exact in double, one rounding
	const float wh = 1.0f + ck, wl = ck - (wh - 1.0f);
	const float pr = u * wh, pe = __fmaf_rn(u, wh, -pr) + u * wl;
	const FF d = two_sum(pr, -1.0f);
	const float st = d.h + (d.l + pe);
	const float ct = sqrt_1m_sq(st);
	// The reference branches four ways on (cos theta > 0.707107, sin theta_k < 0.707107) so that each tangent /
	// cotangent is formed from the well-conditioned quotient (dj_brdf.h:2096-2118).  The four formulas are the same
	// operations on selected operands -- two quotients, one product, one sum or difference, one 1 +- product, one final
	// quotient -- so they are evaluated once, branch-free (random pairs put ~8 lanes of a warp in each branch):
	//   A = ct >= T45: q1 = st / ct (tan theta)      else q1 = ct / st (cot theta)
	//   B = sk <  T45: q2 = sk / ck (tan theta_k)    else q2 = ck / sk (cot theta_k)
	//   A == B:  -+(q1 + q2) / (1 - q1 q2)   (minus when both are tangents)
	//   A != B:  (1 + q1 q2) / (q1 - q2 when A, q2 - q1 otherwise)
	const bool A = ct >= T45; // (double)ct > 0.707107
	const bool B = sk < T45;
	const float q1 = div_lean(A ? st : ct, A ? ct : st);
	const float q2 = div_lean(B ? sk : ck, B ? ck : sk);
	const float p = q1 * q2;
	const bool same = A == B;
	const FF t = two_sum(1.0f, same ? -p : p);
	const float sum = q1 + q2;
	const float nh = same ? (A ? -sum : sum) : t.h, nl = same ? 0.0f : t.l;
	const float dh = same ? t.h : (A ? q1 - q2 : q2 - q1), dl = same ? t.l : 0.0f;
	return div_ff(nh, nl, dh, dl);
}

// What the sampling path derives from the pair's second uniform alone: computed once per pair, outside the material loop.
//   Beckmann: a = qf3_radial(u2) = erfinv(2 u2 - 1)                                          (dj_brdf.h:1954-1957, 1891-1894)
//   GGX:      a = the sign S, b = the rational approximation's quotient pn / qn              (dj_brdf.h:2121-2146)
struct SampleU2 { float a, b; };

// ggx::qf3_radial + qf3_rational_approx, dj_brdf.h:2121-2146 (the two double Horner forms stay in double)
DJB_DEV SampleU2 ggx_qf3_u2(float u)
{
	SampleU2 r;
	if (u < 0.5f) { u = (0.5f - u) * 2.0f; r.a = -1.0f; } // float(2.0 * (0.5 - u)): one rounding either way
	else { u = (u - 0.5f) * 2.0f; r.a = 1.0f; }
	const double du = (double)u;
	const float pn = (float)(du * (du * (du * (-0.365728915865723) + 0.790235037209296) - 0.424965825137544)
	                         + 0.000152998850436920);
	const float qn = (float)(du * (du * (du * (du * 0.169507819808272 - 0.397203533833404) - 0.232500544458471) + 1.0)
	                         - 0.539825872510702);
	r.b = div_lean(pn, qn);
	return r;
}
DJB_DEV float ggx_qf3_lean(SampleU2 su, float qf2)
{
	// alpha = float(sqrt(1.0 + double(qf2 * qf2)))
	const FF v = two_sum(1.0f, qf2 * qf2);
	float alpha;
	if (v.h < 1e30f) {
		const float y = mufu_rsq(v.h);
		float g = v.h * y, hh = 0.5f * y;
		const float rr = __fmaf_rn(-g, hh, 0.5f);
		g = __fmaf_rn(g, rr, g);
		hh = __fmaf_rn(hh, rr, hh);
		alpha = __fmaf_rn(__fmaf_rn(-g, g, v.h) + v.l, hh, g);
	} else {
		alpha = (float)sqrt(1.0 + (double)(qf2 * qf2));
	}
	return su.a * alpha * su.b; // S * alpha * (pn / qn), left to right
}

// djb::erfinv (Giles), dj_brdf.h:691-721: the part after w = -logf((1 - u)(1 + u))
DJB_DEV float erfinv_poly(float w, float u)
{
	float p;
	if (w < 5.0f) {
		w = w - 2.5f;
		p = 2.81022636e-08f;
		p = 3.43273939e-07f + p * w;
		p = -3.5233877e-06f + p * w;
		p = -4.39150654e-06f + p * w;
		p = 0.00021858087f + p * w;
		p = -0.00125372503f + p * w;
		p = -0.00417768164f + p * w;
		p = 0.246640727f + p * w;
		p = 1.50140941f + p * w;
	} else {
		// w = float(sqrt(double(w)) - 3.0): correctly rounded root + its residual, then one rounding
		if (w < 1e30f) {
			const float sh = sqrt_lean(w);
			const float sl = __fmaf_rn(-sh, sh, w) * (0.5f * mufu_rcp(sh));
			w = (sh - 3.0f) + sl;
		} else {
			w = (float)(sqrt((double)w) - 3.0);
		}
		p = -0.000200214257f;
		p = 0.000100950558f + p * w;
		p = 0.00134934322f + p * w;
		p = -0.00367342844f + p * w;
		p = 0.00573950773f + p * w;
		p = -0.0076224613f + p * w;
		p = 0.00943887047f + p * w;
		p = 1.00167406f + p * w;
		p = 2.83297682f + p * w;
	}
	return p * u;
}
DJB_DEV float erfinv_lean(const GlfCtx &GT, float u) { return erfinv_poly(-logf_lean(GT, (1.0f - u) * (1.0f + u)), u); }
// one trip of the quantile search needs erfinv(b) and expf(-erfinv(b)^2): the rare arguments outside the restated branches of
// logf / expf (b = +-1, NaN) are tested once, after both fast evaluations, and redone literally
static __device__ __noinline__ float2 erfinv_exp_literal(float u)
{
	const float ie = erfinv_giles(u);
	return make_float2(ie, expf_cr(-ie * ie));
}
DJB_DEV void erfinv_exp_lean(const GlfCtx &GT, float u, float &ie, float &ex)
{
	const float x1 = (1.0f - u) * (1.0f + u);
	ie = erfinv_poly(-glf_logf(GT.T, GT.H, x1), u);
	const float x2 = -ie * ie;
	ex = glf_expf(GT.T, GT.H, x2);
	if (!(glf_logf_ok(x1) && glf_expf_ok(x2))) {
		const float2 r = erfinv_exp_literal(u);
		ie = r.x;
		ex = r.y;
	}
}

// beckmann::qf2_radial, dj_brdf.h:1897-1952
DJB_DEV float beckmann_qf2_lean(const float2 *__restrict__ T, const GlfCtx &GT, float u, float ck, float sk)
{
	const float sqrt_pi_inv = (float)(1.0 / sqrt(DJB_PI));
	const bool sk_ok = sk > 1e-18f; // sk == 0 (k along the normal): cot = inf, the guarded division gives it
	const float cot = sk_ok ? div_lean(ck, sk) : __fdiv_rn(ck, sk), tan_k = div_lean(sk, ck);
	// c = djb::erf(cot) (A&S 7.1.26, dj_brdf.h:667-688) and
	// normalization = float(1.0 / (double(1 + c) + double(sqrt_pi_inv * tan_k) * exp(double(-cot * cot))))
	// contain the same exponential (erf's own argument is -|cot| |cot|, the same float): one float-float evaluation serves both.
	const float xx = -cot * cot;
	float c, normalization;
	const float B = sqrt_pi_inv * tan_k;
	if (xx > -40.0f) {
		const ExpFF e = exp_ff(T, xx);
		const float sc = pow2i(e.m); // m >= -58
		// |cot| >= 4: (poly t) exp(-cot^2) is below half an ulp of 1.0f, the reference's float(1.0 - ...) is exactly 1
		// (checked exhaustively on the CPU for every float in [3.9, 10]: 1.0f from 3.9195216 on)
		float y = 1.0f;
		if (xx > -16.0f) {
			const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f;
			const float px = 0.3275911f * fabsf(cot);
			const float uh = 1.0f + px;
			const float ul = px <= 1.0f ? px - (uh - 1.0f) : 1.0f - (uh - px);
			const float y0 = rcp_lean(uh);
			const float t = __fmaf_rn(y0, __fmaf_rn(-uh, y0, 1.0f) - ul * y0, y0);
			const float poly = ((((a5 * t + a4) * t) + a3) * t + a2) * t + a1;
			const float pt = poly * t;
			const float mh = pt * e.ph, ml = __fmaf_rn(pt, e.ph, -mh) + pt * e.pl;
			const float Mh = mh * sc, Ml = ml * sc;
			const float dh = 1.0f - Mh, dl = (1.0f - dh) - Mh;
			y = dh + (dl - Ml);
		}
		c = cot < 0.0f ? -y : y;
		const float A = 1.0f + c;
		const float bh = B * e.ph, bl = __fmaf_rn(B, e.ph, -bh) + B * e.pl;
		const FF d = two_sum(A, bh * sc);
		normalization = div_ff(1.0f, 0.0f, d.h, d.l + bl * sc);
	} else if (xx == xx) {
		c = cot < 0.0f ? -1.0f : 1.0f;
		normalization = rcp_lean(1.0f + c); // exp(-cot^2) < 2^-57: invisible next to 1 + c
		if (!(c > 0.0f)) normalization = (float)(1.0 / ((double)(1.0f + c) + (double)B * exp((double)xx)));
	} else {
		c = erf_as(cot);
		normalization = (float)(1.0 / ((double)(1.0f + c) + (double)B * exp((double)xx)));
	}
	float a = -1.0f;
	u = fmax_ref(u, 1e-6f);
	const float fit = 1.0f + ck * (-0.876f + ck * (0.4265f - 0.0594f * ck));
	float b = c - (1.0f + c) * powf_lean(GT, 1.0f - u, fit);
	int it = 0;
	float ie = 0.0f;
	bool converged = false;
	while (++it < 10) {
		if (!(b >= a && b <= c)) b = 0.5f * (a + c);
		float ex;
		erfinv_exp_lean(GT, b, ie, ex);
		const float value = normalization * (1.0f + b + sqrt_pi_inv * tan_k * ex) - u;
		const float derivative = normalization * (1.0f - ie * tan_k);
		if (fabsf(value) < 1e-5f) { converged = true; break; }
		if (value > 0.0f) c = b; else a = b;
		b -= __fdiv_rn(value, derivative);
	}
	// erfinv(max(-0.9999, b)): after a converged search b is the argument `ie` was just computed from
	const float bm = fmax_ref(-0.9999f, b);
	if (converged && bm == b) return ie;
	return erfinv_lean(GT, bm);
}

template <int NDF>
DJB_DEV SampleU2 lean_sample_u2(const GlfCtx &GT, float u2) // u2: already clamped as microfacet::sample does
{
	if (NDF == NDF_GGX) return ggx_qf3_u2(u2);
	SampleU2 r;
	r.a = erfinv_lean(GT, 2.0f * u2 - 1.0f); // float(2.0 * u2 - 1): the product is exact, one rounding
	r.b = 0.0f;
	return r;
}
DJB_DEV float sample_clamp_u(float u) { return sat_ref(u) * 0.99998f + 0.00001f; } // dj_brdf.h:1680-1681

// radial::sample_vp22_std_smith, dj_brdf.h:1818-1846
template <int NDF>
DJB_DEV void lean_std_slopes(const float2 *T, const GlfCtx &GT, float u1, SampleU2 su2, V3 k, float &xs, float &ys)
{
	const float ck = k.z;
	const float sk = k.z < 1.0f ? sqrt_1m_sq(k.z) : 0.0f;
	float tx, ty;
	if (NDF == NDF_GGX) {
		tx = ggx_qf2_lean(u1, ck, sk);
		ty = ggx_qf3_lean(su2, tx);
	} else {
		tx = beckmann_qf2_lean(T, GT, u1, ck, sk);
		ty = su2.a;
	}
	if (sk == 0.0f) {
		xs = tx;
		ys = ty;
	} else {
		const float nrm = inv_sqrt(k.x * k.x + k.y * k.y);
		const float cp = k.x * nrm, sp = k.y * nrm;
		xs = cp * tx - sp * ty;
		ys = sp * tx + cp * ty;
	}
}

// microfacet::sample, dj_brdf.h:1669-1709; u1 clamped, su2 = lean_sample_u2(clamped u2)
template <int NDF>
DJB_DEV V3 lean_sample(const float2 *T, const GlfCtx &GT, const Params &p, float u1, SampleU2 su2, V3 o)
{
	const float oyay = o.y * p.ay;
	const float a = o.x * p.ax + oyay * p.rho;
	const float b = oyay * p.srho;
	const float c = o.z - o.x * p.tx - o.y * p.ty;
	const V3 os = normalize(mk(a, b, c));
	if (os.z > 0.0f) {
		float txm, tym;
		lean_std_slopes<NDF>(T, GT, u1, su2, os, txm, tym);
		const float txh = p.ax * txm + p.tx;
		const float chol = p.rho * txm + p.srho * tym;
		const float tyh = p.ay * chol + p.ty;
		const V3 h = normalize(mk(-txh, -tyh, 1.0f));
		const float k = 2.0f * dot(o, h); // float(2.0 * dot): exact
		return scale(k, h) - o;
	}
	return mk(0.f, 0.f, 1.f);
}

// params block + the per-material constants of the hot loop (staged once per CTA in shared memory)
struct ParamsX {
	Params p;
	float rcp_ax;  // RN(1 / ax)
	float rho_ay;  // rho * ay          (first product of `rho * ay * x`, dj_brdf.h:1583)
	float nrm;     // ax * ay * sqrt(1 - rho^2)   (dj_brdf.h:1578, 1584)
	float rcp_nrm; // RN(1 / nrm)
};
DJB_DEV ParamsX extend_params(const Params &p)
{
	ParamsX x;
	x.p = p;
	x.rcp_ax = __frcp_rn(p.ax);
	x.rho_ay = p.rho * p.ay;
	x.nrm = p.ax * p.ay * p.srho;
	x.rcp_nrm = __frcp_rn(x.nrm);
	return x;
}

// what does not depend on the material
struct PairX {
	V3 i, o, h;
	float sx, sy;       // slopes of h: -h.x / h.z, -h.y / h.z   (dj_brdf.h:1564-1565)
	float c4, rcp_c4;   // h.z^4
	float den, rcp_den; // 4 o.z (eval / evalp) or 4 (i . h) (pdf)
	float inv_iz;       // RN(1 / i.z) (eval)
	float cd;           // sat(o . h)
	float kh;           // o . h (pdf: the visible-normal density's numerator)
	bool facing;        // h.z > 1e-4: the NDF is non-zero (dj_brdf.h:1561)
	bool den_ok;        // den, c4 and their reciprocals are normal numbers: the lean divisions are exact
	bool both_up;       // pdf, 1e-5 tier: i.z > 0, o.z > 0, i.z o.z > 1e-30, |i|, |o| <= 2 (see fast_pdf_try)
};
// A centred lobe of bounded roughness (params::isotropic / elliptic: n = (0, 0, 1), no slope offset): both G1 lie in (0, 1], so
// the shadowing term is positive exactly when both are (fast_pdf_try)
DJB_DEV bool params_centred(const Params &p)
{
	return p.tx == 0.0f && p.ty == 0.0f && p.nx == 0.0f && p.ny == 0.0f && p.nz == 1.0f && p.ax > 0.0f && p.ax <= 1e3f && p.ay > 0.0f &&
	       p.ay <= 1e3f && fabsf(p.rho) < 1.0f;
}

template <int OP>
DJB_DEV PairX make_pair(V3 i, V3 o)
{
	PairX c;
	c.i = i; c.o = o;
	c.h = normalize(i + o);
	c.facing = c.h.z > 1e-4f;
	float rz = __frcp_rn(c.h.z);
	c.sx = div_by(-c.h.x, c.h.z, rz);
	c.sy = div_by(-c.h.y, c.h.z, rz);
	float c2 = c.h.z * c.h.z;
	c.c4 = c2 * c2;
	c.rcp_c4 = __frcp_rn(c.c4);
	c.den = 4.0f * (OP == OP_PDF ? dot(i, c.h) : o.z);
	c.rcp_den = __frcp_rn(c.den);
	c.inv_iz = OP == OP_EVAL ? __frcp_rn(i.z) : 0.0f;
	c.kh = dot(o, c.h);
	c.cd = sat_ref(c.kh);
	const float lo = 1e-28f, hi = 1e30f;
	c.den_ok = fabsf(c.den) > lo && fabsf(c.den) < hi && c.c4 > lo;
	c.both_up = OP == OP_PDF && i.z > 0.0f && o.z > 0.0f && i.z * o.z > 1e-30f && dot(i, i) <= 4.0f && dot(o, o) <= 4.0f;
	return c;
}

// microfacet::sigma, dj_brdf.h:1619-1631
template <int NDF>
DJB_DEV float lean_sigma(const float2 *T, const Params &p, V3 k)
{
	float kyay = k.y * p.ay;
	float a = k.x * p.ax + kyay * p.rho;
	float b = kyay * p.srho;
	float c = k.z - k.x * p.tx - k.y * p.ty;
	float nrm = sqrt_lean(a * a + b * b + c * c);
	float cz = rcp_lean(nrm) * c;
	return nrm * (NDF == NDF_BECKMANN ? beck_sigma_std_lean(T, cz) : sigma_std_radial<NDF_GGX>(cz));
}

// microfacet::g1, dj_brdf.h:1633-1642
template <int NDF>
DJB_DEV float lean_g1(const float2 *T, const Params &p, V3 k)
{
	if (dot(k, mk(p.nx, p.ny, p.nz)) > 0.0f) return div_lean(k.z, lean_sigma<NDF>(T, p, k));
	return 0.0f;
}

// microfacet::gaf, dj_brdf.h:1644-1665
template <int NDF>
DJB_DEV float lean_gaf(const float2 *T, const Params &p, bool shadow, V3 i, V3 o)
{
	float g1o = lean_g1<NDF>(T, p, o);
	if (shadow) {
		float g1i = lean_g1<NDF>(T, p, i);
		float t = g1i * g1o;
		if (t > 0.0f) return div_lean(t, g1i + g1o - t);
		return 0.0f;
	}
	return g1o;
}

// microfacet::ndf + p22, dj_brdf.h:1559-1587, with the per-pair and per-material reciprocals.  In two steps so that the
// Beckmann kernel can test the cheap part (squared standard-space slope radius) on every lane and compact the rest:
// r2 > 103.5 means exp(-r2) / pi rounds to zero, i.e. D == 0 exactly (beck_p22_lean's first test).
DJB_DEV float lean_ndf_r2(const ParamsX &m, const PairX &c)
{
	float x = c.sx - m.p.tx, y = c.sy - m.p.ty;
	float xs = div_by(x, m.p.ax, m.rcp_ax);
	float t1 = m.p.ax * y - m.rho_ay * x;
	float ys = div_by(t1, m.nrm, m.rcp_nrm);
	return xs * xs + ys * ys;
}
template <int NDF>
DJB_DEV float lean_ndf_from_r2(const float2 *T, const ParamsX &m, const PairX &c, float r2)
{
	float pv = NDF == NDF_BECKMANN ? beck_p22_lean(T, r2) : p22_radial<NDF_GGX>(r2);
	if (NDF == NDF_BECKMANN) { // the exponential underflows gradually: small numerators take the lifted division
		if (pv == 0.0f) return 0.0f;
		return div_by_small(div_by_small(pv, m.nrm, m.rcp_nrm), c.c4, c.rcp_c4);
	}
	return div_by(div_by(pv, m.nrm, m.rcp_nrm), c.c4, c.rcp_c4);
}
template <int NDF>
DJB_DEV float lean_ndf(const float2 *T, const ParamsX &m, const PairX &c)
{
	if (!c.facing) return 0.0f;
	return lean_ndf_from_r2<NDF>(T, m, c, lean_ndf_r2(m, c));
}

// F D G / (4 o.z) (evalp, dj_brdf.h:1529-1547); `scale` = 1 / i.z for eval (dj_brdf.h:1551-1555), unused otherwise.
// Split in two so that the Beckmann kernel can compact the expensive second half across a warp (kernels_mf.cu):
//   head: Dn = lean_ndf; if lean_skip(Dn, c) the result is lean_zero<OP>(c) -- D == 0 (h below the 1e-4 gate, or Beckmann's
//         exponential underflowed) makes F D G / (4 o.z) +0 for o.z > 0 whatever G is, so both projected areas are skipped;
//   tail: the shadowing term, the Fresnel term and the quotient, from Dn and the pair.
DJB_DEV bool lean_skip(float Dn, const PairX &c) { return Dn == 0.0f && c.den > 0.0f; }
template <int OP>
DJB_DEV V3 lean_zero(const PairX &c)
{
	const V3 z = mk(0.f, 0.f, 0.f);
	return OP == OP_EVAL ? scale(c.inv_iz, z) : z; // 0 * (1 / i.z): keeps the reference's -0 / NaN for i.z <= 0
}
template <int NDF, int FK, int OP>
DJB_DEV V3 lean_evalp_tail(const float2 *T, const Params &p, const FresnelDev &f, bool shadow, const PairX &c, float Dn)
{
	const float G = lean_gaf<NDF>(T, p, shadow, c.i, c.o);
	if (G > 0.0f) {
		const float num = Dn * G;
		float k;
		if (!c.den_ok) k = __fdiv_rn(num, c.den);
		else k = NDF == NDF_GGX ? div_by(num, c.den, c.rcp_den) : div_by_small(num, c.den, c.rcp_den);
		V3 e = scale(k, fresnel_eval<FK>(f, c.cd));
		return OP == OP_EVAL ? scale(c.inv_iz, e) : e;
	}
	return lean_zero<OP>(c);
}
template <int NDF, int FK, int OP>
DJB_DEV V3 lean_evalp(const float2 *T, const ParamsX &m, const FresnelDev &f, bool shadow, const PairX &c)
{
	const float Dn = lean_ndf<NDF>(T, m, c);
	if (lean_skip(Dn, c)) return lean_zero<OP>(c);
	return lean_evalp_tail<NDF, FK, OP>(T, m.p, f, shadow, c, Dn);
}

// microfacet::pdf, dj_brdf.h:1713-1730 with vndf, dj_brdf.h:1602-1615.  sigma(o) appears in G1(o) and in the
// visible-normal density: evaluated once.  Same head / tail split (vndf == 0: the pdf is +0 whatever G is).
template <int NDF>
DJB_DEV float lean_pdf_tail(const float2 *T, const Params &p, bool shadow, const PairX &c, float Dn)
{
	const float sg_o = lean_sigma<NDF>(T, p, c.o);
	const float rsg_o = rcp_lean(sg_o);
	const float g1o = dot(c.o, mk(p.nx, p.ny, p.nz)) > 0.0f ? div_by(c.o.z, sg_o, rsg_o) : 0.0f;
	float G = g1o;
	if (shadow) { // gaf, dj_brdf.h:1644-1665
		const float g1i = lean_g1<NDF>(T, p, c.i);
		const float t = g1i * g1o;
		G = t > 0.0f ? div_lean(t, g1i + g1o - t) : 0.0f;
	}
	if (G > 0.0f) {
		const float kh = c.kh; // dot(o, h), as make_pair formed it
		float v = 0.0f;
		if (kh > 0.0f) {
			const float num = kh * Dn;
			v = NDF == NDF_GGX ? div_by(num, sg_o, rsg_o) : div_by_small(num, sg_o, rsg_o);
		}
		if (!c.den_ok) return __fdiv_rn(v, c.den);
		return NDF == NDF_GGX ? div_by(v, c.den, c.rcp_den) : div_by_small(v, c.den, c.rcp_den);
	}
	return 0.0f;
}
template <int NDF>
DJB_DEV float lean_pdf(const float2 *T, const ParamsX &m, bool shadow, const PairX &c)
{
	const float Dn = lean_ndf<NDF>(T, m, c);
	if (lean_skip(Dn, c)) return 0.0f;
	return lean_pdf_tail<NDF>(T, m.p, shadow, c, Dn);
}

// microfacet::evalp_is, dj_brdf.h:1734-1765: sample, weight F G / G1(o), pdf = vndf / (4 cos theta_d)
template <int NDF, int FK>
DJB_DEV V3 lean_evalp_is(const float2 *T, const GlfCtx &GT, const ParamsX &m, const FresnelDev &f, bool shadow, float u1,
                         SampleU2 su2, V3 o, V3 &i_out, float &pdf_out)
{
	const Params &p = m.p;
	const V3 i = lean_sample<NDF>(T, GT, p, u1, su2, o);
	pdf_out = 0.0f;
	i_out = mk(0.f, 0.f, 0.f);
	// G1(o) and sigma(o) are shared by the shadowing term, the weight and the visible-normal density
	const float sg_o = lean_sigma<NDF>(T, p, o);
	const float rsg_o = rcp_lean(sg_o);
	const float g1o = dot(o, mk(p.nx, p.ny, p.nz)) > 0.0f ? div_by(o.z, sg_o, rsg_o) : 0.0f;
	float G = g1o;
	if (shadow) {
		const float g1i = lean_g1<NDF>(T, p, i);
		const float t = g1i * g1o;
		G = t > 0.0f ? div_lean(t, g1i + g1o - t) : 0.0f;
	}
	if (G > 0.0f) {
		PairX c; // only the half-vector fields are used by lean_ndf
		c.h = normalize(i + o);
		c.facing = c.h.z > 1e-4f;
		const float rz = rcp_lean(c.h.z);
		c.sx = div_by(-c.h.x, c.h.z, rz);
		c.sy = div_by(-c.h.y, c.h.z, rz);
		const float c2 = c.h.z * c.h.z;
		c.c4 = c2 * c2;
		c.rcp_c4 = rcp_lean(c.c4);
		const float kh = dot(o, c.h);
		const float cd = sat_ref(kh);
		i_out = i;
		float v = 0.0f;
		if (kh > 0.0f) {
			const float num = kh * lean_ndf<NDF>(T, m, c);
			v = NDF == NDF_GGX ? div_by(num, sg_o, rsg_o) : div_by_small(num, sg_o, rsg_o);
		}
		const float den = 4.0f * cd;
		if (!(den > 1e-28f)) pdf_out = __fdiv_rn(v, den);
		else pdf_out = NDF == NDF_GGX ? div_lean(v, den) : div_by_small(v, den, rcp_lean(den));
		return scale(div_by(G, g1o, rcp_lean(g1o)), fresnel_eval<FK>(f, cd));
	}
	return mk(0.f, 0.f, 0.f);
}

// =====================================================================================================================
// The 1e-5 tier of eval / evalp / pdf (BASELINE.json north_star: "match the reference header's eval / pdf to <= 1e-5 relative").
// The functions above reproduce the reference's rounded floats bit for bit; these spend the tolerance instead: MUFU reciprocal /
// reciprocal square root / exp2 without correction steps, FMA contraction, no float-float arithmetic -- about 40 % of the
// instructions.  What is NOT approximated, because the reference's own arithmetic is ill-conditioned there and only mirroring
// it stays within 1e-5 (finding 3):
//   * every gate that decides the zero pattern: h.z > 1e-4, dot(k, n) > 0, G > 0, D == 0 -- formed from the same floats by the
//     same operations as in the exact tier (h, its slopes and the per-pair denominators are the exact tier's: they are computed
//     once per pair, not per material).  One exception, decided by argument instead of evaluation: the `G > 0` gate of pdf over
//     centred lobes (fast_pdf_try);
//   * the squared standard-space slope radius r2 that enters exp(-r2): a relative error e in r2 is an error r2 e in the
//     exponential (r2 goes up to 100), so r2 is lean_ndf_r2, bit-identical to the reference's;
//   * Beckmann results in the gradual-underflow tail (r2 > 78): handed to the exact tier;
//   * s = sqrt(1 - float(c c)) of beckmann::sigma_std_radial keeps the reference's rounded product.
// Measured against the exact tier on the device and against the reference on the CPU: tests/test_gpu_parity.py::test_fast_tier_*.
DJB_DEV float mufu_ex2(float x)
{
	float y;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}
// exp(x) for -87 < x <= 0: x log2(e) as head + tail (the rounding of the product alone would cost 3e-6 at x = -80), 2^head by
// MUFU.EX2 (2 ulp), the tail applied linearly: relative error < 4e-7
DJB_DEV float exp_neg_fast(float x)
{
	const float L2E_H = 0x1.715476p+0f, L2E_L = 0x1.4ae0c0p-26f, LN2 = 0x1.62e430p-1f;
	const float th = x * L2E_H;
	float tl = __fmaf_rn(x, L2E_H, -th);
	tl = __fmaf_rn(x, L2E_L, tl);
	const float r = mufu_ex2(th);
	return __fmaf_rn(r, tl * LN2, r);
}
// beckmann::sigma_std_radial (dj_brdf.h:1871-1879) with djb::erf (A&S 7.1.26 as the reference evaluates it, :667-688)
DJB_DEV float fast_beck_sigma_std(float c)
{
	const float cc = c * c;       // the reference's rounded product: 1 - cc below is exact for cc >= 1/2
	const float v = 1.0f - cc;
	if (!(v > 0.0f)) return 1.0f; // c == 1 (the reference's early-out); c a rounding above 1 cannot happen there
	const float rs = mufu_rsq(v), s = v * rs, nu = c * rs;
	const float x = -nu * nu;
	if (x < -16.5f && c > 0.0f) return c; // same exact shortcut as the exact tier
	const float e = exp_neg_fast(fmaxf(x, -87.0f));
	const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f;
	const float d = __fmaf_rn(0.3275911f, fabsf(nu), 1.0f);
	float t = mufu_rcp(d);
	t = __fmaf_rn(t, __fmaf_rn(-d, t, 1.0f), t);
	float poly = __fmaf_rn(a5, t, a4);
	poly = __fmaf_rn(poly, t, a3);
	poly = __fmaf_rn(poly, t, a2);
	poly = __fmaf_rn(poly, t, a1);
	const float y = __fmaf_rn(-(poly * t), e, 1.0f);
	const float erfv = nu < 0.0f ? -y : y;
	return 0.5f * __fmaf_rn(c, 1.0f + erfv, s * (e * 0x1.20dd74p-1f));
}
// microfacet::sigma (dj_brdf.h:1619-1631) and, for the caller's G1, its reciprocal
template <int NDF>
DJB_DEV float fast_sigma(const Params &p, V3 k)
{
	const float kyay = k.y * p.ay;
	const float a = __fmaf_rn(k.x, p.ax, kyay * p.rho);
	const float b = kyay * p.srho;
	const float c = __fmaf_rn(-k.y, p.ty, __fmaf_rn(-k.x, p.tx, k.z));
	const float n2 = __fmaf_rn(a, a, __fmaf_rn(b, b, c * c));
	const float rn = mufu_rsq(n2), nrm = n2 * rn;
	if (NDF == NDF_GGX) return 0.5f * (nrm + c); // nrm (1 + c / nrm) / 2
	return nrm * fast_beck_sigma_std(fminf(c * rn, 1.0f));
}
// microfacet::gaf (dj_brdf.h:1644-1665); rsg_o_out: 1 / sigma(o) for the pdf.
// `ill` is set when the reference's own expression G1i G1o / (G1i + G1o - G1i G1o) is ill-conditioned: with an off-centre lobe
// (tx_n, ty_n != 0) a G1 can exceed 1 and the denominator cancels, so only the reference's exact floats reproduce its result;
// the caller then takes the exact tier for that query.  For centred lobes G1 <= 1 and the denominator is >= max(G1i, G1o): never.
template <int NDF>
DJB_DEV float fast_gaf(const Params &p, bool shadow, V3 i, V3 o, float &rsg_o_out, bool &ill)
{
	const V3 n = mk(p.nx, p.ny, p.nz);
	const float rsg_o = mufu_rcp(fast_sigma<NDF>(p, o));
	const float g1o = dot(o, n) > 0.0f ? o.z * rsg_o : 0.0f; // the gate: the reference's own dot product
	rsg_o_out = rsg_o;
	ill = false;
	if (!shadow) return g1o;
	const float g1i = dot(i, n) > 0.0f ? i.z * mufu_rcp(fast_sigma<NDF>(p, i)) : 0.0f;
	const float t = g1i * g1o, d = g1i + g1o - t;
	// cancellation (needs a G1 > 1): the quotient amplifies the G1s' relative errors (a few 1e-7 here) by (g1i + g1o + t) / d;
	// up to 8x stays well inside 1e-5.  For G1 <= 1 the denominator is >= both terms, so (g1i + g1o + t) / d <= 3: never ill
	ill = d < 0.125f * (g1i + g1o + t);
	return t > 0.0f ? t * mufu_rcp(d) : 0.0f;
}
// microfacet::ndf (dj_brdf.h:1559-1587) from the exact r2; Beckmann: r2 <= 78 (the caller sends the tail to the exact tier)
template <int NDF>
DJB_DEV float fast_ndf_from_r2(const ParamsX &m, const PairX &c, float r2)
{
	if (NDF == NDF_GGX) {
		const float t = 1.0f + r2;
		return mufu_rcp(((0x1.921fb6p+1f * t) * t) * (m.nrm * c.c4));
	}
	return ((exp_neg_fast(-r2) * 0x1.45f306p-2f) * m.rcp_nrm) * c.rcp_c4;
}
constexpr float FAST_BECK_R2_MAX = 78.0f;
// the part of a query after the cheap D == 0 / tail tests (what the Beckmann kernel compacts across a warp)
// `ill` (fast_gaf): the result is not written by the fast path; the caller evaluates the query with the exact functions -- in
// place (the plain kernels) or, in the compacting Beckmann kernel, by moving the item to the queue of exact work.
template <int NDF, int FK, int OP>
DJB_DEV V3 fast_evalp_try(const ParamsX &m, const FresnelDev &f, bool shadow, const PairX &c, float r2, bool &ill)
{
	float rsg_o;
	const float G = fast_gaf<NDF>(m.p, shadow, c.i, c.o, rsg_o, ill);
	if (G > 0.0f) {
		const float num = fast_ndf_from_r2<NDF>(m, c, r2) * G;
		const float k = c.den_ok ? num * c.rcp_den : __fdiv_rn(num, c.den);
		const V3 e = scale(k, fresnel_eval<FK>(f, c.cd));
		return OP == OP_EVAL ? scale(c.inv_iz, e) : e;
	}
	return lean_zero<OP>(c);
}
// The pdf uses the shadowing term only as a gate (dj_brdf.h:1713-1730: `if (gaf(h, i, o) > 0)`).  `centred` (params_centred, every
// material of the launch) with c.both_up decides it without sigma(i): G1(k) = k.z / sigma(k) with 0 < sigma(k) <= 4e3 for |k| <= 2,
// so both G1 are positive, their product is >= i.z o.z / 1.6e7 > 6e-38 (no underflow), and with G1 <= 1 the denominator
// G1i + G1o - G1i G1o is >= the larger of the two: the reference's G is positive, and never ill-conditioned.
template <int NDF>
DJB_DEV float fast_pdf_try(const ParamsX &m, bool shadow, const PairX &c, float r2, bool &ill, bool centred = false)
{
	float rsg_o;
	float G = 1.0f;
	if (centred && c.both_up) {
		ill = false;
		rsg_o = mufu_rcp(fast_sigma<NDF>(m.p, c.o));
	} else {
		G = fast_gaf<NDF>(m.p, shadow, c.i, c.o, rsg_o, ill);
	}
	if (G > 0.0f) {
		const float v = c.kh > 0.0f ? (c.kh * fast_ndf_from_r2<NDF>(m, c, r2)) * rsg_o : 0.0f;
		return c.den_ok ? v * c.rcp_den : __fdiv_rn(v, c.den);
	}
	return 0.0f;
}
template <int NDF, int FK, int OP>
DJB_DEV V3 fast_evalp_tail(const float2 *T, const ParamsX &m, const FresnelDev &f, bool shadow, const PairX &c, float r2)
{
	bool ill;
	const V3 r = fast_evalp_try<NDF, FK, OP>(m, f, shadow, c, r2, ill);
	if (ill) return lean_evalp_tail<NDF, FK, OP>(T, m.p, f, shadow, c, lean_ndf_from_r2<NDF>(T, m, c, r2));
	return r;
}
template <int NDF>
DJB_DEV float fast_pdf_tail(const float2 *T, const ParamsX &m, bool shadow, const PairX &c, float r2, bool centred = false)
{
	bool ill;
	const float r = fast_pdf_try<NDF>(m, shadow, c, r2, ill, centred);
	if (ill) return lean_pdf_tail<NDF>(T, m.p, shadow, c, lean_ndf_from_r2<NDF>(T, m, c, r2));
	return r;
}
// Beckmann's underflow tail (r2 > 78) takes the exact functions; r2 > 103.5 (D == 0) is their cheap early-out.  The choice is
// made per query from its own operands only, so a result never depends on which other queries share the warp or the batch.
DJB_DEV bool fast_beck_exact_vote(float r2) { return !(r2 <= FAST_BECK_R2_MAX); }
// r2 for GGX in the 1e-5 tier: D = 1 / (pi (1 + r2)^2 ...) turns a relative error e of r2 into at most 2 e, so the contracted
// form (reciprocals instead of correctly rounded quotients) is enough; Beckmann's exp(-r2) needs the exact one (r2 e, r2 <= 100)
template <int NDF>
DJB_DEV float fast_ndf_r2(const ParamsX &m, const PairX &c)
{
	if (NDF != NDF_GGX) return lean_ndf_r2(m, c);
	const float x = c.sx - m.p.tx, y = c.sy - m.p.ty;
	const float xs = x * m.rcp_ax;
	const float ys = __fmaf_rn(m.p.ax, y, -(m.rho_ay * x)) * m.rcp_nrm;
	return __fmaf_rn(xs, xs, ys * ys);
}
// one (pair, material) query of the 1e-5 tier
template <int NDF, int FK, int OP>
DJB_DEV V3 fast_evalp(const float2 *T, const ParamsX &m, const FresnelDev &f, bool shadow, const PairX &c)
{
	if (!c.facing) return c.den > 0.0f ? lean_zero<OP>(c) : lean_evalp<NDF, FK, OP>(T, m, f, shadow, c);
	const float r2 = fast_ndf_r2<NDF>(m, c);
	if (NDF == NDF_BECKMANN && fast_beck_exact_vote(r2)) return lean_evalp<NDF, FK, OP>(T, m, f, shadow, c);
	return fast_evalp_tail<NDF, FK, OP>(T, m, f, shadow, c, r2);
}
template <int NDF>
DJB_DEV float fast_pdf(const float2 *T, const ParamsX &m, bool shadow, const PairX &c, bool centred = false)
{
	if (!c.facing) return c.den > 0.0f ? 0.0f : lean_pdf<NDF>(T, m, shadow, c);
	const float r2 = fast_ndf_r2<NDF>(m, c);
	if (NDF == NDF_BECKMANN && fast_beck_exact_vote(r2)) return lean_pdf<NDF>(T, m, shadow, c);
	return fast_pdf_tail<NDF>(T, m, shadow, c, r2, centred);
}

// =====================================================================================================================
// The 1e-5 tier of sample (the same switch as eval / evalp / pdf above).  The exact tier re-runs glibc's logf / powf / expf
// in double inside the quantile search of beckmann::qf2_radial so that every trip reproduces the reference's floats; this
// tier evaluates the same search, trip for trip, with MUFU.LG2 / EX2 / RCP and FMAs (errors of a few 1e-7 per step).
// What stays exact, because the reference's own arithmetic is ill-conditioned there:
//   * the warped view vector o_std = normalize(a, b, c) and with it the gate o_std.z > 0 (the (0, 0, 1) pattern is
//     identical), cos_theta_k and sin_theta_k = float(sqrt(1 - cos^2)) (a 1-ulp change of cos_theta_k near normal incidence
//     moves sin_theta_k by per cent);
//   * GGX: sin_theta = float(u (1 + cos_theta_k) - 1) and 1 - sin_theta^2 (cancellation for u -> 0, 1);
//   * everything that depends on u2 alone (computed once per pair by lean_sample_u2).
// The search stops at the reference's own criterion |CDF(b) - u| < 1e-5; where the two tiers' values straddle it, this tier
// makes one trip more or less than the reference and the sample differs by the reference's own convergence tolerance.
// A result that is not finite (a division by an exact zero, a search that met b = +-1) is recomputed by the exact tier.
// Measured against the reference / the exact tier: tests/test_gpu_parity.py::test_fast_tier_sample_*.
DJB_DEV float mufu_lg2(float x)
{
	float y;
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}
// 1 / sqrt(x), MUFU.RSQ + one Newton step (relative error ~1e-7)
DJB_DEV float rsq_fast(float x)
{
	const float r = mufu_rsq(x);
	return __fmaf_rn(0.5f * r, __fmaf_rn(-(x * r), r, 1.0f), r);
}
// djb::erfinv (Giles), dj_brdf.h:691-721, with w = -ln((1 - u)(1 + u)) from MUFU.LG2.  (One Horner chain on selected
// coefficients instead of the two branches was measured: 22.7 -> 24.4 ms for Beckmann sampling; the branch stays.)
DJB_DEV float erfinv_fast(float u)
{
	float w = mufu_lg2((1.0f - u) * (1.0f + u)) * -0x1.62e430p-1f;
	float p;
	if (w < 5.0f) {
		w -= 2.5f;
		p = 2.81022636e-08f;
		p = __fmaf_rn(p, w, 3.43273939e-07f);
		p = __fmaf_rn(p, w, -3.5233877e-06f);
		p = __fmaf_rn(p, w, -4.39150654e-06f);
		p = __fmaf_rn(p, w, 0.00021858087f);
		p = __fmaf_rn(p, w, -0.00125372503f);
		p = __fmaf_rn(p, w, -0.00417768164f);
		p = __fmaf_rn(p, w, 0.246640727f);
		p = __fmaf_rn(p, w, 1.50140941f);
	} else {
		w = w * mufu_rsq(w) - 3.0f;
		p = -0.000200214257f;
		p = __fmaf_rn(p, w, 0.000100950558f);
		p = __fmaf_rn(p, w, 0.00134934322f);
		p = __fmaf_rn(p, w, -0.00367342844f);
		p = __fmaf_rn(p, w, 0.00573950773f);
		p = __fmaf_rn(p, w, -0.0076224613f);
		p = __fmaf_rn(p, w, 0.00943887047f);
		p = __fmaf_rn(p, w, 1.00167406f);
		p = __fmaf_rn(p, w, 2.83297682f);
	}
	return p * u;
}
// beckmann::qf2_radial, dj_brdf.h:1897-1952; ck > 0, sk >= 0
DJB_DEV float beckmann_qf2_fast(float u, float ck, float sk)
{
	const float SPI = 0x1.20dd76p-1f; // float(1 / sqrt(pi))
	const float cot = ck * mufu_rcp(sk), tan_k = sk * mufu_rcp(ck); // sk == 0: cot = +inf, handled by the clamps below
	const float xx = -cot * cot;
	const float e = exp_neg_fast(fmaxf(xx, -87.0f));
	float c = 1.0f; // djb::erf(cot): exactly 1 from cot = 3.92 on (exact tier, beckmann_qf2_lean)
	if (xx > -16.0f) {
		const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f;
		const float d = __fmaf_rn(0.3275911f, cot, 1.0f);
		float t = mufu_rcp(d);
		t = __fmaf_rn(t, __fmaf_rn(-d, t, 1.0f), t);
		float poly = __fmaf_rn(a5, t, a4);
		poly = __fmaf_rn(poly, t, a3);
		poly = __fmaf_rn(poly, t, a2);
		poly = __fmaf_rn(poly, t, a1);
		c = __fmaf_rn(-(poly * t), e, 1.0f);
	}
	const float B = SPI * tan_k;
	const float normalization = mufu_rcp(__fmaf_rn(B, e, 1.0f + c));
	float a = -1.0f;
	u = fmaxf(u, 1e-6f);
	const float fit = __fmaf_rn(ck, __fmaf_rn(ck, __fmaf_rn(-0.0594f, ck, 0.4265f), -0.876f), 1.0f);
	float b = __fmaf_rn(-(1.0f + c), mufu_ex2(fit * mufu_lg2(1.0f - u)), c);
	float ie = 0.0f;
	bool converged = false;
	int it = 0;
	while (++it < 10) {
		if (!(b >= a && b <= c)) b = 0.5f * (a + c);
		ie = erfinv_fast(b);
		const float ex = exp_neg_fast(-ie * ie);
		const float value = __fmaf_rn(normalization, __fmaf_rn(B, ex, 1.0f + b), -u);
		const float derivative = normalization * __fmaf_rn(-ie, tan_k, 1.0f);
		if (fabsf(value) < 1e-5f) { converged = true; break; }
		if (value > 0.0f) c = b; else a = b;
		b = __fmaf_rn(-value, mufu_rcp(derivative), b);
	}
	const float bm = fmaxf(-0.9999f, b);
	if (converged && bm == b) return ie;
	return erfinv_fast(bm);
}
// ggx::qf2_radial, dj_brdf.h:2089-2119: the reference's four tangent / cotangent forms are one function,
// -(sin cos_k + cos sin_k) / (cos cos_k - sin sin_k)
DJB_DEV float ggx_qf2_fast(float u, float ck, float sk)
{
	const float wh = 1.0f + ck, wl = ck - (wh - 1.0f);
	const float pr = u * wh, pe = __fmaf_rn(u, wh, -pr) + u * wl;
	const FF d = two_sum(pr, -1.0f);
	const float st = d.h + (d.l + pe); // float(u (1.0 + ck) - 1.0), as the exact tier forms it
	const float cc = st * st;
	const float vh = 1.0f - cc, vl = (1.0f - vh) - cc; // 1 - st^2 exactly (|st| < 1: u is clamped to [1e-5, 0.99999])
	const float y = mufu_rsq(vh);
	const float ct = __fmaf_rn(vl, 0.5f * y, vh * y);
	const float num = __fmaf_rn(st, ck, ct * sk), den = __fmaf_rn(ct, ck, -(st * sk));
	return -num * mufu_rcp(den);
}
template <int NDF>
static __device__ __noinline__ V3 lean_sample_redo(const float2 *T, GlfCtx GT, Params p, float u1, SampleU2 su2, V3 o)
{
	return lean_sample<NDF>(T, GT, p, u1, su2, o);
}
// microfacet::sample, dj_brdf.h:1669-1709; u1 clamped, su2 = lean_sample_u2(clamped u2)
template <int NDF>
DJB_DEV V3 fast_sample(const float2 *T, const GlfCtx &GT, const Params &p, float u1, SampleU2 su2, V3 o)
{
	const float oyay = o.y * p.ay;
	const float a = o.x * p.ax + oyay * p.rho;
	const float b = oyay * p.srho;
	const float c = o.z - o.x * p.tx - o.y * p.ty;
	const V3 os = normalize(mk(a, b, c)); // exact: the gate and (cos, sin) of the view angle
	if (!(os.z > 0.0f)) return mk(0.f, 0.f, 1.f);
	const float ck = os.z;
	const float sk = ck < 1.0f ? sqrt_1m_sq(ck) : 0.0f;
	float tx, ty;
	if (NDF == NDF_GGX) {
		tx = ggx_qf2_fast(u1, ck, sk);
		const float v = __fmaf_rn(tx, tx, 1.0f);
		ty = su2.a * (v * mufu_rsq(v)) * su2.b; // S sqrt(1 + tx^2) (pn / qn)
	} else {
		tx = beckmann_qf2_fast(u1, ck, sk);
		ty = su2.a;
	}
	float xs = tx, ys = ty;
	if (sk != 0.0f) {
		const float nrm = mufu_rsq(__fmaf_rn(os.x, os.x, os.y * os.y));
		const float cp = os.x * nrm, sp = os.y * nrm;
		xs = __fmaf_rn(cp, tx, -(sp * ty));
		ys = __fmaf_rn(sp, tx, cp * ty);
	}
	const float txh = __fmaf_rn(p.ax, xs, p.tx);
	const float tyh = __fmaf_rn(p.ay, __fmaf_rn(p.rho, xs, p.srho * ys), p.ty);
	const float r = rsq_fast(__fmaf_rn(txh, txh, __fmaf_rn(tyh, tyh, 1.0f)));
	const V3 h = mk(-txh * r, -tyh * r, r);
	const float k = 2.0f * __fmaf_rn(o.x, h.x, __fmaf_rn(o.y, h.y, o.z * h.z));
	const V3 i = mk(__fmaf_rn(k, h.x, -o.x), __fmaf_rn(k, h.y, -o.y), __fmaf_rn(k, h.z, -o.z));
	if (!(fabsf(i.x) + fabsf(i.y) + fabsf(i.z) < 1e30f)) return lean_sample_redo<NDF>(T, GT, p, u1, su2, o);
	return i;
}

} // namespace djb200
