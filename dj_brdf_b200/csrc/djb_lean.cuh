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
	bool facing;        // h.z > 1e-4: the NDF is non-zero (dj_brdf.h:1561)
	bool den_ok;        // den, c4 and their reciprocals are normal numbers: the lean divisions are exact
};

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
	c.cd = sat_ref(dot(o, c.h));
	const float lo = 1e-28f, hi = 1e30f;
	c.den_ok = fabsf(c.den) > lo && fabsf(c.den) < hi && c.c4 > lo;
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

// microfacet::ndf + p22, dj_brdf.h:1559-1587, with the per-pair and per-material reciprocals
template <int NDF>
DJB_DEV float lean_ndf(const float2 *T, const ParamsX &m, const PairX &c)
{
	if (!c.facing) return 0.0f;
	float x = c.sx - m.p.tx, y = c.sy - m.p.ty;
	float xs = div_by(x, m.p.ax, m.rcp_ax);
	float t1 = m.p.ax * y - m.rho_ay * x;
	float ys = div_by(t1, m.nrm, m.rcp_nrm);
	float r2 = xs * xs + ys * ys;
	float pv = NDF == NDF_BECKMANN ? beck_p22_lean(T, r2) : p22_radial<NDF_GGX>(r2);
	if (NDF == NDF_BECKMANN) { // the exponential underflows gradually: small numerators take the lifted division
		if (pv == 0.0f) return 0.0f;
		return div_by_small(div_by_small(pv, m.nrm, m.rcp_nrm), c.c4, c.rcp_c4);
	}
	return div_by(div_by(pv, m.nrm, m.rcp_nrm), c.c4, c.rcp_c4);
}

// F D G / (4 o.z) (evalp, dj_brdf.h:1529-1547); `scale` = 1 / i.z for eval (dj_brdf.h:1551-1555), unused otherwise
template <int NDF, int FK, int OP>
DJB_DEV V3 lean_evalp(const float2 *T, const ParamsX &m, const FresnelDev &f, bool shadow, const PairX &c)
{
	const V3 z = mk(0.f, 0.f, 0.f);
	const V3 zero = OP == OP_EVAL ? scale(c.inv_iz, z) : z; // 0 * (1 / i.z): keeps the reference's -0 / NaN for i.z <= 0
	const float Dn = lean_ndf<NDF>(T, m, c);
	// D == 0 (h below the 1e-4 gate, or Beckmann's exponential underflowed): F D G / (4 o.z) is +0 for o.z > 0
	// whatever G is, so the two projected-area evaluations are skipped
	if (Dn == 0.0f && c.den > 0.0f) return zero;
	const float G = lean_gaf<NDF>(T, m.p, shadow, c.i, c.o);
	if (G > 0.0f) {
		const float num = Dn * G;
		float k;
		if (!c.den_ok) k = __fdiv_rn(num, c.den);
		else k = NDF == NDF_GGX ? div_by(num, c.den, c.rcp_den) : div_by_small(num, c.den, c.rcp_den);
		V3 e = scale(k, fresnel_eval<FK>(f, c.cd));
		return OP == OP_EVAL ? scale(c.inv_iz, e) : e;
	}
	return zero;
}

// microfacet::pdf, dj_brdf.h:1713-1730 with vndf, dj_brdf.h:1602-1615.  sigma(o) appears in G1(o) and in the
// visible-normal density: evaluated once.
template <int NDF>
DJB_DEV float lean_pdf(const float2 *T, const ParamsX &m, bool shadow, const PairX &c)
{
	const float Dn = lean_ndf<NDF>(T, m, c);
	if (Dn == 0.0f && c.den > 0.0f) return 0.0f; // vndf == 0: the pdf is +0 whatever G is
	const Params &p = m.p;
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
		const float kh = dot(c.o, c.h);
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

} // namespace djb200
