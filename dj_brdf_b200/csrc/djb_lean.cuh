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
	const float lo = 1e-30f, hi = 1e30f;
	c.den_ok = fabsf(c.den) > lo && fabsf(c.den) < hi && c.c4 > lo;
	return c;
}

// microfacet::sigma, dj_brdf.h:1619-1631
template <int NDF>
DJB_DEV float lean_sigma(const Params &p, V3 k)
{
	float kyay = k.y * p.ay;
	float a = k.x * p.ax + kyay * p.rho;
	float b = kyay * p.srho;
	float c = k.z - k.x * p.tx - k.y * p.ty;
	float nrm = sqrt_lean(a * a + b * b + c * c);
	float cz = rcp_lean(nrm) * c;
	return nrm * sigma_std_radial<NDF>(cz);
}

// microfacet::g1, dj_brdf.h:1633-1642
template <int NDF>
DJB_DEV float lean_g1(const Params &p, V3 k)
{
	if (dot(k, mk(p.nx, p.ny, p.nz)) > 0.0f) return div_lean(k.z, lean_sigma<NDF>(p, k));
	return 0.0f;
}

// microfacet::gaf, dj_brdf.h:1644-1665
template <int NDF>
DJB_DEV float lean_gaf(const Params &p, bool shadow, V3 i, V3 o)
{
	float g1o = lean_g1<NDF>(p, o);
	if (shadow) {
		float g1i = lean_g1<NDF>(p, i);
		float t = g1i * g1o;
		if (t > 0.0f) return div_lean(t, g1i + g1o - t);
		return 0.0f;
	}
	return g1o;
}

// microfacet::ndf + p22, dj_brdf.h:1559-1587, with the per-pair and per-material reciprocals
template <int NDF>
DJB_DEV float lean_ndf(const ParamsX &m, const PairX &c)
{
	if (!c.facing) return 0.0f;
	float x = c.sx - m.p.tx, y = c.sy - m.p.ty;
	float xs = div_by(x, m.p.ax, m.rcp_ax);
	float t1 = m.p.ax * y - m.rho_ay * x;
	float ys = div_by(t1, m.nrm, m.rcp_nrm);
	float pv = p22_radial<NDF>(xs * xs + ys * ys);
	// Beckmann's exp underflows gradually: below the normal range the FMA quotients would round twice
	if (NDF == NDF_BECKMANN && !(pv > 1e-30f)) return __fdiv_rn(__fdiv_rn(pv, m.nrm), c.c4);
	return div_by(div_by(pv, m.nrm, m.rcp_nrm), c.c4, c.rcp_c4);
}

// F D G / (4 o.z) (evalp, dj_brdf.h:1529-1547); `scale` = 1 / i.z for eval (dj_brdf.h:1551-1555), unused otherwise
template <int NDF, int FK, int OP>
DJB_DEV V3 lean_evalp(const ParamsX &m, const FresnelDev &f, bool shadow, const PairX &c)
{
	float G = lean_gaf<NDF>(m.p, shadow, c.i, c.o);
	if (G > 0.0f) {
		float Dn = lean_ndf<NDF>(m, c);
		float num = Dn * G;
		float k = (c.den_ok && (NDF == NDF_GGX || num > 1e-30f)) ? div_by(num, c.den, c.rcp_den) : __fdiv_rn(num, c.den);
		V3 e = scale(k, fresnel_eval<FK>(f, c.cd));
		return OP == OP_EVAL ? scale(c.inv_iz, e) : e;
	}
	V3 z = mk(0.f, 0.f, 0.f);
	return OP == OP_EVAL ? scale(c.inv_iz, z) : z; // 0 * (1 / i.z): keeps the reference's -0 / NaN for i.z <= 0
}

// microfacet::pdf, dj_brdf.h:1713-1730 with vndf, dj_brdf.h:1602-1615
template <int NDF>
DJB_DEV float lean_pdf(const ParamsX &m, bool shadow, const PairX &c)
{
	float G = lean_gaf<NDF>(m.p, shadow, c.i, c.o);
	if (G > 0.0f) {
		float kh = dot(c.o, c.h);
		float v = 0.0f;
		if (kh > 0.0f) {
			float num = kh * lean_ndf<NDF>(m, c), sg = lean_sigma<NDF>(m.p, c.o);
			v = (NDF == NDF_GGX || num > 1e-30f) ? div_lean(num, sg) : __fdiv_rn(num, sg);
		}
		return (c.den_ok && (NDF == NDF_GGX || v > 1e-30f)) ? div_by(v, c.den, c.rcp_den) : __fdiv_rn(v, c.den);
	}
	return 0.0f;
}

} // namespace djb200
