// djb_device.cuh -- device-side numerics of the microfacet / MERL / LEAN paths (sm_100a).
//
// Parity contract (DESIGN.md "Numerics"): results must match the reference header compiled as a
// pinned scalar CPU program (float_t = float, unqualified libm calls bind to the C double
// functions, no FMA contraction).  So every function below keeps the reference's rounding points:
// float storage, double sub-expressions exactly where the reference has a double literal, M_PI or
// a libm call, IEEE division and square root.  This translation unit MUST be compiled with
// -fmad=false (the build enforces it) and without -use_fast_math; CUDA's double sqrt and
// division are IEEE-correct, double exp/acos/atan2/sin/cos are within 1-2 ulp of glibc's, which
// survives the rounding back to float except for ~1e-8 of inputs.
//
// "Same rounding point" does not mean "same instruction".  Where the reference's double detour is ONE IEEE
// operation on float operands -- float(sqrt(double(x))), float(double(a) / double(b)), float(1.0 + double(c)),
// float(1.0 / double(b)) -- the float operation rounds identically (double has >= 2 * 24 + 2 digits, so the
// double rounding is innocuous), and the hot functions below use __fsqrt_rn / __fdiv_rn / __frcp_rn / float
// add directly: no conversions through the quarter-rate XU pipe, no FP64 divide.  Short multi-operation detours
// (1 / sqrt, 1 / (pi t^2)) use a correctly rounded float primitive or a float-float sequence that reproduces
// the rounded float except at double-rounding ties (~1e-8 of inputs, 1 ulp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "djb_glibcf.h"
#include "djb_dmath.cuh"

namespace djb200 {

#define DJB_DEV __device__ __forceinline__
#define DJB_PI 3.14159265358979323846

struct V3 { float x, y, z; };

// device copy of djb200_params / djb::microfacet::params (dj_brdf.h:238-242)
struct Params {
	float nx, ny, nz;
	float a1, a2, phi_a;
	float ax, ay;
	float rho, srho;
	float tx, ty;
};
static_assert(sizeof(Params) == 48, "params block must stay 48 bytes");

enum { NDF_BECKMANN = 0, NDF_GGX = 1 };
enum { FK_IDEAL = 0, FK_SCHLICK = 1, FK_UNPOLARIZED = 2, FK_SGD = 3, FK_SPLINE = 4 };

struct FresnelDev {
	float v[6];
	const float *pts; // device (or shared) pointer to n*3 floats
	int npts;
};

// ---- vec3 algebra, dj_brdf.h:597-637 ---------------------------------------------------------
DJB_DEV V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
DJB_DEV V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
DJB_DEV V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
DJB_DEV V3 scale(float k, V3 a) { return mk(k * a.x, k * a.y, k * a.z); }
DJB_DEV float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
DJB_DEV V3 cross(V3 a, V3 b)
{
	return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// (1.0 / b) rounded to float: the scalar of `vec3 / float_t` (:601)
DJB_DEV float rcp_via_double(float b) { return __frcp_rn(b); } // == (float)(1.0 / (double)b): one IEEE op
// inversesqrt(): 1.0 / sqrt(x) in double, rounded once to float (:612-616)
// two operations in double: equals the correctly rounded float rsqrt except at ties of the double rounding
// (0 differences in 1.5e9 half vectors, djb200_debug_merl_filter_stats)
DJB_DEV float inv_sqrt(float x) { return __frsqrt_rn(x); }
DJB_DEV V3 normalize(V3 v) { return scale(inv_sqrt(dot(v, v)), v); }
// djb::min/max/sat templates (:574-576), including their NaN pass-through
DJB_DEV float fmin_ref(float a, float b) { return a < b ? a : b; }
DJB_DEV float fmax_ref(float a, float b) { return a > b ? a : b; }
DJB_DEV float sat_ref(float x) { return fmin_ref(1.0f, fmax_ref(0.0f, x)); }

// vec3(theta, phi), :589-595
DJB_DEV V3 spherical(float theta, float phi)
{
	double st, ct, sp, cp; // djb_dmath.cuh: (float)sin / (float)cos equal libm's for every float up to 1e5 but one
	sincos_d((double)theta, &st, &ct);
	sincos_d((double)phi, &sp, &cp);
	float s = (float)st;
	return mk((float)((double)s * cp), (float)((double)s * sp), (float)ct);
}

// xyz_to_theta_phi, :650-661
DJB_DEV void to_theta_phi(V3 p, float &theta, float &phi)
{
	double z = (double)p.z;
	if (z > 0.99999) {
		theta = 0.0f;
		phi = 0.0f;
	} else if (z < -0.99999) {
		theta = (float)DJB_PI;
		phi = 0.0f;
	} else {
		theta = (float)acos_d(z);
		phi = (float)atan2_t((double)p.y, (double)p.x, g_dm_table_dev);
	}
}

// ---- special functions ------------------------------------------------------------------------
// djb::erf (A&S 7.1.26, :667-688).  `e` must be exp((double)(-x*x)), shared with the caller.
DJB_DEV float erf_as(float x, double e)
{
	const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f;
	const float a4 = -1.453152027f, a5 = 1.061405429f, p = 0.3275911f;
	float sgn = (x < 0.0f) ? -1.0f : 1.0f;
	x = fabsf(x);
	float t = (float)(1.0 / (1.0 + (double)(p * x)));
	float poly = ((((a5 * t + a4) * t) + a3) * t + a2) * t + a1;
	float y = (float)(1.0 - (double)(poly * t) * e);
	return sgn * y;
}
DJB_DEV float erf_as(float x) { return erf_as(x, exp((double)(-x * x))); }

// single-precision libm calls of the reference (logf/expf/powf, :695, 1917, 1935): glibc's own algorithms, operation
// for operation (djb_glibcf.h), so the results are glibc's results; arguments outside the restated main branches (zero,
// subnormal, negative, NaN, overflowing) are evaluated in double and rounded once.
static __device__ __noinline__ float logf_literal(float x) { return (float)log((double)x); }
static __device__ __noinline__ float expf_literal(float x) { return (float)exp((double)x); }
static __device__ __noinline__ float powf_literal(float x, float y) { return (float)pow((double)x, (double)y); }
DJB_DEV float logf_cr(float x) { return glf_logf_ok(x) ? glf_logf(GlfTablePtr{g_glf_table}, glf_hot(), x) : logf_literal(x); }
DJB_DEV float expf_cr(float x) { return glf_expf_ok(x) ? glf_expf(GlfTablePtr{g_glf_table}, glf_hot(), x) : expf_literal(x); }
DJB_DEV float powf_cr(float x, float y)
{
	if (glf_powf_ok(x, y)) {
		bool ok;
		const float r = glf_powf(GlfTablePtr{g_glf_table}, x, y, ok);
		if (ok) return r;
	}
	return powf_literal(x, y);
}

// djb::erfinv (Giles), :691-721
DJB_DEV float erfinv_giles(float u)
{
	float w = -logf_cr((1.0f - u) * (1.0f + u)), p;
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
		w = (float)(sqrt((double)w) - 3.0);
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

// ---- rotations and the half / difference frame, :754-793 -------------------------------------
DJB_DEV V3 rotate_about(V3 x, V3 axis, float angle)
{
	double sd, cd;
	sincos_d((double)angle, &sd, &cd);
	float c = (float)cd, s = (float)sd;
	V3 out = scale(c, x);
	float t1 = dot(axis, x);
	float t2 = (float)((double)t1 * (1.0 - (double)c));
	out = out + scale(t2, axis);
	out = out + scale(s, cross(axis, x));
	return out;
}

DJB_DEV void io_to_hd(V3 i, V3 o, V3 &h, V3 &d, float &theta_h)
{
	float ph;
	h = normalize(i + o);
	to_theta_phi(h, theta_h, ph);
	V3 tmp = rotate_about(i, mk(0.f, 0.f, 1.f), -ph);
	d = normalize(rotate_about(tmp, mk(0.f, 1.f, 0.f), -theta_h));
}

DJB_DEV void hd_to_io(V3 h, V3 d, V3 &i, V3 &o)
{
	float th, ph;
	to_theta_phi(h, th, ph);
	V3 tmp = rotate_about(d, mk(0.f, 1.f, 0.f), th);
	i = normalize(rotate_about(tmp, mk(0.f, 0.f, 1.f), ph));
	float k = (float)(2.0 * (double)dot(i, h));
	o = normalize(scale(k, h) - i);
}

// ---- Fresnel, :1292-1344 -----------------------------------------------------------------------
DJB_DEV float unpolarized_channel(float c, float n)
{
	float g = (float)sqrt_d((double)(n * n + c * c) - 1.0);
	float t1 = (float)((double)(c * (g + c)) - 1.0);
	float t2 = (float)((double)(c * (g - c)) + 1.0);
	float t3 = (t1 * t1) / (t2 * t2);
	float t4 = ((g - c) * (g - c)) / ((g + c) * (g + c));
	return (float)((0.5 * (double)t4) * (1.0 + (double)t3));
}

// spline::eval<vec3> with uwrap_edge, :1191-1218
DJB_DEV V3 spline_rgb(const float *pts, int n, float u)
{
	float x = u * (float)n - u;
	float ip = truncf(x); // modf(): integral part towards zero, fraction keeps the sign
	float frac = x - ip;
	int i1 = (int)ip, i2 = (int)ip + 1;
	i1 = i1 >= n ? n - 1 : (i1 < 0 ? 0 : i1);
	i2 = i2 >= n ? n - 1 : (i2 < 0 ? 0 : i2);
	V3 p1 = mk(pts[3 * i1], pts[3 * i1 + 1], pts[3 * i1 + 2]);
	V3 p2 = mk(pts[3 * i2], pts[3 * i2 + 1], pts[3 * i2 + 2]);
	return p1 + scale(frac, p2 - p1);
}

// spline::eval<float_t> with uwrap_edge (tabulated p22 / sigma / cdf / qf), :1207-1218
DJB_DEV float spline_f(const float *pts, int n, float u)
{
	float x = u * (float)n - u;
	float ip = truncf(x);
	float frac = x - ip;
	int i1 = (int)ip, i2 = (int)ip + 1;
	i1 = i1 >= n ? n - 1 : (i1 < 0 ? 0 : i1);
	i2 = i2 >= n ? n - 1 : (i2 < 0 ? 0 : i2);
	float p1 = pts[i1], p2 = pts[i2];
	return p1 + frac * (p2 - p1);
}

template <int FK>
DJB_DEV V3 fresnel_eval(const FresnelDev &f, float c)
{
	if (FK == FK_SCHLICK) {
		float c1 = 1.0f - c, c2 = c1 * c1, c5 = c2 * c2 * c1; // float(1.0 - c): one rounding either way
		return mk(f.v[0] + c5 * (1.0f - f.v[0]), f.v[1] + c5 * (1.0f - f.v[1]), f.v[2] + c5 * (1.0f - f.v[2]));
	} else if (FK == FK_UNPOLARIZED) {
		return mk(unpolarized_channel(c, f.v[0]), unpolarized_channel(c, f.v[1]), unpolarized_channel(c, f.v[2]));
	} else if (FK == FK_SGD) {
		float pw = pow5_one_minus(c); // (float)pow(1.0 - (double)c, 5.0) for every float c (tests/cpp/dmath_check.cpp)
		return mk((f.v[0] - c * f.v[3]) + pw * (1.0f - f.v[0]), (f.v[1] - c * f.v[4]) + pw * (1.0f - f.v[1]),
		          (f.v[2] - c * f.v[5]) + pw * (1.0f - f.v[2]));
	} else if (FK == FK_SPLINE) {
		float u = acos_coord_pi(c); // (float)(2.0 * acos((double)c) / M_PI) for every float c (tests/cpp/dmath_check.cpp)
		return spline_rgb(f.pts, f.npts, u);
	}
	return mk(1.0f, 1.0f, 1.0f);
}

// ---- standard radial distributions ------------------------------------------------------------
template <int NDF>
DJB_DEV float p22_radial(float r2)
{
	if (NDF == NDF_GGX) { // :2056-2060: t = float(1.0 + r2); float(1.0 / (M_PI * t * t)) with the products in double
		float t = 1.0f + r2;
		// float-float: D = pi * t^2 (t^2 exact as q + ql, pi = PI_H + PI_L + 2^-48 remainder), then one Newton
		// step on the correctly rounded float reciprocal of its head gives 1 / D to ~2^-47 before the final rounding
		const float PI_H = 3.14159274101257324f, PI_L = -8.74227765734758578e-08f;
		float q = t * t, ql = __fmaf_rn(t, t, -q);
		float Dh = PI_H * q;
		float De = __fmaf_rn(PI_H, q, -Dh) + PI_H * ql + PI_L * q;
		float y0 = __frcp_rn(Dh);
		float r = __fmaf_rn(-Dh, y0, 1.0f) - De * y0;
		return __fmaf_rn(y0, r, y0);
	}
	return (float)(exp((double)(-r2)) / DJB_PI); // :1866-1869
}

template <int NDF>
DJB_DEV float sigma_std_radial(float c)
{
	if (NDF == NDF_GGX) return (1.0f + c) * 0.5f; // :2062-2065, float((1.0 + c) / 2.0): one rounding either way
	// beckmann, :1871-1879
	if (c == 1.0f) return 1.0f;
	float s = (float)sqrt_d(1.0 - (double)(c * c));
	float nu = c / s;
	double e = exp((double)(-nu * nu)); // also the exp() inside djb::erf: (-|nu|)*|nu| == (-nu)*nu
	float tmp = (float)(e * (double)inv_sqrt((float)DJB_PI));
	return (float)(((double)c * (1.0 + (double)erf_as(nu, e)) + (double)(s * tmp)) / 2.0);
}

// microfacet::sigma, :1619-1631 (only the z of the normalised warped direction is consumed)
template <int NDF>
DJB_DEV float mf_sigma(const Params &p, V3 k)
{
	float a = k.x * p.ax + k.y * p.ay * p.rho;
	float b = k.y * p.ay * p.srho;
	float c = k.z - k.x * p.tx - k.y * p.ty;
	float nrm = __fsqrt_rn(a * a + b * b + c * c); // == (float)sqrt((double)(...)): one IEEE op
	float cz = rcp_via_double(nrm) * c;
	return nrm * sigma_std_radial<NDF>(cz);
}

// :1633-1642
template <int NDF>
DJB_DEV float mf_g1(const Params &p, V3 k)
{
	float test = dot(k, mk(p.nx, p.ny, p.nz));
	if (test > 0.0f) return k.z / mf_sigma<NDF>(p, k);
	return 0.0f;
}

// :1644-1665
template <int NDF>
DJB_DEV float mf_gaf(const Params &p, bool shadow, V3 i, V3 o)
{
	float g1o = mf_g1<NDF>(p, o);
	if (shadow) {
		float g1i = mf_g1<NDF>(p, i);
		float t = g1i * g1o;
		if (t > 0.0f) return t / (g1i + g1o - t);
		return 0.0f;
	}
	return g1o;
}

// :1574-1587
template <int NDF>
DJB_DEV float mf_p22(const Params &p, float x, float y)
{
	x -= p.tx;
	y -= p.ty;
	float nrm = p.ax * p.ay * p.srho;
	float xs = x / p.ax;
	float t1 = p.ax * y - p.rho * p.ay * x;
	float ys = t1 / nrm; // tmp2 of the reference is the same product as nrm
	return p22_radial<NDF>(xs * xs + ys * ys) / nrm;
}

// :1559-1570
template <int NDF>
DJB_DEV float mf_ndf(const Params &p, V3 h)
{
	if (h.z > 1e-4f) {
		float c2 = h.z * h.z, c4 = c2 * c2;
		float sx = -h.x / h.z, sy = -h.y / h.z;
		return mf_p22<NDF>(p, sx, sy) / c4;
	}
	return 0.0f;
}

// :1602-1615
template <int NDF>
DJB_DEV float mf_vndf(const Params &p, V3 h, V3 k)
{
	float kh = dot(k, h);
	if (kh > 0.0f) return kh * mf_ndf<NDF>(p, h) / mf_sigma<NDF>(p, k);
	return 0.0f;
}

// microfacet::evalp, :1529-1547 -- h = normalize(i + o) is supplied by the caller because it does
// not depend on the params block
template <int NDF, int FK>
DJB_DEV V3 mf_evalp(const Params &p, const FresnelDev &f, bool shadow, V3 i, V3 o, V3 h)
{
	float G = mf_gaf<NDF>(p, shadow, i, o);
	if (G > 0.0f) {
		float cd = sat_ref(dot(o, h));
		V3 Fr = fresnel_eval<FK>(f, cd);
		float Dn = mf_ndf<NDF>(p, h);
		return scale(__fdiv_rn(Dn * G, 4.0f * o.z), Fr); // 4 * o.z is exact; the quotient of two floats rounds once
	}
	return mk(0.f, 0.f, 0.f);
}

// :1713-1730 (beckmann and ggx both take the Smith-VNDF branch)
template <int NDF>
DJB_DEV float mf_pdf(const Params &p, bool shadow, V3 i, V3 o, V3 h)
{
	float G = mf_gaf<NDF>(p, shadow, i, o);
	if (G > 0.0f) return __fdiv_rn(mf_vndf<NDF>(p, h, o), 4.0f * dot(i, h));
	return 0.0f;
}

// ---- visible-normal sampling -------------------------------------------------------------------
// ggx::qf2_radial, :2089-2119
DJB_DEV float ggx_qf2(float u, float ck, float sk)
{
	float st = (float)((double)u * (1.0 + (double)ck) - 1.0);
	float ct = (float)sqrt_d(1.0 - (double)(st * st));
	if ((double)ct > 0.707107) {
		float tt = st / ct;
		if ((double)sk < 0.707107) {
			float tk = sk / ck;
			return (float)((double)(-(tt + tk)) / (1.0 - (double)(tt * tk)));
		} else {
			float kk = ck / sk;
			return (float)((1.0 + (double)(tt * kk)) / (double)(tt - kk));
		}
	} else {
		float cot = ct / st;
		if ((double)sk < 0.707107) {
			float tk = sk / ck;
			return (float)((1.0 + (double)(tk * cot)) / (double)(tk - cot));
		} else {
			float kk = ck / sk;
			return (float)((double)(cot + kk) / (1.0 - (double)(cot * kk)));
		}
	}
}

// ggx::qf3_radial + qf3_rational_approx, :2121-2146
DJB_DEV float ggx_qf3(float u, float qf2)
{
	float alpha = (float)sqrt_d(1.0 + (double)(qf2 * qf2));
	float S;
	if ((double)u < 0.5) {
		u = (float)(2.0 * (0.5 - (double)u));
		S = -1.0f;
	} else {
		u = (float)(2.0 * ((double)u - 0.5));
		S = 1.0f;
	}
	double du = (double)u;
	float pn = (float)(du * (du * (du * (-0.365728915865723) + 0.790235037209296) - 0.424965825137544)
	                   + 0.000152998850436920);
	float qn = (float)(du * (du * (du * (du * 0.169507819808272 - 0.397203533833404) - 0.232500544458471) + 1.0)
	                   - 0.539825872510702);
	return S * alpha * (pn / qn);
}

// beckmann::qf2_radial, :1897-1952
DJB_DEV float beckmann_qf2(float u, float ck, float sk)
{
	const float sqrt_pi_inv = (float)(1.0 / sqrt(DJB_PI));
	float cot = ck / sk, tan_k = sk / ck;
	float a = -1.0f, c = erf_as(cot);
	u = fmax_ref(u, 1e-6f);
	float fit = 1.0f + ck * (-0.876f + ck * (0.4265f - 0.0594f * ck));
	float b = c - (1.0f + c) * powf_cr(1.0f - u, fit);
	float normalization =
		(float)(1.0 / ((double)(1.0f + c) + (double)(sqrt_pi_inv * tan_k) * exp((double)(-cot * cot))));
	int it = 0;
	while (++it < 10) {
		if (!(b >= a && b <= c)) b = 0.5f * (a + c);
		float ie = erfinv_giles(b);
		float value = normalization * (1.0f + b + sqrt_pi_inv * tan_k * expf_cr(-ie * ie)) - u;
		float derivative = normalization * (1.0f - ie * tan_k);
		if (fabsf(value) < 1e-5f) break;
		if (value > 0.0f) c = b; else a = b;
		b -= value / derivative;
	}
	return erfinv_giles(fmax_ref(-0.9999f, b));
}

// radial::sample_vp22_std_smith, :1818-1846
template <int NDF>
DJB_DEV void sample_std_slopes(float u1, float u2, V3 k, float &xs, float &ys)
{
	float ck = k.z;
	float sk = k.z < 1.0f ? (float)sqrt_d(1.0 - (double)(k.z * k.z)) : 0.0f;
	float tx, ty;
	if (NDF == NDF_GGX) {
		tx = ggx_qf2(u1, ck, sk);
		ty = ggx_qf3(u2, tx);
	} else {
		tx = beckmann_qf2(u1, ck, sk);
		ty = erfinv_giles((float)(2.0 * (double)u2 - 1.0)); // qf3_radial -> qf1, :1891-1894
	}
	if (sk == 0.0f) {
		xs = tx;
		ys = ty;
	} else {
		float nrm = inv_sqrt(k.x * k.x + k.y * k.y);
		float cp = k.x * nrm, sp = k.y * nrm;
		xs = cp * tx - sp * ty;
		ys = sp * tx + cp * ty;
	}
}

// microfacet::sample, :1669-1709
template <int NDF>
DJB_DEV V3 mf_sample(const Params &p, float u1, float u2, V3 o)
{
	u1 = sat_ref(u1) * 0.99998f + 0.00001f;
	u2 = sat_ref(u2) * 0.99998f + 0.00001f;
	float a = o.x * p.ax + o.y * p.ay * p.rho;
	float b = o.y * p.ay * p.srho;
	float c = o.z - o.x * p.tx - o.y * p.ty;
	V3 os = normalize(mk(a, b, c));
	if (os.z > 0.0f) {
		float txm, tym;
		sample_std_slopes<NDF>(u1, u2, os, txm, tym);
		float txh = p.ax * txm + p.tx;
		float chol = p.rho * txm + p.srho * tym;
		float tyh = p.ay * chol + p.ty;
		V3 h = normalize(mk(-txh, -tyh, 1.0f));
		float k = (float)(2.0 * (double)dot(o, h));
		return scale(k, h) - o;
	}
	return mk(0.f, 0.f, 1.f);
}

// microfacet::evalp_is, :1734-1765
template <int NDF, int FK>
DJB_DEV V3 mf_evalp_is(const Params &p, const FresnelDev &f, bool shadow, float u1, float u2, V3 o,
                       V3 &i_out, float &pdf_out)
{
	V3 i = mf_sample<NDF>(p, u1, u2, o);
	V3 h = normalize(i + o);
	float G = mf_gaf<NDF>(p, shadow, i, o);
	pdf_out = 0.0f;
	i_out = mk(0.f, 0.f, 0.f);
	if (G > 0.0f) {
		float cd = sat_ref(dot(o, h));
		i_out = i;
		V3 Fr = fresnel_eval<FK>(f, cd);
		float g1 = mf_g1<NDF>(p, o);
		pdf_out = __fdiv_rn(mf_vndf<NDF>(p, h, o), 4.0f * cd);
		return scale(G / g1, Fr);
	}
	return mk(0.f, 0.f, 0.f);
}

// ---- params construction on the device (E10 / L2), dj_brdf.h:1378-1393, 1437-1474 ---------------
// (float)sqrt(0.5 * (double)x) for a float x: 0.5 x is a float unless x is subnormal-small, and the double rounding of a square
// root is innocuous (djb_device.cuh header)
DJB_DEV float sqrt_half(float x) { return fabsf(x) >= 1e-30f ? __fsqrt_rn(0.5f * x) : (float)sqrt(0.5 * (double)x); }
DJB_DEV void params_from_pdf(float ax, float ay, float rho, float tx, float ty, Params &p)
{
	p.ax = ax;
	p.ay = ay;
	p.rho = rho;
	p.srho = (float)sqrt_d(1.0 - (double)(rho * rho));
	float qx = ax * ax, qy = ay * ay;
	float cov = (float)((double)(rho * ax * ay) * 2.0);
	float t1 = qx + qy, t2 = qx - qy;
	float t3 = __fsqrt_rn(t2 * t2 + cov * cov); // == (float)sqrt((double)(...)): the argument is a float
	p.a1 = sqrt_half(t1 + t3);
	p.a2 = sqrt_half(t1 - t3);
	p.phi_a = (cov != 0.0f) ? (float)atan_t((double)((qx - qy - t3) / cov), g_dm_table_dev) : 0.0f;
	p.tx = tx;
	p.ty = ty;
	V3 n = normalize(mk(-tx, -ty, 1.0f));
	p.nx = n.x;
	p.ny = n.y;
	p.nz = n.z;
}

// params::elliptic on the device, dj_brdf.h:1355-1376, 1422-1426, 1451-1459 (same rounding points as capi.cu's elliptic_h)
DJB_DEV void params_elliptic_dev(float a1, float a2, float phi, Params &p)
{
	double sd, cd;
	sincos((double)phi, &sd, &cd); // the library's: sincos_d here costs the fused LEAN shading kernel 8 registers and 6 % (measured)
	const float c = (float)cd, s = (float)sd;
	const float c2 = (float)(2.0 * (double)c * (double)c - 1.0);
	const float q1 = a1 * a1, q2 = a2 * a2, t1 = q1 + q2, t2 = q1 - q2;
	p.a1 = a1;
	p.a2 = a2;
	p.phi_a = phi;
	p.ax = sqrt_half(t1 + t2 * c2);
	p.ay = sqrt_half(t1 - t2 * c2);
	p.rho = (q2 - q1) * c * s / (p.ax * p.ay);
	p.srho = (float)sqrt_d(1.0 - (double)(p.rho * p.rho));
	p.tx = 0.0f;
	p.ty = 0.0f;
	const V3 n = normalize(mk(-0.0f, -0.0f, 1.0f));
	p.nx = n.x;
	p.ny = n.y;
	p.nz = n.z;
}

// beckmann::lrep_to_params, :1976-1990
DJB_DEV void lrep_to_params(float E1, float E2, float E3, float E4, float E5, Params &p)
{
	float t1 = fmax_ref(0.0f, E3 - E1 * E1);
	float t2 = fmax_ref(0.0f, E4 - E2 * E2);
	double sx = sqrt_d(2.0 * (double)t1), sy = sqrt_d(2.0 * (double)t2); // t >= 0; 0 and the tiny / huge ones take the library
	float ax = (float)(1e-5 > sx ? 1e-5 : sx);
	float ay = (float)(1e-5 > sy ? 1e-5 : sy);
	float rho = 2.0f * (E5 - E1 * E2) / (ax * ay);
	rho = fmin_ref(0.99f, fmax_ref(-0.99f, rho));
	params_from_pdf(ax, ay, rho, E1, E2, p);
}

// The per-shading-point parameter construction of the LEAN-filtering plugin (mitsuba/dj_beckmannconductor.cpp:283-314):
// base roughness ellipse -> params -> lrep; LEAN texel (E1..E5, biased) -> lrep, scaled by dmapscale; sum; -> params.
struct LeanShadingCfg {
	float bias, dmap_scale;
	int lean_filtering;
};
DJB_DEV void lean_shading_params(const LeanShadingCfg &c, float a1, float a2, float phi, float E1, float E2, float E3, float E4,
                                 float E5, Params &out)
{
	Params base;
	params_elliptic_dev(a1, a2, phi, base);
	E1 -= c.bias;
	E2 -= c.bias;
	E5 -= c.bias * c.bias;
	if (!c.lean_filtering) { // naive MIP mapping: second moments rebuilt from the filtered means
		E3 = E1 * E1;
		E4 = E2 * E2;
		E5 = E1 * E2;
	}
	const float sc = c.dmap_scale, sc2 = sc * sc; // lrep::operator*=, dj_brdf.h:2020-2031
	E1 *= sc; E2 *= sc; E3 *= sc2; E4 *= sc2; E5 *= sc2;
	// params_to_lrep(base), dj_brdf.h:1965-1974
	const float r1 = base.tx, r2 = base.ty;
	const float r3 = 0.5f * base.ax * base.ax + base.tx * base.tx;
	const float r4 = 0.5f * base.ay * base.ay + base.ty * base.ty;
	const float r5 = 0.5f * base.rho * base.ax * base.ay + base.tx * base.ty;
	// lrep1 + lrep2, dj_brdf.h:1992-1999
	lrep_to_params(E1 + r1, E2 + r2, E3 + r3 + 2.0f * E1 * r1, E4 + r4 + 2.0f * E2 * r2, E5 + r5 + E1 * r2 + E2 * r1, out);
}

// ---- MERL index arithmetic, :906-957 (IEEE double mul/div/sqrt + truncation: bit exact) ---------
DJB_DEV int merl_theta_half_index(float th)
{
	if (th <= 0.0f) return 0;
	float deg = (float)(((double)th / (DJB_PI / 2.0)) * 90.0);
	float t = deg * 90.0f;
	t = __fsqrt_rn(t); // == (float)sqrt((double)t)
	int r = (int)t;
	return r < 0 ? 0 : (r >= 90 ? 89 : r);
}
DJB_DEV int merl_theta_diff_index(float td)
{
	int t = (int)((double)td / (DJB_PI * 0.5) * 90.0);
	return t < 0 ? 0 : (t < 89 ? t : 89);
}
DJB_DEV int merl_phi_diff_index(float pd)
{
	if (pd < 0.0f) pd = (float)((double)pd + DJB_PI);
	int t = (int)((double)pd / DJB_PI * 360.0 / 2.0);
	return t < 0 ? 0 : (t < 179 ? t : 179);
}
DJB_DEV int merl_cell(V3 i, V3 o)
{
	V3 h, d;
	float th, td, pd;
	io_to_hd(i, o, h, d, th); // merl::eval recomputes theta_h from h: same value
	to_theta_phi(d, td, pd);
	return merl_phi_diff_index(pd) + merl_theta_diff_index(td) * 180 + merl_theta_half_index(th) * 16200;
}

// ---- UTIA lookup, dj_brdf.h:1063-1157 (table already normalised and cast to float at upload) -----
constexpr int UT_NTI = 6, UT_NPI = 48, UT_NTV = 6, UT_NPV = 48;
constexpr int UT_CELLS = 3 * UT_NTI * UT_NPI * UT_NTV * UT_NPV;

// x^y for x > 0 as exp(y log x): within ~1e-12 relative of pow() for the exponents the presets hold (|y log x| < 3e4), i.e. the
// same float after the final rounding in all but ~1e-5 of cases, at less than half of pow()'s double operations
DJB_DEV double pow_pos(double x, double y) { return exp_d(y * log_d(x)); } // djb_dmath.cuh: constant-bank polynomials

// floor(x / d) for a float x >= 0 and d = 15 or 7.5 (utia's cell index, dj_brdf.h:1090-1100, is (int)floor((double)x / d)): the
// float product with 1 / d may land on the wrong side of a multiple of d, the remainder x - q d -- exact in float, x and q d
// share x's grid -- puts it back
DJB_DEV int floor_div(float x, float d, float inv_d)
{
	int q = (int)floorf(x * inv_d);
	const float r = __fmaf_rn(-d, (float)q, x);
	if (r < 0.0f) --q;
	else if (r >= d) ++q;
	return q;
}

// Device layout of the UTIA table: per (theta_i, phi_i, theta_v, phi_v) cell the three channels of the cell (lo) and of its phi_v
// neighbour, wrapped (hi): 32 bytes, fetched by one 256-bit load (sm_100's LDG.256)
#ifndef DJB200_UTIA_ENTRY_DEFINED
#define DJB200_UTIA_ENTRY_DEFINED
struct __align__(32) UtiaEntry { float4 lo, hi; };
#endif
DJB_DEV UtiaEntry utia_load(const UtiaEntry *p)
{
	UtiaEntry v;
	asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
	    : "=f"(v.lo.x), "=f"(v.lo.y), "=f"(v.lo.z), "=f"(v.lo.w), "=f"(v.hi.x), "=f"(v.hi.y), "=f"(v.hi.z), "=f"(v.hi.w)
	    : "l"(p));
	return v;
}
// utia::eval, dj_brdf.h:1063-1157.  `T`: the djb_dmath.cuh table (shared memory in the query kernel)
DJB_DEV V3 utia_eval1(const UtiaEntry *__restrict__ tab, V3 i, V3 o, const double *T)
{
	const float r2d = (float)(180.0 / DJB_PI);
	float ti = (float)((double)r2d * acos_d((double)i.z)), to = (float)((double)r2d * acos_d((double)o.z));
	float pi = (float)((double)r2d * atan2_t((double)i.y, (double)i.x, T));
	float po = (float)((double)r2d * atan2_t((double)o.y, (double)o.x, T));
	if (ti >= 90.0f || to >= 90.0f) return mk(0.f, 0.f, 0.f);
	// (float)((double)x +- 360.0) is the float sum: a double holds the sum of two floats to more than 2 x 24 + 2 bits
	while (pi < 0.0f) pi = pi + 360.0f;
	while (po < 0.0f) po = po + 360.0f;
	while (pi >= 360.0f) pi = pi - 360.0f;
	while (po >= 360.0f) po = po - 360.0f;
	int iti[2], itv[2], ipi[2], ipv[1];
	if (ti >= 0.0f && to >= 0.0f) { // always, for finite directions
		iti[0] = floor_div(ti, 15.0f, 1.0f / 15.0f);
		itv[0] = floor_div(to, 15.0f, 1.0f / 15.0f);
		ipi[0] = floor_div(pi, 7.5f, 1.0f / 7.5f);
		ipv[0] = floor_div(po, 7.5f, 1.0f / 7.5f);
	} else {
		iti[0] = (int)floor((double)ti / 15.0);
		itv[0] = (int)floor((double)to / 15.0);
		ipi[0] = (int)floor((double)pi / 7.5);
		ipv[0] = (int)floor((double)po / 7.5);
	}
	iti[1] = iti[0] + 1;
	if (iti[0] > UT_NTI - 2) { iti[0] = UT_NTI - 2; iti[1] = UT_NTI - 1; }
	itv[1] = itv[0] + 1;
	if (itv[0] > UT_NTV - 2) { itv[0] = UT_NTV - 2; itv[1] = UT_NTV - 1; }
	ipi[1] = ipi[0] + 1;
	if (ipi[1] == UT_NPI) ipi[1] = 0;
	float sum, wti[2], wtv[2], wpi[2], wpv[2];
	// (float)(15.0 * k), (float)(7.5 * k): multiples of 7.5 up to 367.5 are floats, so the float product is the same number
	wti[1] = ti - 15.0f * (float)iti[0]; wti[0] = 15.0f * (float)iti[1] - ti;
	sum = wti[0] + wti[1]; wti[0] /= sum; wti[1] /= sum;
	wtv[1] = to - 15.0f * (float)itv[0]; wtv[0] = 15.0f * (float)itv[1] - to;
	sum = wtv[0] + wtv[1]; wtv[0] /= sum; wtv[1] /= sum;
	// the phi weights use the unwrapped upper index (dj_brdf.h:1110-1117 computes them before the wrap)
	const int ipi1 = ipi[0] + 1, ipv1 = ipv[0] + 1;
	wpi[1] = pi - 7.5f * (float)ipi[0]; wpi[0] = 7.5f * (float)ipi1 - pi;
	sum = wpi[0] + wpi[1]; wpi[0] /= sum; wpi[1] /= sum;
	wpv[1] = po - 7.5f * (float)ipv[0]; wpv[0] = 7.5f * (float)ipv1 - po;
	sum = wpv[0] + wpv[1]; wpv[0] /= sum; wpv[1] /= sum;
	// An entry holds its own cell and its phi_v neighbour (wrapped): the two innermost taps of the reference's loop nest are one
	// 256-bit load, 8 loads per query.  The table (2.65 MB) lives in L2; a warp's 32 taps fall in 32 different lines, so the number of
	// load wavefronts, not bytes, is what the query pays for.
	const int nc = UT_NPV * UT_NTV;
	float rgb[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
	for (int t = 0; t < 8; ++t) { // the reference's order of the 16 taps (theta_i, theta_v, phi_i, phi_v nested), per channel
		const int a = t >> 2, b = (t >> 1) & 1, c = t & 1;
		const UtiaEntry v = utia_load(tab + (nc * (UT_NPI * iti[a] + ipi[c]) + UT_NPV * itv[b] + ipv[0]));
		const float w3 = wti[a] * wtv[b] * wpi[c], w0 = w3 * wpv[0], w1 = w3 * wpv[1];
		rgb[0] += w0 * v.lo.x; rgb[1] += w0 * v.lo.y; rgb[2] += w0 * v.lo.z;
		rgb[0] += w1 * v.hi.x; rgb[1] += w1 * v.hi.y; rgb[2] += w1 * v.hi.z;
	}
#pragma unroll
	for (int isp = 0; isp < 3; ++isp) {
		float acc = rgb[isp];
		if (acc >= 0.0375f) // (double)acc > 0.0375: the float 0.0375f lies above the double 0.0375 and no float lies between
			acc = (float)pow_pos_t(div_core((double)(float)((double)acc + 0.055), 1.055), (double)2.4f, T); // base > 0.0875
		else
			acc /= 12.92f;
		rgb[isp] = acc * 100.0f;
	}
	return mk(fmax_ref(0.0f, rgb[0]), fmax_ref(0.0f, rgb[1]), fmax_ref(0.0f, rgb[2]));
}


// ---- SGD / ABC analytic BRDFs, dj_brdf.h:3416-3499, 3608-3668 -----------------------------------------
// Coefficients stay doubles (they are doubles in the reference's tables); the per-channel helpers run in double and
// their results are narrowed where the reference's vec3::from_raw narrows them.
// m: djb200_sgd_data.ch, i.e. [3][11] = rhoD rhoS alpha p f0 f1 kap lambda c k theta0 per channel
DJB_DEV double sgd_g1_ch(double acos_kz, const double *m) // sgd__g1, :3415-3422
{
	const double t1 = acos_kz - m[10];
	// theta <= theta0: pow(0, k) = 0, exp(c 0) = 1, 1 + lambda 0 = 1 exactly (k > 0, c and lambda finite in every preset)
	if (!(t1 > 0.0) && m[9] > 0.0) return t1 == t1 ? 1.0 : t1;
	double t3 = 1.0 + m[7] * (1.0 - exp_d(m[8] * pow_pos(t1, m[9])));
	t3 = 0.0 > t3 ? 0.0 : t3;
	return 1.0 < t3 ? 1.0 : t3;
}
// the general path: every special case of the reference (zero / negative / non-finite intermediate values) through the library
static __device__ __noinline__ V3 sgd_eval1_general(const double *__restrict__ m, float iz, float oz, float hz, V3 Fr)
{
	const float f3[3] = {Fr.x, Fr.y, Fr.z};
	float out[3];
	const double ai = acos((double)iz), ao = acos((double)oz);
	const double ch = (double)hz, c2 = ch * ch, t2 = (1.0 - c2) / c2;
	const double inv_pi = 1.0 / DJB_PI;
	const float r1 = rcp_via_double(iz * oz), r2 = rcp_via_double((float)DJB_PI);
	for (int c = 0; c < 3; ++c) {
		const double *mc = m + 11 * c;
		const float g1i = (float)sgd_g1_ch(ai, mc), g1o = (float)sgd_g1_ch(ao, mc);
		const double ax = mc[2] + t2 / mc[2];
		// sgd__ndf, :3424-3432: (kap exp(-ax) / pi) / (ax^p c2 c2), the two exponentials merged into one
		const float nd = ax > 0.0 ? (float)((mc[6] * exp_d(-ax - mc[3] * log_d(ax)) * inv_pi) / (c2 * c2))
		                          : (float)((mc[6] * exp(-ax) * inv_pi) / (pow(ax, mc[3]) * c2 * c2));
		const float fdg = (f3[c] * nd) * (g1i * g1o);
		out[c] = r2 * ((float)mc[0] + r1 * ((float)mc[1] * fdg));
	}
	return mk(out[0], out[1], out[2]);
}
// what the plain path of sgd_eval1 asks of a material (the same for every pair of a launch: the kernel evaluates it once)
DJB_DEV bool sgd_material_plain(const double *__restrict__ m)
{
	bool ok = true;
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		const double al = fabs(m[11 * c + 2]);
		ok = ok && al >= 1e-30 && al <= 1e30 && fabs(m[11 * c + 6]) < 1e100 && fabs(m[11 * c + 8]) < 1e200 && m[11 * c + 9] > 0.0;
	}
	return ok;
}
// sgd__g1 of one direction for the three channels at once (straight-line code: the three chains of double operations interleave).
// `ok` is cleared when a value leaves the range of the table-driven exp / log (the caller then takes the general path).
// Requires sgd_material_plain: k > 0, so theta <= theta0 gives exactly 1 (see sgd_g1_ch), and |c| < 1e200.
DJB_DEV void sgd_g1_x3(double acos_kz, const double *__restrict__ m, const double *T, float *g1, bool &ok)
{
	double t1[3], w[3];
	bool pos[3], any = false;
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		t1[c] = acos_kz - m[11 * c + 10];
		pos[c] = t1[c] > 0.0;
		any = any || pos[c];
	}
	g1[0] = g1[1] = g1[2] = 1.0f;
	if (!any) return;
	// pow(t1, k) = exp(k log t1); where that is < 1e-300, c times it is 0 to the exponential that follows (|c| < 1e200): exp_t_sat's
	// "some number < 1e-300" changes nothing.  An exponential of less than 1e-300: 1 - it is 1 either way.
	bool okl = true;
#pragma unroll
	for (int c = 0; c < 3; ++c) w[c] = m[11 * c + 9] * log_t_core(pos[c] ? t1[c] : 1.0, T); // t1 > 0 is a normal number here
#pragma unroll
	for (int c = 0; c < 3; ++c) w[c] = m[11 * c + 8] * exp_t_sat(pos[c] ? w[c] : 0.0, T, okl);
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		double t3 = 1.0 + m[11 * c + 7] * (1.0 - exp_t_sat(pos[c] ? w[c] : 0.0, T, okl));
		t3 = 0.0 > t3 ? 0.0 : t3;
		t3 = 1.0 < t3 ? 1.0 : t3;
		if (pos[c]) g1[c] = (float)t3;
	}
	ok = ok && okl;
}
DJB_DEV V3 sgd_eval1(const double *__restrict__ m, V3 i, V3 o, const double *T, bool plain) // sgd::eval, :3454-3469
{
	if (!(i.z > 0.0f && o.z > 0.0f)) return mk(0.f, 0.f, 0.f);
	const V3 h = normalize(i + o);
	FresnelDev fr;
	fr.pts = nullptr;
	fr.npts = 0;
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		fr.v[c] = (float)m[11 * c + 4];
		fr.v[3 + c] = (float)m[11 * c + 5];
	}
	const V3 Fr = fresnel_eval<FK_SGD>(fr, sat_ref(dot(i, h)));
	const float f3[3] = {Fr.x, Fr.y, Fr.z};
	float out[3];
	// the plain case: every logarithm of a positive normal number, every exponential within (-1e6, 700), alpha and cos^4 in float
	// range, both cosines at most 1 -- 9 logarithms and 15 exponentials through the table-driven forms, the divisions through
	// one Newton step
	const double ch = (double)h.z, c2 = ch * ch, c4 = c2 * c2;
	bool ok = plain && c4 >= 1e-30 && i.z <= 1.0f && o.z <= 1.0f; // c2 <= 1
	float g1i[3], g1o[3], nd[3];
	if (ok) {
		const double ai = acos_d((double)i.z), ao = acos_d((double)o.z);
		sgd_g1_x3(ai, m, T, g1i, ok);
		sgd_g1_x3(ao, m, T, g1o, ok);
		const double t2 = div_core(1.0 - c2, c2);
		double y = dm_rcp_seed(c4); // 1 / c4 to full precision: two Newton steps
		y = dm_fma(dm_fma(-c4, y, 1.0), y, y);
		y = dm_fma(dm_fma(-c4, y, 1.0), y, y);
		const double s = y * (1.0 / DJB_PI);
		double ax[3], e[3];
		bool lok[3];
#pragma unroll
		for (int c = 0; c < 3; ++c) {
			ax[c] = m[11 * c + 2] + div_core(t2, m[11 * c + 2]);
			lok[c] = log_t_ok(ax[c]);
			ok = ok && lok[c];
		}
#pragma unroll
		for (int c = 0; c < 3; ++c) e[c] = -ax[c] - m[11 * c + 3] * log_t_core(lok[c] ? ax[c] : 1.0, T);
#pragma unroll
		for (int c = 0; c < 3; ++c) // an exponential < 1e-300 times kap / (pi cos^4) < 1e130 (checked) is 0 as a float either way
			nd[c] = (float)(m[11 * c + 6] * exp_t_sat(e[c], T, ok) * s);
	}
	if (!ok) return sgd_eval1_general(m, i.z, o.z, h.z, Fr);
	const float r1 = rcp_via_double(i.z * o.z), r2 = rcp_via_double((float)DJB_PI);
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		const float fdg = (f3[c] * nd[c]) * (g1i[c] * g1o[c]);
		out[c] = r2 * ((float)m[11 * c] + r1 * ((float)m[11 * c + 1] * fdg));
	}
	return mk(out[0], out[1], out[2]);
}
// m: kD[3] A[3] B C ior (djb200_abc_data)
DJB_DEV V3 abc_eval1(const double *__restrict__ m, V3 i, V3 o, const double *T) // abc::eval, :3633-3647
{
	if (!(i.z > 0.0f && o.z > 0.0f)) return mk(0.f, 0.f, 0.f);
	const V3 h = normalize(i + o);
	const float Fc = unpolarized_channel(sat_ref(dot(i, h)), (float)m[8]);
	const float g1_i = fmin_ref(1.0f, 2.0f * (h.z * i.z / dot(h, i))); // abc::gaf, :3649-3655
	const float g1_o = fmin_ref(1.0f, 2.0f * (h.z * o.z / dot(h, o)));
	const float G = fmin_ref(g1_i, g1_o);
	const double base = 1.0 + m[6] * (1.0 - (double)h.z);
	const double den = base > 0.0 ? pow_pos_t(base, m[7], T) : pow(base, m[7]); // abc__ndf, :3608-3613
	const float r1 = rcp_via_double((float)DJB_PI), r2 = rcp_via_double((float)(DJB_PI * (double)i.z * (double)o.z));
	float out[3];
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		const float nd = (float)(m[3 + c] / den);
		out[c] = r1 * (float)m[c] + r2 * ((Fc * nd) * G);
	}
	return mk(out[0], out[1], out[2]);
}

// merl::eval on the uploaded cells (scaled float4 per cell), dj_brdf.h:987-1024
DJB_DEV V3 merl_eval1(const float4 *__restrict__ cells, V3 i, V3 o)
{
	float4 v = __ldg(cells + merl_cell(i, o));
	if (v.x < 0.0f || v.y < 0.0f || v.z < 0.0f) return mk(0.f, 0.f, 0.f);
	return mk(v.x, v.y, v.z);
}

// Fresnel term with a run-time kind (uniform across a launch, so the switch never diverges)
DJB_DEV V3 fresnel_rt(int kind, const FresnelDev &f, float c)
{
	switch (kind) {
	case FK_SCHLICK: return fresnel_eval<FK_SCHLICK>(f, c);
	case FK_UNPOLARIZED: return fresnel_eval<FK_UNPOLARIZED>(f, c);
	case FK_SGD: return fresnel_eval<FK_SGD>(f, c);
	case FK_SPLINE: return fresnel_eval<FK_SPLINE>(f, c);
	default: return mk(1.f, 1.f, 1.f);
	}
}

} // namespace djb200
