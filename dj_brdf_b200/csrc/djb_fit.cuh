// djb_fit.cuh -- device pieces shared by the isotropic and anisotropic power-iteration fits
// (SURVEY.md rows F1-F9): the fit input ("source") evaluation and the tabulated microfacet BRDF that
// the reference evaluates while it builds its Fresnel table (dj_brdf.h:2151-2211, 2583-2701).
#pragma once
#include "djb_device.cuh"
#include "../../include/djb200.h"

namespace djb200 {

// what a fit reads: brdf.eval(i, o) with the reference's NULL user_param (=> params::standard())
struct FitSourceDev {
	int kind;             // djb200_source_kind
	const float4 *merl;   // scaled cells
	const UtiaEntry *utia; // normalised float table (djb_device.cuh: UtiaEntry)
	int ndf, shadow, fresnel_kind;
	FresnelDev fr;        // fr.pts: device pointer
	double coef[33];      // SGD: djb200_sgd_data.ch flattened; ABC: djb200_abc_data (9 values)
};

DJB_DEV Params standard_params()
{
	Params p; // params::standard() = elliptic(1, 1, 0): every derived value is exact
	p.nx = -0.0f; p.ny = -0.0f; p.nz = 1.0f;
	p.a1 = 1.0f; p.a2 = 1.0f; p.phi_a = 0.0f;
	p.ax = 1.0f; p.ay = 1.0f;
	p.rho = 0.0f; p.srho = 1.0f;
	p.tx = 0.0f; p.ty = 0.0f;
	return p;
}

template <int NDF>
DJB_DEV V3 mf_eval_rt(const Params &p, int fk, const FresnelDev &f, bool shadow, V3 i, V3 o)
{
	V3 h = normalize(i + o);
	float G = mf_gaf<NDF>(p, shadow, i, o);
	V3 e = mk(0.f, 0.f, 0.f);
	if (G > 0.0f) {
		float cd = sat_ref(dot(o, h));
		V3 Fr = fresnel_rt(fk, f, cd);
		float Dn = mf_ndf<NDF>(p, h);
		e = scale((float)((double)(Dn * G) / (4.0 * (double)o.z)), Fr);
	}
	return scale(rcp_via_double(i.z), e);
}

DJB_DEV V3 source_eval(const FitSourceDev &s, V3 i, V3 o)
{
	if (s.kind == DJB200_SOURCE_MERL) return merl_eval1(s.merl, i, o);
	if (s.kind == DJB200_SOURCE_UTIA) return utia_eval1(s.utia, i, o, g_dm_table_dev);
	if (s.kind == DJB200_SOURCE_SGD) return sgd_eval1(s.coef, i, o, g_dm_table_dev, sgd_material_plain(s.coef));
	if (s.kind == DJB200_SOURCE_ABC) return abc_eval1(s.coef, i, o, g_dm_table_dev);
	const Params p = standard_params();
	if (s.ndf == NDF_GGX) return mf_eval_rt<NDF_GGX>(p, s.fresnel_kind, s.fr, s.shadow != 0, i, o);
	return mf_eval_rt<NDF_BECKMANN>(p, s.fresnel_kind, s.fr, s.shadow != 0, i, o);
}

// vec3::intensity, dj_brdf.h:69
DJB_DEV float intensity(V3 v) { return (0.2126f * v.x + 0.7152f * v.y) + 0.0722f * v.z; }

// spline::uwrap_repeat (dj_brdf.h:1183-1189) without its loops: the same residue for every int, a single compare in the
// common case, and no multi-second spin when a degenerate (inf / NaN) table coordinate saturates the index
DJB_DEV int wrap_repeat(int i, int n)
{
	if ((unsigned)i < (unsigned)n) return i;
	i %= n;
	return i < 0 ? i + n : i;
}

// spline::eval2d<float_t>(uwrap_edge, u1, uwrap_repeat, u2), dj_brdf.h:1220-1247
DJB_DEV float spline2d_f(const float *pts, int w, int h, float u1, float u2)
{
	float x1 = u1 * (float)w - u1;
	float ip1 = truncf(x1), frac1 = x1 - ip1;
	int i1 = (int)ip1, i2 = (int)ip1 + 1;
	i1 = i1 >= w ? w - 1 : (i1 < 0 ? 0 : i1);
	i2 = i2 >= w ? w - 1 : (i2 < 0 ? 0 : i2);
	float x2 = u2 * (float)h - u2;
	float ip2 = truncf(x2), frac2 = x2 - ip2;
	int j1 = (int)ip2, j2 = (int)ip2 + 1;
	j1 = wrap_repeat(j1, h);
	j2 = wrap_repeat(j2, h);
	float p1 = pts[i1 + w * j1], p2 = pts[i2 + w * j1], p3 = pts[i1 + w * j2], p4 = pts[i2 + w * j2];
	float t1 = p1 + frac1 * (p2 - p1);
	float t2 = p3 + frac1 * (p4 - p3);
	return t1 + frac2 * (t2 - t1);
}

// djb::tabular (radial tables), dj_brdf.h:2151-2163
// The table coordinates are functions of one float: djb_dmath.cuh's versions give the reference's (libm's) float for every argument
// (acos_coord: all but one of 2.1e9; tests/cpp/dmath_check.cpp).  `T`: the djb_dmath.cuh table -- shared memory in the query kernels,
// g_dm_table_dev (global memory) in the fit kernels.
struct TabIso {
	const float *p22, *sigma;
	int n;
	const double *T;
	DJB_DEV float p22_radial(float r2) const
	{
		float r = __fsqrt_rn(r2); // == (float)sqrt((double)r2)
		return spline_f(p22, n, atan_coord(r, T));
	}
	DJB_DEV float p22_std(float x, float y) const { return p22_radial(x * x + y * y); }
	DJB_DEV float sigma_std(V3 k) const
	{
		return spline_f(sigma, n, acos_coord(k.z));
	}
};

// djb::tabular_anisotropic (theta x phi tables), dj_brdf.h:2178-2211
struct TabAniso {
	const float *p22, *sigma;
	int w, h; // elevation_res, azimuthal_res
	const double *T;
	DJB_DEV float p22_theta_phi(float theta, float phi) const
	{
		if ((double)phi < 0.0) phi = (float)((double)phi + 2.0 * DJB_PI);
		float u1 = (float)((double)theta * 2.0 / DJB_PI);
		float u2 = (float)((double)phi * 0.5 / DJB_PI);
		return spline2d_f(p22, w, h, u1, u2);
	}
	DJB_DEV float p22_std(float x, float y) const
	{
		float theta = (float)atan_t(sqrt_d((double)(x * x + y * y)), T);
		float phi = (float)atan2_t((double)(-y), (double)(-x), T);
		return p22_theta_phi(theta, phi);
	}
	DJB_DEV float sigma_std(V3 k) const
	{
		float theta = (float)acos_d((double)k.z);
		float phi = (float)atan2_t((double)k.y, (double)k.x, T);
		if ((double)phi < 0.0) phi = (float)((double)phi + 2.0 * DJB_PI);
		float u1 = (float)((double)theta * 2.0 / DJB_PI);
		float u2 = (float)((double)phi * 0.5 / DJB_PI);
		return spline2d_f(sigma, w, h, u1, u2);
	}
};

// the generic microfacet queries (dj_brdf.h:1559-1665) on a tabulated distribution
template <class T>
DJB_DEV float tab_sigma(const T &t, const Params &p, V3 k)
{
	float a = k.x * p.ax + k.y * p.ay * p.rho;
	float b = k.y * p.ay * p.srho;
	float c = k.z - k.x * p.tx - k.y * p.ty;
	float nrm = __fsqrt_rn(a * a + b * b + c * c); // == (float)sqrt((double)(...))
	V3 ks = scale(rcp_via_double(nrm), mk(a, b, c));
	return nrm * t.sigma_std(ks);
}

template <class T>
DJB_DEV float tab_g1(const T &t, const Params &p, V3 k)
{
	if (dot(k, mk(p.nx, p.ny, p.nz)) > 0.0f) return k.z / tab_sigma(t, p, k);
	return 0.0f;
}

template <class T>
DJB_DEV float tab_ndf(const T &t, const Params &p, V3 h)
{
	if (h.z > 1e-4f) {
		float c2 = h.z * h.z, c4 = c2 * c2;
		float x = -h.x / h.z, y = -h.y / h.z;
		x -= p.tx;
		y -= p.ty;
		float nrm = p.ax * p.ay * p.srho;
		float xs = x / p.ax;
		float t1 = p.ax * y - p.rho * p.ay * x;
		float ys = t1 / nrm;
		return (t.p22_std(xs, ys) / nrm) / c4;
	}
	return 0.0f;
}

// microfacet::eval with fresnel::ideal (the state of a tabular object while its Fresnel table is built)
template <class T>
DJB_DEV float tab_eval_ideal(const T &t, const Params &p, bool shadow, V3 i, V3 o)
{
	V3 h = normalize(i + o);
	float g1o = tab_g1(t, p, o), G = g1o;
	if (shadow) {
		float g1i = tab_g1(t, p, i);
		float tmp = g1i * g1o;
		G = tmp > 0.0f ? tmp / (g1i + g1o - tmp) : 0.0f;
	}
	float e = 0.0f;
	if (G > 0.0f) {
		float Dn = tab_ndf(t, p, h);
		e = 1.0f * __fdiv_rn(Dn * G, 4.0f * o.z); // == (float)((double)(Dn * G) / (4.0 * (double)o.z))
	}
	return rcp_via_double(i.z) * e;
}

// one theta_d bin of compute_fresnel (dj_brdf.h:2583-2641 and 2643-2701 are the same loop)
template <class T>
DJB_DEV V3 fresnel_bin(const T &t, const FitSourceDev &src, bool shadow, int i, int cnt)
{
	const Params sp = standard_params();
	const float phi_d = (float)(DJB_PI * 0.5), phi_h = 0.0f;
	float tt = (float)i / (float)cnt;
	float theta_d = (float)((double)tt * DJB_PI * 0.5);
	V3 f = mk(0.f, 0.f, 0.f);
	int c0 = 0, c1 = 0, c2 = 0;
	float theta_h = 0.0f;
	const V3 dir_d = spherical(theta_d, phi_d);
	for (int j = 0; (double)theta_h < DJB_PI * 0.5 - (double)theta_d; ++j) {
		float t1 = (float)j / (float)cnt;
		theta_h = (float)((double)(t1 * t1) * DJB_PI * 0.5);
		if ((double)theta_h > DJB_PI * 0.5) continue;
		V3 dir_h = spherical(theta_h, phi_h), dir_i, dir_o;
		hd_to_io(dir_h, dir_d, dir_i, dir_o);
		dir_i = mk(0.f, 0.f, 1.f);
		V3 fr1 = source_eval(src, dir_i, dir_o);
		float fr2 = tab_eval_ideal(t, sp, shadow, dir_i, dir_o); // ideal Fresnel: r == g == b
		if ((double)fr2 > 1e-4) {
			f.x += fr1.x / fr2; ++c0;
			f.y += fr1.y / fr2; ++c1;
			f.z += fr1.z / fr2; ++c2;
		}
	}
	V3 out;
	out.x = c0 == 0 ? 1.0f : fmin_ref(1.0f, f.x / (float)c0);
	out.y = c1 == 0 ? 1.0f : fmin_ref(1.0f, f.y / (float)c1);
	out.z = c2 == 0 ? 1.0f : fmin_ref(1.0f, f.z / (float)c2);
	return out;
}

} // namespace djb200
