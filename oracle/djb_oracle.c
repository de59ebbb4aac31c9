/* oracle/djb_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See djb_oracle.h.
 *
 * Conventions used below:
 *   F(x)   -- round to float (an assignment to a float_t in the reference)
 *   D(x)   -- promote to double (a double literal / M_PI / unqualified libm call in the reference)
 * Plain `float op float` expressions are IEEE single operations (x86-64 SSE, FLT_EVAL_METHOD 0),
 * and the file is compiled with -ffp-contract=off so no a*b+c is fused.
 */
#include "djb_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define F(x) ((float)(x))
#define D(x) ((double)(x))
#define ORC_PI 3.14159265358979323846 /* M_PI */
#define ORC_API __attribute__((visibility("default")))

typedef struct v3 { float x, y, z; } v3;

/* ---------------------------------------------------------------------------------------------
 * small vector algebra: dj_brdf.h:597-637 */
static inline v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_ld(const float *p, int64_t k) { return v3_make(p[3 * k], p[3 * k + 1], p[3 * k + 2]); }
static inline void v3_st(float *p, int64_t k, v3 v) { p[3 * k] = v.x; p[3 * k + 1] = v.y; p[3 * k + 2] = v.z; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(float k, v3 a) { return v3_make(k * a.x, k * a.y, k * a.z); } /* :599-600 */
/* a / b  ==  (1.0 / b) * a, the reciprocal formed in double then rounded (:601) */
static inline v3 v3_div(v3 a, float b) { return v3_scale(F(1.0 / D(b)), a); }
static inline float v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; } /* :618-621 */
static inline v3 v3_cross(v3 a, v3 b) /* :623-628 */
{
	return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float inv_sqrt(float x) { return F(1.0 / sqrt(D(x))); } /* :612-616 */
static inline v3 v3_normalize(v3 v) { return v3_scale(inv_sqrt(v3_dot(v, v)), v); } /* :630-637 */
/* vec3(theta, phi), :589-595 */
static inline v3 v3_spherical(float theta, float phi)
{
	float s = F(sin(D(theta)));
	return v3_make(F(D(s) * cos(D(phi))), F(D(s) * sin(D(phi))), F(cos(D(theta))));
}
/* djb::min / max / sat templates, :574-576 (note the NaN behaviour of the ternaries) */
static inline float f_min(float a, float b) { return a < b ? a : b; }
static inline float f_max(float a, float b) { return a > b ? a : b; }
static inline float f_sat(float x) { return f_min(1.0f, f_max(0.0f, x)); }

/* :650-661 */
static inline void to_theta_phi(v3 p, float *theta, float *phi)
{
	if (D(p.z) > 0.99999) {
		*theta = 0.0f;
		*phi = 0.0f;
	} else if (D(p.z) < -0.99999) {
		*theta = F(ORC_PI);
		*phi = 0.0f;
	} else {
		*theta = F(acos(D(p.z)));
		*phi = F(atan2(D(p.y), D(p.x)));
	}
}

/* ---------------------------------------------------------------------------------------------
 * special functions */
/* djb::erf, Abramowitz-Stegun 7.1.26, :667-688 */
static float as_erf(float x)
{
	const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f;
	const float a4 = -1.453152027f, a5 = 1.061405429f, p = 0.3275911f;
	int sign = (x < 0) ? -1 : 1;
	x = F(fabs(D(x)));
	float t = F(1.0 / (1.0 + D(p * x)));
	float poly = ((((a5 * t + a4) * t) + a3) * t + a2) * t + a1;
	float y = F(1.0 - D(poly * t) * exp(D(-x * x)));
	return (float)sign * y;
}

/* djb::erfinv (Giles), :691-721 -- all single precision except one double sqrt */
static float giles_erfinv(float u)
{
	float w = -logf((1.0f - u) * (1.0f + u)), p;
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
		w = F(sqrt(D(w)) - D(3.0f));
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

/* ---------------------------------------------------------------------------------------------
 * rotations and the half/difference frame, :754-793 */
static v3 rotate_about(v3 x, v3 axis, float angle)
{
	float c = F(cos(D(angle))), s = F(sin(D(angle)));
	v3 out = v3_scale(c, x);
	float t1 = v3_dot(axis, x);
	float t2 = F(D(t1) * (1.0 - D(c)));
	out = v3_add(out, v3_scale(t2, axis));
	out = v3_add(out, v3_scale(s, v3_cross(axis, x)));
	return out;
}

static void io_to_hd(v3 i, v3 o, v3 *h, v3 *d)
{
	float th, ph;
	*h = v3_normalize(v3_add(i, o));
	to_theta_phi(*h, &th, &ph);
	v3 tmp = rotate_about(i, v3_make(0, 0, 1), -ph);
	*d = v3_normalize(rotate_about(tmp, v3_make(0, 1, 0), -th));
}

static void hd_to_io(v3 h, v3 d, v3 *i, v3 *o)
{
	float th, ph;
	to_theta_phi(h, &th, &ph);
	v3 tmp = rotate_about(d, v3_make(0, 1, 0), th);
	*i = v3_normalize(rotate_about(tmp, v3_make(0, 0, 1), ph));
	/* 2.0 * dot(i,h) * h - i : the scalar is formed in double, then rounded when it meets the vec3 */
	float k = F(2.0 * D(v3_dot(*i, h)));
	*o = v3_normalize(v3_sub(v3_scale(k, h), *i));
}

/* ---------------------------------------------------------------------------------------------
 * Fresnel terms, :1292-1344 */
static float unpolarized_channel(float c, float n) /* :1292-1303 */
{
	float g = F(sqrt(D(n * n + c * c) - 1.0));
	float t1 = F(D(c * (g + c)) - 1.0);
	float t2 = F(D(c * (g - c)) + 1.0);
	float t3 = (t1 * t1) / (t2 * t2);
	float t4 = ((g - c) * (g - c)) / ((g + c) * (g + c));
	return F((0.5 * D(t4)) * (1.0 + D(t3)));
}

/* spline::eval with uwrap_edge on vec3 points, :1191-1218 */
static v3 spline_eval_v3(const float *pts, int n, float u)
{
	double ip;
	float frac = F(modf(D(u * (float)n - u), &ip));
	int i1 = (int)ip, i2 = (int)ip + 1;
	if (i1 >= n) i1 = n - 1; else if (i1 < 0) i1 = 0;
	if (i2 >= n) i2 = n - 1; else if (i2 < 0) i2 = 0;
	v3 p1 = v3_ld(pts, i1), p2 = v3_ld(pts, i2);
	return v3_add(p1, v3_scale(frac, v3_sub(p2, p1)));
}

static v3 fresnel_eval(const orc_fresnel *fr, float c)
{
	switch (fr ? fr->kind : ORC_F_IDEAL) {
	case ORC_F_SCHLICK: { /* :1320-1328 */
		float c1 = F(1.0 - D(c)), c2 = c1 * c1, c5 = c2 * c2 * c1;
		v3 f0 = v3_make(fr->v[0], fr->v[1], fr->v[2]);
		return v3_add(f0, v3_scale(c5, v3_sub(v3_make(1, 1, 1), f0)));
	}
	case ORC_F_UNPOLARIZED: /* :1305-1314 */
		return v3_make(unpolarized_channel(c, fr->v[0]), unpolarized_channel(c, fr->v[1]),
		               unpolarized_channel(c, fr->v[2]));
	case ORC_F_SGD: { /* :1330-1336: f0 - c*f1 + pow(1-c,5)*(1-f0), the pow in double */
		v3 f0 = v3_make(fr->v[0], fr->v[1], fr->v[2]), f1 = v3_make(fr->v[3], fr->v[4], fr->v[5]);
		float pw = F(pow(1.0 - D(c), 5.0));
		return v3_add(v3_sub(f0, v3_scale(c, f1)), v3_scale(pw, v3_sub(v3_make(1, 1, 1), f0)));
	}
	case ORC_F_SPLINE: { /* :1338-1344 */
		float u = F(2.0 * acos(D(c)) / ORC_PI);
		return spline_eval_v3(fr->pts, fr->npts, u);
	}
	default: return v3_make(1, 1, 1); /* ideal, :167 */
	}
}

/* ---------------------------------------------------------------------------------------------
 * params, :1355-1474 */
static void ellipse_to_pdf(float a1, float a2, float phi, float *ax, float *ay, float *rho)
{
	float c = F(cos(D(phi))), s = F(sin(D(phi)));
	float c2 = F(2.0 * D(c) * D(c) - D(1.0f));
	float q1 = a1 * a1, q2 = a2 * a2, t1 = q1 + q2, t2 = q1 - q2;
	*ax = F(sqrt(0.5 * D(t1 + t2 * c2)));
	*ay = F(sqrt(0.5 * D(t1 - t2 * c2)));
	*rho = (q2 - q1) * c * s / ((*ax) * (*ay));
}

static void pdf_to_ellipse(float ax, float ay, float rho, float *a1, float *a2, float *phi)
{
	float qx = ax * ax, qy = ay * ay;
	float cov = F(D(rho * ax * ay) * 2.0);
	float t1 = qx + qy, t2 = qx - qy;
	float t3 = F(sqrt(D(t2 * t2 + cov * cov)));
	*a1 = F(sqrt(0.5 * D(t1 + t3)));
	*a2 = F(sqrt(0.5 * D(t1 - t3)));
	*phi = (D(cov) != 0.0) ? F(atan(D((qx - qy - t3) / cov))) : 0.0f;
}

static void set_location(orc_params *p, float tx, float ty)
{
	p->tx = tx;
	p->ty = ty;
	v3 n = v3_normalize(v3_make(-tx, -ty, 1.0f));
	p->n[0] = n.x; p->n[1] = n.y; p->n[2] = n.z;
}

ORC_API void orc_params_elliptic(float a1, float a2, float phi_a, orc_params *out)
{
	out->a1 = a1; out->a2 = a2; out->phi_a = phi_a;
	ellipse_to_pdf(a1, a2, phi_a, &out->ax, &out->ay, &out->rho);
	out->srho = F(sqrt(1.0 - D(out->rho * out->rho)));
	set_location(out, 0.0f, 0.0f);
}

ORC_API void orc_params_pdfparams(float ax, float ay, float rho, float tx, float ty, orc_params *out)
{
	out->ax = ax; out->ay = ay; out->rho = rho;
	out->srho = F(sqrt(1.0 - D(rho * rho)));
	pdf_to_ellipse(ax, ay, rho, &out->a1, &out->a2, &out->phi_a);
	set_location(out, tx, ty);
}

/* ---------------------------------------------------------------------------------------------
 * standard (radial) distributions */
/* spline::eval<float_t> with uwrap_edge, :1191-1218 */
float orc__spline_eval_f(const float *pts, int n, float u)
{
	double ip;
	float frac = F(modf(D(u * (float)n - u), &ip));
	int i1 = (int)ip, i2 = (int)ip + 1;
	if (i1 >= n) i1 = n - 1; else if (i1 < 0) i1 = 0;
	if (i2 >= n) i2 = n - 1; else if (i2 < 0) i2 = 0;
	float p1 = pts[i1], p2 = pts[i2];
	return p1 + frac * (p2 - p1);
}

/* djb::tabular as a third radial family (ndf == ORC_NDF_TABULAR): the tables of the fit that is
 * being built on this thread (djb_oracle_fit.c), :2151-2163 */
#define ORC_NDF_TABULAR 2
#define ORC_NDF_TABULAR_ANISO 3
static __thread const float *t_tab_p22, *t_tab_sigma, *t_tab_qf;
static __thread int t_tab_np22, t_tab_nsigma; /* isotropic: table lengths; anisotropic: w (elevation), h (azimuth) */
void orc__set_tabular(const float *p22, int np22, const float *sigma, int nsigma)
{
	t_tab_p22 = p22; t_tab_np22 = np22; t_tab_sigma = sigma; t_tab_nsigma = nsigma;
}
void orc__set_tabular_qf(const float *qf) { t_tab_qf = qf; }
/* tabular_anisotropic's quantile tables: qf1 has n_qf1 entries (normally azim_res), qf2 is elev_res x azim_res */
static __thread const float *t_tab_qf1, *t_tab_qf2;
static __thread int t_tab_nqf1;
void orc__set_tabular_aniso_qf(const float *qf1, int n_qf1, const float *qf2)
{
	t_tab_qf1 = qf1; t_tab_nqf1 = n_qf1; t_tab_qf2 = qf2;
}

/* spline::uwrap_repeat, :1183-1189: the residue its two loops arrive at, computed directly (a degenerate coordinate casts to
 * INT_MIN, and the loops then take seconds per call unless the compiler replaces them, as gcc -O3 does in the reference build) */
static inline int wrap_repeat(int i, int n)
{
	i %= n;
	return i < 0 ? i + n : i;
}

/* spline::eval2d<float_t>(uwrap_edge, u1, uwrap_repeat, u2), :1220-1247 */
float orc__spline_eval2d_f(const float *pts, int w, int h, float u1, float u2)
{
	double ip1, ip2;
	float frac1 = F(modf(D(u1 * (float)w - u1), &ip1));
	int i1 = (int)ip1, i2 = (int)ip1 + 1;
	if (i1 >= w) i1 = w - 1; else if (i1 < 0) i1 = 0;
	if (i2 >= w) i2 = w - 1; else if (i2 < 0) i2 = 0;
	float frac2 = F(modf(D(u2 * (float)h - u2), &ip2));
	int j1 = (int)ip2, j2 = (int)ip2 + 1;
	j1 = wrap_repeat(j1, h);
	j2 = wrap_repeat(j2, h);
	float p1 = pts[i1 + w * j1], p2 = pts[i2 + w * j1], p3 = pts[i1 + w * j2], p4 = pts[i2 + w * j2];
	float t1 = p1 + frac1 * (p2 - p1);
	float t2 = p3 + frac1 * (p4 - p3);
	return t1 + frac2 * (t2 - t1);
}

/* tabular_anisotropic::p22_std_theta_phi, :2185-2197 */
float orc__aniso_p22_theta_phi(float theta, float phi)
{
	if (D(phi) < 0.0) phi = F(D(phi) + 2.0 * ORC_PI);
	float u1 = F(D(theta) * 2.0 / ORC_PI);
	float u2 = F(D(phi) * 0.5 / ORC_PI);
	return orc__spline_eval2d_f(t_tab_p22, t_tab_np22, t_tab_nsigma, u1, u2);
}

static float p22_radial(int ndf, float r2)
{
	if (ndf == ORC_NDF_TABULAR) { /* tabular::p22_radial, :2151-2156 */
		float r = F(sqrt(D(r2)));
		float u = F(sqrt(2.0 * atan(D(r)) / D(F(ORC_PI))));
		return orc__spline_eval_f(t_tab_p22, t_tab_np22, u);
	}
	if (ndf == ORC_NDF_GGX) { /* :2056-2060 */
		float t = F(1.0 + D(r2));
		return F(1.0 / (ORC_PI * D(t) * D(t)));
	}
	return F(exp(D(-r2)) / ORC_PI); /* :1866-1869 */
}

static float sigma_std_radial(int ndf, float c)
{
	if (ndf == ORC_NDF_TABULAR) { /* tabular::sigma_std_radial, :2158-2162 */
		float u = F(2.0 * acos(D(c)) / D(F(ORC_PI)));
		return orc__spline_eval_f(t_tab_sigma, t_tab_nsigma, u);
	}
	if (ndf == ORC_NDF_GGX) return F((1.0 + D(c)) / 2.0); /* :2062-2065 */
	/* beckmann, :1871-1879 */
	if (D(c) == 1.0) return 1.0f;
	float s = F(sqrt(1.0 - D(c * c)));
	float nu = c / s;
	float tmp = F(exp(D(-nu * nu)) * D(inv_sqrt(F(ORC_PI))));
	return F((D(c) * (1.0 + D(as_erf(nu))) + D(s * tmp)) / 2.0);
}

/* the two virtuals of djb::microfacet: radial::p22_std / sigma_std (:1796-1804) forward to the
 * radial functions; tabular_anisotropic has its own (:2178-2183, 2199-2211) */
static float p22_std(int ndf, float x, float y)
{
	if (ndf == ORC_NDF_TABULAR_ANISO) {
		float theta = F(atan(sqrt(D(x * x + y * y))));
		float phi = F(atan2(D(-y), D(-x)));
		return orc__aniso_p22_theta_phi(theta, phi);
	}
	return p22_radial(ndf, x * x + y * y);
}

static float sigma_std(int ndf, v3 k)
{
	if (ndf == ORC_NDF_TABULAR_ANISO) {
		float theta = F(acos(D(k.z)));
		float phi = F(atan2(D(k.y), D(k.x)));
		if (D(phi) < 0.0) phi = F(D(phi) + 2.0 * ORC_PI);
		float u1 = F(D(theta) * 2.0 / ORC_PI);
		float u2 = F(D(phi) * 0.5 / ORC_PI);
		return orc__spline_eval2d_f(t_tab_sigma, t_tab_np22, t_tab_nsigma, u1, u2);
	}
	return sigma_std_radial(ndf, k.z);
}

/* microfacet::sigma, :1619-1631 */
static float mf_sigma(int ndf, const orc_params *p, v3 k)
{
	float a = k.x * p->ax + k.y * p->ay * p->rho;
	float b = k.y * p->ay * p->srho;
	float c = k.z - k.x * p->tx - k.y * p->ty;
	float nrm = F(sqrt(D(a * a + b * b + c * c)));
	v3 ks = v3_div(v3_make(a, b, c), nrm);
	return nrm * sigma_std(ndf, ks);
}

/* :1633-1642 */
static float mf_g1(int ndf, const orc_params *p, v3 k)
{
	float test = v3_dot(k, v3_make(p->n[0], p->n[1], p->n[2]));
	if (D(test) > 0.0) return k.z / mf_sigma(ndf, p, k);
	return 0.0f;
}

/* :1644-1665 */
static float mf_gaf(int ndf, int shadow, const orc_params *p, v3 i, v3 o)
{
	float g1o = mf_g1(ndf, p, o);
	if (shadow) {
		float g1i = mf_g1(ndf, p, i);
		float t = g1i * g1o;
		if (D(t) > 0.0) return t / (g1i + g1o - t);
		return 0.0f;
	}
	return g1o;
}

/* :1574-1587 */
static float mf_p22(int ndf, const orc_params *p, float x, float y)
{
	x -= p->tx;
	y -= p->ty;
	float nrm = p->ax * p->ay * p->srho;
	float xs = x / p->ax;
	float t1 = p->ax * y - p->rho * p->ay * x;
	float t2 = p->ax * p->ay * p->srho;
	float ys = t1 / t2;
	return p22_std(ndf, xs, ys) / nrm;
}

/* :1559-1570 */
static float mf_ndf(int ndf, const orc_params *p, v3 h)
{
	if (h.z > 1e-4f) {
		float c2 = h.z * h.z, c4 = c2 * c2;
		float sx = -h.x / h.z, sy = -h.y / h.z;
		return mf_p22(ndf, p, sx, sy) / c4;
	}
	return 0.0f;
}

/* :1602-1615 */
static float mf_vndf(int ndf, const orc_params *p, v3 h, v3 k)
{
	float kh = v3_dot(k, h);
	if (D(kh) > 0.0) return kh * mf_ndf(ndf, p, h) / mf_sigma(ndf, p, k);
	return 0.0f;
}

/* :1529-1547 */
static v3 mf_evalp(int ndf, const orc_fresnel *fr, int shadow, const orc_params *p, v3 i, v3 o)
{
	v3 h = v3_normalize(v3_add(i, o));
	float G = mf_gaf(ndf, shadow, p, i, o);
	if (D(G) > 0.0) {
		float cd = f_sat(v3_dot(o, h));
		v3 Fr = fresnel_eval(fr, cd);
		float Dn = mf_ndf(ndf, p, h);
		return v3_scale(F(D(Dn * G) / (4.0 * D(o.z))), Fr);
	}
	return v3_make(0, 0, 0);
}

/* :1551-1555 */
static v3 mf_eval(int ndf, const orc_fresnel *fr, int shadow, const orc_params *p, v3 i, v3 o)
{
	return v3_div(mf_evalp(ndf, fr, shadow, p, i, o), i.z);
}

/* :1713-1730 (beckmann and ggx both support Smith VNDF sampling) */
static float mf_pdf(int ndf, int shadow, const orc_params *p, v3 i, v3 o)
{
	v3 h = v3_normalize(v3_add(i, o));
	float G = mf_gaf(ndf, shadow, p, i, o);
	if (D(G) > 0.0) {
		/* tabular::supports_smith_vndf_sampling() is false (:413): the pdf of normal-map sampling, :1724-1725 */
		if (ndf == ORC_NDF_TABULAR || ndf == ORC_NDF_TABULAR_ANISO)
			return F(D(h.z * mf_ndf(ndf, p, h)) / (4.0 * D(v3_dot(i, h))));
		return F(D(mf_vndf(ndf, p, h, o)) / (4.0 * D(v3_dot(i, h))));
	}
	return 0.0f;
}

/* ---------------------------------------------------------------------------------------------
 * visible-normal sampling */
/* ggx::qf2_radial, :2089-2119 */
static float ggx_qf2(float u, float ck, float sk)
{
	float st = F(D(u) * (1.0 + D(ck)) - 1.0);
	float ct = F(sqrt(1.0 - D(st * st)));
	if (D(ct) > 0.707107) {
		float tt = st / ct;
		if (D(sk) < 0.707107) {
			float tk = sk / ck;
			return F(D(-(tt + tk)) / (1.0 - D(tt * tk)));
		} else {
			float kk = ck / sk;
			return F((1.0 + D(tt * kk)) / D(tt - kk));
		}
	} else {
		float cot = ct / st;
		if (D(sk) < 0.707107) {
			float tk = sk / ck;
			return F((1.0 + D(tk * cot)) / D(tk - cot));
		} else {
			float kk = ck / sk;
			return F(D(cot + kk) / (1.0 - D(cot * kk)));
		}
	}
}

/* ggx::qf3_radial + qf3_rational_approx, :2121-2146 */
static float ggx_qf3(float u, float qf2)
{
	float alpha = F(sqrt(1.0 + D(qf2 * qf2)));
	float S;
	if (D(u) < 0.5) {
		u = F(2.0 * (0.5 - D(u)));
		S = -1.0f;
	} else {
		u = F(2.0 * (D(u) - 0.5));
		S = 1.0f;
	}
	double du = D(u);
	float p = F(du * (du * (du * (-0.365728915865723) + 0.790235037209296) - 0.424965825137544)
	            + 0.000152998850436920);
	float q = F(du * (du * (du * (du * 0.169507819808272 - 0.397203533833404) - 0.232500544458471) + 1)
	            - 0.539825872510702);
	return S * alpha * (p / q);
}

/* beckmann::qf2_radial, :1897-1952 */
static float beckmann_qf2(float u, float ck, float sk)
{
	const float sqrt_pi_inv = F(1. / sqrt(ORC_PI));
	float cot = ck / sk, tan_k = sk / ck;
	float a = -1, c = as_erf(cot);
	u = f_max(u, 1e-6f);
	float fit = 1 + ck * (-0.876f + ck * (0.4265f - 0.0594f * ck));
	float b = c - (1 + c) * powf(1 - u, fit);
	float normalization = F(1 / (D(1 + c) + D(sqrt_pi_inv * tan_k) * exp(D(-cot * cot))));
	int it = 0;
	while (++it < 10) {
		if (!(b >= a && b <= c)) b = 0.5f * (a + c);
		float ie = giles_erfinv(b);
		float value = normalization * (1 + b + sqrt_pi_inv * tan_k * expf(-ie * ie)) - u;
		float derivative = normalization * (1 - ie * tan_k);
		if (fabs(D(value)) < D(1e-5f)) break;
		if (value > 0) c = b; else a = b;
		b -= value / derivative;
	}
	return giles_erfinv(f_max(-0.9999f, b));
}

/* radial::sample_vp22_std_smith, :1818-1846 */
static void sample_std_slopes(int ndf, float u1, float u2, v3 k, float *xs, float *ys)
{
	if (ndf == ORC_NDF_TABULAR_ANISO) { /* tabular_anisotropic::sample_vp22_std_nmap, :2826-2837, with qf1 / qf2, :2780-2784, 2814-2824 */
		float phi = F(D(orc__spline_eval_f(t_tab_qf1, t_tab_nqf1, u1)) * 2.0 * ORC_PI);
		float uphi = F(D(phi) / (2.0 * ORC_PI));
		float theta = F(D(orc__spline_eval2d_f(t_tab_qf2, t_tab_np22, t_tab_nsigma, u2, uphi)) * 0.5 * ORC_PI);
		float tan_theta = F(tan(D(theta)));
		*xs = F(D(-tan_theta) * cos(D(phi)));
		*ys = F(D(-tan_theta) * sin(D(phi)));
		return;
	}
	if (ndf == ORC_NDF_TABULAR) { /* radial::sample_vp22_std_nmap, :1806-1816, with tabular::qf_radial, :2172-2176 */
		float phi_h = F(D(u1) * ORC_PI * 2.0);
		float qf = orc__spline_eval_f(t_tab_qf, t_tab_np22, u2);
		float r_h = F(tan(D(qf * F(ORC_PI) / 2.0f)));
		*xs = F(D(r_h) * cos(D(phi_h)));
		*ys = F(D(r_h) * sin(D(phi_h)));
		return;
	}
	float ck = k.z;
	float sk = D(k.z) < 1.0 ? F(sqrt(1.0 - D(k.z * k.z))) : 0.0f;
	float tx, ty;
	if (ndf == ORC_NDF_GGX) {
		tx = ggx_qf2(u1, ck, sk);
		ty = ggx_qf3(u2, tx);
	} else {
		tx = beckmann_qf2(u1, ck, sk);
		ty = giles_erfinv(F(2.0 * D(u2) - 1.0)); /* qf3_radial -> qf1, :1891-1894, 1954-1957 */
	}
	if (D(sk) == 0.0) {
		*xs = tx;
		*ys = ty;
	} else {
		float nrm = inv_sqrt(k.x * k.x + k.y * k.y);
		float cp = k.x * nrm, sp = k.y * nrm;
		*xs = cp * tx - sp * ty;
		*ys = sp * tx + cp * ty;
	}
}

/* microfacet::sample, :1669-1709 */
static v3 mf_sample(int ndf, const orc_params *p, float u1, float u2, v3 o)
{
	u1 = f_sat(u1) * 0.99998f + 0.00001f;
	u2 = f_sat(u2) * 0.99998f + 0.00001f;
	float a = o.x * p->ax + o.y * p->ay * p->rho;
	float b = o.y * p->ay * p->srho;
	float c = o.z - o.x * p->tx - o.y * p->ty;
	v3 os = v3_normalize(v3_make(a, b, c));
	if (D(os.z) > 0.0) {
		float txm, tym;
		sample_std_slopes(ndf, u1, u2, os, &txm, &tym);
		float txh = p->ax * txm + p->tx;
		float chol = p->rho * txm + p->srho * tym;
		float tyh = p->ay * chol + p->ty;
		v3 h = v3_normalize(v3_make(-txh, -tyh, 1.0f));
		float k = F(2.0 * D(v3_dot(o, h)));
		return v3_sub(v3_scale(k, h), o);
	}
	return v3_make(0, 0, 1);
}

/* microfacet::evalp_is, :1734-1765 */
static v3 mf_evalp_is(int ndf, const orc_fresnel *fr, int shadow, const orc_params *p,
                      float u1, float u2, v3 o, v3 *i_out, float *pdf_out)
{
	v3 i = mf_sample(ndf, p, u1, u2, o);
	v3 h = v3_normalize(v3_add(i, o));
	float G = mf_gaf(ndf, shadow, p, i, o);
	*pdf_out = 0.0f;
	if (D(G) > 0.0) {
		float cd = f_sat(v3_dot(o, h));
		*i_out = i;
		if (ndf == ORC_NDF_TABULAR || ndf == ORC_NDF_TABULAR_ANISO) { /* :1753-1756 */
			float pdf_ = F(D(h.z * mf_ndf(ndf, p, h)) / (4.0 * D(cd)));
			*pdf_out = pdf_;
			return v3_div(mf_evalp(ndf, fr, shadow, p, i, o), pdf_);
		}
		v3 Fr = fresnel_eval(fr, cd);
		float g1 = mf_g1(ndf, p, o);
		*pdf_out = F(D(mf_vndf(ndf, p, h, o)) / (4.0 * D(cd)));
		return v3_scale(G / g1, Fr);
	}
	return v3_make(0, 0, 0);
}

/* ---------------------------------------------------------------------------------------------
 * threading helper */
typedef void (*range_fn)(void *ctx, int64_t b, int64_t e);
typedef struct { range_fn fn; void *ctx; int64_t b, e; } range_job;
static void *range_tramp(void *arg)
{
	range_job *j = (range_job *)arg;
	j->fn(j->ctx, j->b, j->e);
	return NULL;
}
void orc_parallel_ranges(int64_t n, int nthreads, range_fn fn, void *ctx)
{
	if (nthreads <= 1 || n < 2 * (int64_t)nthreads) { fn(ctx, 0, n); return; }
	if (nthreads > 256) nthreads = 256;
	pthread_t th[256];
	range_job jobs[256];
	int64_t chunk = (n + nthreads - 1) / nthreads;
	int started = 0;
	for (int t = 0; t < nthreads; ++t) {
		int64_t b = t * chunk, e = b + chunk < n ? b + chunk : n;
		if (b >= e) break;
		jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].b = b; jobs[t].e = e;
		pthread_create(&th[t], NULL, range_tramp, &jobs[t]);
		++started;
	}
	for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

typedef struct {
	int op, ndf, shadow;
	const orc_fresnel *F;
	orc_params P;
	const float *a, *b;
	float *o0, *o1, *o2;
	const float *tab_p22, *tab_sigma, *tab_qf; /* ORC_NDF_TABULAR / ORC_NDF_TABULAR_ANISO only */
	const float *tab_qf1, *tab_qf2;            /* ORC_NDF_TABULAR_ANISO sampling */
	int tab_nqf1;
	int tab_res, tab_ar;                       /* radial: table length; anisotropic: elevation x azimuth resolution */
} mf_ctx;

static void mf_range(void *vctx, int64_t s, int64_t e)
{
	mf_ctx *c = (mf_ctx *)vctx;
	if (c->ndf == ORC_NDF_TABULAR) { /* the tables are thread-local state */
		orc__set_tabular(c->tab_p22, c->tab_res, c->tab_sigma, c->tab_res);
		orc__set_tabular_qf(c->tab_qf);
	} else if (c->ndf == ORC_NDF_TABULAR_ANISO) {
		orc__set_tabular(c->tab_p22, c->tab_res, c->tab_sigma, c->tab_ar);
		orc__set_tabular_aniso_qf(c->tab_qf1, c->tab_nqf1, c->tab_qf2);
	}
	for (int64_t k = s; k < e; ++k) {
		switch (c->op) {
		case 0: v3_st(c->o0, k, mf_eval(c->ndf, c->F, c->shadow, &c->P, v3_ld(c->a, k), v3_ld(c->b, k))); break;
		case 1: v3_st(c->o0, k, mf_evalp(c->ndf, c->F, c->shadow, &c->P, v3_ld(c->a, k), v3_ld(c->b, k))); break;
		case 2: c->o0[k] = mf_pdf(c->ndf, c->shadow, &c->P, v3_ld(c->a, k), v3_ld(c->b, k)); break;
		case 3: v3_st(c->o0, k, mf_sample(c->ndf, &c->P, c->a[2 * k], c->a[2 * k + 1], v3_ld(c->b, k))); break;
		case 4: {
			v3 iv = v3_make(0, 0, 0);
			float pdf;
			v3 w = mf_evalp_is(c->ndf, c->F, c->shadow, &c->P, c->a[2 * k], c->a[2 * k + 1], v3_ld(c->b, k), &iv, &pdf);
			v3_st(c->o0, k, w);
			v3_st(c->o1, k, iv);
			c->o2[k] = pdf;
		} break;
		}
	}
}

static void mf_run(int op, int ndf, const orc_fresnel *F, int shadow, const orc_params *P,
                   const float *a, const float *b, int64_t n, float *o0, float *o1, float *o2, int nthreads)
{
	mf_ctx c;
	memset(&c, 0, sizeof c);
	c.op = op; c.ndf = ndf; c.shadow = shadow; c.F = F;
	if (P) c.P = *P; else orc_params_elliptic(1.0f, 1.0f, 0.0f, &c.P); /* params::standard(), :1412-1415 */
	c.a = a; c.b = b; c.o0 = o0; c.o1 = o1; c.o2 = o2;
	orc_parallel_ranges(n, nthreads, mf_range, &c);
}

ORC_API void orc_microfacet_eval(int ndf, const orc_fresnel *F, int shadow, const orc_params *P,
                                 const float *wi, const float *wo, int64_t n, float *out3, int nthreads)
{ mf_run(0, ndf, F, shadow, P, wi, wo, n, out3, NULL, NULL, nthreads); }
ORC_API void orc_microfacet_evalp(int ndf, const orc_fresnel *F, int shadow, const orc_params *P,
                                  const float *wi, const float *wo, int64_t n, float *out3, int nthreads)
{ mf_run(1, ndf, F, shadow, P, wi, wo, n, out3, NULL, NULL, nthreads); }
ORC_API void orc_microfacet_pdf(int ndf, const orc_fresnel *F, int shadow, const orc_params *P,
                                const float *wi, const float *wo, int64_t n, float *out1, int nthreads)
{ mf_run(2, ndf, F, shadow, P, wi, wo, n, out1, NULL, NULL, nthreads); }
ORC_API void orc_microfacet_sample(int ndf, const orc_fresnel *F, int shadow, const orc_params *P,
                                   const float *u2, const float *wo, int64_t n, float *out3, int nthreads)
{ mf_run(3, ndf, F, shadow, P, u2, wo, n, out3, NULL, NULL, nthreads); }
ORC_API void orc_microfacet_evalp_is(int ndf, const orc_fresnel *F, int shadow, const orc_params *P,
                                     const float *u2, const float *wo, int64_t n,
                                     float *out_w3, float *out_i3, float *out_pdf, int nthreads)
{ mf_run(4, ndf, F, shadow, P, u2, wo, n, out_w3, out_i3, out_pdf, nthreads); }

/* the public component queries of djb::microfacet (:258-272, 1559-1665): what = 0 ndf(h = a), 1 gaf(h = a, i = b, o = c),
 * 2 g1(h = a, k = b), 3 sigma(k = a), 4 p22(x = a.x, y = a.y), 5 vp22(x, y = a.xy, k = b), 6 vndf(h = a, k = b),
 * 7 fresnel(cos = a.x) -> rgb */
ORC_API void orc_microfacet_component(int ndf, const orc_fresnel *Fr, int shadow, const orc_params *P, int what,
                                      const float *a, const float *b, const float *c, int64_t n, float *out)
{
	orc_params std_p;
	if (!P) { orc_params_elliptic(1.0f, 1.0f, 0.0f, &std_p); P = &std_p; }
	for (int64_t k = 0; k < n; ++k) {
		v3 va = v3_ld(a, k), vb = b ? v3_ld(b, k) : v3_make(0, 0, 1), vc = c ? v3_ld(c, k) : v3_make(0, 0, 1);
		switch (what) {
		case 0: out[k] = mf_ndf(ndf, P, va); break;
		case 1: out[k] = mf_gaf(ndf, shadow, P, vb, vc); break;
		case 2: out[k] = mf_g1(ndf, P, vb); break;
		case 3: out[k] = mf_sigma(ndf, P, va); break;
		case 4: out[k] = mf_p22(ndf, P, va.x, va.y); break;
		case 5: { /* vp22, :1589-1598 */
			v3 h = v3_normalize(v3_make(-va.x, -va.y, 1.0f));
			float jacobian = h.z * h.z * h.z;
			out[k] = jacobian * mf_vndf(ndf, P, h, vb);
		} break;
		case 6: out[k] = mf_vndf(ndf, P, va, vb); break;
		default: v3_st(out, k, fresnel_eval(Fr, va.x)); break;
		}
	}
}

/* djb::tabular as an evaluable / samplable BRDF (the tables come from orc_fit_tabular): the microfacet queries of
 * dj_brdf.h:1529-1765 on the tabulated radial distribution.  op: 0 eval, 1 evalp, 2 pdf, 3 sample, 4 evalp_is */
ORC_API void orc_tabular_query(int op, const float *p22, const float *sigma, const float *qf, int res,
                               const orc_fresnel *F, int shadow, const orc_params *P, const float *a, const float *b,
                               int64_t n, float *o0, float *o1, float *o2, int nthreads)
{
	mf_ctx c;
	memset(&c, 0, sizeof c);
	c.op = op; c.ndf = ORC_NDF_TABULAR; c.shadow = shadow; c.F = F;
	if (P) c.P = *P; else orc_params_elliptic(1.0f, 1.0f, 0.0f, &c.P);
	c.a = a; c.b = b; c.o0 = o0; c.o1 = o1; c.o2 = o2;
	c.tab_p22 = p22; c.tab_sigma = sigma; c.tab_qf = qf; c.tab_res = res;
	orc_parallel_ranges(n, nthreads, mf_range, &c);
	orc__set_tabular(NULL, 0, NULL, 0);
	orc__set_tabular_qf(NULL);
}

/* djb::tabular_anisotropic as an evaluable BRDF (eval, evalp, pdf; its sampling tables are not restated):
 * op 0 eval, 1 evalp, 2 pdf on the elev_res x azim_res tables of orc_fit_tabular_anisotropic */
ORC_API void orc_tabular_aniso_query(int op, const float *p22, const float *sigma, int elev_res, int azim_res,
                                     const orc_fresnel *F, int shadow, const orc_params *P, const float *wi,
                                     const float *wo, int64_t n, float *o0, int nthreads)
{
	mf_ctx c;
	memset(&c, 0, sizeof c);
	if (op < 0 || op > 2) return;
	c.op = op; c.ndf = ORC_NDF_TABULAR_ANISO; c.shadow = shadow; c.F = F;
	if (P) c.P = *P; else orc_params_elliptic(1.0f, 1.0f, 0.0f, &c.P);
	c.a = wi; c.b = wo; c.o0 = o0;
	c.tab_p22 = p22; c.tab_sigma = sigma; c.tab_res = elev_res; c.tab_ar = azim_res;
	orc_parallel_ranges(n, nthreads, mf_range, &c);
	orc__set_tabular(NULL, 0, NULL, 0);
}

/* sample (op 3) / evalp_is (op 4) of djb::tabular_anisotropic with the quantile tables of orc_aniso_sampling_tables */
ORC_API void orc_tabular_aniso_sample_query(int op, const float *p22, const float *sigma, const float *qf1, int n_qf1,
                                            const float *qf2, int elev_res, int azim_res, const orc_fresnel *F,
                                            int shadow, const orc_params *P, const float *u, const float *wo,
                                            int64_t n, float *o0, float *o1, float *o2, int nthreads)
{
	mf_ctx c;
	memset(&c, 0, sizeof c);
	if (op < 3 || op > 4) return;
	c.op = op; c.ndf = ORC_NDF_TABULAR_ANISO; c.shadow = shadow; c.F = F;
	if (P) c.P = *P; else orc_params_elliptic(1.0f, 1.0f, 0.0f, &c.P);
	c.a = u; c.b = wo; c.o0 = o0; c.o1 = o1; c.o2 = o2;
	c.tab_p22 = p22; c.tab_sigma = sigma; c.tab_res = elev_res; c.tab_ar = azim_res;
	c.tab_qf1 = qf1; c.tab_nqf1 = n_qf1; c.tab_qf2 = qf2;
	orc_parallel_ranges(n, nthreads, mf_range, &c);
	orc__set_tabular(NULL, 0, NULL, 0);
	orc__set_tabular_aniso_qf(NULL, 0, NULL);
}

/* used by djb_oracle_fit.c */
void orc__mf_eval1(int ndf, const orc_fresnel *F, int shadow, const orc_params *P,
                   const float *wi, const float *wo, float *out3)
{
	v3_st(out3, 0, mf_eval(ndf, F, shadow, P, v3_ld(wi, 0), v3_ld(wo, 0)));
}

ORC_API void orc_io_to_hd(const float *wi, const float *wo, int64_t n, float *h3, float *d3)
{
	for (int64_t k = 0; k < n; ++k) {
		v3 h, d;
		io_to_hd(v3_ld(wi, k), v3_ld(wo, k), &h, &d);
		v3_st(h3, k, h);
		v3_st(d3, k, d);
	}
}

ORC_API void orc_hd_to_io(const float *h3, const float *d3, int64_t n, float *wi, float *wo)
{
	for (int64_t k = 0; k < n; ++k) {
		v3 i, o;
		hd_to_io(v3_ld(h3, k), v3_ld(d3, k), &i, &o);
		v3_st(wi, k, i);
		v3_st(wo, k, o);
	}
}

/* ---------------------------------------------------------------------------------------------
 * MERL, :906-1024 */
static int merl_th_index(float th) /* :906-920 */
{
	if (D(th) <= 0.0) return 0;
	float deg = F((D(th) / (ORC_PI / 2.0)) * 90);
	float t = deg * 90.0f;
	t = F(sqrt(D(t)));
	int r = (int)t;
	if (r < 0) r = 0; else if (r >= 90) r = 89;
	return r;
}
static int merl_td_index(float td) /* :926-936 */
{
	int t = (int)(D(td) / (ORC_PI * 0.5) * 90);
	return t < 0 ? 0 : (t < 89 ? t : 89);
}
static int merl_pd_index(float pd) /* :940-957 */
{
	if (D(pd) < 0.0) pd = F(D(pd) + ORC_PI);
	int t = (int)(D(pd) / ORC_PI * 360 / 2);
	return t < 0 ? 0 : (t < 179 ? t : 179);
}
static int merl_cell(v3 i, v3 o)
{
	v3 h, d;
	float th, ph, td, pd;
	io_to_hd(i, o, &h, &d);
	to_theta_phi(h, &th, &ph);
	to_theta_phi(d, &td, &pd);
	return merl_pd_index(pd) + merl_td_index(td) * 180 + merl_th_index(th) * 16200;
}
static v3 merl_eval1(const double *tab, v3 i, v3 o) /* :987-1024 */
{
	int c = merl_cell(i, o);
	v3 rgb;
	rgb.x = F(tab[c] * (1.00 / 1500.0));
	rgb.y = F(tab[c + 1458000] * (1.15 / 1500.0));
	rgb.z = F(tab[c + 2916000] * (1.66 / 1500.0));
	if (D(rgb.x) < 0.0 || D(rgb.y) < 0.0 || D(rgb.z) < 0.0) return v3_make(0, 0, 0);
	return rgb;
}

typedef struct { const double *tab; const float *wi, *wo; float *out; int32_t *idx; } merl_ctx;
static void merl_range(void *vctx, int64_t s, int64_t e)
{
	merl_ctx *c = (merl_ctx *)vctx;
	for (int64_t k = s; k < e; ++k) {
		if (c->idx) c->idx[k] = merl_cell(v3_ld(c->wi, k), v3_ld(c->wo, k));
		else v3_st(c->out, k, merl_eval1(c->tab, v3_ld(c->wi, k), v3_ld(c->wo, k)));
	}
}
ORC_API void orc_merl_index(const float *wi, const float *wo, int64_t n, int32_t *idx, int nthreads)
{
	merl_ctx c = {NULL, wi, wo, NULL, idx};
	orc_parallel_ranges(n, nthreads, merl_range, &c);
}
ORC_API void orc_merl_eval(const double *table, const float *wi, const float *wo, int64_t n,
                           float *out3, int nthreads)
{
	merl_ctx c = {table, wi, wo, out3, NULL};
	orc_parallel_ranges(n, nthreads, merl_range, &c);
}
void orc__merl_eval1(const double *tab, const float *wi, const float *wo, float *out3)
{
	v3_st(out3, 0, merl_eval1(tab, v3_ld(wi, 0), v3_ld(wo, 0)));
}

/* ---------------------------------------------------------------------------------------------
 * UTIA, :1039-1177 */
#define UT_NTI 6
#define UT_NPI 48
#define UT_NTV 6
#define UT_NPV 48
ORC_API void orc_utia_normalize(double *t) /* :1162-1177 */
{
	int cnt = 3 * UT_NTI * UT_NPI * UT_NTV * UT_NPV;
	float k = 1.f / 140.f;
	for (int i = 0; i < cnt; ++i) t[i] = (0.0 > t[i] ? 0.0 : t[i]);
	for (int i = 0; i < cnt; ++i) t[i] *= D(k);
}

static v3 utia_eval1(const double *tab, v3 i, v3 o) /* :1063-1157 */
{
	float r2d = F(180.0 / ORC_PI);
	float ti = F(D(r2d) * acos(D(i.z))), to = F(D(r2d) * acos(D(o.z)));
	float pi = F(D(r2d) * atan2(D(i.y), D(i.x))), po = F(D(r2d) * atan2(D(o.y), D(o.x)));
	if (D(ti) >= 90.0 || D(to) >= 90.0) return v3_make(0, 0, 0);
	while (D(pi) < 0.0) pi = F(D(pi) + 360.0);
	while (D(po) < 0.0) po = F(D(po) + 360.0);
	while (pi >= 360) pi = F(D(pi) - 360.0);
	while (po >= 360) po = F(D(po) - 360.0);
	int iti[2], itv[2], ipi[2], ipv[2];
	iti[0] = (int)floor(D(ti) / 15.0); iti[1] = iti[0] + 1;
	if (iti[0] > UT_NTI - 2) { iti[0] = UT_NTI - 2; iti[1] = UT_NTI - 1; }
	itv[0] = (int)floor(D(to) / 15.0); itv[1] = itv[0] + 1;
	if (itv[0] > UT_NTV - 2) { itv[0] = UT_NTV - 2; itv[1] = UT_NTV - 1; }
	ipi[0] = (int)floor(D(pi) / 7.5); ipi[1] = ipi[0] + 1;
	ipv[0] = (int)floor(D(po) / 7.5); ipv[1] = ipv[0] + 1;
	float sum, wti[2], wtv[2], wpi[2], wpv[2];
	wti[1] = ti - F(15.0 * iti[0]); wti[0] = F(15.0 * iti[1]) - ti;
	sum = wti[0] + wti[1]; wti[0] /= sum; wti[1] /= sum;
	wtv[1] = to - F(15.0 * itv[0]); wtv[0] = F(15.0 * itv[1]) - to;
	sum = wtv[0] + wtv[1]; wtv[0] /= sum; wtv[1] /= sum;
	wpi[1] = pi - F(7.5 * ipi[0]); wpi[0] = F(7.5 * ipi[1]) - pi;
	sum = wpi[0] + wpi[1]; wpi[0] /= sum; wpi[1] /= sum;
	wpv[1] = po - F(7.5 * ipv[0]); wpv[0] = F(7.5 * ipv[1]) - po;
	sum = wpv[0] + wpv[1]; wpv[0] /= sum; wpv[1] /= sum;
	if (ipi[1] == UT_NPI) ipi[1] = 0;
	if (ipv[1] == UT_NPV) ipv[1] = 0;
	int nc = UT_NPV * UT_NTV, nr = UT_NPI * UT_NTI;
	float rgb[3];
	for (int isp = 0; isp < 3; ++isp) {
		rgb[isp] = 0.0f;
		for (int a = 0; a < 2; ++a)
		for (int b = 0; b < 2; ++b)
		for (int c = 0; c < 2; ++c)
		for (int d = 0; d < 2; ++d) {
			float w = wti[a] * wtv[b] * wpi[c] * wpv[d];
			int idx = isp * nr * nc + nc * (UT_NPI * iti[a] + ipi[c]) + UT_NPV * itv[b] + ipv[d];
			rgb[isp] += w * F(tab[idx]);
		}
		if (D(rgb[isp]) > 0.0375)
			rgb[isp] = F(pow(D(F(D(rgb[isp]) + 0.055)) / 1.055, D(2.4f)));
		else
			rgb[isp] /= 12.92f;
		rgb[isp] *= 100.0f;
	}
	return v3_make(f_max(0.0f, rgb[0]), f_max(0.0f, rgb[1]), f_max(0.0f, rgb[2]));
}

typedef struct { const double *tab; const float *wi, *wo; float *out; } utia_ctx;
static void utia_range(void *vctx, int64_t s, int64_t e)
{
	utia_ctx *c = (utia_ctx *)vctx;
	for (int64_t k = s; k < e; ++k) v3_st(c->out, k, utia_eval1(c->tab, v3_ld(c->wi, k), v3_ld(c->wo, k)));
}
ORC_API void orc_utia_eval(const double *table, const float *wi, const float *wo, int64_t n,
                           float *out3, int nthreads)
{
	utia_ctx c = {table, wi, wo, out3};
	orc_parallel_ranges(n, nthreads, utia_range, &c);
}
void orc__utia_eval1(const double *tab, const float *wi, const float *wo, float *out3)
{
	v3_st(out3, 0, utia_eval1(tab, v3_ld(wi, 0), v3_ld(wo, 0)));
}

/* ---------------------------------------------------------------------------------------------
 * SGD and ABC analytic BRDFs, :3416-3499 and :3608-3668.  Coefficients are doubles (the reference's static tables);
 * the per-channel helper functions run in double and their results are narrowed by vec3::from_raw. */
static v3 sgd_eval1(const orc_sgd *m, v3 i, v3 o) /* :3454-3469 */
{
	if (!(D(i.z) > 0.0 && D(o.z) > 0.0)) return v3_make(0, 0, 0);
	v3 h = v3_normalize(v3_add(i, o));
	v3 Ks = v3_make(F(m->rhoS[0]), F(m->rhoS[1]), F(m->rhoS[2]));
	v3 Kd = v3_make(F(m->rhoD[0]), F(m->rhoD[1]), F(m->rhoD[2]));
	orc_fresnel fr;
	fr.kind = ORC_F_SGD; /* fresnel::sgd(from_raw(f0), from_raw(f1)), :3443-3444 */
	for (int c = 0; c < 3; ++c) { fr.v[c] = F(m->f0[c]); fr.v[3 + c] = F(m->f1[c]); }
	fr.pts = NULL; fr.npts = 0;
	v3 Fr = fresnel_eval(&fr, f_sat(v3_dot(i, h)));
	float g1i[3], g1o[3], nd[3];
	for (int c = 0; c < 3; ++c) {
		/* sgd__g1, :3415-3422 */
		double ti = acos(D(i.z)) - m->theta0[c], to = acos(D(o.z)) - m->theta0[c];
		ti = 0.0 > ti ? 0.0 : ti;
		to = 0.0 > to ? 0.0 : to;
		double vi = 1.0 + m->lambda[c] * (1.0 - exp(m->c[c] * pow(ti, m->k[c])));
		double vo = 1.0 + m->lambda[c] * (1.0 - exp(m->c[c] * pow(to, m->k[c])));
		vi = 0.0 > vi ? 0.0 : vi; vi = 1.0 < vi ? 1.0 : vi;
		vo = 0.0 > vo ? 0.0 : vo; vo = 1.0 < vo ? 1.0 : vo;
		g1i[c] = F(vi);
		g1o[c] = F(vo);
		/* sgd__ndf, :3424-3432 */
		const double inv_pi = 1.0 / ORC_PI;
		double ch = D(h.z), c2 = ch * ch, t2 = (1.0 - c2) / c2, ax = m->alpha[c] + t2 / m->alpha[c];
		nd[c] = F((m->kap[c] * exp(-ax) * inv_pi) / (pow(ax, m->p[c]) * c2 * c2));
	}
	v3 G = v3_make(g1i[0] * g1o[0], g1i[1] * g1o[1], g1i[2] * g1o[2]);
	v3 Dn = v3_make(nd[0], nd[1], nd[2]);
	v3 FDG = v3_make((Fr.x * Dn.x) * G.x, (Fr.y * Dn.y) * G.y, (Fr.z * Dn.z) * G.z);
	v3 spec = v3_div(v3_make(Ks.x * FDG.x, Ks.y * FDG.y, Ks.z * FDG.z), i.z * o.z);
	return v3_div(v3_add(Kd, spec), F(ORC_PI));
}

static v3 abc_eval1(const orc_abc *m, v3 i, v3 o) /* :3633-3647 */
{
	if (!(D(i.z) > 0.0 && D(o.z) > 0.0)) return v3_make(0, 0, 0);
	v3 h = v3_normalize(v3_add(i, o));
	v3 Kd = v3_make(F(m->kD[0]), F(m->kD[1]), F(m->kD[2]));
	float ior = F(m->ior); /* fresnel::unpolarized(vec3(ior)), :3623 */
	float Fc = unpolarized_channel(f_sat(v3_dot(i, h)), ior);
	/* abc::gaf, :3649-3655 */
	float g1_i = f_min(1.0f, 2.0f * (h.z * i.z / v3_dot(h, i)));
	float g1_o = f_min(1.0f, 2.0f * (h.z * o.z / v3_dot(h, o)));
	float G = f_min(g1_i, g1_o);
	float nd[3];
	for (int c = 0; c < 3; ++c) /* abc__ndf, :3608-3613 */
		nd[c] = F(m->A[c] / pow(1.0 + m->B * (1.0 - D(h.z)), m->C));
	v3 FDG = v3_make((Fc * nd[0]) * G, (Fc * nd[1]) * G, (Fc * nd[2]) * G);
	float den = F(ORC_PI * D(i.z) * D(o.z));
	return v3_add(v3_div(Kd, F(ORC_PI)), v3_div(FDG, den));
}

typedef struct { const orc_sgd *sgd; const orc_abc *abc; const float *wi, *wo; float *out; } analytic_ctx;
static void analytic_range(void *vctx, int64_t s, int64_t e)
{
	analytic_ctx *c = (analytic_ctx *)vctx;
	for (int64_t k = s; k < e; ++k)
		v3_st(c->out, k, c->sgd ? sgd_eval1(c->sgd, v3_ld(c->wi, k), v3_ld(c->wo, k))
		                        : abc_eval1(c->abc, v3_ld(c->wi, k), v3_ld(c->wo, k)));
}
ORC_API void orc_sgd_eval(const orc_sgd *m, const float *wi, const float *wo, int64_t n, float *out3, int nthreads)
{
	analytic_ctx c = {m, NULL, wi, wo, out3};
	orc_parallel_ranges(n, nthreads, analytic_range, &c);
}
ORC_API void orc_abc_eval(const orc_abc *m, const float *wi, const float *wo, int64_t n, float *out3, int nthreads)
{
	analytic_ctx c = {NULL, m, wi, wo, out3};
	orc_parallel_ranges(n, nthreads, analytic_range, &c);
}

/* ---------------------------------------------------------------------------------------------
 * LEAN */
ORC_API void orc_lrep_to_params(const float *E5, int64_t n, orc_params *out) /* :1976-1990 */
{
	for (int64_t k = 0; k < n; ++k) {
		const float *E = E5 + 5 * k;
		float t1 = f_max(0.0f, E[2] - E[0] * E[0]);
		float t2 = f_max(0.0f, E[3] - E[1] * E[1]);
		double sx = sqrt(2.0 * D(t1)), sy = sqrt(2.0 * D(t2));
		float ax = F(1e-5 > sx ? 1e-5 : sx);
		float ay = F(1e-5 > sy ? 1e-5 : sy);
		float rho = 2.0f * (E[4] - E[0] * E[1]) / (ax * ay);
		rho = f_min(0.99f, f_max(-0.99f, rho));
		orc_params_pdfparams(ax, ay, rho, E[0], E[1], &out[k]);
	}
}

ORC_API void orc_params_to_lrep(const orc_params *p, int64_t n, float *E5) /* :1965-1974 */
{
	for (int64_t k = 0; k < n; ++k) {
		const orc_params *q = &p[k];
		float *E = E5 + 5 * k;
		E[0] = q->tx;
		E[1] = q->ty;
		E[2] = 0.5f * q->ax * q->ax + q->tx * q->tx;
		E[3] = 0.5f * q->ay * q->ay + q->ty * q->ty;
		E[4] = 0.5f * q->rho * q->ax * q->ay + q->tx * q->ty;
	}
}

/* The per-shading-point parameter construction of mitsuba/dj_beckmannconductor.cpp:283-314 (the same block is repeated
 * in pdf, :338-361, and sample, :379-400): params::elliptic(alpha1, alpha2, alphaAngle); the LEAN texel (E1..E5) minus
 * its bias, as a lrep (or, without LEAN filtering, the lrep rebuilt from the means); lrep1 *= dmapscale (:2020-2031);
 * params_to_lrep(params) (:1965-1974); lrep1 + lrep2 (:1992-1999); lrep_to_params (:1976-1990).
 * alpha: n x 3 when alpha_per_pair, else 3 floats. */
ORC_API void orc_lean_shading_params(float bias, float dmap_scale, int lean_filtering, int alpha_per_pair,
                                     const float *alpha, const float *E5, int64_t n, orc_params *out)
{
	for (int64_t k = 0; k < n; ++k) {
		const float *a = alpha + (alpha_per_pair ? 3 * k : 0), *E = E5 + 5 * k;
		orc_params base;
		orc_params_elliptic(a[0], a[1], a[2], &base);
		float E1 = E[0], E2 = E[1], E3 = E[2], E4 = E[3], E5v = E[4];
		E1 -= bias;
		E2 -= bias;
		E5v -= bias * bias;
		if (!lean_filtering) { E3 = E1 * E1; E4 = E2 * E2; E5v = E1 * E2; }
		float sc2 = dmap_scale * dmap_scale;
		E1 *= dmap_scale; E2 *= dmap_scale; E3 *= sc2; E4 *= sc2; E5v *= sc2;
		float r[5], s[5];
		orc_params_to_lrep(&base, 1, r);
		s[0] = E1 + r[0];
		s[1] = E2 + r[1];
		s[2] = E3 + r[2] + 2.0f * E1 * r[0];
		s[3] = E4 + r[3] + 2.0f * E2 * r[1];
		s[4] = E5v + r[4] + E1 * r[1] + E2 * r[0];
		orc_lrep_to_params(s, 1, out + k);
	}
}


/* utils/nmap2leanmap.cpp:18-54 (bias == 0) and nmap2leanmap_biased.cpp:23-63 (bias == 25).
 * Planar layout c*W*H + y*W + x on both sides (CImg.h:10146-10149). */
ORC_API void orc_nmap2leanmap(const uint8_t *nmap, int w, int h, float base_roughness, float bias,
                              float *l1, float *l2)
{
	size_t plane = (size_t)w * h;
	for (size_t px = 0; px < plane; ++px) {
		float t1 = ((float)nmap[px] / 255.f) * 2.0f - 1.0f;
		float t2 = ((float)nmap[plane + px] / 255.f) * 2.0f - 1.0f;
		float t3 = ((float)nmap[2 * plane + px] / 255.f);
		float sx = -t1 / t3, sy = -t2 / t3;
		float br = 0.5f * base_roughness * base_roughness;
		l1[px] = bias != 0.0f ? sx + bias : sx;
		l1[plane + px] = bias != 0.0f ? sy + bias : sy;
		l1[2 * plane + px] = 1.f;
		l1[3 * plane + px] = 1.f;
		l2[px] = sx * sx + br;
		l2[plane + px] = sy * sy + br;
		l2[2 * plane + px] = bias != 0.0f ? sx * sy + bias * bias : sx * sy;
		l2[3 * plane + px] = 1.f;
	}
}

/* ---------------------------------------------------------------------------------------------
 * the power-iteration fits share this file's static helpers */
#include "djb_oracle_fit.c"

/* djb::erf (A&S 7.1.26, :667-688) on an array: lets tests check exhaustively the ranges where device code replaces it
 * by its saturated value */
ORC_API void orc_erf(const float *x, int64_t n, float *out)
{
	for (int64_t k = 0; k < n; ++k) out[k] = as_erf(x[k]);
}

/* ---------------------------------------------------------------------------------------------
 * the public scalar queries of djb::radial (:307-310): what = 0 p22_radial(r^2), 1 sigma_std_radial(cos), 2 cdf_radial(r),
 * 3 qf_radial(u); family = ORC_NDF_BECKMANN / ORC_NDF_GGX / 2 (tabular: the four fitted tables of length res) */
ORC_API void orc_radial_query(int family, int what, const float *p22, const float *sigma, const float *qf, const float *cdf,
                              int res, const float *x, int64_t n, float *out)
{
	if (family == ORC_NDF_TABULAR) orc__set_tabular(p22, res, sigma, res);
	for (int64_t k = 0; k < n; ++k) {
		float v = x[k], r;
		if (what == 0) r = p22_radial(family, v);
		else if (what == 1) r = sigma_std_radial(family, v);
		else if (family == ORC_NDF_TABULAR) {
			if (what == 2) { /* tabular::cdf_radial, :2164-2169 */
				float u = F(atan(D(v)) * D(2.0f) / D(F(ORC_PI)));
				if (u < 0.0f) u = 0.0f;
				r = orc__spline_eval_f(cdf, res, F(sqrt(D(u))));
			} else { /* tabular::qf_radial, :2171-2176 */
				float q = orc__spline_eval_f(qf, res, v);
				r = F(tan(D(q * F(ORC_PI) / 2.0f)));
			}
		} else if (family == ORC_NDF_GGX) {
			if (what == 2) { float t = v * v; r = F(D(t) / (1.0 + D(t))); } /* :2067-2071 */
			else r = F(sqrt(D(v) / (1.0 - D(v))));                          /* :2073-2076 */
		} else {
			if (what == 2) r = F(1.0 - exp(D(-v * v)));                     /* :1881-1884 */
			else r = F(sqrt(-log(1.0 - D(v))));                             /* :1886-1889 */
		}
		out[k] = r;
	}
	if (family == ORC_NDF_TABULAR) orc__set_tabular(NULL, 0, NULL, 0);
}

/* ---------------------------------------------------------------------------------------------
 * dmap2nmap, utils/dmap2nmap.cpp:13-44: central differences of an 8-bit displacement map with CImg's atXY() clamping at
 * the borders, slopes scaled by (size / 2) * scale, unit normal packed as 8-bit RGB (truncating casts).
 * That file includes <math.h> through CImg.h, so its sqrt(float) is the float overload (unlike inside dj_brdf.h). */
ORC_API void orc_dmap2nmap(const uint8_t *dmap, int w, int h, float scale, uint8_t *nmap)
{
	const size_t plane = (size_t)w * h;
	for (int j = 0; j < h; ++j)
		for (int i = 0; i < w; ++i) {
			int il = i - 1 < 0 ? 0 : i - 1, ir = i + 1 >= w ? w - 1 : i + 1;
			int jt = j - 1 < 0 ? 0 : j - 1, jb = j + 1 >= h ? h - 1 : j + 1;
			float z_l = (float)dmap[il + (size_t)j * w] / 255.f, z_r = (float)dmap[ir + (size_t)j * w] / 255.f;
			float z_b = (float)dmap[i + (size_t)jb * w] / 255.f, z_t = (float)dmap[i + (size_t)jt * w] / 255.f;
			float slope_x = (float)w * 0.5f * scale * (z_r - z_l);
			float slope_y = (float)h * 0.5f * scale * (z_t - z_b);
			float nrm_sqr = 1.f + slope_x * slope_x + slope_y * slope_y;
			float nrm_inv = F(1.0 / D(sqrtf(nrm_sqr)));
			float nx = -slope_x * nrm_inv, ny = -slope_y * nrm_inv, nz = nrm_inv;
			float tmp1 = F(0.5 * D(nx) + 0.5), tmp2 = F(0.5 * D(ny) + 0.5);
			size_t px = i + (size_t)j * w;
			nmap[px] = (uint8_t)(tmp1 * 255);
			nmap[plane + px] = (uint8_t)(tmp2 * 255);
			nmap[2 * plane + px] = (uint8_t)(nz * 255);
		}
}
