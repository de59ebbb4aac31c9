// djb200_facade.hpp -- C++ host side of the B200 BRDF engine: the reference's `djb::` class surface
// (jdupuy/dj_brdf, dj_brdf.h:62-535) re-created on top of the C-ABI in djb200.h.
//
// A host program written against dj_brdf.h keeps its source: same class names, constructors, virtual
// signatures, `microfacet::params` factories, fit constructors / accessors and LEAN `lrep` algebra.
// What changes is where the numbers are computed: every query goes through libdjb200.so to the GPU.
//   * scalar calls (`eval(i, o, &params)`) are batches of one -- correct, but one kernel launch each;
//   * the added `*_batch` members are the intended path: arrays of directions in, arrays out, optionally
//     already in device memory (`djb::device` tag) and under M parameter blocks at once.
// There is no CPU implementation behind this header: if libdjb200.so or the GPU is missing, calls throw
// djb::exc (the reference's own exception type, dj_brdf.h:54-59).
//
// Header-only; C++11; link with -ldjb200.
#ifndef DJB200_FACADE_HPP
#define DJB200_FACADE_HPP

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <exception>
#include <mutex>
#include <string>
#include <vector>

#include "djb200.h"

#ifndef DJB_ASSERT
#include <cassert>
#define DJB_ASSERT(x) assert(x)
#endif

namespace djb {

typedef float float_t; // DJB_USE_DOUBLE_PRECISION is not supported: the kernels mirror the float build

// dj_brdf.h:54-59
class exc : public std::exception {
	std::string m_str;
public:
	exc(const char *fmt, ...)
	{
		char buf[512];
		va_list ap;
		va_start(ap, fmt);
		vsnprintf(buf, sizeof buf, fmt, ap);
		va_end(ap);
		m_str = buf;
	}
	virtual ~exc() throw() {}
	const char *what() const throw() { return m_str.c_str(); }
};

namespace detail {
inline void check(djb200_status s)
{
	if (s != DJB200_OK) throw exc("djb_error: %s", djb200_last_error());
}
} // namespace detail

enum memory_space { host = DJB200_MEM_HOST, device = DJB200_MEM_DEVICE };

// dj_brdf.h:62-71.  Packed 12 bytes: an array of vec3 is the float[3] layout the C-ABI takes.
struct vec3 {
	static vec3 from_raw(const double *v) { return vec3((float_t)v[0], (float_t)v[1], (float_t)v[2]); }
	static vec3 from_raw(const float *v) { return vec3(v[0], v[1], v[2]); }
	static const float_t *to_raw(const vec3 &v) { return &v.x; }
	explicit vec3(float_t s = 0) : x(s), y(s), z(s) {}
	vec3(float_t x, float_t y, float_t z) : x(x), y(y), z(z) {}
	explicit vec3(float_t theta, float_t phi)
	{
		float_t s = (float_t)std::sin((double)theta);
		x = (float_t)((double)s * std::cos((double)phi));
		y = (float_t)((double)s * std::sin((double)phi));
		z = (float_t)std::cos((double)theta);
	}
	float_t intensity() const { return (float_t)0.2126 * x + (float_t)0.7152 * y + (float_t)0.0722 * z; }
	float_t x, y, z;
};
static_assert(sizeof(vec3) == 12, "vec3 must stay packed");
inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3 &a, const vec3 &b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(float_t a, const vec3 &b) { return vec3(a * b.x, a * b.y, a * b.z); }
inline vec3 operator*(const vec3 &a, float_t b) { return b * a; }
inline vec3 operator/(const vec3 &a, float_t b) { return (float_t)(1.0 / (double)b) * a; }
inline vec3 operator/(const vec3 &a, const vec3 &b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 &operator+=(vec3 &a, const vec3 &b) { a = a + b; return a; }
inline vec3 &operator-=(vec3 &a, const vec3 &b) { a = a - b; return a; }
inline vec3 &operator*=(vec3 &a, const vec3 &b) { a = a * b; return a; }
inline vec3 &operator*=(vec3 &a, float_t b) { a = a * b; return a; }
inline vec3 &operator/=(vec3 &a, float_t b) { a = a / b; return a; }
inline float_t dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 normalize(const vec3 &v) { return (float_t)(1.0 / std::sqrt((double)dot(v, v))) * v; }

// ---------------------------------------------------------------------------------------------------
// dj_brdf.h:74-109
class brdf {
public:
	virtual vec3 eval(const vec3 &i, const vec3 &o, const void *user_param = NULL) const = 0;
	virtual vec3 eval_hd(const vec3 &h, const vec3 &d, const void *user_param = NULL) const
	{
		vec3 i, o;
		hd_to_io(h, d, &i, &o);
		return eval(i, o, user_param);
	}
	virtual vec3 evalp(const vec3 &i, const vec3 &o, const void *user_param = NULL) const
	{
		return eval(i, o, user_param) * i.z;
	}
	virtual vec3 evalp_hd(const vec3 &h, const vec3 &d, const void *user_param = NULL) const
	{
		vec3 i, o;
		hd_to_io(h, d, &i, &o);
		return eval(i, o, user_param) * i.z;
	}
	// default importance sampling: cosine-weighted hemisphere (dj_brdf.h:819-845)
	virtual vec3 evalp_is(float_t u1, float_t u2, const vec3 &o, vec3 *i, float_t *pdf, const void *user_param = NULL) const
	{
		const vec3 i_ = sample(u1, u2, o, user_param);
		float_t pdf_ = this->pdf(i_, o);
		if (i) (*i) = i_;
		if (pdf) (*pdf) = pdf_;
		return evalp(i_, o, user_param) / pdf_;
	}
	virtual vec3 sample(float_t u1, float_t u2, const vec3 &, const void * = NULL) const
	{
		float_t x, y;
		concentric(u1, u2, &x, &y);
		return vec3(x, y, (float_t)std::sqrt(1.0 - (double)(x * x) - (double)(y * y)));
	}
	virtual float_t pdf(const vec3 &i, const vec3 &, const void * = NULL) const { return (float_t)((double)i.z / M_PI); }

	// batched queries (added): n pairs, results in `out`; `where` says if the arrays are host or device memory
	virtual void eval_batch(const vec3 *i, const vec3 *o, size_t n, vec3 *out, const void *user_param = NULL,
	                        memory_space where = host, void *stream = NULL) const = 0;

	// Rusinkiewicz frame, dj_brdf.h:771-793
	static void io_to_hd(const vec3 &i, const vec3 &o, vec3 *h, vec3 *d)
	{
		detail::check(djb200_io_to_hd(&i.x, &o.x, 1, &h->x, &d->x, DJB200_MEM_HOST, NULL));
	}
	static void hd_to_io(const vec3 &h, const vec3 &d, vec3 *i, vec3 *o)
	{
		detail::check(djb200_hd_to_io(&h.x, &d.x, 1, &i->x, &o->x, DJB200_MEM_HOST, NULL));
	}
	static void io_to_hd_batch(const vec3 *i, const vec3 *o, size_t n, vec3 *h, vec3 *d, memory_space where = host,
	                           void *stream = NULL)
	{
		detail::check(djb200_io_to_hd(&i->x, &o->x, (int64_t)n, &h->x, &d->x, where, stream));
	}

	brdf() {}
	virtual ~brdf() {}
#if 1 // non-copyable, dj_brdf.h:104-108
private:
	brdf(const brdf &);
	brdf &operator=(const brdf &);
#endif
protected:
	static void concentric(float_t u1, float_t u2, float_t *x, float_t *y) // Shirley-Chiu, dj_brdf.h:726-752
	{
		float_t a = 2 * u1 - 1, b = 2 * u2 - 1, r, phi;
		if (a == 0 && b == 0) { *x = *y = 0; return; }
		// the reference's rounding points: phi and the two products are double expressions rounded once
		if (a * a > b * b) { r = a; phi = (float_t)((M_PI / 4.0) * (double)(b / a)); }
		else { r = b; phi = (float_t)((M_PI / 2.0) - (double)(a / b) * (M_PI / 4.0)); }
		*x = (float_t)((double)r * std::cos((double)phi));
		*y = (float_t)((double)r * std::sin((double)phi));
	}
};

// ---------------------------------------------------------------------------------------------------
// dj_brdf.h:149-207: Fresnel terms.  Here they are descriptors handed to the kernels.
namespace fresnel {
// utilities, dj_brdf.h:151-154, 1255-1289 (host arithmetic; the reference's rounding points)
inline void ior_to_f0(float_t ior, float_t *f0)
{
	DJB_ASSERT(ior > 0.0 && "Invalid ior");
	DJB_ASSERT(f0 && "Null output ptr");
	float_t tmp = (float_t)(((double)ior - 1.0) / ((double)ior + 1.0));
	*f0 = tmp * tmp;
}
inline void ior_to_f0(const vec3 &ior, vec3 *f0)
{
	DJB_ASSERT(f0 && "Null output ptr");
	ior_to_f0(ior.x, &f0->x);
	ior_to_f0(ior.y, &f0->y);
	ior_to_f0(ior.z, &f0->z);
}
inline void f0_to_ior(float_t f0, float_t *ior)
{
	DJB_ASSERT(ior && "Null output ptr");
	if ((double)f0 == 1.0) {
		*ior = 1.0;
	} else {
		float_t sqrt_f0 = (float_t)std::sqrt((double)f0);
		*ior = (float_t)((1.0 + (double)sqrt_f0) / (1.0 - (double)sqrt_f0));
	}
}
inline void f0_to_ior(const vec3 &f0, vec3 *ior)
{
	DJB_ASSERT(ior && "Null output ptr");
	f0_to_ior(f0.x, &ior->x);
	f0_to_ior(f0.y, &ior->y);
	f0_to_ior(f0.z, &ior->z);
}
inline vec3 ior_to_f0(const vec3 &ior) // convenience form (added)
{
	vec3 f0;
	ior_to_f0(ior, &f0);
	return f0;
}
class impl {
public:
	virtual ~impl() {}
	virtual impl *copy() const = 0;
	virtual void describe(djb200_fresnel *f) const = 0;
	// F(cos theta_d), dj_brdf.h:160 (one device query; microfacet::component_batch evaluates many)
	vec3 eval(float_t cos_theta_d) const
	{
		djb200_microfacet d;
		memset(&d, 0, sizeof d);
		d.ndf = DJB200_NDF_GGX;
		d.shadow = 1;
		describe(&d.fresnel);
		const float a[3] = {cos_theta_d, 0, 0};
		vec3 r;
		detail::check(djb200_microfacet_component(&d, NULL, DJB200_COMP_FRESNEL, a, NULL, NULL, 1, &r.x, DJB200_MEM_HOST, NULL));
		return r;
	}
};
class ideal : public impl {
public:
	impl *copy() const { return new ideal(); }
	void describe(djb200_fresnel *f) const { memset(f, 0, sizeof *f); f->kind = DJB200_FRESNEL_IDEAL; }
};
class schlick : public impl {
	vec3 f0;
public:
	explicit schlick(const vec3 &f0) : f0(f0) {}
	impl *copy() const { return new schlick(f0); }
	void describe(djb200_fresnel *f) const
	{
		memset(f, 0, sizeof *f);
		f->kind = DJB200_FRESNEL_SCHLICK;
		f->v[0] = f0.x; f->v[1] = f0.y; f->v[2] = f0.z;
	}
};
class unpolarized : public impl {
	vec3 ior;
public:
	explicit unpolarized(const vec3 &ior) : ior(ior) {}
	impl *copy() const { return new unpolarized(ior); }
	void describe(djb200_fresnel *f) const
	{
		memset(f, 0, sizeof *f);
		f->kind = DJB200_FRESNEL_UNPOLARIZED;
		f->v[0] = ior.x; f->v[1] = ior.y; f->v[2] = ior.z;
	}
};
class sgd : public impl {
	vec3 f0, f1;
public:
	sgd(const vec3 &f0, const vec3 &f1) : f0(f0), f1(f1) {}
	impl *copy() const { return new sgd(f0, f1); }
	void describe(djb200_fresnel *f) const
	{
		memset(f, 0, sizeof *f);
		f->kind = DJB200_FRESNEL_SGD;
		f->v[0] = f0.x; f->v[1] = f0.y; f->v[2] = f0.z; f->v[3] = f1.x; f->v[4] = f1.y; f->v[5] = f1.z;
	}
};
class spline : public impl {
	std::vector<vec3> m_points;
public:
	explicit spline(const std::vector<vec3> &points) : m_points(points) {}
	impl *copy() const { return new spline(m_points); }
	const std::vector<vec3> &get_points() const { return m_points; }
	void describe(djb200_fresnel *f) const
	{
		memset(f, 0, sizeof *f);
		f->kind = DJB200_FRESNEL_SPLINE;
		f->points = m_points.empty() ? NULL : &m_points[0].x;
		f->n_points = (int32_t)m_points.size();
	}
};
} // namespace fresnel

// ---------------------------------------------------------------------------------------------------
// dj_brdf.h:210-298
class microfacet : public brdf {
public:
	class params { // 48 bytes, same layout as the reference (dj_brdf.h:238-242) and as djb200_params
		friend class microfacet;
		djb200_params m;
	public:
		static params standard() { return params(); }
		static params isotropic(float_t a) { return params(a, a, 0); }
		static params elliptic(float_t a1, float_t a2, float_t phi_a = 0.0) { return params(a1, a2, phi_a); }
		static params pdfparams(float_t ax, float_t ay, float_t rho = 0.0, float_t tx_n = 0.0, float_t ty_n = 0.0)
		{
			return params(ax, ay, rho, tx_n, ty_n);
		}
		void set_ellipse(float_t a1, float_t a2, float_t phi_a = 0.0)
		{
			djb200_params t;
			detail::check(djb200_params_elliptic(a1, a2, phi_a, &t));
			t.tx_n = m.tx_n; t.ty_n = m.ty_n; // location is kept (dj_brdf.h:1451-1459)
			memcpy(t.n, m.n, sizeof t.n);
			m = t;
		}
		void set_pdfparams(float_t ax, float_t ay, float_t rho = 0.0, float_t tx_n = 0.0, float_t ty_n = 0.0)
		{
			detail::check(djb200_params_pdfparams(ax, ay, rho, tx_n, ty_n, &m));
		}
		void set_location(float_t tx_n, float_t ty_n) // dj_brdf.h:1437-1442 (the mean normal comes from the host factory)
		{
			djb200_params t;
			detail::check(djb200_params_pdfparams(m.ax, m.ay, m.rho, tx_n, ty_n, &t));
			m.tx_n = tx_n; m.ty_n = ty_n;
			memcpy(m.n, t.n, sizeof m.n);
		}
		void set_location(const vec3 &n) // dj_brdf.h:1444-1449
		{
			m.n[0] = n.x; m.n[1] = n.y; m.n[2] = n.z;
			m.tx_n = -n.x / n.z;
			m.ty_n = -n.y / n.z;
		}
		void get_ellipse(float_t *a1, float_t *a2, float_t *phi_a = NULL) const
		{
			if (a1) *a1 = m.a1;
			if (a2) *a2 = m.a2;
			if (phi_a) *phi_a = m.phi_a;
		}
		void get_pdfparams(float_t *ax, float_t *ay, float_t *rho = NULL, float_t *tx_n = NULL, float_t *ty_n = NULL) const
		{
			if (ax) *ax = m.ax;
			if (ay) *ay = m.ay;
			if (rho) *rho = m.rho;
			if (tx_n) *tx_n = m.tx_n;
			if (ty_n) *ty_n = m.ty_n;
		}
		void get_location(float_t *tx_n, float_t *ty_n) const { *tx_n = m.tx_n; *ty_n = m.ty_n; }
		void get_location(vec3 *n) const { *n = vec3(m.n[0], m.n[1], m.n[2]); }
		params(float_t a1 = 1.0, float_t a2 = 1.0, float_t phi_a = 0.0)
		{
			DJB_ASSERT(a1 > 0.0 && a2 > 0.0 && "Invalid ellipse radii");
			detail::check(djb200_params_elliptic(a1, a2, phi_a, &m));
		}
		params(float_t ax, float_t ay, float_t rho, float_t tx_n, float_t ty_n)
		{
			detail::check(djb200_params_pdfparams(ax, ay, rho, tx_n, ty_n, &m));
		}
		const djb200_params *raw() const { return &m; }
	};
	static_assert(sizeof(params) == 48, "params must stay layout-compatible with the reference");

	virtual ~microfacet() { delete m_fresnel; }
	// BRDF interface (dj_brdf.h:247-256): batches of one
	vec3 eval(const vec3 &i, const vec3 &o, const void *user_param = NULL) const
	{
		vec3 r;
		eval_batch(&i, &o, 1, &r, user_param);
		return r;
	}
	vec3 evalp(const vec3 &i, const vec3 &o, const void *user_param = NULL) const
	{
		vec3 r;
		evalp_batch(&i, &o, 1, &r, user_param);
		return r;
	}
	vec3 sample(float_t u1, float_t u2, const vec3 &o, const void *user_param = NULL) const
	{
		float_t u[2] = {u1, u2};
		vec3 r;
		sample_batch(u, &o, 1, &r, user_param);
		return r;
	}
	float_t pdf(const vec3 &i, const vec3 &o, const void *user_param = NULL) const
	{
		float_t r;
		pdf_batch(&i, &o, 1, &r, user_param);
		return r;
	}
	vec3 evalp_is(float_t u1, float_t u2, const vec3 &o, vec3 *i, float_t *pdf, const void *user_param = NULL) const
	{
		float_t u[2] = {u1, u2}, p = 0;
		vec3 w, iv;
		evalp_is_batch(u, &o, 1, &w, &iv, &p, user_param);
		// the reference writes *i only inside `if (G > 0)` (dj_brdf.h:1749-1764); the kernel returns the zero vector as the
		// direction exactly when G <= 0 (a sampled direction has unit length otherwise)
		if (i && (iv.x != 0 || iv.y != 0 || iv.z != 0)) *i = iv;
		if (pdf) *pdf = p;
		return w;
	}

	// batched interface: `user_param` is NULL (standard), or points at `n_params` blocks.  With BROADCAST every
	// pair is evaluated under every block (out: [n_params][n]); with PER_PAIR block k belongs to pair k.
	void eval_batch(const vec3 *i, const vec3 *o, size_t n, vec3 *out, const void *user_param = NULL,
	                memory_space where = host, void *stream = NULL) const
	{
		eval_batch(i, o, n, out, reinterpret_cast<const params *>(user_param), user_param ? 1 : 0, DJB200_PARAMS_BROADCAST, where,
		           stream);
	}
	void eval_batch(const vec3 *i, const vec3 *o, size_t n, vec3 *out, const params *p, size_t n_params,
	                djb200_params_layout layout, memory_space where = host, void *stream = NULL) const
	{
		dispatch(0, raw(p), p ? (int64_t)n_params : 0, layout, &i->x, &o->x, n, &out->x, NULL, NULL, where, stream);
	}
	void evalp_batch(const vec3 *i, const vec3 *o, size_t n, vec3 *out, const void *user_param = NULL, size_t n_params = 1,
	                 djb200_params_layout layout = DJB200_PARAMS_BROADCAST, memory_space where = host, void *stream = NULL) const
	{
		dispatch(1, raw(user_param), user_param ? (int64_t)n_params : 0, layout, &i->x, &o->x, n, &out->x, NULL, NULL, where, stream);
	}
	void pdf_batch(const vec3 *i, const vec3 *o, size_t n, float_t *out, const void *user_param = NULL, size_t n_params = 1,
	               djb200_params_layout layout = DJB200_PARAMS_BROADCAST, memory_space where = host, void *stream = NULL) const
	{
		dispatch(2, raw(user_param), user_param ? (int64_t)n_params : 0, layout, &i->x, &o->x, n, out, NULL, NULL, where, stream);
	}
	void sample_batch(const float_t *u12, const vec3 *o, size_t n, vec3 *out_i, const void *user_param = NULL,
	                  size_t n_params = 1, djb200_params_layout layout = DJB200_PARAMS_BROADCAST, memory_space where = host,
	                  void *stream = NULL) const
	{
		dispatch(3, raw(user_param), user_param ? (int64_t)n_params : 0, layout, u12, &o->x, n, &out_i->x, NULL, NULL, where, stream);
	}
	void evalp_is_batch(const float_t *u12, const vec3 *o, size_t n, vec3 *out_weight, vec3 *out_i, float_t *out_pdf,
	                    const void *user_param = NULL, size_t n_params = 1,
	                    djb200_params_layout layout = DJB200_PARAMS_BROADCAST, memory_space where = host,
	                    void *stream = NULL) const
	{
		dispatch(4, raw(user_param), user_param ? (int64_t)n_params : 0, layout, u12, &o->x, n, out_weight ? &out_weight->x : NULL,
		         out_i ? &out_i->x : NULL, out_pdf, where, stream);
	}

	// LEAN-filtered shading, wavefront style (added): what mitsuba/dj_beckmannconductor.cpp:283-319, 338-366, 379-410 do
	// per shading point -- (alpha1, alpha2, alphaAngle) and the LEAN texel E1..E5 -> params -> query -- for n points in one
	// fused device pass.  E5: n x 5 moments as fetched from the two LEAN maps (still biased); alpha3: n x 3 when the
	// roughness is textured, else NULL and cfg.alpha applies to every point.
	static djb200_lean_shading lean_config(float_t alpha1, float_t alpha2, float_t alpha_angle, float_t bias = 25,
	                                       float_t dmap_scale = 1, bool lean_filtering = true)
	{
		djb200_lean_shading c;
		c.bias = bias; c.dmap_scale = dmap_scale; c.lean_filtering = lean_filtering ? 1 : 0; c.alpha_per_pair = 0;
		c.alpha[0] = alpha1; c.alpha[1] = alpha2; c.alpha[2] = alpha_angle;
		return c;
	}
	static void lean_params_batch(djb200_lean_shading cfg, const float_t *alpha3, const float_t *E5, size_t n, params *out,
	                              memory_space where = host, void *stream = NULL)
	{
		cfg.alpha_per_pair = alpha3 ? 1 : 0;
		detail::check(djb200_lean_shading_params(&cfg, alpha3, E5, (int64_t)n, const_cast<djb200_params *>(out->raw()), where, stream));
	}
	void evalp_lean_batch(djb200_lean_shading cfg, const float_t *alpha3, const float_t *E5, const vec3 *i, const vec3 *o,
	                      size_t n, vec3 *out, memory_space where = host, void *stream = NULL) const
	{
		cfg.alpha_per_pair = alpha3 ? 1 : 0;
		djb200_microfacet d = describe();
		detail::check(djb200_lean_shading_evalp(&d, &cfg, alpha3, E5, &i->x, &o->x, (int64_t)n, &out->x, where, stream));
	}
	void pdf_lean_batch(djb200_lean_shading cfg, const float_t *alpha3, const float_t *E5, const vec3 *i, const vec3 *o,
	                    size_t n, float_t *out, memory_space where = host, void *stream = NULL) const
	{
		cfg.alpha_per_pair = alpha3 ? 1 : 0;
		djb200_microfacet d = describe();
		detail::check(djb200_lean_shading_pdf(&d, &cfg, alpha3, E5, &i->x, &o->x, (int64_t)n, out, where, stream));
	}
	void evalp_is_lean_batch(djb200_lean_shading cfg, const float_t *alpha3, const float_t *E5, const float_t *u12,
	                         const vec3 *o, size_t n, vec3 *out_weight, vec3 *out_i, float_t *out_pdf,
	                         memory_space where = host, void *stream = NULL) const
	{
		cfg.alpha_per_pair = alpha3 ? 1 : 0;
		djb200_microfacet d = describe();
		detail::check(djb200_lean_shading_evalp_is(&d, &cfg, alpha3, E5, u12, &o->x, (int64_t)n, out_weight ? &out_weight->x : NULL,
		                                           out_i ? &out_i->x : NULL, out_pdf, where, stream));
	}

	// the public component queries, dj_brdf.h:258-272 (batches of one; component_batch takes n)
	vec3 fresnel(float_t cos_theta_d) const
	{
		vec3 a(cos_theta_d, 0, 0), r;
		component_batch(DJB200_COMP_FRESNEL, &a, NULL, NULL, 1, &r.x, NULL);
		return r;
	}
	float_t ndf(const vec3 &h, const params &p = params::standard()) const { return component1(DJB200_COMP_NDF, &h, NULL, NULL, p); }
	float_t gaf(const vec3 &h, const vec3 &i, const vec3 &o, const params &p = params::standard()) const
	{
		return component1(DJB200_COMP_GAF, &h, &i, &o, p);
	}
	float_t g1(const vec3 &h, const vec3 &k, const params &p = params::standard()) const { return component1(DJB200_COMP_G1, &h, &k, NULL, p); }
	float_t sigma(const vec3 &k, const params &p = params::standard()) const { return component1(DJB200_COMP_SIGMA, &k, NULL, NULL, p); }
	float_t p22(float_t x, float_t y, const params &p = params::standard()) const
	{
		vec3 a(x, y, 0);
		return component1(DJB200_COMP_P22, &a, NULL, NULL, p);
	}
	float_t vp22(float_t x, float_t y, const vec3 &k, const params &p = params::standard()) const
	{
		vec3 a(x, y, 0);
		return component1(DJB200_COMP_VP22, &a, &k, NULL, p);
	}
	float_t vndf(const vec3 &h, const vec3 &k, const params &p = params::standard()) const { return component1(DJB200_COMP_VNDF, &h, &k, NULL, p); }
	void component_batch(djb200_component what, const vec3 *a, const vec3 *b, const vec3 *c, size_t n, float_t *out,
	                     const params *p = NULL, memory_space where = host, void *stream = NULL) const
	{
		if (ndf_id() < 0) throw exc("djb_error: component queries are available on the analytic families (ggx, beckmann)");
		djb200_microfacet d = describe();
		detail::check(djb200_microfacet_component(&d, p ? p->raw() : NULL, what, &a->x, b ? &b->x : NULL, c ? &c->x : NULL, (int64_t)n,
		                                          out, where, stream));
	}

	virtual bool supports_smith_vndf_sampling() const = 0;
	// dj_brdf.h:275-276, 1783-1791: the base versions throw, and no class of the reference overrides them
	virtual float_t qf2(float_t, const vec3 &) const { throw exc("djb_error: Not Implemented"); }
	virtual float_t qf3(float_t, const vec3 &, float_t) const { throw exc("djb_error: Not Implemented"); }
	void set_shadow(bool shadow) { m_shadow = shadow; ++m_fresnel_rev; }
	void set_fresnel(const fresnel::impl &f)
	{
		delete m_fresnel;
		m_fresnel = f.copy();
		++m_fresnel_rev; // table-backed subclasses re-upload their device handle
	}
	int get_shadow() const { return m_shadow; }
	const fresnel::impl &get_fresnel() const { return *m_fresnel; }
	// construction state as the C-ABI sees it
	djb200_microfacet describe() const
	{
		djb200_microfacet d;
		d.ndf = ndf_id();
		d.shadow = m_shadow ? 1 : 0;
		m_fresnel->describe(&d.fresnel);
		return d;
	}

protected:
	microfacet(const fresnel::impl &f = fresnel::ideal(), bool shadow = true) : m_fresnel(f.copy()), m_shadow(shadow), m_fresnel_rev(0) {}
	virtual int ndf_id() const = 0;
	// The Fresnel term of a table-backed BRDF travels as `n` spline points: the fitted spline as it is, fresnel::ideal as
	// a table of ones (a lerp between ones is exactly 1); other terms cannot be tabulated without changing results.
	void fresnel_points(size_t n, std::vector<float> *out) const
	{
		djb200_fresnel d;
		m_fresnel->describe(&d);
		if (d.kind == DJB200_FRESNEL_SPLINE && (size_t)d.n_points == n) out->assign(d.points, d.points + 3 * n);
		else if (d.kind == DJB200_FRESNEL_IDEAL) out->assign(3 * n, 1.0f);
		else throw exc("djb_error: a tabulated BRDF takes its fitted Fresnel spline or fresnel::ideal");
	}
	// one C-ABI call; op: 0 eval, 1 evalp, 2 pdf, 3 sample, 4 evalp_is.  tabular overrides this with its own entry points.
	virtual void dispatch(int op, const djb200_params *p, int64_t n_params, djb200_params_layout layout, const float *a,
	                      const float *b, size_t n, float *o0, float *o1, float *o2, memory_space where, void *stream) const
	{
		djb200_microfacet d = describe();
		djb200_status st = DJB200_ERR_INVALID_ARGUMENT;
		switch (op) {
		case 0: st = djb200_microfacet_eval(&d, p, n_params, layout, a, b, (int64_t)n, o0, where, stream); break;
		case 1: st = djb200_microfacet_evalp(&d, p, n_params, layout, a, b, (int64_t)n, o0, where, stream); break;
		case 2: st = djb200_microfacet_pdf(&d, p, n_params, layout, a, b, (int64_t)n, o0, where, stream); break;
		case 3: st = djb200_microfacet_sample(&d, p, n_params, layout, a, b, (int64_t)n, o0, where, stream); break;
		case 4: st = djb200_microfacet_evalp_is(&d, p, n_params, layout, a, b, (int64_t)n, o0, o1, o2, where, stream); break;
		}
		detail::check(st);
	}
	static const djb200_params *raw(const void *p) { return reinterpret_cast<const djb200_params *>(p); }
	float_t component1(djb200_component what, const vec3 *a, const vec3 *b, const vec3 *c, const params &p) const
	{
		float_t r;
		component_batch(what, a, b, c, 1, &r, &p);
		return r;
	}
	const fresnel::impl *m_fresnel;
	bool m_shadow;
	unsigned m_fresnel_rev;
};

// dj_brdf.h:300-324: the radial families and their public scalar queries (what tests/plot_qf.cpp / plot_cdf.cpp tabulate)
class radial : public microfacet {
public:
	float_t p22_radial(float_t r_sqr) const { return radial_query(DJB200_RADIAL_P22, r_sqr); }
	float_t sigma_std_radial(float_t cos_theta_k) const { return radial_query(DJB200_RADIAL_SIGMA_STD, cos_theta_k); }
	float_t cdf_radial(float_t r) const { return radial_query(DJB200_RADIAL_CDF, r); }
	float_t qf_radial(float_t u) const { return radial_query(DJB200_RADIAL_QF, u); }
	// dj_brdf.h:311-314, 1848-1860: "Not Implemented" unless the family overrides them (ggx and beckmann do)
	virtual float_t qf2_radial(float_t, float_t, float_t) const { throw exc("djb_error: Not Implemented"); }
	virtual float_t qf3_radial(float_t, float_t) const { throw exc("djb_error: Not Implemented"); }
	// batched (added): n arguments per call
	void radial_query_batch(djb200_radial_what what, const float_t *x, size_t n, float_t *out, memory_space where = host,
	                        void *stream = NULL) const
	{
		detail::check(djb200_radial_query(what, ndf_id(), radial_handle(), x, (int64_t)n, out, where, stream));
	}
protected:
	radial(const fresnel::impl &f = fresnel::ideal(), bool shadow = true) : microfacet(f, shadow) {}
	virtual const djb200_tabular *radial_handle() const { return NULL; } // tabular: its device tables
	// ggx / beckmann ::qf1, qf2_radial, qf3_radial (dj_brdf.h:366-369, 384-389): the pieces `sample` is made of
	float_t quantile_query(int what, float_t a, float_t b = 0, float_t c = 0) const
	{
		float_t r;
		detail::check(djb200_quantile_query(ndf_id(), what, &a, &b, &c, 1, &r, DJB200_MEM_HOST, NULL));
		return r;
	}
private:
	float_t radial_query(djb200_radial_what what, float_t x) const
	{
		float_t r;
		radial_query_batch(what, &x, 1, &r);
		return r;
	}
};

// dj_brdf.h:374-391
class ggx : public radial {
public:
	ggx(const fresnel::impl &f = fresnel::ideal(), bool shadow = true) : radial(f, shadow) {}
	bool supports_smith_vndf_sampling() const { return true; }
	float_t qf1(float_t u) const { return quantile_query(DJB200_MEMBER_QF1, u); }
	float_t qf2_radial(float_t u, float_t cos_theta_k, float_t sin_theta_k) const
	{
		return quantile_query(DJB200_MEMBER_QF2_RADIAL, u, cos_theta_k, sin_theta_k);
	}
	float_t qf3_radial(float_t u, float_t qf2) const { return quantile_query(DJB200_MEMBER_QF3_RADIAL, u, qf2); }
	// batched (added): unused argument arrays may be NULL
	void quantile_query_batch(int what, const float_t *a, const float_t *b, const float_t *c, size_t n, float_t *out,
	                          memory_space where = host, void *stream = NULL) const
	{
		detail::check(djb200_quantile_query(ndf_id(), what, a, b, c, (int64_t)n, out, where, stream));
	}
protected:
	int ndf_id() const { return DJB200_NDF_GGX; }
};

// dj_brdf.h:327-371
class beckmann : public radial {
public:
	// LEAN / LEADR linear representation: five slope moments (dj_brdf.h:330-356, 1959-2051)
	class lrep {
		friend class beckmann;
		float_t m_E1, m_E2, m_E3, m_E4, m_E5;
	public:
		lrep(float_t E1 = 0, float_t E2 = 0, float_t E3 = 1, float_t E4 = 1, float_t E5 = 0) // defaults of dj_brdf.h:335-337
		    : m_E1(E1), m_E2(E2), m_E3(E3), m_E4(E4), m_E5(E5) {}
		lrep operator+(const lrep &r) const
		{
			return lrep(m_E1 + r.m_E1, m_E2 + r.m_E2, m_E3 + r.m_E3 + (float_t)2.0 * m_E1 * r.m_E1,
			            m_E4 + r.m_E4 + (float_t)2.0 * m_E2 * r.m_E2, m_E5 + r.m_E5 + m_E1 * r.m_E2 + m_E2 * r.m_E1);
		}
		lrep operator*(float_t sc) const
		{
			DJB_ASSERT(sc >= (float_t)0.0 && "Invalid scale");
			float_t s2 = sc * sc;
			return lrep(m_E1 * sc, m_E2 * sc, m_E3 * s2, m_E4 * s2, m_E5 * s2);
		}
		lrep &operator+=(const lrep &r)
		{ // in-place update order of the reference: E1/E2 are advanced first (dj_brdf.h:2011-2020)
			m_E1 += r.m_E1;
			m_E2 += r.m_E2;
			m_E3 += r.m_E3 + (float_t)2.0 * m_E1 * r.m_E1;
			m_E4 += r.m_E4 + (float_t)2.0 * m_E2 * r.m_E2;
			m_E5 += r.m_E5 + m_E1 * r.m_E2 + m_E2 * r.m_E1;
			return *this;
		}
		lrep &operator*=(float_t sc) { return *this = *this * sc; }
		void shear(float_t tx, float_t ty)
		{
			m_E1 += tx; m_E2 += ty; m_E3 += tx * tx; m_E4 += ty * ty; m_E5 += tx * ty;
		}
		void scale(float_t x, float_t y)
		{
			m_E1 *= x; m_E2 *= y; m_E3 *= x * x; m_E4 *= y * y; m_E5 *= x * y;
		}
		const float_t *raw() const { return &m_E1; }
	};
	beckmann(const fresnel::impl &f = fresnel::ideal(), bool shadow = true) : radial(f, shadow) {}
	bool supports_smith_vndf_sampling() const { return true; }
	float_t qf1(float_t u) const { return quantile_query(DJB200_MEMBER_QF1, u); }
	float_t qf2_radial(float_t u, float_t cos_theta_k, float_t sin_theta_k) const
	{
		return quantile_query(DJB200_MEMBER_QF2_RADIAL, u, cos_theta_k, sin_theta_k);
	}
	float_t qf3_radial(float_t u, float_t qf2) const { return quantile_query(DJB200_MEMBER_QF3_RADIAL, u, qf2); }
	// batched (added): unused argument arrays may be NULL
	void quantile_query_batch(int what, const float_t *a, const float_t *b, const float_t *c, size_t n, float_t *out,
	                          memory_space where = host, void *stream = NULL) const
	{
		detail::check(djb200_quantile_query(ndf_id(), what, a, b, c, (int64_t)n, out, where, stream));
	}
	static void params_to_lrep(const microfacet::params &p, lrep *l)
	{
		DJB_ASSERT(l && "Null output ptr");
		detail::check(djb200_params_to_lrep(p.raw(), 1, &l->m_E1, DJB200_MEM_HOST, NULL));
	}
	static void lrep_to_params(const lrep &l, microfacet::params *p)
	{
		DJB_ASSERT(p && "Null output ptr");
		detail::check(djb200_lrep_to_params(l.raw(), 1, const_cast<djb200_params *>(p->raw()), DJB200_MEM_HOST, NULL));
	}
	// per-texel batch (mitsuba/dj_beckmannconductor.cpp:295-314): n x 5 moments -> n params blocks
	static void lrep_to_params_batch(const float_t *E, size_t n, microfacet::params *out, memory_space where = host,
	                                 void *stream = NULL)
	{
		detail::check(djb200_lrep_to_params(E, (int64_t)n, const_cast<djb200_params *>(out->raw()), where, stream));
	}
protected:
	int ndf_id() const { return DJB200_NDF_BECKMANN; }
};
static_assert(sizeof(beckmann::lrep) == 20, "lrep is five packed floats");

// ---------------------------------------------------------------------------------------------------
// dj_brdf.h:126-133
class merl : public brdf {
	djb200_merl *m_h;
public:
	explicit merl(const char *path_to_merl_binary) : m_h(NULL) { detail::check(djb200_merl_load(path_to_merl_binary, &m_h)); }
	// samples: the three planes of doubles of a .binary file already in memory
	explicit merl(const std::vector<double> &samples) : m_h(NULL)
	{
		if (samples.size() != (size_t)3 * 90 * 90 * 180) throw exc("djb_error: MERL sample array must hold 3*90*90*180 doubles");
		detail::check(djb200_merl_create(&samples[0], &m_h));
	}
	~merl() { djb200_merl_destroy(m_h); }
	vec3 eval(const vec3 &i, const vec3 &o, const void * = NULL) const
	{
		vec3 r;
		eval_batch(&i, &o, 1, &r);
		return r;
	}
	void eval_batch(const vec3 *i, const vec3 *o, size_t n, vec3 *out, const void * = NULL, memory_space where = host,
	                void *stream = NULL) const
	{
		detail::check(djb200_merl_eval(m_h, &i->x, &o->x, (int64_t)n, &out->x, where, stream));
	}
	const djb200_merl *handle() const { return m_h; }
};

// dj_brdf.h:136-146
class utia : public brdf {
	djb200_utia *m_h;
public:
	explicit utia(const char *filename) : m_h(NULL) { detail::check(djb200_utia_load(filename, &m_h)); }
	~utia() { djb200_utia_destroy(m_h); }
	vec3 eval(const vec3 &i, const vec3 &o, const void * = NULL) const
	{
		vec3 r;
		eval_batch(&i, &o, 1, &r);
		return r;
	}
	void eval_batch(const vec3 *i, const vec3 *o, size_t n, vec3 *out, const void * = NULL, memory_space where = host,
	                void *stream = NULL) const
	{
		detail::check(djb200_utia_eval(m_h, &i->x, &o->x, (int64_t)n, &out->x, where, stream));
	}
	const djb200_utia *handle() const { return m_h; }
};

// dj_brdf.h:481-511: shifted gamma distribution BRDF of a MERL material (coefficients looked up by name)
class sgd : public brdf {
	djb200_sgd_data m_data;
	fresnel::impl *m_fresnel;
public:
	explicit sgd(const char *name) : m_fresnel(NULL)
	{
		detail::check(djb200_sgd_preset(name, &m_data)); // throws "djb_error: No SGD parameters for <name>" like :3449
		m_fresnel = new fresnel::sgd(
			vec3((float_t)m_data.ch[0][DJB200_SGD_F0], (float_t)m_data.ch[1][DJB200_SGD_F0], (float_t)m_data.ch[2][DJB200_SGD_F0]),
			vec3((float_t)m_data.ch[0][DJB200_SGD_F1], (float_t)m_data.ch[1][DJB200_SGD_F1], (float_t)m_data.ch[2][DJB200_SGD_F1]));
	}
	~sgd() { delete m_fresnel; }
	vec3 eval(const vec3 &i, const vec3 &o, const void * = NULL) const
	{
		vec3 r;
		eval_batch(&i, &o, 1, &r);
		return r;
	}
	void eval_batch(const vec3 *i, const vec3 *o, size_t n, vec3 *out, const void * = NULL, memory_space where = host,
	                void *stream = NULL) const
	{
		detail::check(djb200_sgd_eval(&m_data, &i->x, &o->x, (int64_t)n, &out->x, where, stream));
	}
	// the per-channel terms eval is made of (dj_brdf.h:506-509, 3471-3499)
	vec3 ndf(const vec3 &h) const { return member(DJB200_MEMBER_NDF, &h.x, NULL, NULL); }
	vec3 gaf(const vec3 &h, const vec3 &i, const vec3 &o) const { return member(DJB200_MEMBER_GAF, &h.x, &i.x, &o.x); }
	vec3 g1(const vec3 &k) const { return member(DJB200_MEMBER_G1, &k.x, NULL, NULL); }
	vec3 fresnel(float_t cos_theta_d) const { return member(DJB200_MEMBER_FRESNEL, &cos_theta_d, NULL, NULL); }
	void member_batch(int what, const float_t *a, const float_t *b, const float_t *c, size_t n, vec3 *out,
	                  memory_space where = host, void *stream = NULL) const
	{
		detail::check(djb200_sgd_member(&m_data, what, a, b, c, (int64_t)n, &out->x, where, stream));
	}
	const fresnel::impl &get_fresnel() const { return *m_fresnel; }
	const djb200_sgd_data *data() const { return &m_data; }
private:
	vec3 member(int what, const float_t *a, const float_t *b, const float_t *c) const
	{
		vec3 r;
		member_batch(what, a, b, c, 1, &r);
		return r;
	}
};

// dj_brdf.h:514-535: ABC distribution BRDF of a MERL material
class abc : public brdf {
	djb200_abc_data m_data;
	fresnel::impl *m_fresnel;
public:
	explicit abc(const char *name) : m_fresnel(NULL)
	{
		detail::check(djb200_abc_preset(name, &m_data));
		m_fresnel = new fresnel::unpolarized(vec3((float_t)m_data.ior));
	}
	~abc() { delete m_fresnel; }
	vec3 eval(const vec3 &i, const vec3 &o, const void * = NULL) const
	{
		vec3 r;
		eval_batch(&i, &o, 1, &r);
		return r;
	}
	void eval_batch(const vec3 *i, const vec3 *o, size_t n, vec3 *out, const void * = NULL, memory_space where = host,
	                void *stream = NULL) const
	{
		detail::check(djb200_abc_eval(&m_data, &i->x, &o->x, (int64_t)n, &out->x, where, stream));
	}
	// dj_brdf.h:531-533, 3649-3668
	vec3 ndf(const vec3 &h) const { return member3(DJB200_MEMBER_NDF, &h.x); }
	float_t gaf(const vec3 &h, const vec3 &i, const vec3 &o) const
	{
		float_t r;
		detail::check(djb200_abc_member(&m_data, DJB200_MEMBER_GAF, &h.x, &i.x, &o.x, 1, &r, DJB200_MEM_HOST, NULL));
		return r;
	}
	vec3 fresnel(float_t cos_theta_d) const { return member3(DJB200_MEMBER_FRESNEL, &cos_theta_d); }
	// batched (added): out is n x 3 floats, n x 1 for DJB200_MEMBER_GAF
	void member_batch(int what, const float_t *a, const float_t *b, const float_t *c, size_t n, float_t *out,
	                  memory_space where = host, void *stream = NULL) const
	{
		detail::check(djb200_abc_member(&m_data, what, a, b, c, (int64_t)n, out, where, stream));
	}
	const fresnel::impl &get_fresnel() const { return *m_fresnel; }
	const djb200_abc_data *data() const { return &m_data; }
private:
	vec3 member3(int what, const float_t *a) const
	{
		vec3 r;
		member_batch(what, a, NULL, NULL, 1, &r.x);
		return r;
	}
};

namespace detail {
inline djb200_source describe_source(const brdf &b)
{
	djb200_source s;
	memset(&s, 0, sizeof s);
	if (const merl *m = dynamic_cast<const merl *>(&b)) { s.kind = DJB200_SOURCE_MERL; s.merl = m->handle(); }
	else if (const utia *u = dynamic_cast<const utia *>(&b)) { s.kind = DJB200_SOURCE_UTIA; s.utia = u->handle(); }
	else if (dynamic_cast<const ggx *>(&b) || dynamic_cast<const beckmann *>(&b)) {
		s.kind = DJB200_SOURCE_MICROFACET;
		s.microfacet = static_cast<const microfacet &>(b).describe();
	}
	else if (const sgd *g = dynamic_cast<const sgd *>(&b)) { s.kind = DJB200_SOURCE_SGD; s.sgd = g->data(); }
	else if (const abc *a = dynamic_cast<const abc *>(&b)) { s.kind = DJB200_SOURCE_ABC; s.abc = a->data(); }
	else throw exc("djb_error: this BRDF type cannot be fitted on the device (merl, utia, sgd, abc, ggx, beckmann can)");
	return s;
}
} // namespace detail

// The device-resident copy of a table-backed BRDF (djb::tabular, djb::tabular_anisotropic), created on first use and again
// after set_fresnel / set_shadow.  The reference's query methods are const and data-race free, and its Mitsuba plugins call
// pdf() / sample() on one shared object from every render thread (mitsuba/dj_abc.cpp:77, 89), so the const path here only READS a
// published (handle, revision) pair; creation is serialised by a mutex, and a superseded handle is never destroyed while the
// object lives (another thread may still be launching on it): it is retired and freed by the destructor.
namespace detail {
class device_tables {
	struct node { djb200_tabular *h; unsigned rev; };
	mutable std::atomic<node *> m_cur;
	mutable std::mutex m_mutex;
	mutable std::vector<node *> m_retired;
	device_tables(const device_tables &);
	device_tables &operator=(const device_tables &);
public:
	device_tables() : m_cur(NULL) {}
	~device_tables()
	{
		if (node *n = m_cur.load()) m_retired.push_back(n);
		for (size_t k = 0; k < m_retired.size(); ++k) {
			djb200_tabular_destroy(m_retired[k]->h);
			delete m_retired[k];
		}
	}
	// the handle built for revision `rev`; `make` creates one (called at most once per revision, under the lock)
	template <class Make>
	const djb200_tabular *get(unsigned rev, const Make &make) const
	{
		node *n = m_cur.load(std::memory_order_acquire);
		if (n && n->rev == rev) return n->h;
		std::lock_guard<std::mutex> lock(m_mutex);
		n = m_cur.load(std::memory_order_acquire);
		if (n && n->rev == rev) return n->h;
		node *fresh = new node;
		fresh->h = NULL;
		fresh->rev = rev;
		try { fresh->h = make(); } catch (...) { delete fresh; throw; }
		if (n) m_retired.push_back(n);
		m_cur.store(fresh, std::memory_order_release);
		return fresh->h;
	}
};
} // namespace detail

// ---------------------------------------------------------------------------------------------------
// dj_brdf.h:394-425: the isotropic "power iteration" fit, and -- as in the reference -- a microfacet BRDF of its own: the
// tables are built on the GPU, uploaded once as a device-resident handle, and eval / evalp / pdf / sample / evalp_is run on
// them (normal-map sampling: tabular does not support Smith VNDF sampling, dj_brdf.h:413).
class tabular : public radial {
	std::vector<float_t> m_p22, m_sigma, m_cdf, m_qf, m_residuals;
	std::vector<vec3> m_fresnel_pts;
	float_t m_alpha_beckmann, m_alpha_ggx;
	detail::device_tables m_device;
	tabular() : radial() {}
public:
	tabular(const brdf &source, int resolution, bool shadow = true, int iterations = 4) : radial(fresnel::ideal(), shadow)
	{
		const brdf *src = &source;
		std::vector<tabular *> self(1, this);
		run(&src, 1, resolution, shadow, iterations, self);
	}
	// many materials in one device pass (one CTA per material)
	static std::vector<tabular *> fit_batch(const std::vector<const brdf *> &sources, int resolution, bool shadow = true,
	                                        int iterations = 4)
	{
		std::vector<tabular *> out;
		for (size_t k = 0; k < sources.size(); ++k) out.push_back(new tabular());
		if (!sources.empty()) run(&sources[0], sources.size(), resolution, shadow, iterations, out);
		return out;
	}
	static microfacet::params fit_beckmann_parameters(const tabular &tab) { return microfacet::params::isotropic(tab.m_alpha_beckmann); }
	static microfacet::params fit_ggx_parameters(const tabular &tab) { return microfacet::params::isotropic(tab.m_alpha_ggx); }
	const std::vector<float_t> &get_p22v() const { return m_p22; }
	const std::vector<float_t> &get_sigmav() const { return m_sigma; }
	const std::vector<float_t> &get_cdfv() const { return m_cdf; }
	const std::vector<float_t> &get_qfv() const { return m_qf; }
	const std::vector<float_t> &get_residuals() const { return m_residuals; }
	bool supports_smith_vndf_sampling() const { return false; }

protected:
	int ndf_id() const { return -1; }
	void dispatch(int op, const djb200_params *p, int64_t n_params, djb200_params_layout layout, const float *a, const float *b,
	              size_t n, float *o0, float *o1, float *o2, memory_space where, void *stream) const
	{
		const djb200_tabular *h = upload();
		djb200_status st = DJB200_ERR_INVALID_ARGUMENT;
		switch (op) {
		case 0: st = djb200_tabular_eval(h, p, n_params, layout, a, b, (int64_t)n, o0, where, stream); break;
		case 1: st = djb200_tabular_evalp(h, p, n_params, layout, a, b, (int64_t)n, o0, where, stream); break;
		case 2: st = djb200_tabular_pdf(h, p, n_params, layout, a, b, (int64_t)n, o0, where, stream); break;
		case 3: st = djb200_tabular_sample(h, p, n_params, layout, a, b, (int64_t)n, o0, where, stream); break;
		case 4: st = djb200_tabular_evalp_is(h, p, n_params, layout, a, b, (int64_t)n, o0, o1, o2, where, stream); break;
		}
		detail::check(st);
	}
	const djb200_tabular *radial_handle() const { return upload(); }
	struct make_handle {
		const tabular *t;
		djb200_tabular *operator()() const
		{
			std::vector<float> fr;
			t->fresnel_points(t->m_p22.size(), &fr);
			djb200_tabular_fit f;
			memset(&f, 0, sizeof f);
			f.res = (int32_t)t->m_p22.size();
			f.p22 = const_cast<float *>(&t->m_p22[0]); f.sigma = const_cast<float *>(&t->m_sigma[0]);
			f.cdf = const_cast<float *>(&t->m_cdf[0]); f.qf = const_cast<float *>(&t->m_qf[0]);
			f.fresnel = &fr[0];
			djb200_tabular *h = NULL;
			detail::check(djb200_tabular_create(&f, t->m_shadow ? 1 : 0, &h));
			return h;
		}
	};
	// the device handle for the current Fresnel term / shadowing switch (uploaded by the constructor; again after a setter)
	const djb200_tabular *upload() const
	{
		make_handle mk = {this};
		return m_device.get(m_fresnel_rev, mk);
	}

private:
	static void run(const brdf *const *sources, size_t n, int res, bool shadow, int iterations, std::vector<tabular *> &out)
	{
		DJB_ASSERT(res > 2 && "Invalid Resolution");
		std::vector<djb200_source> src(n);
		std::vector<djb200_tabular_fit> fit(n);
		for (size_t k = 0; k < n; ++k) {
			tabular &t = *out[k];
			src[k] = detail::describe_source(*sources[k]);
			t.m_p22.assign(res, 0); t.m_sigma.assign(res, 0); t.m_cdf.assign(res, 0); t.m_qf.assign(res, 0);
			t.m_residuals.assign(iterations > 0 ? iterations : 1, 0);
			t.m_fresnel_pts.assign(res, vec3(0));
			t.m_shadow = shadow;
			fit[k].res = res;
			fit[k].p22 = &t.m_p22[0]; fit[k].sigma = &t.m_sigma[0]; fit[k].cdf = &t.m_cdf[0]; fit[k].qf = &t.m_qf[0];
			fit[k].fresnel = &t.m_fresnel_pts[0].x;
			fit[k].residuals = &t.m_residuals[0];
		}
		detail::check(djb200_fit_tabular(&src[0], (int32_t)n, res, shadow ? 1 : 0, iterations, &fit[0], NULL));
		for (size_t k = 0; k < n; ++k) {
			out[k]->m_alpha_beckmann = fit[k].alpha_beckmann;
			out[k]->m_alpha_ggx = fit[k].alpha_ggx;
			out[k]->set_fresnel(fresnel::spline(out[k]->m_fresnel_pts)); // get_fresnel() returns the fitted spline
			out[k]->upload(); // eagerly: the const query path then only reads the handle
		}
	}
};

// dj_brdf.h:428-478: the anisotropic fit (elevation x azimuth tables + 5-parameter Beckmann / GGX fits) and, as in the
// reference, a microfacet BRDF of its own: eval / evalp / pdf on the tables, sample / evalp_is by normal-map sampling through
// the marginal / conditional quantile tables, which the device builds when the handle is created.
class tabular_anisotropic : public microfacet {
	std::vector<float_t> m_p22, m_sigma, m_residuals;
	std::vector<vec3> m_fresnel_pts;
	float_t m_beckmann[5], m_ggx[5];
	int m_elevation_res, m_azimuthal_res;
	detail::device_tables m_device;
public:
	tabular_anisotropic(const brdf &source, int elevation_res, int azimuthal_res, bool shadow = true, int iterations = 4)
	    : microfacet(fresnel::ideal(), shadow), m_elevation_res(elevation_res), m_azimuthal_res(azimuthal_res)
	{
		DJB_ASSERT(elevation_res > 1 && azimuthal_res > 1 && "Invalid Resolution");
		djb200_source src = detail::describe_source(source);
		size_t tab = (size_t)elevation_res * azimuthal_res;
		m_p22.assign(tab, 0); m_sigma.assign(tab, 0);
		m_residuals.assign(iterations > 0 ? iterations : 1, 0);
		m_fresnel_pts.assign(elevation_res, vec3(0));
		djb200_tabular_anisotropic_fit fit;
		memset(&fit, 0, sizeof fit);
		fit.elev_res = elevation_res; fit.azim_res = azimuthal_res;
		fit.p22 = &m_p22[0]; fit.sigma = &m_sigma[0]; fit.fresnel = &m_fresnel_pts[0].x; fit.residuals = &m_residuals[0];
		detail::check(djb200_fit_tabular_anisotropic(&src, 1, elevation_res, azimuthal_res, shadow ? 1 : 0, iterations, &fit, NULL));
		memcpy(m_beckmann, fit.beckmann, sizeof m_beckmann);
		memcpy(m_ggx, fit.ggx, sizeof m_ggx);
		set_fresnel(fresnel::spline(m_fresnel_pts)); // get_fresnel() returns the fitted spline (dj_brdf.h:2700)
		upload(); // eagerly: the const query path then only reads the handle
	}
	static microfacet::params fit_beckmann_parameters(const tabular_anisotropic &t)
	{
		return microfacet::params::pdfparams(t.m_beckmann[0], t.m_beckmann[1], t.m_beckmann[2], t.m_beckmann[3], t.m_beckmann[4]);
	}
	static microfacet::params fit_ggx_parameters(const tabular_anisotropic &t)
	{
		return microfacet::params::pdfparams(t.m_ggx[0], t.m_ggx[1], t.m_ggx[2], t.m_ggx[3], t.m_ggx[4]);
	}
	const std::vector<float_t> &get_p22v(int *w = NULL, int *h = NULL) const
	{
		if (w) *w = m_elevation_res;
		if (h) *h = m_azimuthal_res;
		return m_p22;
	}
	const std::vector<float_t> &get_sigmav(int *w = NULL, int *h = NULL) const
	{
		if (w) *w = m_elevation_res;
		if (h) *h = m_azimuthal_res;
		return m_sigma;
	}
	const std::vector<float_t> &get_residuals() const { return m_residuals; }
	bool supports_smith_vndf_sampling() const { return false; }
	// the public table queries, dj_brdf.h:450-455, 2766-2824
	float_t pdf1(float_t phi) const { return table_query(DJB200_MEMBER_PDF1, phi, 0); }
	float_t pdf2(float_t theta, float_t phi) const
	{
		DJB_ASSERT(theta >= 0.0 && "Invalid Angle");
		return table_query(DJB200_MEMBER_PDF2, theta, phi);
	}
	float_t cdf1(float_t phi) const { return table_query(DJB200_MEMBER_CDF1, phi, 0); }
	float_t cdf2(float_t theta, float_t phi) const
	{
		DJB_ASSERT(theta >= 0.0 && "Invalid Angle");
		return table_query(DJB200_MEMBER_CDF2, theta, phi);
	}
	float_t qf1(float_t u1) const
	{
		DJB_ASSERT(u1 >= 0.0 && u1 <= 1.0 && "Invalid Variate");
		return table_query(DJB200_MEMBER_TQF1, u1, 0);
	}
	float_t qf2(float_t u, float_t phi) const
	{
		DJB_ASSERT(u >= 0.0 && u <= 1.0 && "Invalid Variate");
		return table_query(DJB200_MEMBER_TQF2, u, phi);
	}
	// batched (added): b may be NULL for the one-argument members
	void table_query_batch(int what, const float_t *a, const float_t *b, size_t n, float_t *out, memory_space where = host,
	                       void *stream = NULL) const
	{
		detail::check(djb200_tabular_anisotropic_query(upload(), what, a, b, (int64_t)n, out, where, stream));
	}
	// the six sampling tables behind pdf1 / cdf1 / qf1 / pdf2 / cdf2 / qf2 (dj_brdf.h:2766-2824); any pointer may be NULL
	void get_sampling_tables(std::vector<float_t> *pdf1, std::vector<float_t> *cdf1, std::vector<float_t> *qf1,
	                         std::vector<float_t> *pdf2, std::vector<float_t> *cdf2, std::vector<float_t> *qf2) const
	{
		const djb200_tabular *handle = upload();
		std::vector<float_t> *v[6] = {pdf1, cdf1, qf1, pdf2, cdf2, qf2};
		float *p[6];
		for (int k = 0; k < 6; ++k) {
			if (v[k]) v[k]->assign(k < 3 ? (size_t)m_azimuthal_res : m_p22.size(), 0);
			p[k] = v[k] ? &(*v[k])[0] : NULL;
		}
		int32_t counts[2];
		detail::check(djb200_tabular_anisotropic_sampling_tables(handle, p[0], p[1], p[2], p[3], p[4], p[5], counts));
		if (qf1) qf1->resize(counts[0]);
		if (qf2) qf2->resize(counts[1]);
	}

protected:
	int ndf_id() const { return -1; }
	struct make_handle {
		const tabular_anisotropic *t;
		djb200_tabular *operator()() const
		{
			std::vector<float> fr;
			t->fresnel_points((size_t)t->m_elevation_res, &fr);
			djb200_tabular_anisotropic_fit f;
			memset(&f, 0, sizeof f);
			f.elev_res = t->m_elevation_res; f.azim_res = t->m_azimuthal_res;
			f.p22 = const_cast<float *>(&t->m_p22[0]); f.sigma = const_cast<float *>(&t->m_sigma[0]);
			f.fresnel = &fr[0];
			djb200_tabular *h = NULL;
			detail::check(djb200_tabular_anisotropic_create(&f, t->m_shadow ? 1 : 0, &h));
			return h;
		}
	};
	const djb200_tabular *upload() const
	{
		make_handle mk = {this};
		return m_device.get(m_fresnel_rev, mk);
	}
	float_t table_query(int what, float_t a, float_t b) const
	{
		float_t r;
		table_query_batch(what, &a, &b, 1, &r);
		return r;
	}
	void dispatch(int op, const djb200_params *p, int64_t n_params, djb200_params_layout layout, const float *a, const float *b,
	              size_t n, float *o0, float *o1, float *o2, memory_space where, void *stream) const
	{
		const djb200_tabular *h = upload();
		djb200_status st = DJB200_ERR_INVALID_ARGUMENT;
		switch (op) {
		case 0: st = djb200_tabular_eval(h, p, n_params, layout, a, b, (int64_t)n, o0, where, stream); break;
		case 1: st = djb200_tabular_evalp(h, p, n_params, layout, a, b, (int64_t)n, o0, where, stream); break;
		case 2: st = djb200_tabular_pdf(h, p, n_params, layout, a, b, (int64_t)n, o0, where, stream); break;
		case 3: st = djb200_tabular_sample(h, p, n_params, layout, a, b, (int64_t)n, o0, where, stream); break;
		case 4: st = djb200_tabular_evalp_is(h, p, n_params, layout, a, b, (int64_t)n, o0, o1, o2, where, stream); break;
		}
		detail::check(st);
	}
};

// utils/nmap2leanmap.cpp:18-54 (bias = 0) and nmap2leanmap_biased.cpp:23-63 (bias = 25) on raw planar buffers
inline void nmap2leanmap(const uint8_t *nmap_planar_rgb, int w, int h, float_t base_roughness, float_t bias,
                         float_t *leanmap_1, float_t *leanmap_2, memory_space where = host, void *stream = NULL)
{
	detail::check(djb200_nmap_to_leanmap(nmap_planar_rgb, w, h, base_roughness, bias, leanmap_1, leanmap_2, where, stream));
}

// utils/dmap2nmap.cpp:13-44 on raw buffers: 8-bit displacement map [h][w] -> planar 8-bit normal map [3][h][w]
inline void dmap2nmap(const uint8_t *dmap, int w, int h, uint8_t *nmap_planar_rgb, float scale = 0.1f, memory_space where = host,
                      void *stream = NULL)
{
	detail::check(djb200_dmap_to_nmap(dmap, w, h, scale, nmap_planar_rgb, where, stream));
}

} // namespace djb
#endif // DJB200_FACADE_HPP
