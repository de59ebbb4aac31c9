// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" batch wrappers around the UNMODIFIED reference header, which is
// compiled in place from /root/reference (never copied into this repo).  The result,
// oracle/_ref/libdjbref.so, is (a) the ground truth the C restatement in
// oracle/djb_oracle.c is pinned against, (b) the generator of tests/golden/*, and
// (c) the "reference" CPU baseline timed by bench.py.
//
// Build recipe: oracle/Makefile.  The TU is pinned as SURVEY.md section 0 finding 2
// demands: only <c...> headers before the reference (no <math.h>, so the unqualified
// sqrt/exp/cos/... calls inside namespace djb bind to the C double functions), and
// g++ -O3 -ffp-contract=off -DNVERBOSE (no -march=native, no -ffast-math).
//
// Nothing in the product path (dj_brdf_b200/, include/) may link or load this file.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <string>
#include <thread>
#include <exception>

#define DJ_BRDF_IMPLEMENTATION 1
#include "dj_brdf.h" // found through -I/root/reference

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

// run fn(begin, end) on nthreads disjoint index ranges (all reference queries are const)
template <typename F>
void parallel_ranges(int64_t n, int nthreads, F fn)
{
	if (nthreads <= 1 || n < 2 * (int64_t)nthreads) {
		fn((int64_t)0, n);
		return;
	}
	std::vector<std::thread> pool;
	int64_t chunk = (n + nthreads - 1) / nthreads;
	for (int t = 0; t < nthreads; ++t) {
		int64_t b = t * chunk, e = b + chunk < n ? b + chunk : n;
		if (b >= e) break;
		pool.emplace_back([=]() { fn(b, e); });
	}
	for (size_t t = 0; t < pool.size(); ++t) pool[t].join();
}

inline djb::vec3 ld3(const float *p, int64_t k) { return djb::vec3(p[3 * k], p[3 * k + 1], p[3 * k + 2]); }
inline void st3(float *p, int64_t k, const djb::vec3 &v) { p[3 * k] = v.x; p[3 * k + 1] = v.y; p[3 * k + 2] = v.z; }

djb::fresnel::impl *make_fresnel(int kind, const float *d, int nd)
{
	switch (kind) {
	case 0: return new djb::fresnel::ideal();
	case 1: return new djb::fresnel::schlick(djb::vec3(d[0], d[1], d[2]));
	case 2: return new djb::fresnel::unpolarized(djb::vec3(d[0], d[1], d[2]));
	case 3: return new djb::fresnel::sgd(djb::vec3(d[0], d[1], d[2]), djb::vec3(d[3], d[4], d[5]));
	case 4: {
		std::vector<djb::vec3> pts;
		for (int i = 0; i < nd / 3; ++i) pts.push_back(djb::vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
		return new djb::fresnel::spline(pts);
	}
	}
	return NULL;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// params (dj_brdf.h:213-243, 1399-1474): the object is 12 floats; exported raw.
REF_API int ref_sizeof_params() { return (int)sizeof(djb::microfacet::params); }

REF_API void ref_params_elliptic(float a1, float a2, float phi_a, float *out12)
{
	djb::microfacet::params p = djb::microfacet::params::elliptic(a1, a2, phi_a);
	memcpy(out12, &p, sizeof(p));
}

REF_API void ref_params_pdfparams(float ax, float ay, float rho, float tx, float ty, float *out12)
{
	djb::microfacet::params p = djb::microfacet::params::pdfparams(ax, ay, rho, tx, ty);
	memcpy(out12, &p, sizeof(p));
}

// ---------------------------------------------------------------------------------------------
// object factories; every handle is a djb::brdf*
REF_API void *ref_microfacet_create(int ndf, int fresnel_kind, const float *fdata, int nf, int shadow)
{
	djb::fresnel::impl *f = make_fresnel(fresnel_kind, fdata, nf);
	if (!f) return NULL;
	djb::brdf *b = NULL;
	if (ndf == 0) b = new djb::beckmann(*f, shadow != 0);
	else if (ndf == 1) b = new djb::ggx(*f, shadow != 0);
	delete f;
	return b;
}

REF_API void *ref_merl_open(const char *path)
{
	try { return static_cast<djb::brdf *>(new djb::merl(path)); } catch (std::exception &e) {
		fprintf(stderr, "ref_merl_open: %s\n", e.what());
		return NULL;
	}
}

REF_API void *ref_utia_open(const char *path)
{
	try { return static_cast<djb::brdf *>(new djb::utia(path)); } catch (std::exception &e) {
		fprintf(stderr, "ref_utia_open: %s\n", e.what());
		return NULL;
	}
}

REF_API void *ref_sgd_create(const char *name)
{
	try { return static_cast<djb::brdf *>(new djb::sgd(name)); } catch (std::exception &) { return NULL; }
}

REF_API void *ref_abc_create(const char *name)
{
	try { return static_cast<djb::brdf *>(new djb::abc(name)); } catch (std::exception &) { return NULL; }
}

REF_API void ref_brdf_destroy(void *h) { delete static_cast<djb::brdf *>(h); }

// ---------------------------------------------------------------------------------------------
// batched queries through the reference's virtual interface (dj_brdf.h:74-109)
// params12: NULL (=> reference default) or one 48-byte params block shared by the batch.
REF_API void ref_brdf_eval(void *h, const float *params12, const float *wi, const float *wo,
                           int64_t n, float *out3, int nthreads)
{
	const djb::brdf *b = static_cast<djb::brdf *>(h);
	parallel_ranges(n, nthreads, [=](int64_t s, int64_t e) {
		for (int64_t k = s; k < e; ++k) st3(out3, k, b->eval(ld3(wi, k), ld3(wo, k), params12));
	});
}

REF_API void ref_brdf_evalp(void *h, const float *params12, const float *wi, const float *wo,
                            int64_t n, float *out3, int nthreads)
{
	const djb::brdf *b = static_cast<djb::brdf *>(h);
	parallel_ranges(n, nthreads, [=](int64_t s, int64_t e) {
		for (int64_t k = s; k < e; ++k) st3(out3, k, b->evalp(ld3(wi, k), ld3(wo, k), params12));
	});
}

REF_API void ref_brdf_pdf(void *h, const float *params12, const float *wi, const float *wo,
                          int64_t n, float *out1, int nthreads)
{
	const djb::brdf *b = static_cast<djb::brdf *>(h);
	parallel_ranges(n, nthreads, [=](int64_t s, int64_t e) {
		for (int64_t k = s; k < e; ++k) out1[k] = b->pdf(ld3(wi, k), ld3(wo, k), params12);
	});
}

REF_API void ref_brdf_sample(void *h, const float *params12, const float *u2, const float *wo,
                             int64_t n, float *out3, int nthreads)
{
	const djb::brdf *b = static_cast<djb::brdf *>(h);
	parallel_ranges(n, nthreads, [=](int64_t s, int64_t e) {
		for (int64_t k = s; k < e; ++k)
			st3(out3, k, b->sample(u2[2 * k], u2[2 * k + 1], ld3(wo, k), params12));
	});
}

REF_API void ref_brdf_evalp_is(void *h, const float *params12, const float *u2, const float *wo,
                               int64_t n, float *out_w3, float *out_i3, float *out_pdf, int nthreads)
{
	const djb::brdf *b = static_cast<djb::brdf *>(h);
	parallel_ranges(n, nthreads, [=](int64_t s, int64_t e) {
		for (int64_t k = s; k < e; ++k) {
			// the reference leaves *i untouched when G <= 0 (dj_brdf.h:1749-1764): pre-fill with 0
			djb::vec3 i(0);
			djb::float_t pdf = 0;
			djb::vec3 w = b->evalp_is(u2[2 * k], u2[2 * k + 1], ld3(wo, k), &i, &pdf, params12);
			st3(out_w3, k, w);
			st3(out_i3, k, i);
			out_pdf[k] = pdf;
		}
	});
}

// microfacet component queries (dj_brdf.h:258-272), for stage-level parity
REF_API void ref_microfacet_components(void *h, const float *params12, const float *wi, const float *wo,
                                       int64_t n, float *out_ndf, float *out_gaf, float *out_sigma_o)
{
	const djb::microfacet *m = dynamic_cast<djb::microfacet *>(static_cast<djb::brdf *>(h));
	if (!m) return;
	djb::microfacet::params p = djb::microfacet::params::standard();
	if (params12) memcpy(&p, params12, sizeof(p));
	for (int64_t k = 0; k < n; ++k) {
		djb::vec3 i = ld3(wi, k), o = ld3(wo, k);
		djb::vec3 hv = djb::normalize(i + o);
		out_ndf[k] = m->ndf(hv, p);
		out_gaf[k] = m->gaf(hv, i, o, p);
		out_sigma_o[k] = m->sigma(o, p);
	}
}

// all eight public component queries (dj_brdf.h:258-272): what = 0 ndf(h = a), 1 gaf(a, b, c), 2 g1(a, b), 3 sigma(a),
// 4 p22(a.x, a.y), 5 vp22(a.x, a.y, b), 6 vndf(a, b), 7 fresnel(a.x) -> rgb
REF_API int ref_microfacet_component(void *h, const float *params12, int what, const float *a, const float *b, const float *c,
                                     int64_t n, float *out)
{
	const djb::microfacet *m = dynamic_cast<djb::microfacet *>(static_cast<djb::brdf *>(h));
	if (!m) return -1;
	djb::microfacet::params p = djb::microfacet::params::standard();
	if (params12) memcpy(&p, params12, sizeof(p));
	for (int64_t k = 0; k < n; ++k) {
		djb::vec3 va = ld3(a, k), vb = b ? ld3(b, k) : djb::vec3(0, 0, 1), vc = c ? ld3(c, k) : djb::vec3(0, 0, 1);
		switch (what) {
		case 0: out[k] = m->ndf(va, p); break;
		case 1: out[k] = m->gaf(va, vb, vc, p); break;
		case 2: out[k] = m->g1(va, vb, p); break;
		case 3: out[k] = m->sigma(va, p); break;
		case 4: out[k] = m->p22(va.x, va.y, p); break;
		case 5: out[k] = m->vp22(va.x, va.y, vb, p); break;
		case 6: out[k] = m->vndf(va, vb, p); break;
		default: st3(out, k, m->fresnel(va.x)); break;
		}
	}
	return 0;
}

// ---------------------------------------------------------------------------------------------
// MERL cell index exactly as merl::eval forms it (dj_brdf.h:990-1002); uses the reference's own
// static helpers, which are visible because this is the implementation TU.
REF_API void ref_merl_index(const float *wi, const float *wo, int64_t n, int32_t *idx, int nthreads)
{
	parallel_ranges(n, nthreads, [=](int64_t s, int64_t e) {
		for (int64_t k = s; k < e; ++k) {
			djb::vec3 hv, dv;
			djb::float_t th, ph, td, pd;
			djb::brdf::io_to_hd(ld3(wi, k), ld3(wo, k), &hv, &dv);
			djb::xyz_to_theta_phi(hv, &th, &ph);
			djb::xyz_to_theta_phi(dv, &td, &pd);
			idx[k] = djb::phi_diff_index(pd)
			       + djb::theta_diff_index(td) * MERL_SAMPLING_RES_PHI_D / 2
			       + djb::theta_half_index(th) * MERL_SAMPLING_RES_PHI_D / 2 * MERL_SAMPLING_RES_THETA_D;
		}
	});
}

REF_API void ref_io_to_hd(const float *wi, const float *wo, int64_t n, float *h3, float *d3)
{
	for (int64_t k = 0; k < n; ++k) {
		djb::vec3 hv, dv;
		djb::brdf::io_to_hd(ld3(wi, k), ld3(wo, k), &hv, &dv);
		st3(h3, k, hv);
		st3(d3, k, dv);
	}
}

REF_API void ref_hd_to_io(const float *h3, const float *d3, int64_t n, float *wi, float *wo)
{
	for (int64_t k = 0; k < n; ++k) {
		djb::vec3 iv, ov;
		djb::brdf::hd_to_io(ld3(h3, k), ld3(d3, k), &iv, &ov);
		st3(wi, k, iv);
		st3(wo, k, ov);
	}
}

// ---------------------------------------------------------------------------------------------
// LEAN algebra (dj_brdf.h:1965-1990)
REF_API void ref_lrep_to_params(const float *E5, int64_t n, float *out12)
{
	for (int64_t k = 0; k < n; ++k) {
		const float *E = E5 + 5 * k;
		djb::beckmann::lrep l(E[0], E[1], E[2], E[3], E[4]);
		djb::microfacet::params p;
		djb::beckmann::lrep_to_params(l, &p);
		memcpy(out12 + 12 * k, &p, sizeof(p));
	}
}

REF_API void ref_params_to_lrep(const float *params12, int64_t n, float *E5)
{
	for (int64_t k = 0; k < n; ++k) {
		djb::microfacet::params p;
		memcpy(&p, params12 + 12 * k, sizeof(p));
		djb::beckmann::lrep l;
		djb::beckmann::params_to_lrep(p, &l);
		memcpy(E5 + 5 * k, &l, 5 * sizeof(float));
	}
}

// The statements of mitsuba/dj_beckmannconductor.cpp:283-314 on the reference's own classes (Float = float,
// BIAS = 25 there; a parameter here so that unbiased maps can be checked too).
REF_API void ref_lean_shading_params(float bias, float dmap_scale, int lean_filtering, int alpha_per_pair,
                                     const float *alpha, const float *E5, int64_t n, float *out12)
{
	for (int64_t k = 0; k < n; ++k) {
		const float *a = alpha + (alpha_per_pair ? 3 * k : 0), *E = E5 + 5 * k;
		djb::microfacet::params params = djb::microfacet::params::elliptic(a[0], a[1], a[2]);
		float E1 = E[0], E2 = E[1], E3 = E[2], E4 = E[3], E5v = E[4];
		const float BIAS = bias;
		E1 -= BIAS;
		E2 -= BIAS;
		E5v -= BIAS * BIAS;
		djb::beckmann::lrep lrep1, lrep2;
		if (lean_filtering) lrep1 = djb::beckmann::lrep(E1, E2, E3, E4, E5v);
		else lrep1 = djb::beckmann::lrep(E1, E2, E1 * E1, E2 * E2, E1 * E2);
		lrep1 *= dmap_scale;
		djb::beckmann::params_to_lrep(params, &lrep2);
		djb::beckmann::lrep_to_params(lrep1 + lrep2, &params);
		memcpy(out12 + 12 * k, &params, sizeof(params));
	}
}

// ---------------------------------------------------------------------------------------------
// isotropic fit (dj_brdf.h:2215-2236, 3133-3184).  Returns a djb::brdf* that is a djb::tabular.
REF_API void *ref_tabular_create(void *src, int res, int shadow)
{
	const djb::brdf *b = static_cast<djb::brdf *>(src);
	return static_cast<djb::brdf *>(new djb::tabular(*b, res, shadow != 0));
}

// out: p22[res], sigma[res], cdf[res], qf[res], fresnel[3*res], alpha[2] = (beckmann, ggx)
REF_API int ref_tabular_get(void *tabh, float *p22, float *sigma, float *cdf, float *qf,
                            float *fresnel3, float *alpha2)
{
	const djb::tabular *t = dynamic_cast<djb::tabular *>(static_cast<djb::brdf *>(tabh));
	if (!t) return -1;
	int n = (int)t->get_p22v().size();
	if (p22) memcpy(p22, &t->get_p22v()[0], sizeof(float) * n);
	if (sigma) memcpy(sigma, &t->get_sigmav()[0], sizeof(float) * t->get_sigmav().size());
	if (cdf) memcpy(cdf, &t->get_cdfv()[0], sizeof(float) * t->get_cdfv().size());
	if (qf) memcpy(qf, &t->get_qfv()[0], sizeof(float) * t->get_qfv().size());
	if (fresnel3) {
		const djb::fresnel::spline *s = dynamic_cast<const djb::fresnel::spline *>(&t->get_fresnel());
		if (!s) return -2;
		for (size_t i = 0; i < s->get_points().size(); ++i) st3(fresnel3, (int64_t)i, s->get_points()[i]);
	}
	if (alpha2) {
		djb::float_t a, dummy;
		djb::tabular::fit_beckmann_parameters(*t).get_ellipse(&a, &dummy, NULL);
		alpha2[0] = a;
		djb::tabular::fit_ggx_parameters(*t).get_ellipse(&a, &dummy, NULL);
		alpha2[1] = a;
	}
	return n;
}

// ---------------------------------------------------------------------------------------------
// anisotropic fit (dj_brdf.h:2238-2273, 3186-3307)
REF_API void *ref_tabular_anisotropic_create(void *src, int elev_res, int azim_res, int shadow)
{
	const djb::brdf *b = static_cast<djb::brdf *>(src);
	return static_cast<djb::brdf *>(new djb::tabular_anisotropic(*b, elev_res, azim_res, shadow != 0));
}

// out: p22[w*h], sigma[w*h], fresnel[3*w], beckmann5 / ggx5 = (ax, ay, rho, tx, ty)
REF_API int ref_tabular_anisotropic_get(void *tabh, float *p22, float *sigma, float *fresnel3,
                                        float *beckmann5, float *ggx5)
{
	const djb::tabular_anisotropic *t =
		dynamic_cast<djb::tabular_anisotropic *>(static_cast<djb::brdf *>(tabh));
	if (!t) return -1;
	int w, h;
	const std::vector<djb::float_t> &pv = t->get_p22v(&w, &h);
	if (p22) memcpy(p22, &pv[0], sizeof(float) * pv.size());
	const std::vector<djb::float_t> &sv = t->get_sigmav(NULL, NULL);
	if (sigma) memcpy(sigma, &sv[0], sizeof(float) * sv.size());
	if (fresnel3) {
		const djb::fresnel::spline *s = dynamic_cast<const djb::fresnel::spline *>(&t->get_fresnel());
		if (!s) return -2;
		for (size_t i = 0; i < s->get_points().size(); ++i) st3(fresnel3, (int64_t)i, s->get_points()[i]);
	}
	if (beckmann5)
		djb::tabular_anisotropic::fit_beckmann_parameters(*t).get_pdfparams(
			beckmann5, beckmann5 + 1, beckmann5 + 2, beckmann5 + 3, beckmann5 + 4);
	if (ggx5)
		djb::tabular_anisotropic::fit_ggx_parameters(*t).get_pdfparams(
			ggx5, ggx5 + 1, ggx5 + 2, ggx5 + 3, ggx5 + 4);
	return (int)pv.size();
}

// the four public scalar queries of djb::radial (dj_brdf.h:307-310): what = 0 p22_radial, 1 sigma_std_radial, 2 cdf_radial,
// 3 qf_radial
REF_API int ref_radial_query(void *h, int what, const float *x, int n, float *out)
{
	const djb::radial *r = dynamic_cast<djb::radial *>(static_cast<djb::brdf *>(h));
	if (!r) return -1;
	for (int k = 0; k < n; ++k) {
		switch (what) {
		case 0: out[k] = r->p22_radial(x[k]); break;
		case 1: out[k] = r->sigma_std_radial(x[k]); break;
		case 2: out[k] = r->cdf_radial(x[k]); break;
		case 3: out[k] = r->qf_radial(x[k]); break;
		default: return -2;
		}
	}
	return 0;
}

// radial queries used by tests/plot_qf.cpp and tests/plot_cdf.cpp of the reference
REF_API int ref_radial_qf_cdf(void *h, const float *x, int n, float *qf_out, float *cdf_out)
{
	const djb::radial *r = dynamic_cast<djb::radial *>(static_cast<djb::brdf *>(h));
	if (!r) return -1;
	for (int k = 0; k < n; ++k) {
		if (qf_out) qf_out[k] = r->qf_radial(x[k]);
		if (cdf_out) cdf_out[k] = r->cdf_radial(x[k]);
	}
	return 0;
}

// the remaining public scalar members (dj_brdf.h:366-369, 384-389, 506-509, 531-533).  what: 0 qf1(a), 1 qf2_radial(a, b, c),
// 2 qf3_radial(a, b) on a ggx / beckmann handle; 20 ndf(a3), 21 gaf(a3, b3, c3), 22 g1(a3), 23 fresnel(a) on an sgd / abc
// handle (vec3 arguments as n x 3; vec3 results n x 3; abc::gaf returns a scalar, n x 1).  Returns floats written per item.
REF_API int ref_member_query(void *h, int what, const float *a, const float *b, const float *c, int n, float *out)
{
	djb::brdf *base = static_cast<djb::brdf *>(h);
	if (what <= 2) {
		const djb::beckmann *bk = dynamic_cast<djb::beckmann *>(base);
		const djb::ggx *gg = dynamic_cast<djb::ggx *>(base);
		if (!bk && !gg) return -1;
		for (int k = 0; k < n; ++k) {
			if (what == 0) out[k] = bk ? bk->qf1(a[k]) : gg->qf1(a[k]);
			else if (what == 1) out[k] = bk ? bk->qf2_radial(a[k], b[k], c[k]) : gg->qf2_radial(a[k], b[k], c[k]);
			else out[k] = bk ? bk->qf3_radial(a[k], b[k]) : gg->qf3_radial(a[k], b[k]);
		}
		return 1;
	}
	if (const djb::sgd *s = dynamic_cast<djb::sgd *>(base)) {
		for (int k = 0; k < n; ++k) {
			djb::vec3 r;
			switch (what) {
			case 20: r = s->ndf(ld3(a, k)); break;
			case 21: r = s->gaf(ld3(a, k), ld3(b, k), ld3(c, k)); break;
			case 22: r = s->g1(ld3(a, k)); break;
			case 23: r = s->fresnel(a[k]); break;
			default: return -2;
			}
			st3(out, k, r);
		}
		return 3;
	}
	if (const djb::abc *s = dynamic_cast<djb::abc *>(base)) {
		for (int k = 0; k < n; ++k) {
			if (what == 21) { out[k] = s->gaf(ld3(a, k), ld3(b, k), ld3(c, k)); continue; }
			djb::vec3 r;
			if (what == 20) r = s->ndf(ld3(a, k));
			else if (what == 23) r = s->fresnel(a[k]);
			else return -2;
			st3(out, k, r);
		}
		return what == 21 ? 1 : 3;
	}
	return -1;
}
