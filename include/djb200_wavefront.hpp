// djb200_wavefront.hpp -- wavefront-style BSDF adapters: the per-shading-point work of the reference's Mitsuba plugins
// done for whole arrays of shading records in one device pass each.
//
// A Mitsuba plugin is called once per BSDFSamplingRecord; a wavefront renderer (or a batched light-transport pass) has
// thousands of records in flight.  These classes take the records as structure-of-arrays views and return what the
// plugin's eval() / pdf() / sample() compute *before* Mitsuba's own factors (fresnelConductorExact, specularReflectance),
// which stay in the renderer.  Conventions are the plugins': bRec.wi is the viewer direction `o` of dj_brdf.h, bRec.wo
// the light direction `i` (mitsuba/dj_brdf.cpp:360-361).
#ifndef DJB200_WAVEFRONT_HPP
#define DJB200_WAVEFRONT_HPP

#include "djb200_facade.hpp"

namespace djb {
namespace wavefront {

// n shading records as arrays (host or device memory, `where` says which)
struct records {
	size_t n;
	const vec3 *wi;        // bRec.wi
	const vec3 *wo;        // bRec.wo (eval / pdf)
	const float_t *u;      // n x 2 uniforms (sample)
	const float_t *alpha3; // n x (alpha1, alpha2, alphaAngle) from the roughness textures, or NULL: constants
	const float_t *lean5;  // n x (E1..E5) as fetched from leanmap1.rg / leanmap2.rgb, or NULL: no LEAN maps
	memory_space where;
	void *stream;
	records() : n(0), wi(NULL), wo(NULL), u(NULL), alpha3(NULL), lean5(NULL), where(host), stream(NULL) {}
};

// mitsuba/dj_beckmannconductor.cpp: Beckmann conductor with LEAN-filtered normal / displacement maps
class beckmann_conductor {
	beckmann m_brdf;
	djb200_lean_shading m_cfg;
public:
	// alpha1 / alpha2 / alpha_angle: the plugin's constant roughness (used where records.alpha3 is NULL);
	// the other arguments are its "leanFiltering" and "dmapscale" properties and the BIAS constant (:300)
	beckmann_conductor(float_t alpha1, float_t alpha2, float_t alpha_angle, bool lean_filtering = true, float_t dmap_scale = 1,
	                   float_t bias = 25, const fresnel::impl &f = fresnel::ideal())
	    : m_brdf(f), m_cfg(microfacet::lean_config(alpha1, alpha2, alpha_angle, bias, dmap_scale, lean_filtering))
	{
	}
	// eval(), :283-319: fr_cos for every record
	void eval(const records &r, vec3 *out) const
	{
		m_brdf.evalp_lean_batch(m_cfg, r.alpha3, r.lean5, r.wo, r.wi, r.n, out, r.where, r.stream);
	}
	// pdf(), :338-366
	void pdf(const records &r, float_t *out) const
	{
		m_brdf.pdf_lean_batch(m_cfg, r.alpha3, r.lean5, r.wo, r.wi, r.n, out, r.where, r.stream);
	}
	// sample(), :379-410: importance weight fr_cos / pdf, the sampled bRec.wo and its pdf
	void sample(const records &r, vec3 *out_weight, vec3 *out_wo, float_t *out_pdf) const
	{
		m_brdf.evalp_is_lean_batch(m_cfg, r.alpha3, r.lean5, r.u, r.wi, r.n, out_weight, out_wo, out_pdf, r.where, r.stream);
	}
};

// mitsuba/dj_brdf.cpp with textured roughness: params::elliptic per record (:353-357), then evalp / pdf / evalp_is.
// `brdf` is any microfacet BRDF of the facade (ggx, beckmann, tabular, tabular_anisotropic).
class rough_microfacet {
	const microfacet &m_brdf;
public:
	explicit rough_microfacet(const microfacet &b) : m_brdf(b) {}
	// params blocks for the records' (alpha1, alpha2, alphaAngle): the elliptic factory is host arithmetic
	static void make_params(const float_t *alpha3, size_t n, std::vector<microfacet::params> *out)
	{
		out->clear();
		out->reserve(n);
		for (size_t k = 0; k < n; ++k) out->push_back(microfacet::params::elliptic(alpha3[3 * k], alpha3[3 * k + 1], alpha3[3 * k + 2]));
	}
	void eval(const records &r, const microfacet::params *per_record, vec3 *out) const
	{
		m_brdf.evalp_batch(r.wo, r.wi, r.n, out, per_record, r.n, DJB200_PARAMS_PER_PAIR, r.where, r.stream);
	}
	void pdf(const records &r, const microfacet::params *per_record, float_t *out) const
	{
		m_brdf.pdf_batch(r.wo, r.wi, r.n, out, per_record, r.n, DJB200_PARAMS_PER_PAIR, r.where, r.stream);
	}
	void sample(const records &r, const microfacet::params *per_record, vec3 *out_weight, vec3 *out_wo, float_t *out_pdf) const
	{
		m_brdf.evalp_is_batch(r.u, r.wi, r.n, out_weight, out_wo, out_pdf, per_record, r.n, DJB200_PARAMS_PER_PAIR, r.where, r.stream);
	}
};

} // namespace wavefront
} // namespace djb
#endif // DJB200_WAVEFRONT_HPP
