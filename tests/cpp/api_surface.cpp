// tests/cpp/api_surface.cpp -- compile-only: every public member of the reference's djb:: interface that the facade
// provides, pinned to the reference's exact signature (pointer-to-member casts).  The same file compiles against the
// reference header (-DUSE_REFERENCE) and against include/compat/dj_brdf.h, so a signature that drifts fails the build.
#ifdef USE_REFERENCE
#define DJ_BRDF_IMPLEMENTATION 1
#endif
#include "dj_brdf.h"

using namespace djb;
typedef microfacet::params P;

template <class T> static void is(T) {}
#define MEMBER(expr, ...) is<__VA_ARGS__>(expr);

int main()
{
	// brdf (dj_brdf.h:74-109)
	MEMBER(&brdf::eval, vec3 (brdf::*)(const vec3 &, const vec3 &, const void *) const)
	MEMBER(&brdf::eval_hd, vec3 (brdf::*)(const vec3 &, const vec3 &, const void *) const)
	MEMBER(&brdf::evalp, vec3 (brdf::*)(const vec3 &, const vec3 &, const void *) const)
	MEMBER(&brdf::evalp_hd, vec3 (brdf::*)(const vec3 &, const vec3 &, const void *) const)
	MEMBER(&brdf::evalp_is, vec3 (brdf::*)(float_t, float_t, const vec3 &, vec3 *, float_t *, const void *) const)
	MEMBER(&brdf::sample, vec3 (brdf::*)(float_t, float_t, const vec3 &, const void *) const)
	MEMBER(&brdf::pdf, float_t (brdf::*)(const vec3 &, const vec3 &, const void *) const)
	MEMBER(&brdf::io_to_hd, void (*)(const vec3 &, const vec3 &, vec3 *, vec3 *))
	MEMBER(&brdf::hd_to_io, void (*)(const vec3 &, const vec3 &, vec3 *, vec3 *))
	// fresnel (dj_brdf.h:149-207)
	MEMBER(&fresnel::ior_to_f0, void (*)(float_t, float_t *))
	MEMBER(&fresnel::ior_to_f0, void (*)(const vec3 &, vec3 *))
	MEMBER(&fresnel::f0_to_ior, void (*)(float_t, float_t *))
	MEMBER(&fresnel::f0_to_ior, void (*)(const vec3 &, vec3 *))
	MEMBER(&fresnel::impl::eval, vec3 (fresnel::impl::*)(float_t) const)
	MEMBER(&fresnel::impl::copy, fresnel::impl *(fresnel::impl::*)() const)
	MEMBER(&fresnel::spline::get_points, const std::vector<vec3> &(fresnel::spline::*)() const)
	{
		fresnel::ideal a;
		fresnel::unpolarized b(vec3(1.5f));
		fresnel::schlick c(vec3(0.04f));
		fresnel::sgd d(vec3(0.1f), vec3(0.2f));
		fresnel::spline e(std::vector<vec3>(4, vec3(1)));
		(void)a; (void)b; (void)c; (void)d; (void)e;
	}
	// microfacet::params (dj_brdf.h:213-243)
	MEMBER(&P::standard, P (*)())
	MEMBER(&P::isotropic, P (*)(float_t))
	MEMBER(&P::elliptic, P (*)(float_t, float_t, float_t))
	MEMBER(&P::pdfparams, P (*)(float_t, float_t, float_t, float_t, float_t))
	MEMBER(&P::set_ellipse, void (P::*)(float_t, float_t, float_t))
	MEMBER(&P::set_pdfparams, void (P::*)(float_t, float_t, float_t, float_t, float_t))
	MEMBER(&P::set_location, void (P::*)(float_t, float_t))
	MEMBER(&P::set_location, void (P::*)(const vec3 &))
	MEMBER(&P::get_ellipse, void (P::*)(float_t *, float_t *, float_t *) const)
	MEMBER(&P::get_pdfparams, void (P::*)(float_t *, float_t *, float_t *, float_t *, float_t *) const)
	MEMBER(&P::get_location, void (P::*)(float_t *, float_t *) const)
	MEMBER(&P::get_location, void (P::*)(vec3 *) const)
	{
		P a, b(0.1f, 0.2f, 0.3f), c(0.1f, 0.2f, 0.0f, 0.0f, 0.0f);
		(void)a; (void)b; (void)c;
	}
	// microfacet (dj_brdf.h:245-298)
	MEMBER(&microfacet::eval, vec3 (microfacet::*)(const vec3 &, const vec3 &, const void *) const)
	MEMBER(&microfacet::evalp, vec3 (microfacet::*)(const vec3 &, const vec3 &, const void *) const)
	MEMBER(&microfacet::sample, vec3 (microfacet::*)(float_t, float_t, const vec3 &, const void *) const)
	MEMBER(&microfacet::pdf, float_t (microfacet::*)(const vec3 &, const vec3 &, const void *) const)
	MEMBER(&microfacet::evalp_is, vec3 (microfacet::*)(float_t, float_t, const vec3 &, vec3 *, float_t *, const void *) const)
	MEMBER(&microfacet::fresnel, vec3 (microfacet::*)(float_t) const)
	MEMBER(&microfacet::ndf, float_t (microfacet::*)(const vec3 &, const P &) const)
	MEMBER(&microfacet::gaf, float_t (microfacet::*)(const vec3 &, const vec3 &, const vec3 &, const P &) const)
	MEMBER(&microfacet::g1, float_t (microfacet::*)(const vec3 &, const vec3 &, const P &) const)
	MEMBER(&microfacet::sigma, float_t (microfacet::*)(const vec3 &, const P &) const)
	MEMBER(&microfacet::p22, float_t (microfacet::*)(float_t, float_t, const P &) const)
	MEMBER(&microfacet::vp22, float_t (microfacet::*)(float_t, float_t, const vec3 &, const P &) const)
	MEMBER(&microfacet::vndf, float_t (microfacet::*)(const vec3 &, const vec3 &, const P &) const)
	MEMBER(&microfacet::supports_smith_vndf_sampling, bool (microfacet::*)() const)
	MEMBER(&microfacet::qf2, float_t (microfacet::*)(float_t, const vec3 &) const)
	MEMBER(&microfacet::qf3, float_t (microfacet::*)(float_t, const vec3 &, float_t) const)
	MEMBER(&microfacet::set_shadow, void (microfacet::*)(bool))
	MEMBER(&microfacet::set_fresnel, void (microfacet::*)(const fresnel::impl &))
	MEMBER(&microfacet::get_shadow, int (microfacet::*)() const)
	MEMBER(&microfacet::get_fresnel, const fresnel::impl &(microfacet::*)() const)
	// radial families (dj_brdf.h:300-391)
	MEMBER(&radial::p22_radial, float_t (radial::*)(float_t) const)
	MEMBER(&radial::sigma_std_radial, float_t (radial::*)(float_t) const)
	MEMBER(&radial::cdf_radial, float_t (radial::*)(float_t) const)
	MEMBER(&radial::qf_radial, float_t (radial::*)(float_t) const)
	MEMBER(&radial::qf2_radial, float_t (radial::*)(float_t, float_t, float_t) const)
	MEMBER(&radial::qf3_radial, float_t (radial::*)(float_t, float_t) const)
	MEMBER(&beckmann::qf1, float_t (beckmann::*)(float_t) const)
	MEMBER(&beckmann::qf2_radial, float_t (beckmann::*)(float_t, float_t, float_t) const)
	MEMBER(&beckmann::qf3_radial, float_t (beckmann::*)(float_t, float_t) const)
	MEMBER(&ggx::qf1, float_t (ggx::*)(float_t) const)
	MEMBER(&ggx::qf2_radial, float_t (ggx::*)(float_t, float_t, float_t) const)
	MEMBER(&ggx::qf3_radial, float_t (ggx::*)(float_t, float_t) const)
	MEMBER(&beckmann::params_to_lrep, void (*)(const P &, beckmann::lrep *))
	MEMBER(&beckmann::lrep_to_params, void (*)(const beckmann::lrep &, P *))
	MEMBER(&beckmann::lrep::operator+, beckmann::lrep (beckmann::lrep::*)(const beckmann::lrep &) const)
	MEMBER(&beckmann::lrep::operator*, beckmann::lrep (beckmann::lrep::*)(float_t) const)
	MEMBER(&beckmann::lrep::operator+=, beckmann::lrep &(beckmann::lrep::*)(const beckmann::lrep &))
	MEMBER(&beckmann::lrep::operator*=, beckmann::lrep &(beckmann::lrep::*)(float_t))
	MEMBER(&beckmann::lrep::shear, void (beckmann::lrep::*)(float_t, float_t))
	MEMBER(&beckmann::lrep::scale, void (beckmann::lrep::*)(float_t, float_t))
	// fits (dj_brdf.h:394-478)
	MEMBER(&tabular::fit_beckmann_parameters, P (*)(const tabular &))
	MEMBER(&tabular::fit_ggx_parameters, P (*)(const tabular &))
	MEMBER(&tabular::get_p22v, const std::vector<float_t> &(tabular::*)() const)
	MEMBER(&tabular::get_sigmav, const std::vector<float_t> &(tabular::*)() const)
	MEMBER(&tabular::get_cdfv, const std::vector<float_t> &(tabular::*)() const)
	MEMBER(&tabular::get_qfv, const std::vector<float_t> &(tabular::*)() const)
	MEMBER(&tabular_anisotropic::fit_beckmann_parameters, P (*)(const tabular_anisotropic &))
	MEMBER(&tabular_anisotropic::fit_ggx_parameters, P (*)(const tabular_anisotropic &))
	MEMBER(&tabular_anisotropic::get_p22v, const std::vector<float_t> &(tabular_anisotropic::*)(int *, int *) const)
	MEMBER(&tabular_anisotropic::get_sigmav, const std::vector<float_t> &(tabular_anisotropic::*)(int *, int *) const)
	MEMBER(&tabular_anisotropic::pdf1, float_t (tabular_anisotropic::*)(float_t) const)
	MEMBER(&tabular_anisotropic::pdf2, float_t (tabular_anisotropic::*)(float_t, float_t) const)
	MEMBER(&tabular_anisotropic::cdf1, float_t (tabular_anisotropic::*)(float_t) const)
	MEMBER(&tabular_anisotropic::cdf2, float_t (tabular_anisotropic::*)(float_t, float_t) const)
	MEMBER(&tabular_anisotropic::qf1, float_t (tabular_anisotropic::*)(float_t) const)
	MEMBER(&tabular_anisotropic::qf2, float_t (tabular_anisotropic::*)(float_t, float_t) const)
	// data-driven BRDFs (dj_brdf.h:126-146, 481-535): construction is by file / material name
	MEMBER(&merl::eval, vec3 (merl::*)(const vec3 &, const vec3 &, const void *) const)
	MEMBER(&utia::eval, vec3 (utia::*)(const vec3 &, const vec3 &, const void *) const)
	MEMBER(&sgd::eval, vec3 (sgd::*)(const vec3 &, const vec3 &, const void *) const)
	MEMBER(&abc::eval, vec3 (abc::*)(const vec3 &, const vec3 &, const void *) const)
	MEMBER(&sgd::ndf, vec3 (sgd::*)(const vec3 &) const)
	MEMBER(&sgd::gaf, vec3 (sgd::*)(const vec3 &, const vec3 &, const vec3 &) const)
	MEMBER(&sgd::g1, vec3 (sgd::*)(const vec3 &) const)
	MEMBER(&sgd::fresnel, vec3 (sgd::*)(float_t) const)
	MEMBER(&abc::ndf, vec3 (abc::*)(const vec3 &) const)
	MEMBER(&abc::gaf, float_t (abc::*)(const vec3 &, const vec3 &, const vec3 &) const)
	MEMBER(&abc::fresnel, vec3 (abc::*)(float_t) const)
	MEMBER(&sgd::get_fresnel, const fresnel::impl &(sgd::*)() const)
	MEMBER(&abc::get_fresnel, const fresnel::impl &(abc::*)() const)
	// vec3 (dj_brdf.h:62-71, 597-637)
	{
		const double d[3] = {1, 2, 3};
		const float f[3] = {1, 2, 3};
		vec3 a = vec3::from_raw(d), b = vec3::from_raw(f), c(0.3f, 0.2f);
		const float_t *r = vec3::to_raw(a);
		vec3 e = a + b - c * b / b * 2.0f + 2.0f * a / 3.0f;
		e += a; e *= b; e *= 2.0f;
		(void)r; (void)e.intensity(); (void)dot(a, b); (void)normalize(a);
	}
	return 0;
}
