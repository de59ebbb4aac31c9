// tests/cpp/params_host_check.cpp -- the host-only part of the djb:: interface (params factories / setters / getters, LEAN
// lrep algebra): built once against the reference header and once against the facade, the two outputs must be identical
// text (hex floats).  No device call is made, so this runs without a GPU.
#include <cstdio>
#ifdef USE_REFERENCE
#define DJ_BRDF_IMPLEMENTATION 1
#endif
#include "dj_brdf.h"

static void show(const char *tag, const djb::microfacet::params &p)
{
	float a1, a2, ph, ax, ay, rho, tx, ty;
	djb::vec3 n;
	p.get_ellipse(&a1, &a2, &ph);
	p.get_pdfparams(&ax, &ay, &rho, &tx, &ty);
	p.get_location(&n);
	printf("%s %a %a %a | %a %a %a %a %a | %a %a %a\n", tag, a1, a2, ph, ax, ay, rho, tx, ty, n.x, n.y, n.z);
}

int main()
{
	show("standard", djb::microfacet::params::standard());
	show("isotropic", djb::microfacet::params::isotropic(0.37f));
	for (int k = 0; k < 40; ++k) {
		float a1 = 0.02f + 0.019f * k, a2 = 0.8f - 0.017f * k, phi = 0.08f * k - 0.7f;
		djb::microfacet::params p = djb::microfacet::params::elliptic(a1, a2, phi);
		show("elliptic", p);
		p.set_location(0.01f * k - 0.2f, 0.3f - 0.02f * k);
		show("set_location", p);
		p.set_location(djb::vec3(0.01f * k, -0.015f * k, 0.9f));
		show("set_location_n", p);
		p.set_ellipse(a2, a1, -phi);
		show("set_ellipse", p);
		djb::microfacet::params q = djb::microfacet::params::pdfparams(a1, a2, 0.02f * k - 0.4f, 0.1f, -0.05f);
		show("pdfparams", q);
		q.set_pdfparams(a2, a1, 0.4f - 0.02f * k);
		show("set_pdfparams", q);
	}
	// LEAN algebra on the host: combine, scale, shear (dj_brdf.h:1992-2051); printed through the raw moments
	for (int k = 0; k < 20; ++k) {
		djb::beckmann::lrep a(0.01f * k, -0.02f * k, 0.3f + 0.01f * k, 0.2f + 0.02f * k, 0.005f * k);
		djb::beckmann::lrep b(-0.03f * k, 0.015f * k, 0.1f + 0.03f * k, 0.4f - 0.01f * k, -0.002f * k);
		djb::beckmann::lrep c = a + b, d = c * (0.5f + 0.1f * k), e = a;
		e += b;
		e *= 1.5f;
		e.shear(0.1f, -0.2f);
		e.scale(1.1f, 0.9f);
		const djb::beckmann::lrep *all[3] = {&c, &d, &e};
		for (int j = 0; j < 3; ++j) {
			const float *m = reinterpret_cast<const float *>(all[j]); // five packed floats on both sides
			printf("lrep %a %a %a %a %a\n", m[0], m[1], m[2], m[3], m[4]);
		}
	}
	// Fresnel utilities, dj_brdf.h:151-154
	for (int k = 1; k < 30; ++k) {
		float f0, ior;
		djb::fresnel::ior_to_f0(1.0f + 0.11f * k, &f0);
		djb::fresnel::f0_to_ior(0.033f * k, &ior);
		djb::vec3 v0, v1;
		djb::fresnel::ior_to_f0(djb::vec3(1.1f + 0.05f * k, 1.5f, 2.4f + 0.1f * k), &v0);
		djb::fresnel::f0_to_ior(djb::vec3(0.02f * k, 1.0f, 0.5f), &v1);
		printf("fresnel %a %a | %a %a %a | %a %a %a\n", f0, ior, v0.x, v0.y, v0.z, v1.x, v1.y, v1.z);
	}
	// brdf::sample / brdf::pdf defaults (cosine-weighted concentric warp, dj_brdf.h:726-752, 828-843), through a BRDF class that
	// keeps them (sgd; its constructor only looks the preset up)
	{
		djb::sgd sg("gold-metallic-paint");
		const djb::vec3 up(0, 0, 1);
		for (int k = 0; k <= 16; ++k)
			for (int j = 0; j <= 16; ++j) {
				const float u1 = (k == 16) ? 0.5f : (k + 0.37f) / 16.0f, u2 = (j == 16) ? 0.5f : (j + 0.81f) / 16.0f;
				const djb::vec3 s = sg.sample(u1, u2, up);
				printf("sample %a %a %a pdf %a\n", s.x, s.y, s.z, sg.pdf(s, up));
			}
	}
	djb::beckmann::lrep dflt;
	const float *m = reinterpret_cast<const float *>(&dflt);
	printf("lrep default %a %a %a %a %a\n", m[0], m[1], m[2], m[3], m[4]);
	return 0;
}
