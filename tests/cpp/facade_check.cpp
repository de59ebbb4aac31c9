// tests/cpp/facade_check.cpp -- drives the C++ facade (include/djb200_facade.hpp) the way a host program written
// against dj_brdf.h would, on inputs the Python test wrote, and dumps the results for comparison with the oracle.
//   facade_check <in.bin> <out.bin>
// in.bin : int32 n, then wi[n][3], wo[n][3], u[n][2] (float32)
// out.bin: for ndf in (ggx, beckmann): eval[n][3] (batch), pdf[n] (batch), sample[n][3] (batch), eval_scalar[16][3],
//          eval16[2][n][3] (two params blocks, BROADCAST); then lrep round trip [5]; then djb::tabular(ggx, 90) as a BRDF:
//          eval[n][3], sample[n][3], eval_scalar[3], alpha (beckmann, ggx)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "dj_brdf.h" // include/compat: the reference's header name

static void put(FILE *f, const void *p, size_t bytes) { if (fwrite(p, 1, bytes, f) != bytes) abort(); }

int main(int argc, char **argv)
{
	if (argc != 3) return 2;
	FILE *fi = fopen(argv[1], "rb");
	if (!fi) return 3;
	int n = 0;
	if (fread(&n, 4, 1, fi) != 1) return 4;
	std::vector<djb::vec3> wi(n), wo(n), out(n);
	std::vector<float> u(2 * n), pdf(n);
	if (fread(&wi[0], 12, n, fi) != (size_t)n || fread(&wo[0], 12, n, fi) != (size_t)n || fread(&u[0], 8, n, fi) != (size_t)n) return 5;
	fclose(fi);
	FILE *fo = fopen(argv[2], "wb");
	try {
		djb::microfacet::params P = djb::microfacet::params::elliptic(0.1f, 0.4f, 0.7f);
		djb::microfacet::params two[2] = {djb::microfacet::params::isotropic(0.3f),
		                                  djb::microfacet::params::pdfparams(0.3f, 0.2f, 0.4f, 0.1f, -0.2f)};
		djb::ggx ggx(djb::fresnel::schlick(djb::vec3(0.9f, 0.5f, 0.2f)));
		djb::beckmann beckmann(djb::fresnel::schlick(djb::vec3(0.9f, 0.5f, 0.2f)));
		const djb::microfacet *brdfs[2] = {&ggx, &beckmann};
		for (int b = 0; b < 2; ++b) {
			const djb::microfacet &m = *brdfs[b];
			m.eval_batch(&wi[0], &wo[0], n, &out[0], &P);
			put(fo, &out[0], 12 * (size_t)n);
			m.pdf_batch(&wi[0], &wo[0], n, &pdf[0], &P);
			put(fo, &pdf[0], 4 * (size_t)n);
			m.sample_batch(&u[0], &wo[0], n, &out[0], &P);
			put(fo, &out[0], 12 * (size_t)n);
			const djb::brdf &as_base = m; // the reference's virtual interface, one pair per call
			for (int k = 0; k < 16; ++k) {
				djb::vec3 e = as_base.eval(wi[k], wo[k], &P);
				put(fo, &e, 12);
			}
			std::vector<djb::vec3> out2(2 * (size_t)n);
			m.eval_batch(&wi[0], &wo[0], n, &out2[0], two, 2, DJB200_PARAMS_BROADCAST);
			put(fo, &out2[0], 24 * (size_t)n);
		}
		djb::beckmann::lrep l;
		djb::beckmann::params_to_lrep(two[1], &l);
		l += djb::beckmann::lrep(0.01f, 0.02f, 0.03f, 0.04f, 0.001f);
		djb::microfacet::params back;
		djb::beckmann::lrep_to_params(l, &back);
		float q[5];
		back.get_pdfparams(&q[0], &q[1], &q[2], &q[3], &q[4]);
		put(fo, q, sizeof q);
		// djb::tabular: fitted on the GPU from an analytic GGX, then used as a BRDF through the base-class interface
		djb::ggx plain;
		djb::tabular tab(plain, 90);
		const djb::brdf &tb = tab;
		tb.eval_batch(&wi[0], &wo[0], n, &out[0]);
		put(fo, &out[0], 12 * (size_t)n);
		tab.sample_batch(&u[0], &wo[0], n, &out[0]);
		put(fo, &out[0], 12 * (size_t)n);
		djb::vec3 e0 = tb.eval(wi[0], wo[0]);
		put(fo, &e0, 12);
		float ab[2], dummy;
		djb::tabular::fit_beckmann_parameters(tab).get_ellipse(&ab[0], &dummy, NULL);
		djb::tabular::fit_ggx_parameters(tab).get_ellipse(&ab[1], &dummy, NULL);
		put(fo, ab, sizeof ab);
		// the scalar members behind sample / eval (dj_brdf.h:366-369, 384-389, 450-455, 506-509, 531-533), one call per item
		const int nm = n < 64 ? n : 64;
		djb::sgd sgd("gold-metallic-paint");
		djb::abc abc("gold-metallic-paint");
		djb::beckmann bplain;
		djb::tabular_anisotropic ta(bplain, 8, 10);
		for (int k = 0; k < nm; ++k) {
			const float c = wo[k].z, sn = sqrtf(1.0f - c * c), u1 = u[2 * k], u2 = u[2 * k + 1];
			float q[8];
			q[0] = ggx.qf1(u1); q[1] = ggx.qf2_radial(u1, c, sn); q[2] = ggx.qf3_radial(u2, q[1]);
			q[3] = beckmann.qf1(u1); q[4] = beckmann.qf2_radial(u1, c, sn); q[5] = beckmann.qf3_radial(u2, q[4]);
			q[6] = abc.gaf(wi[k], wi[k], wo[k]);
			q[7] = 0;
			put(fo, q, sizeof q);
			djb::vec3 v[7] = {sgd.ndf(wi[k]), sgd.gaf(wi[k], wi[k], wo[k]), sgd.g1(wo[k]), sgd.fresnel(c), abc.ndf(wi[k]), abc.fresnel(c),
			                  djb::vec3(0)};
			put(fo, v, sizeof v);
			const float phi = 6.2f * u2, theta = 1.5f * u1;
			float t[6] = {ta.pdf1(phi), ta.cdf1(phi), ta.qf1(u1), ta.pdf2(theta, phi), ta.cdf2(theta, phi), ta.qf2(u1, phi)};
			put(fo, t, sizeof t);
		}
		int threw = 0;
		try { ((const djb::microfacet &)ggx).qf2(0.5f, wo[0]); } catch (const djb::exc &) { ++threw; }
		try { ((const djb::radial &)tab).qf2_radial(0.5f, 0.5f, 0.5f); } catch (const djb::exc &) { ++threw; }
		float thr = (float)threw;
		put(fo, &thr, 4);
	} catch (const std::exception &e) {
		fprintf(stderr, "%s\n", e.what());
		return 1;
	}
	fclose(fo);
	return 0;
}
