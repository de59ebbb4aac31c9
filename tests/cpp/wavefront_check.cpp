// tests/cpp/wavefront_check.cpp -- the wavefront adapters of include/djb200_wavefront.hpp on the records file format of
// plugin_driver.cpp; tests/test_plugins.py compares the output with the scalar plugins (reference build).
//   wavefront_check <mode> <records.bin> <out.bin>     mode: lean | naive_mip | beckmann_textured
// out.bin: n x 11 floats like plugin_driver (eval rgb, pdf, sample weight rgb, sampled wo, sample pdf)
#include <cstdio>
#include <cstring>
#include <vector>

#include "djb200_wavefront.hpp"

// the plugins read their roughness textures through Spectrum::average() (mitsuba/dj_brdf.cpp:354-356): for a grey texel
// v that is (v + v + v) * (1 / 3), which is not always v in float -- a renderer hands the adapters the same number
static float spectrum_average(float v) { return (v + v + v) * (1.0f / 3.0f); }

int main(int argc, char **argv)
{
	if (argc != 4) return 2;
	try {
		FILE *f = fopen(argv[2], "rb");
		int n = 0;
		if (!f || fread(&n, 4, 1, f) != 1) return 2;
		std::vector<float> rec((size_t)n * 16), out((size_t)n * 11, 0.0f);
		if (fread(rec.data(), 64, n, f) != (size_t)n) return 2;
		fclose(f);
		std::vector<djb::vec3> wi(n), wo(n), ev(n), w(n), swo(n);
		std::vector<float> u(2 * n), a3(3 * n), e5(5 * n), pdf(n), spdf(n);
		for (int k = 0; k < n; ++k) {
			const float *r = &rec[(size_t)k * 16];
			wi[k] = djb::vec3(r[0], r[1], r[2]);
			wo[k] = djb::vec3(r[3], r[4], r[5]);
			u[2 * k] = r[6]; u[2 * k + 1] = r[7];
			for (int c = 0; c < 3; ++c) a3[3 * k + c] = spectrum_average(r[8 + c]);
			memcpy(&e5[5 * k], r + 11, 20);
		}
		djb::wavefront::records R;
		R.n = n; R.wi = wi.data(); R.wo = wo.data(); R.u = u.data();
		const std::string mode = argv[1];
		if (mode == "lean" || mode == "naive_mip") {
			const bool lean = mode == "lean";
			R.alpha3 = lean ? a3.data() : NULL; // the "naive_mip" config of the test keeps constant roughness
			R.lean5 = e5.data();
			djb::wavefront::beckmann_conductor b(spectrum_average(0.1f), spectrum_average(0.1f), 0.0f, lean, lean ? 1.5f : 1.0f);
			b.eval(R, ev.data());
			b.pdf(R, pdf.data());
			b.sample(R, w.data(), swo.data(), spdf.data());
		} else if (mode == "beckmann_textured") {
			djb::beckmann g;
			djb::wavefront::rough_microfacet b(g);
			std::vector<djb::microfacet::params> P;
			djb::wavefront::rough_microfacet::make_params(a3.data(), n, &P);
			b.eval(R, P.data(), ev.data());
			b.pdf(R, P.data(), pdf.data());
			b.sample(R, P.data(), w.data(), swo.data(), spdf.data());
		} else {
			return 2;
		}
		for (int k = 0; k < n; ++k) {
			float *o = &out[(size_t)k * 11];
			o[0] = ev[k].x; o[1] = ev[k].y; o[2] = ev[k].z; o[3] = pdf[k];
			o[4] = w[k].x; o[5] = w[k].y; o[6] = w[k].z;
			o[7] = swo[k].x; o[8] = swo[k].y; o[9] = swo[k].z; o[10] = spdf[k];
		}
		f = fopen(argv[3], "wb");
		if (!f || fwrite(out.data(), 44, n, f) != (size_t)n) return 2;
		fclose(f);
	} catch (std::exception &e) {
		fprintf(stderr, "wavefront_check: %s\n", e.what());
		return 1;
	}
	return 0;
}
