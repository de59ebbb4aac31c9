// tests/cpp/plugin_driver.cpp -- TEST INFRASTRUCTURE.  Drives ONE unmodified Mitsuba plugin source of the reference
// (mitsuba/dj_*.cpp, pulled in through -DPLUGIN_SOURCE=...) against the mock Mitsuba API of tests/cpp/mock_mitsuba.
// The same program is built twice by oracle/Makefile: with -I/root/reference (the reference's own dj_brdf.h, CPU) and with
// -Iinclude/compat (the facade over libdjb200.so, GPU); tests/test_plugins.py feeds both the same records and compares.
//
//   plugin_driver <config.txt> <records.bin> <out.bin>
//   config.txt : one property per line -- "string name value" | "float name v" | "bool name 0|1" | "spectrum name r g b"
//   records.bin: int32 n, then n x 16 floats: wi[3] wo[3] u[2] alpha1 alpha2 alphaAngle E1 E2 E3 E4 E5
//                (NaN in alpha1 / E1 = "no texture value for this record": the plugin's constants apply)
//   out.bin    : n x 11 floats: eval rgb, pdf, sample weight rgb, sampled wo xyz, sample pdf
#include PLUGIN_SOURCE
#include "mock_impl.cpp"

#include <fstream>

using namespace mitsuba;

int main(int argc, char **argv)
{
	if (argc != 4) {
		fprintf(stderr, "usage: %s config.txt records.bin out.bin\n", argv[0]);
		return 2;
	}
	try {
		Properties props("plugin");
		std::ifstream cfg(argv[1]);
		std::string kind, name;
		bool textured = false;
		while (cfg >> kind >> name) {
			if (kind == "string") { std::string v; cfg >> v; props.setString(name, v); }
			else if (kind == "float") { float v; cfg >> v; props.setFloat(name, v); }
			else if (kind == "bool") { int v; cfg >> v; props.setBoolean(name, v != 0); }
			else if (kind == "spectrum") { Spectrum s; cfg >> s[0] >> s[1] >> s[2]; props.setSpectrum(name, s); }
			else if (kind == "driver" && name == "textured") { int v; cfg >> v; textured = v != 0; }
			else { fprintf(stderr, "bad config line: %s %s\n", kind.c_str(), name.c_str()); return 2; }
		}
		BSDF *bsdf = static_cast<BSDF *>(CreateInstance(props));
		bsdf->incRef();
		// nested texture elements, as a scene file would attach them (dj_brdf.cpp:436-452, dj_beckmannconductor.cpp:430-452)
		static const char *roles[] = {"alpha1", "alpha2", "alphaAngle", "leanmap1", "leanmap2"};
		if (textured) {
			for (int k = 0; k < 5; ++k) {
				if (k >= 3 && std::string(GetDescription()) != "Rough conductor BRDF") break;
				ConstantSpectrumTexture *t = new ConstantSpectrumTexture(Spectrum(k < 2 ? 0.1f : 0.0f));
				t->setRole(roles[k]);
				bsdf->addChild(roles[k], t);
			}
		}
		bsdf->configure();

		FILE *f = fopen(argv[2], "rb");
		if (!f) { fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
		int n = 0;
		if (fread(&n, 4, 1, f) != 1) return 2;
		std::vector<float> rec((size_t)n * 16), out((size_t)n * 11, 0.0f);
		if (fread(rec.data(), 64, n, f) != (size_t)n) return 2;
		fclose(f);
		for (int k = 0; k < n; ++k) {
			const float *r = &rec[(size_t)k * 16];
			float *o = &out[(size_t)k * 11];
			std::map<std::string, Spectrum> ovr;
			if (r[8] == r[8]) {
				ovr["alpha1"] = Spectrum(r[8]);
				ovr["alpha2"] = Spectrum(r[9]);
				ovr["alphaAngle"] = Spectrum(r[10]);
			}
			if (r[11] == r[11]) {
				Spectrum a, b;
				a.fromLinearRGB(r[11], r[12], 0.0f);
				b.fromLinearRGB(r[13], r[14], r[15]);
				ovr["leanmap1"] = a;
				ovr["leanmap2"] = b;
			}
			Intersection its;
			its.overrides = &ovr;
			BSDFSamplingRecord bRec(its, Vector(r[0], r[1], r[2]), Vector(r[3], r[4], r[5]));
			Spectrum e = bsdf->eval(bRec, ESolidAngle);
			o[0] = e[0]; o[1] = e[1]; o[2] = e[2];
			o[3] = bsdf->pdf(bRec, ESolidAngle);
			BSDFSamplingRecord sRec(its, Vector(r[0], r[1], r[2]), Vector(0, 0, 1));
			Float pdf = 0;
			Spectrum w = bsdf->sample(sRec, pdf, Point2(r[6], r[7]));
			o[4] = w[0]; o[5] = w[1]; o[6] = w[2];
			o[7] = sRec.wo.x; o[8] = sRec.wo.y; o[9] = sRec.wo.z;
			o[10] = pdf;
		}
		f = fopen(argv[3], "wb");
		if (!f || fwrite(out.data(), 44, n, f) != (size_t)n) { fprintf(stderr, "cannot write %s\n", argv[3]); return 2; }
		fclose(f);
		bsdf->decRef();
	} catch (std::exception &e) {
		fprintf(stderr, "plugin_driver: %s\n", e.what());
		return 1;
	}
	return 0;
}
