// merl_params_batch -- Beckmann and GGX roughness of MERL materials, all files fitted in ONE device pass.
// Same output format as the reference's examples/merl_params.cpp (params.txt: "name beckmann ggx", %.3f), but the
// materials are uploaded first and djb::tabular::fit_batch() runs the whole batch on the GPU (one CTA per material).
//
//   g++ -O2 -std=c++11 -I../include merl_params_batch.cpp -L../dj_brdf_b200 -ldjb200 -Wl,-rpath,../dj_brdf_b200 -o merl_params_batch
//   ./merl_params_batch [-o params.txt] [-i iterations] a.binary b.binary ...
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "djb200_facade.hpp"

int main(int argc, char **argv)
{
	const char *out_path = "params.txt";
	int iterations = 4; // the reference's km.eigenvector(4)
	std::vector<const char *> files;
	for (int i = 1; i < argc; ++i) {
		if (!strcmp(argv[i], "-h")) {
			printf("%s [-o params.txt] [-i iterations] merl1.binary merl2.binary ...\n", argv[0]);
			return EXIT_SUCCESS;
		} else if (!strcmp(argv[i], "-o") && i + 1 < argc) out_path = argv[++i];
		else if (!strcmp(argv[i], "-i") && i + 1 < argc) iterations = atoi(argv[++i]);
		else files.push_back(argv[i]);
	}
	if (files.empty()) return EXIT_SUCCESS;
	try {
		std::vector<djb::merl *> tables;
		std::vector<const djb::brdf *> sources;
		for (size_t k = 0; k < files.size(); ++k) {
			tables.push_back(new djb::merl(files[k]));
			sources.push_back(tables.back());
		}
		std::vector<djb::tabular *> fits = djb::tabular::fit_batch(sources, 90, true, iterations);
		FILE *pf = fopen(out_path, "w");
		if (!pf) throw djb::exc("djb_error: cannot write %s", out_path);
		fprintf(pf, "# MERL Beckmann GGX\n");
		for (size_t k = 0; k < files.size(); ++k) {
			float beckmann, ggx, dummy;
			djb::tabular::fit_beckmann_parameters(*fits[k]).get_ellipse(&beckmann, &dummy, NULL);
			djb::tabular::fit_ggx_parameters(*fits[k]).get_ellipse(&ggx, &dummy, NULL);
			std::string name(files[k]);
			size_t slash = name.find_last_of('/');
			if (slash != std::string::npos) name = name.substr(slash + 1);
			name = name.substr(0, name.find('.'));
			fprintf(pf, "%s %.3f %.3f\n", name.c_str(), beckmann, ggx);
			printf("%s beckmann %.9g ggx %.9g\n", name.c_str(), beckmann, ggx);
		}
		fclose(pf);
		for (size_t k = 0; k < files.size(); ++k) { delete fits[k]; delete tables[k]; }
	} catch (const std::exception &e) {
		fprintf(stderr, "%s\n", e.what());
		return EXIT_FAILURE;
	}
	return EXIT_SUCCESS;
}
