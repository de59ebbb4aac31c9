// oracle/ref_open.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The same harness as ref_harness.cpp around the same UNMODIFIED reference header, compiled a second time with the
// access specifiers opened (class -> struct, private / protected -> public; every standard header is included first so
// that only the reference is affected).  This does not change a single instruction of the reference's arithmetic; it
// only lets the harness read tables that have no accessor: tabular_anisotropic's m_pdf1 / m_cdf1 / m_qf1 / m_pdf2 /
// m_cdf2 / m_qf2 (dj_brdf.h:431-436), which pin the oracle's restatement of dj_brdf.h:2848-3103 stage by stage.
// Output: oracle/_ref/libdjbref_open.so (oracle/Makefile).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <exception>
#include <fstream>
#include <iostream>
#include <stdint.h>
#include <string>
#include <thread>
#include <vector>

#define class struct
#define private public
#define protected public
#include "ref_harness.cpp"
#undef class
#undef private
#undef protected

// out arrays: azim_res floats (1-D tables) / elev_res * azim_res floats (2-D tables), zero-filled beyond the vector's
// size; sizes6 = the six vector sizes in the order pdf1, cdf1, qf1, pdf2, cdf2, qf2
REF_API int ref_tabular_anisotropic_sampling_tables(void *tabh, float *pdf1, float *cdf1, float *qf1, float *pdf2,
                                                    float *cdf2, float *qf2, int *sizes6)
{
	const djb::tabular_anisotropic *t = dynamic_cast<djb::tabular_anisotropic *>(static_cast<djb::brdf *>(tabh));
	if (!t) return -1;
	const std::vector<djb::float_t> *v[6] = {&t->m_pdf1, &t->m_cdf1, &t->m_qf1, &t->m_pdf2, &t->m_cdf2, &t->m_qf2};
	float *o[6] = {pdf1, cdf1, qf1, pdf2, cdf2, qf2};
	const size_t cap[6] = {(size_t)t->m_azimuthal_res, (size_t)t->m_azimuthal_res, (size_t)t->m_azimuthal_res,
	                       (size_t)t->m_elevation_res * t->m_azimuthal_res, (size_t)t->m_elevation_res * t->m_azimuthal_res,
	                       (size_t)t->m_elevation_res * t->m_azimuthal_res};
	for (int k = 0; k < 6; ++k) {
		sizes6[k] = (int)v[k]->size();
		if (!o[k]) continue;
		memset(o[k], 0, sizeof(float) * cap[k]);
		size_t n = v[k]->size() < cap[k] ? v[k]->size() : cap[k];
		if (n) memcpy(o[k], &(*v[k])[0], sizeof(float) * n);
	}
	return 0;
}

// the public table queries (dj_brdf.h:2766-2824): what = 0 pdf1(phi), 1 cdf1(phi), 2 qf1(u), 3 pdf2(theta, phi),
// 4 cdf2(theta, phi), 5 qf2(u, phi); a = first argument, b = second argument (ignored for 0..2)
REF_API int ref_tabular_anisotropic_lookup(void *tabh, int what, const float *a, const float *b, int n, float *out)
{
	const djb::tabular_anisotropic *t = dynamic_cast<djb::tabular_anisotropic *>(static_cast<djb::brdf *>(tabh));
	if (!t) return -1;
	for (int k = 0; k < n; ++k) {
		switch (what) {
		case 0: out[k] = t->pdf1(a[k]); break;
		case 1: out[k] = t->cdf1(a[k]); break;
		case 2: out[k] = t->qf1(a[k]); break;
		case 3: out[k] = t->pdf2(a[k], b[k]); break;
		case 4: out[k] = t->cdf2(a[k], b[k]); break;
		case 5: out[k] = t->qf2(a[k], b[k]); break;
		default: return -2;
		}
	}
	return 0;
}
