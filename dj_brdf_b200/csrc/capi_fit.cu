// capi_fit.cu -- C-ABI entry points of the power-iteration fits (include/djb200.h "fits"): argument
// checks, source descriptors -> device, workspaces, result download.  Numerics: kernels_fit.cu.
#include <cstring>
#include <vector>

#include "djb_fit.cuh"
#include "djb_internal.h"

using namespace djb200;

namespace {

struct DevBuf { // RAII for cudaMalloc
	void *p = nullptr;
	~DevBuf() { if (p) cudaFree(p); }
	cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
	template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

// host source descriptors -> device array (spline Fresnel points are uploaded into `spline_store`)
djb200_status build_sources(const djb200_source *sources, int32_t n, std::vector<FitSourceDev> &out,
                            std::vector<DevBuf> &spline_store)
{
	out.resize(n);
	spline_store.resize(n);
	int dev = 0;
	cudaGetDevice(&dev);
	for (int32_t k = 0; k < n; ++k) {
		const djb200_source &s = sources[k];
		FitSourceDev d;
		memset(&d, 0, sizeof d);
		d.kind = s.kind;
		switch (s.kind) {
		case DJB200_SOURCE_MERL:
			if (!s.merl) return fail(DJB200_ERR_INVALID_ARGUMENT, "source %d: merl handle is NULL", k);
			if (s.merl->device != dev)
				return fail(DJB200_ERR_INVALID_ARGUMENT, "source %d: table lives on device %d, current device is %d", k,
				            s.merl->device, dev);
			d.merl = s.merl->cells;
			break;
		case DJB200_SOURCE_UTIA:
			if (!s.utia) return fail(DJB200_ERR_INVALID_ARGUMENT, "source %d: utia handle is NULL", k);
			if (s.utia->device != dev)
				return fail(DJB200_ERR_INVALID_ARGUMENT, "source %d: table lives on device %d, current device is %d", k,
				            s.utia->device, dev);
			d.utia = s.utia->table;
			break;
		case DJB200_SOURCE_MICROFACET: {
			const djb200_microfacet &m = s.microfacet;
			if (m.ndf != DJB200_NDF_BECKMANN && m.ndf != DJB200_NDF_GGX)
				return fail(DJB200_ERR_INVALID_ARGUMENT, "source %d: unknown ndf %d", k, m.ndf);
			if (m.fresnel.kind < 0 || m.fresnel.kind > DJB200_FRESNEL_SPLINE)
				return fail(DJB200_ERR_INVALID_ARGUMENT, "source %d: unknown fresnel kind %d", k, m.fresnel.kind);
			d.ndf = m.ndf;
			d.shadow = m.shadow;
			d.fresnel_kind = m.fresnel.kind;
			memcpy(d.fr.v, m.fresnel.v, sizeof d.fr.v);
			if (m.fresnel.kind == DJB200_FRESNEL_SPLINE) {
				if (!m.fresnel.points || m.fresnel.n_points < 1)
					return fail(DJB200_ERR_INVALID_ARGUMENT, "source %d: spline fresnel needs points", k);
				size_t bytes = sizeof(float) * 3 * (size_t)m.fresnel.n_points;
				cudaError_t e = spline_store[k].alloc(bytes);
				if (e == cudaSuccess) e = cudaMemcpy(spline_store[k].p, m.fresnel.points, bytes, cudaMemcpyHostToDevice);
				if (e != cudaSuccess) return cuda_fail(e, "fresnel spline upload");
				d.fr.pts = spline_store[k].as<float>();
				d.fr.npts = m.fresnel.n_points;
			}
		} break;
		default: return fail(DJB200_ERR_INVALID_ARGUMENT, "source %d: unknown kind %d", k, s.kind);
		}
		out[k] = d;
	}
	return DJB200_OK;
}

} // namespace

extern "C" {

djb200_status djb200_fit_tabular(const djb200_source *sources, int32_t n_sources, int32_t res, int32_t shadow,
                                 int32_t iterations, djb200_tabular_fit *results, void *stream)
{
	if (n_sources < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "negative source count");
	if (res <= 2) return fail(DJB200_ERR_INVALID_ARGUMENT, "Invalid Resolution"); // DJB_ASSERT, dj_brdf.h:2218
	if (res > 1024) return fail(DJB200_ERR_UNSUPPORTED, "resolution %d > 1024 does not fit one CTA's shared memory", res);
	if (iterations < 1) return fail(DJB200_ERR_INVALID_ARGUMENT, "need at least one power iteration");
	if (n_sources == 0) return DJB200_OK;
	if (!sources || !results) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	for (int32_t k = 0; k < n_sources; ++k) {
		const djb200_tabular_fit &r = results[k];
		if (r.res != res) return fail(DJB200_ERR_INVALID_ARGUMENT, "result %d: res field %d != %d", k, r.res, res);
		if (!r.p22 || !r.sigma || !r.cdf || !r.qf || !r.fresnel)
			return fail(DJB200_ERR_INVALID_ARGUMENT, "result %d: NULL output array", k);
	}
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;

	std::vector<FitSourceDev> src;
	std::vector<DevBuf> splines;
	rs = build_sources(sources, n_sources, src, splines);
	if (rs != DJB200_OK) return rs;

	cudaStream_t st = (cudaStream_t)stream;
	const size_t n = (size_t)n_sources, cnt = (size_t)res - 1;
	DevBuf d_src, d_K, d_grid, d_out, d_resid;
	// outputs packed per kind: p22 | sigma | cdf | qf (n x res each) | fresnel (n x res x 3) | alpha (n x 2)
	const size_t per_kind = n * (size_t)res;
	const size_t out_floats = 4 * per_kind + 3 * per_kind + 2 * n;
#define FCU(call)                                                \
	do {                                                         \
		cudaError_t e__ = (call);                                \
		if (e__ != cudaSuccess) return cuda_fail(e__, #call);    \
	} while (0)
	FCU(d_src.alloc(sizeof(FitSourceDev) * n));
	FCU(d_K.alloc(sizeof(double) * n * cnt * cnt));
	FCU(d_grid.alloc(sizeof(float) * n * 180 * 90));
	FCU(d_out.alloc(sizeof(float) * out_floats));
	FCU(d_resid.alloc(sizeof(float) * n * (size_t)iterations));
	FCU(cudaMemcpyAsync(d_src.p, src.data(), sizeof(FitSourceDev) * n, cudaMemcpyHostToDevice, st));
	float *o = d_out.as<float>();
	float *o_p22 = o, *o_sigma = o + per_kind, *o_cdf = o + 2 * per_kind, *o_qf = o + 3 * per_kind;
	float *o_fres = o + 4 * per_kind, *o_alpha = o_fres + 3 * per_kind;
	FCU(launch_fit_tabular(d_src.as<FitSourceDev>(), n_sources, res, shadow, iterations, d_K.as<double>(),
	                       d_grid.as<float>(), o_p22, o_sigma, o_cdf, o_qf, o_fres, o_alpha, d_resid.as<float>(), st));
	std::vector<float> h(out_floats), hres(n * (size_t)iterations);
	FCU(cudaMemcpyAsync(h.data(), o, sizeof(float) * out_floats, cudaMemcpyDeviceToHost, st));
	FCU(cudaMemcpyAsync(hres.data(), d_resid.p, sizeof(float) * hres.size(), cudaMemcpyDeviceToHost, st));
	FCU(cudaStreamSynchronize(st));
#undef FCU
	for (size_t k = 0; k < n; ++k) {
		djb200_tabular_fit &r = results[k];
		memcpy(r.p22, h.data() + k * res, sizeof(float) * res);
		memcpy(r.sigma, h.data() + per_kind + k * res, sizeof(float) * res);
		memcpy(r.cdf, h.data() + 2 * per_kind + k * res, sizeof(float) * res);
		memcpy(r.qf, h.data() + 3 * per_kind + k * res, sizeof(float) * res);
		memcpy(r.fresnel, h.data() + 4 * per_kind + 3 * k * res, sizeof(float) * 3 * res);
		r.alpha_beckmann = h[7 * per_kind + 2 * k];
		r.alpha_ggx = h[7 * per_kind + 2 * k + 1];
		if (r.residuals) memcpy(r.residuals, hres.data() + k * iterations, sizeof(float) * iterations);
	}
	return DJB200_OK;
}

djb200_status djb200_fit_tabular_anisotropic(const djb200_source *, int32_t, int32_t, int32_t, int32_t, int32_t,
                                             djb200_tabular_anisotropic_fit *, void *)
{
	return fail(DJB200_ERR_UNSUPPORTED, "anisotropic fit: not built yet");
}

} // extern "C"
