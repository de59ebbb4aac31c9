// capi_fit.cu -- C-ABI entry points of the power-iteration fits (include/djb200.h "fits"): argument
// checks, source descriptors -> device, workspaces, result download.  Numerics: kernels_fit.cu.
#include <dlfcn.h>
#include <nccl.h> // types and prototypes only: the library is dlopen()ed at run time, libdjb200.so does not link NCCL

#include <cstring>
#include <mutex>
#include <vector>

#include "djb_fit.cuh"
#include "djb_internal.h"

using namespace djb200;

namespace {

// Device blocks of the fit handles / workspaces.  cudaMalloc + cudaFree cost 0.1 - 0.4 ms each and an anisotropic fit owns
// seventeen blocks for 1.3 ms of kernels (measured: 2.1 - 6.6 ms per call depending on the allocator's state), so released blocks
// are kept per host thread and handed out again (same device, at most twice the wanted size); 256 MB of them at most, freed at
// thread exit.  A block goes back only after the device is idle (what cudaFree would have waited for too).
struct BlockCache {
	struct Block { void *p; size_t bytes; int device; };
	std::vector<Block> free_;
	size_t total = 0;
	~BlockCache() { for (Block &b : free_) cudaFree(b.p); }
	void *take(size_t bytes, int device, size_t *got)
	{
		int best = -1;
		for (int k = 0; k < (int)free_.size(); ++k)
			if (free_[k].device == device && free_[k].bytes >= bytes && free_[k].bytes <= 2 * bytes + 4096 &&
			    (best < 0 || free_[k].bytes < free_[best].bytes))
				best = k;
		if (best < 0) return nullptr;
		void *p = free_[best].p;
		*got = free_[best].bytes;
		total -= free_[best].bytes;
		free_.erase(free_.begin() + best);
		return p;
	}
	void give(void *p, size_t bytes, int device)
	{
		int cur = -1;
		cudaGetDevice(&cur);
		// a block of another device than the current one is simply freed (the synchronisation below covers the current device only)
		if (cur != device || total + bytes > (256u << 20) || free_.size() >= 256) { cudaFree(p); return; }
		cudaDeviceSynchronize(); // nothing in flight may still use the block when it is handed out again
		free_.push_back({p, bytes, device});
		total += bytes;
	}
};
thread_local BlockCache t_blocks;

struct DevBuf { // RAII for a device block
	void *p = nullptr;
	size_t bytes = 0;
	int device = 0;
	DevBuf() = default;
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes), device(o.device) { o.p = nullptr; }
	DevBuf &operator=(DevBuf &&o) noexcept
	{
		if (this != &o) { release(); p = o.p; bytes = o.bytes; device = o.device; o.p = nullptr; }
		return *this;
	}
	~DevBuf() { release(); }
	void release()
	{
		if (p) t_blocks.give(p, bytes, device);
		p = nullptr;
	}
	cudaError_t alloc(size_t want)
	{
		release();
		want = (want + 255) & ~(size_t)255;
		if (!want) want = 256;
		cudaGetDevice(&device);
		p = t_blocks.take(want, device, &bytes);
		if (p) return cudaSuccess;
		bytes = want;
		return cudaMalloc(&p, want);
	}
	template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

// grow-only device scratch kept per host thread between fit calls (cudaMalloc / cudaFree per call cost more than
// the fit kernel itself); freed at thread exit
struct Scratch {
	void *p = nullptr;
	size_t bytes = 0;
	int device = -1;
	~Scratch() { if (p) cudaFree(p); }
	cudaError_t reserve(size_t want)
	{
		int dev = 0;
		cudaGetDevice(&dev);
		if (dev == device && want <= bytes) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr;
		bytes = 0;
		device = dev;
		cudaError_t e = cudaMalloc(&p, want ? want : 1);
		if (e == cudaSuccess) bytes = want;
		return e;
	}
	template <class T> T *as() { return reinterpret_cast<T *>(p); }
};
thread_local Scratch t_fit_src, t_fit_K, t_fit_ws, t_fit_out, t_fit_resid, t_fit_ndfgrid;

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
		case DJB200_SOURCE_SGD:
			if (!s.sgd) return fail(DJB200_ERR_INVALID_ARGUMENT, "source %d: sgd coefficients are NULL", k);
			memcpy(d.coef, s.sgd->ch, sizeof(double) * 33);
			break;
		case DJB200_SOURCE_ABC:
			if (!s.abc) return fail(DJB200_ERR_INVALID_ARGUMENT, "source %d: abc coefficients are NULL", k);
			for (int c = 0; c < 3; ++c) {
				d.coef[c] = s.abc->kD[c];
				d.coef[3 + c] = s.abc->A[c];
			}
			d.coef[6] = s.abc->B;
			d.coef[7] = s.abc->C;
			d.coef[8] = s.abc->ior;
			break;
		default: return fail(DJB200_ERR_INVALID_ARGUMENT, "source %d: unknown kind %d", k, s.kind);
		}
		out[k] = d;
	}
	return DJB200_OK;
}

} // namespace

extern "C" {

// Core of the batched isotropic fit: results stay packed on the device in `*out_dev` (per-thread scratch, valid until this thread's
// next fit call): p22 | sigma | cdf | qf (n x res each) | fresnel (n x res x 3) | alpha (n x 2); residuals n x iterations.
static djb200_status fit_tabular_device(const djb200_source *sources, int32_t n_sources, int32_t res, int32_t shadow,
                                        int32_t iterations, cudaStream_t st, float **out_dev, float **resid_dev)
{
	if (n_sources < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "negative source count");
	if (res <= 2) return fail(DJB200_ERR_INVALID_ARGUMENT, "Invalid Resolution"); // DJB_ASSERT, dj_brdf.h:2218
	if (fit_tabular_smem_bytes(res) > 227 * 1024)
		return fail(DJB200_ERR_UNSUPPORTED, "resolution %d does not fit one CTA's shared memory (max ~1500)", res);
	if (iterations < 1) return fail(DJB200_ERR_INVALID_ARGUMENT, "need at least one power iteration");
	if (!sources) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;

	std::vector<FitSourceDev> src;
	std::vector<DevBuf> splines;
	rs = build_sources(sources, n_sources, src, splines);
	if (rs != DJB200_OK) return rs;

	const size_t n = (size_t)n_sources, cnt = (size_t)res - 1;
	Scratch &d_src = t_fit_src, &d_K = t_fit_K, &d_grid = t_fit_ws, &d_out = t_fit_out, &d_resid = t_fit_resid, &d_ndfgrid = t_fit_ndfgrid;
	const size_t per_kind = n * (size_t)res;
	const size_t out_floats = djb200_fit_tabular_packed_floats(n_sources, res);
#define FCU(call)                                                \
	do {                                                         \
		cudaError_t e__ = (call);                                \
		if (e__ != cudaSuccess) return cuda_fail(e__, #call);    \
	} while (0)
	FCU(d_src.reserve(sizeof(FitSourceDev) * n));
	FCU(d_K.reserve(sizeof(double) * n * cnt * cnt));
	FCU(d_grid.reserve(sizeof(float4) * n * cnt * (cnt + 2))); // Fresnel ratio workspace
	FCU(d_out.reserve(sizeof(float) * out_floats));
	FCU(d_resid.reserve(sizeof(float) * n * (size_t)iterations));
	FCU(cudaMemcpyAsync(d_src.p, src.data(), sizeof(FitSourceDev) * n, cudaMemcpyHostToDevice, st));
	float *o = d_out.as<float>();
	float *o_p22 = o, *o_sigma = o + per_kind, *o_cdf = o + 2 * per_kind, *o_qf = o + 3 * per_kind;
	float *o_fres = o + 4 * per_kind, *o_alpha = o_fres + 3 * per_kind;
	float *grid_ws = nullptr;
	if (fit_tabular_parts(n_sources, res) > 1) { // small batch: the split mode passes the NDF grid through global memory
		FCU(d_ndfgrid.reserve(sizeof(float) * n * 180 * 90));
		grid_ws = d_ndfgrid.as<float>();
	}
	FCU(launch_fit_tabular(d_src.as<FitSourceDev>(), n_sources, res, shadow, iterations, d_K.as<double>(),
	                       d_grid.as<float4>(), grid_ws, o_p22, o_sigma, o_cdf, o_qf, o_fres, o_alpha, d_resid.as<float>(), st));
	if (!splines.empty()) FCU(cudaStreamSynchronize(st)); // spline uploads are freed when this function returns
	*out_dev = o;
	*resid_dev = d_resid.as<float>();
	return DJB200_OK;
}

djb200_status djb200_debug_fit_parts(int parts)
{
	if (parts != 0 && parts != 1 && (parts < 3 || parts > 8)) return fail(DJB200_ERR_INVALID_ARGUMENT, "parts must be 0, 1 or 3..8");
	g_fit_parts.store(parts);
	return DJB200_OK;
}

djb200_status djb200_debug_fit_phase_clocks(int64_t out_clocks[10])
{
	if (!out_clocks) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	long long c[10];
	cudaError_t e = cudaDeviceSynchronize();
	if (e == cudaSuccess) e = fit_phase_clocks(c);
	if (e != cudaSuccess) return cuda_fail(e, "fit_phase_clocks");
	for (int k = 0; k < 10; ++k) out_clocks[k] = (int64_t)c[k];
	return DJB200_OK;
}

int64_t djb200_fit_tabular_packed_floats(int32_t n_sources, int32_t res)
{
	return n_sources < 0 || res < 0 ? 0 : (int64_t)n_sources * ((int64_t)7 * res + 2);
}

djb200_status djb200_fit_tabular_packed(const djb200_source *sources, int32_t n_sources, int32_t res, int32_t shadow,
                                        int32_t iterations, float *out, float *residuals, int mem, void *stream)
{
	if (n_sources == 0) return DJB200_OK;
	if (!out) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	if (mem != DJB200_MEM_HOST && mem != DJB200_MEM_DEVICE) return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown memory space %d", mem);
	cudaStream_t st = (cudaStream_t)stream;
	float *o = nullptr, *r = nullptr;
	djb200_status rs = fit_tabular_device(sources, n_sources, res, shadow, iterations, st, &o, &r);
	if (rs != DJB200_OK) return rs;
	const cudaMemcpyKind kind = mem == DJB200_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
	FCU(cudaMemcpyAsync(out, o, sizeof(float) * (size_t)djb200_fit_tabular_packed_floats(n_sources, res), kind, st));
	if (residuals) FCU(cudaMemcpyAsync(residuals, r, sizeof(float) * (size_t)n_sources * (size_t)iterations, kind, st));
	if (mem == DJB200_MEM_HOST) FCU(cudaStreamSynchronize(st)); // host results are ready on return; device results are stream ordered
	return DJB200_OK;
}

djb200_status djb200_fit_tabular(const djb200_source *sources, int32_t n_sources, int32_t res, int32_t shadow,
                                 int32_t iterations, djb200_tabular_fit *results, void *stream)
{
	if (n_sources == 0) return DJB200_OK;
	if (!sources || !results) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	if (n_sources < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "negative source count");
	if (res <= 2) return fail(DJB200_ERR_INVALID_ARGUMENT, "Invalid Resolution"); // DJB_ASSERT, dj_brdf.h:2218
	if (iterations < 1) return fail(DJB200_ERR_INVALID_ARGUMENT, "need at least one power iteration");
	for (int32_t k = 0; k < n_sources; ++k) {
		const djb200_tabular_fit &r = results[k];
		if (r.res != res) return fail(DJB200_ERR_INVALID_ARGUMENT, "result %d: res field %d != %d", k, r.res, res);
		if (!r.p22 || !r.sigma || !r.cdf || !r.qf || !r.fresnel)
			return fail(DJB200_ERR_INVALID_ARGUMENT, "result %d: NULL output array", k);
	}
	const size_t n = (size_t)n_sources, per_kind = n * (size_t)res;
	std::vector<float> h((size_t)djb200_fit_tabular_packed_floats(n_sources, res)), hres(n * (size_t)iterations);
	djb200_status rs = djb200_fit_tabular_packed(sources, n_sources, res, shadow, iterations, h.data(), hres.data(),
	                                             DJB200_MEM_HOST, stream);
	if (rs != DJB200_OK) return rs;
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

// ---- anisotropic fit ---------------------------------------------------------------------------------
struct djb200_aniso_fit {
	int er, ar, shadow, n, device;
	FitSourceDev src;
	DevBuf spline, rowpre, colpre, colrcp, ones, p22, sigma, fresnel, terms, scale, pre_f, pre_d, params;
};

#define ACU(call)                                                \
	do {                                                         \
		cudaError_t e__ = (call);                                \
		if (e__ != cudaSuccess) return cuda_fail(e__, #call);    \
	} while (0)

djb200_status djb200_aniso_fit_create(const djb200_source *source, int32_t elev_res, int32_t azim_res, int32_t shadow,
                                      void *stream, djb200_aniso_fit **out)
{
	if (!source || !out) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	if (elev_res <= 1 || azim_res <= 1) return fail(DJB200_ERR_INVALID_ARGUMENT, "Invalid Resolution"); // dj_brdf.h:2244
	if ((int64_t)(elev_res - 1) * azim_res > (1 << 24)) return fail(DJB200_ERR_UNSUPPORTED, "resolution too large");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	std::vector<FitSourceDev> src;
	std::vector<DevBuf> splines;
	rs = build_sources(source, 1, src, splines);
	if (rs != DJB200_OK) return rs;
	djb200_aniso_fit *f = new djb200_aniso_fit;
	f->er = elev_res; f->ar = azim_res; f->shadow = shadow; f->n = (elev_res - 1) * azim_res;
	cudaGetDevice(&f->device);
	f->src = src[0];
	f->spline = std::move(splines[0]); // the handle keeps the spline points alive
	const size_t n = (size_t)f->n, tab = (size_t)elev_res * azim_res;
	cudaError_t e = f->rowpre.alloc(sizeof(float4) * n);
	if (e == cudaSuccess) e = f->colpre.alloc(sizeof(float4) * n);
	if (e == cudaSuccess) e = f->colrcp.alloc(sizeof(float) * n);
	if (e == cudaSuccess) e = f->ones.alloc(sizeof(double) * n);
	if (e == cudaSuccess) e = f->p22.alloc(sizeof(float) * tab);
	if (e == cudaSuccess) e = f->sigma.alloc(sizeof(float) * tab);
	if (e == cudaSuccess) e = f->fresnel.alloc(sizeof(float) * 3 * elev_res);
	if (e == cudaSuccess) e = f->terms.alloc(sizeof(float) * 7 * 512 * 128);
	if (e == cudaSuccess) e = f->scale.alloc(sizeof(float) * 4);
	if (e == cudaSuccess) e = f->pre_f.alloc(sizeof(float) * aniso_sigma_pre_floats(azim_res));
	if (e == cudaSuccess) e = f->pre_d.alloc(sizeof(double) * aniso_sigma_pre_doubles(azim_res));
	if (e == cudaSuccess) e = f->params.alloc(sizeof(float) * 10);
	if (e == cudaSuccess)
		e = aniso_launch_pre(f->src, elev_res, azim_res, f->rowpre.as<float4>(), f->colpre.as<float4>(), f->colrcp.as<float>(),
		                     f->ones.as<double>(), (cudaStream_t)stream);
	if (e != cudaSuccess) { delete f; return cuda_fail(e, "aniso fit setup"); }
	*out = f;
	return DJB200_OK;
}

djb200_status djb200_aniso_fit_destroy(djb200_aniso_fit *f)
{
	delete f;
	return DJB200_OK;
}

int64_t djb200_aniso_fit_size(const djb200_aniso_fit *f) { return f ? f->n : 0; }

djb200_status djb200_aniso_fit_matvec(djb200_aniso_fit *f, const double *v_in, double *v_out, int64_t row0, int64_t row1,
                                      void *stream)
{
	if (!f || !v_out) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	if (row0 < 0 || row1 > f->n || row0 > row1) return fail(DJB200_ERR_INVALID_ARGUMENT, "row range [%lld, %lld) outside [0, %d)", (long long)row0, (long long)row1, f->n);
	ACU(aniso_launch_matvec(f->er, f->ar, f->rowpre.as<float4>(), f->colpre.as<float4>(), f->colrcp.as<float>(),
	                        v_in ? v_in : f->ones.as<double>(),
	                        v_out, (int)row0, (int)row1, (cudaStream_t)stream));
	return DJB200_OK;
}

djb200_status djb200_aniso_fit_set_iterate(djb200_aniso_fit *f, const double *v, void *stream)
{
	if (!f || !v) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	ACU(aniso_launch_p22(f->er, f->ar, v, f->p22.as<float>(), f->terms.as<float>(), f->scale.as<float>(), (cudaStream_t)stream));
	return DJB200_OK;
}

djb200_status djb200_aniso_fit_sigma(djb200_aniso_fit *f, float *sigma_rows, int64_t row0, int64_t row1, void *stream)
{
	if (!f || !sigma_rows) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	if (row0 < 0 || row1 > f->n || row0 > row1) return fail(DJB200_ERR_INVALID_ARGUMENT, "row range outside [0, %d)", f->n);
	ACU(aniso_launch_sigma(f->er, f->ar, f->p22.as<float>(), f->pre_f.as<float>(), f->pre_d.as<double>(), sigma_rows, (int)row0,
	                       (int)row1, (cudaStream_t)stream));
	return DJB200_OK;
}

djb200_status djb200_aniso_fit_finish(djb200_aniso_fit *f, const float *sigma_rows, void *stream)
{
	if (!f || !sigma_rows) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	ACU(aniso_launch_finish(f->src, f->er, f->ar, f->shadow, f->p22.as<float>(), sigma_rows, f->sigma.as<float>(),
	                        f->fresnel.as<float>(), f->terms.as<float>(), f->params.as<float>(), f->params.as<float>() + 5,
	                        (cudaStream_t)stream));
	return DJB200_OK;
}

djb200_status djb200_aniso_fit_download(djb200_aniso_fit *f, djb200_tabular_anisotropic_fit *r, void *stream)
{
	if (!f || !r) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	if (r->elev_res != f->er || r->azim_res != f->ar) return fail(DJB200_ERR_INVALID_ARGUMENT, "result resolution mismatch");
	if (!r->p22 || !r->sigma || !r->fresnel) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL output array");
	cudaStream_t st = (cudaStream_t)stream;
	const size_t tab = (size_t)f->er * f->ar;
	float params[10];
	ACU(cudaMemcpyAsync(r->p22, f->p22.p, sizeof(float) * tab, cudaMemcpyDeviceToHost, st));
	ACU(cudaMemcpyAsync(r->sigma, f->sigma.p, sizeof(float) * tab, cudaMemcpyDeviceToHost, st));
	ACU(cudaMemcpyAsync(r->fresnel, f->fresnel.p, sizeof(float) * 3 * f->er, cudaMemcpyDeviceToHost, st));
	ACU(cudaMemcpyAsync(params, f->params.p, sizeof params, cudaMemcpyDeviceToHost, st));
	ACU(cudaStreamSynchronize(st));
	memcpy(r->beckmann, params, sizeof(float) * 5);
	memcpy(r->ggx, params + 5, sizeof(float) * 5);
	return DJB200_OK;
}

// ---- the exchange step of a fit whose rows span GPUs: NCCL all-gather over NVLink, inside the library ----------------
// One process per GPU; the caller distributes a unique id (djb200_comm_unique_id on rank 0, broadcast by whatever the host
// program uses -- torch.distributed in dj_brdf_b200/fit_sharded.py, MPI, a file) and every rank creates its communicator.
} // extern "C"

namespace {
struct NcclApi {
	void *handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
};
NcclApi &nccl_api()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		// the soname: a process that already holds an NCCL (PyTorch's) gets that same library back.  (The reverse order -- this
		// load first, a PyTorch built against a newer NCCL imported afterwards -- would hand PyTorch the wrong library, so nothing
		// in libdjb200.so loads NCCL until a communicator is asked for.)
		api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!api.handle) return;
		api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
		api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
		api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
		api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
		api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
		api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
	});
	return api;
}
djb200_status nccl_fail(ncclResult_t r, const char *what)
{
	return fail(DJB200_ERR_CUDA, "%s: NCCL error %d (%s)", what, (int)r, nccl_api().GetErrorString ? nccl_api().GetErrorString(r) : "?");
}
} // namespace

struct djb200_comm {
	ncclComm_t comm;
	int rank, world, device;
};

extern "C" {

djb200_status djb200_comm_unique_id(void *out128)
{
	if (!out128) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	static_assert(sizeof(ncclUniqueId) == DJB200_COMM_ID_BYTES, "unique id size");
	NcclApi &N = nccl_api();
	if (!N.ok) return fail(DJB200_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
	ncclUniqueId id;
	ncclResult_t r = N.GetUniqueId(&id);
	if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
	memcpy(out128, &id, sizeof id);
	return DJB200_OK;
}

djb200_status djb200_comm_create(const void *unique_id128, int32_t world, int32_t rank, djb200_comm **out)
{
	if (!unique_id128 || !out) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	if (world < 1 || rank < 0 || rank >= world) return fail(DJB200_ERR_INVALID_ARGUMENT, "rank %d outside a world of %d", rank, world);
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	NcclApi &N = nccl_api();
	if (!N.ok) return fail(DJB200_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
	ncclUniqueId id;
	memcpy(&id, unique_id128, sizeof id);
	djb200_comm *c = new djb200_comm;
	c->rank = rank;
	c->world = world;
	cudaGetDevice(&c->device);
	ncclResult_t r = N.CommInitRank(&c->comm, world, id, rank);
	if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
	*out = c;
	return DJB200_OK;
}

djb200_status djb200_comm_destroy(djb200_comm *c)
{
	if (c) {
		nccl_api().CommDestroy(c->comm);
		delete c;
	}
	return DJB200_OK;
}

// matrix::eigenvector (dj_brdf.h:2467-2480) + the stages after it, with this rank computing rows [row0, row1) of every
// matrix-vector product and of the projected-area table; the iterate (n doubles) and the table rows (n floats) are
// all-gathered in place after each stage.  comm == NULL: one GPU, no exchange.  Every rank ends with the full result in `f`.
djb200_status djb200_aniso_fit_run(djb200_aniso_fit *f, djb200_comm *comm, int32_t iterations, float *residuals_host,
                                   float *timing_ms, void *stream)
{
	if (!f) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	if (iterations < 1) return fail(DJB200_ERR_INVALID_ARGUMENT, "need at least one power iteration");
	cudaStream_t st = (cudaStream_t)stream;
	const int world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
	if (comm && comm->device != f->device) return fail(DJB200_ERR_INVALID_ARGUMENT, "communicator and fit live on different devices");
	const int64_t n = f->n, chunk = (n + world - 1) / world;
	const int64_t row0 = rank * chunk < n ? rank * chunk : n, row1 = (rank + 1) * chunk < n ? (rank + 1) * chunk : n;
	DevBuf va, vb, srows, resid;
	cudaError_t e = va.alloc(sizeof(double) * world * chunk); // padded to whole chunks: the all-gather is in place
	if (e == cudaSuccess) e = vb.alloc(sizeof(double) * world * chunk);
	if (e == cudaSuccess) e = srows.alloc(sizeof(float) * world * chunk);
	if (e == cudaSuccess) e = resid.alloc(sizeof(float) * iterations);
	if (e != cudaSuccess) return cuda_fail(e, "aniso fit workspace");
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}; // total begin / end, one exchange begin / end
	struct EvGuard { cudaEvent_t *e; ~EvGuard() { for (int k = 0; k < 4; ++k) if (e[k]) cudaEventDestroy(e[k]); } } guard{ev};
	if (timing_ms) for (int k = 0; k < 4; ++k) ACU(cudaEventCreate(&ev[k]));
	float exchange_ms = 0.0f;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> xev; // one pair per exchange, read after the final synchronisation
	auto exchange = [&](void *buf, size_t count, ncclDataType_t type, size_t elem) -> djb200_status {
		if (world == 1) return DJB200_OK;
		cudaEvent_t a = nullptr, b = nullptr;
		if (timing_ms) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
		// NCCL is only touched when there is a communicator (it was loaded to create one): a single-GPU fit never loads it
		ncclResult_t r = nccl_api().AllGather((const char *)buf + (size_t)rank * count * elem, buf, count, type, comm->comm, st);
		if (timing_ms) { cudaEventRecord(b, st); xev.push_back({a, b}); }
		return r == ncclSuccess ? DJB200_OK : nccl_fail(r, "ncclAllGather");
	};
	if (timing_ms) ACU(cudaEventRecord(ev[0], st));
	djb200_status rs = DJB200_OK;
	const double *vin = nullptr;
	double *bufs[2] = {va.as<double>(), vb.as<double>()};
	for (int it = 0; it < iterations && rs == DJB200_OK; ++it) {
		double *vout = bufs[it & 1];
		rs = djb200_aniso_fit_matvec(f, vin, vout, row0, row1, stream);
		if (rs == DJB200_OK) rs = exchange(vout, (size_t)chunk, ncclDouble, sizeof(double));
		if (rs == DJB200_OK && residuals_host) {
			e = aniso_launch_residual((int)n, vin ? vin : f->ones.as<double>(), vout, resid.as<float>() + it, st);
			if (e != cudaSuccess) rs = cuda_fail(e, "residual kernel");
		}
		vin = vout;
	}
	if (rs == DJB200_OK) rs = djb200_aniso_fit_set_iterate(f, vin, stream);
	if (rs == DJB200_OK) rs = djb200_aniso_fit_sigma(f, srows.as<float>(), row0, row1, stream);
	if (rs == DJB200_OK) rs = exchange(srows.as<float>(), (size_t)chunk, ncclFloat, sizeof(float));
	if (rs == DJB200_OK) rs = djb200_aniso_fit_finish(f, srows.as<float>(), stream);
	if (rs == DJB200_OK && timing_ms) {
		e = cudaEventRecord(ev[1], st);
		if (e != cudaSuccess) rs = cuda_fail(e, "event record");
	}
	if (rs == DJB200_OK && residuals_host) {
		e = cudaMemcpyAsync(residuals_host, resid.p, sizeof(float) * iterations, cudaMemcpyDeviceToHost, st);
		if (e != cudaSuccess) rs = cuda_fail(e, "residual download");
	}
	e = cudaStreamSynchronize(st); // the workspaces above are freed on return
	if (e != cudaSuccess && rs == DJB200_OK) rs = cuda_fail(e, "aniso fit");
	for (auto &p : xev) {
		float ms = 0.0f;
		if (rs == DJB200_OK && cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) exchange_ms += ms;
		cudaEventDestroy(p.first);
		cudaEventDestroy(p.second);
	}
	if (rs == DJB200_OK && timing_ms) {
		timing_ms[0] = 0.0f;
		cudaEventElapsedTime(&timing_ms[0], ev[0], ev[1]);
		timing_ms[1] = exchange_ms;
	}
	return rs;
}

// the whole fit on one GPU: the stages above over the full row range, materials one after the other (each
// stage is a grid-wide launch, so one material already fills the device)
djb200_status djb200_fit_tabular_anisotropic(const djb200_source *sources, int32_t n_sources, int32_t elev_res,
                                             int32_t azim_res, int32_t shadow, int32_t iterations,
                                             djb200_tabular_anisotropic_fit *results, void *stream)
{
	if (n_sources < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "negative source count");
	if (elev_res <= 1 || azim_res <= 1) return fail(DJB200_ERR_INVALID_ARGUMENT, "Invalid Resolution");
	if (iterations < 1) return fail(DJB200_ERR_INVALID_ARGUMENT, "need at least one power iteration");
	if (n_sources == 0) return DJB200_OK;
	if (!sources || !results) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	for (int32_t k = 0; k < n_sources; ++k) {
		djb200_aniso_fit *f = nullptr;
		djb200_status rs = djb200_aniso_fit_create(sources + k, elev_res, azim_res, shadow, stream, &f);
		if (rs != DJB200_OK) return rs;
		rs = djb200_aniso_fit_run(f, nullptr, iterations, results[k].residuals, nullptr, stream);
		if (rs == DJB200_OK) rs = djb200_aniso_fit_download(f, results + k, stream);
		delete f;
		if (rs != DJB200_OK) return rs;
	}
	return DJB200_OK;
}

} // extern "C"
