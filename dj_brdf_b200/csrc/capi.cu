// capi.cu -- the C-ABI of libdjb200.so (include/djb200.h): argument checking, error reporting,
// host-side params factories, device-resident table handles and the host<->device staging
// pipeline.  All numerics live in the kernel translation units; there is no CPU implementation
// of any bulk path in this library.
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "djb_internal.h"

namespace djb200 {

std::atomic<uint64_t> g_kernel_launches{0};

static thread_local std::string t_error;

djb200_status fail(djb200_status s, const char *fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	t_error = buf;
	return s;
}

djb200_status cuda_fail(cudaError_t e, const char *what)
{
	if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
		return fail(DJB200_ERR_NO_DEVICE, "%s: no CUDA device (%s); libdjb200 has no CPU fallback", what,
		            cudaGetErrorString(e));
	if (e == cudaErrorMemoryAllocation) return fail(DJB200_ERR_OUT_OF_MEMORY, "%s: %s", what, cudaGetErrorString(e));
	return fail(DJB200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define CU(call)                                                  \
	do {                                                          \
		cudaError_t e__ = (call);                                 \
		if (e__ != cudaSuccess) return cuda_fail(e__, #call);     \
	} while (0)

int sm_count()
{
	static thread_local int cached_dev = -1, cached = 0;
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess) return 148;
	if (dev != cached_dev) {
		cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
		cached_dev = dev;
	}
	return cached > 0 ? cached : 148;
}

djb200_status require_device()
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
	if (n <= 0) return fail(DJB200_ERR_NO_DEVICE, "no CUDA device; libdjb200 has no CPU fallback");
	return DJB200_OK;
}

// ---- host-side params factories (dj_brdf.h:1355-1474), same rounding points as the reference ----
static inline float inv_sqrt_h(float x) { return (float)(1.0 / std::sqrt((double)x)); }

static void set_location_h(djb200_params *p, float tx, float ty)
{
	p->tx_n = tx;
	p->ty_n = ty;
	float x = -tx, y = -ty, z = 1.0f;
	float k = inv_sqrt_h((x * x + y * y) + z * z);
	p->n[0] = k * x;
	p->n[1] = k * y;
	p->n[2] = k * z;
}

static void elliptic_h(float a1, float a2, float phi, djb200_params *p)
{
	float c = (float)std::cos((double)phi), s = (float)std::sin((double)phi);
	float c2 = (float)(2.0 * (double)c * (double)c - 1.0);
	float q1 = a1 * a1, q2 = a2 * a2, t1 = q1 + q2, t2 = q1 - q2;
	p->a1 = a1;
	p->a2 = a2;
	p->phi_a = phi;
	p->ax = (float)std::sqrt(0.5 * (double)(t1 + t2 * c2));
	p->ay = (float)std::sqrt(0.5 * (double)(t1 - t2 * c2));
	p->rho = (q2 - q1) * c * s / (p->ax * p->ay);
	p->sqrt_one_minus_rho2 = (float)std::sqrt(1.0 - (double)(p->rho * p->rho));
	set_location_h(p, 0.0f, 0.0f);
}

static void pdfparams_h(float ax, float ay, float rho, float tx, float ty, djb200_params *p)
{
	p->ax = ax;
	p->ay = ay;
	p->rho = rho;
	p->sqrt_one_minus_rho2 = (float)std::sqrt(1.0 - (double)(rho * rho));
	float qx = ax * ax, qy = ay * ay;
	float cov = (float)((double)(rho * ax * ay) * 2.0);
	float t1 = qx + qy, t2 = qx - qy;
	float t3 = (float)std::sqrt((double)(t2 * t2 + cov * cov));
	p->a1 = (float)std::sqrt(0.5 * (double)(t1 + t3));
	p->a2 = (float)std::sqrt(0.5 * (double)(t1 - t3));
	p->phi_a = (cov != 0.0f) ? (float)std::atan((double)((qx - qy - t3) / cov)) : 0.0f;
	set_location_h(p, tx, ty);
}

// ---- staging pipeline for DJB200_MEM_HOST -------------------------------------------------------
struct BulkIn { const void *host; size_t item; };                 // one item per pair
struct BulkOut { void *host; size_t item; };                      // one item per (material, pair)

// Per host thread: three device staging slots + streams, kept between calls (a renderer calls the C-ABI in a
// loop; cudaMalloc / cudaFree per call would serialise the device).  Grow-only; djb200_release_cache() or thread
// exit frees it.
struct StagingArena {
	int device = -1;
	size_t slot_bytes = 0, aux_bytes = 0;
	void *buf[3] = {nullptr, nullptr, nullptr};
	void *aux = nullptr; // small per-call descriptors (params blocks, Fresnel spline points)
	cudaStream_t st[3] = {nullptr, nullptr, nullptr};
	cudaEvent_t aux_ready = nullptr;
	// upload `bytes` of descriptors to aux + offset; every slot stream waits for it
	cudaError_t upload_aux(const void *host, size_t offset, size_t bytes)
	{
		cudaError_t e = cudaMemcpyAsync((char *)aux + offset, host, bytes, cudaMemcpyHostToDevice, st[0]);
		if (e == cudaSuccess) e = cudaEventRecord(aux_ready, st[0]);
		for (int s = 1; s < 3 && e == cudaSuccess; ++s) e = cudaStreamWaitEvent(st[s], aux_ready, 0);
		return e;
	}
	cudaError_t reserve_aux(size_t bytes)
	{
		if (bytes <= aux_bytes) return cudaSuccess;
		if (aux) cudaFree(aux);
		aux = nullptr;
		aux_bytes = 0;
		size_t want = bytes < 65536 ? 65536 : bytes;
		cudaError_t e = cudaMalloc(&aux, want);
		if (e == cudaSuccess) aux_bytes = want;
		return e;
	}
	void release()
	{
		if (aux) cudaFree(aux);
		aux = nullptr;
		aux_bytes = 0;
		if (aux_ready) cudaEventDestroy(aux_ready);
		aux_ready = nullptr;
		for (int s = 0; s < 3; ++s) {
			if (buf[s]) cudaFree(buf[s]);
			if (st[s]) cudaStreamDestroy(st[s]);
			buf[s] = nullptr;
			st[s] = nullptr;
		}
		slot_bytes = 0;
		device = -1;
	}
	cudaError_t reserve(size_t bytes)
	{
		int dev = 0;
		cudaError_t e = cudaGetDevice(&dev);
		if (e != cudaSuccess) return e;
		if (dev != device) release();
		device = dev;
		for (int s = 0; s < 3; ++s)
			if (!st[s] && (e = cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking)) != cudaSuccess) return e;
		if (!aux_ready && (e = cudaEventCreateWithFlags(&aux_ready, cudaEventDisableTiming)) != cudaSuccess) return e;
		if (bytes <= slot_bytes) return cudaSuccess;
		for (int s = 0; s < 3; ++s) {
			if (buf[s]) cudaFree(buf[s]);
			buf[s] = nullptr;
		}
		slot_bytes = 0;
		for (int s = 0; s < 3; ++s)
			if ((e = cudaMalloc(&buf[s], bytes)) != cudaSuccess) return e;
		slot_bytes = bytes;
		return cudaSuccess;
	}
	~StagingArena() { release(); }
};
static thread_local StagingArena t_arena;

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
// device bytes per staging slot of the host-buffer pipeline (DJB200_CHUNK_MB, default 256); read per call so tests can shrink it
static size_t staging_budget_bytes()
{
	const char *env_mb = getenv("DJB200_CHUNK_MB");
	return (size_t)(env_mb && atoi(env_mb) > 0 ? atoi(env_mb) : 256) << 20;
}
constexpr int64_t MIN_CHUNK_PAIRS = 4096;

// body(dev_in[], dev_out[], chunk_n, stream): enqueue the kernels for one chunk; outputs use
// out_stride = chunk_n.  reps = number of material blocks each output holds per pair.
template <class Body>
static djb200_status host_pipeline(int64_t n, const std::vector<BulkIn> &ins, const std::vector<BulkOut> &outs,
                                   int64_t reps, Body body)
{
	if (n <= 0) return DJB200_OK;
	size_t per_pair = 0;
	for (auto &i : ins) per_pair += i.item;
	for (auto &o : outs) per_pair += o.item * (size_t)reps;
	static const bool trace = getenv("DJB200_TRACE") != nullptr;
	const size_t budget = staging_budget_bytes();
	auto t_begin = std::chrono::steady_clock::now();
	int64_t chunk = (int64_t)(budget / (per_pair ? per_pair : 1));
	chunk = chunk < MIN_CHUNK_PAIRS ? MIN_CHUNK_PAIRS : chunk; // callers tile over materials so that this floor stays inside the budget
	chunk &= ~(int64_t)3;
	if (chunk > n) chunk = n;
	const int SLOTS = 3;
	size_t slot_bytes = 0;
	for (auto &i : ins) slot_bytes += align_up(i.item * (size_t)chunk);
	for (auto &o : outs) slot_bytes += align_up(o.item * (size_t)chunk * (size_t)reps);
	StagingArena &A = t_arena;
	cudaError_t e0 = A.reserve(slot_bytes);
	if (e0 != cudaSuccess) {
		A.release();
		return cuda_fail(e0, "staging buffers");
	}
	djb200_status rc = DJB200_OK;
#define PCU(call)                                                                                          \
	do {                                                                                                   \
		cudaError_t e__ = (call);                                                                          \
		if (e__ != cudaSuccess) {                                                                          \
			rc = cuda_fail(e__, #call);                                                                    \
			for (int s__ = 0; s__ < SLOTS; ++s__) cudaStreamSynchronize(A.st[s__]);                        \
			return rc;                                                                                     \
		}                                                                                                  \
	} while (0)

	auto t_alloc = std::chrono::steady_clock::now();
	std::vector<void *> din(ins.size()), dout(outs.size());
	int c = 0;
	for (int64_t off = 0; off < n; off += chunk, ++c) {
		const int si = c % SLOTS;
		cudaStream_t st = A.st[si];
		char *base = (char *)A.buf[si];
		size_t pos = 0;
		for (size_t k = 0; k < ins.size(); ++k) { din[k] = base + pos; pos += align_up(ins[k].item * (size_t)chunk); }
		for (size_t k = 0; k < outs.size(); ++k) { dout[k] = base + pos; pos += align_up(outs[k].item * (size_t)chunk * (size_t)reps); }
		int64_t cn = n - off < chunk ? n - off : chunk;
		for (size_t k = 0; k < ins.size(); ++k)
			PCU(cudaMemcpyAsync(din[k], (const char *)ins[k].host + (size_t)off * ins[k].item, ins[k].item * (size_t)cn,
			                    cudaMemcpyHostToDevice, st));
		cudaError_t e = body(din, dout, cn, st);
		if (e != cudaSuccess) {
			rc = cuda_fail(e, "kernel launch");
			for (int s = 0; s < SLOTS; ++s) cudaStreamSynchronize(A.st[s]);
			return rc;
		}
		for (size_t k = 0; k < outs.size(); ++k) {
			// `reps` rows of cn items: device rows are cn items apart, host rows n items apart (material-major output)
			if (reps == 1)
				PCU(cudaMemcpyAsync((char *)outs[k].host + (size_t)off * outs[k].item, dout[k], outs[k].item * (size_t)cn,
				                    cudaMemcpyDeviceToHost, st));
			else if ((size_t)n * outs[k].item <= ((size_t)1 << 31) - 1)
				PCU(cudaMemcpy2DAsync((char *)outs[k].host + (size_t)off * outs[k].item, (size_t)n * outs[k].item, dout[k],
				                      (size_t)cn * outs[k].item, (size_t)cn * outs[k].item, (size_t)reps,
				                      cudaMemcpyDeviceToHost, st));
			else // host pitch beyond what a pitched copy accepts (cudaDeviceProp::memPitch): one 1-D copy per material row
				for (int64_t r = 0; r < reps; ++r)
					PCU(cudaMemcpyAsync((char *)outs[k].host + ((size_t)r * (size_t)n + (size_t)off) * outs[k].item,
					                    (const char *)dout[k] + (size_t)r * (size_t)cn * outs[k].item, outs[k].item * (size_t)cn,
					                    cudaMemcpyDeviceToHost, st));
		}
	}
	auto t_enq = std::chrono::steady_clock::now();
	for (int s = 0; s < SLOTS; ++s) PCU(cudaStreamSynchronize(A.st[s]));
#undef PCU
	if (trace) {
		auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
			return std::chrono::duration<double, std::milli>(b - a).count();
		};
		fprintf(stderr, "[djb200] host_pipeline n=%lld chunk=%lld: reserve %.2f ms, enqueue %.2f ms, drain %.2f ms\n",
		        (long long)n, (long long)chunk, ms(t_begin, t_alloc), ms(t_alloc, t_enq),
		        ms(t_enq, std::chrono::steady_clock::now()));
	}
	return rc;
}

// small host descriptor -> device scratch (stream ordered).  The device's default memory pool gives freed blocks back to
// the OS at every synchronisation unless told otherwise, which turns the next cudaMallocAsync into a driver allocation
// (0.5 - 3 ms measured per DEVICE-memory call after a sync: more than a 2e7-pair table-BRDF kernel).  A small release
// threshold, set once per device, keeps the few hundred bytes of descriptors cached in the pool.
static void keep_descriptor_pool_warm()
{
	static std::atomic<unsigned> done_mask{0};
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return;
	if (done_mask.load(std::memory_order_relaxed) & (1u << dev)) return;
	cudaMemPool_t pool;
	if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
		uint64_t cur = 0, want = 8ull << 20;
		if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur) == cudaSuccess && cur < want)
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &want);
	}
	done_mask.fetch_or(1u << dev, std::memory_order_relaxed);
}
static cudaError_t upload_small(const void *host, size_t bytes, void **dev, cudaStream_t st)
{
	keep_descriptor_pool_warm();
	cudaError_t e = cudaMallocAsync(dev, bytes, st);
	if (e != cudaSuccess) return e;
	return cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, st);
}

// `tab` != NULL: the queries run on a tabulated BRDF (kernels_tabular.cu) and `mf` is ignored
static djb200_status microfacet_call(int op, const djb200_microfacet *mf, const djb200_params *params,
                                     int64_t n_params, int layout, const float *a, const float *b, int64_t n,
                                     float *out0, float *out1, float *out2, int mem, void *stream,
                                     const djb200_tabular *tab = nullptr)
{
	djb200_microfacet tab_desc;
	if (tab) { // only the shadowing flag of the descriptor is used on this path
		memset(&tab_desc, 0, sizeof tab_desc);
		tab_desc.ndf = DJB200_NDF_GGX;
		tab_desc.shadow = tab->shadow;
		mf = &tab_desc;
		int dev = 0;
		cudaGetDevice(&dev);
		if (dev != tab->device) return fail(DJB200_ERR_INVALID_ARGUMENT, "tabular handle lives on device %d, current device is %d", tab->device, dev);
	}
	auto launch = [tab](const MfLaunch &X, cudaStream_t s) {
		if (!tab) return launch_microfacet(X, s);
		if (tab->azim_res > 0) return launch_tabular_aniso_query(tab->tables, tab->res, tab->azim_res, tab->n_qf1, X, s);
		return launch_tabular_query(tab->tables, tab->res, X, s);
	};
	if (!mf) return fail(DJB200_ERR_INVALID_ARGUMENT, "microfacet descriptor is NULL");
	if (mf->ndf != DJB200_NDF_BECKMANN && mf->ndf != DJB200_NDF_GGX)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown ndf %d", mf->ndf);
	if (mf->fresnel.kind < 0 || mf->fresnel.kind > DJB200_FRESNEL_SPLINE)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown fresnel kind %d", mf->fresnel.kind);
	if (mf->fresnel.kind == DJB200_FRESNEL_SPLINE && (!mf->fresnel.points || mf->fresnel.n_points < 1))
		return fail(DJB200_ERR_INVALID_ARGUMENT, "spline fresnel needs points");
	if (n < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "negative pair count");
	if (layout != DJB200_PARAMS_BROADCAST && layout != DJB200_PARAMS_PER_PAIR)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown params layout %d", layout);
	if (mem != DJB200_MEM_HOST && mem != DJB200_MEM_DEVICE)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown memory space %d", mem);
	// NULL params => params::standard(), like the reference's NULL user_param (dj_brdf.h:1532-1534)
	djb200_params standard;
	if (!params) {
		elliptic_h(1.0f, 1.0f, 0.0f, &standard);
		params = &standard;
		n_params = 1;
		layout = DJB200_PARAMS_BROADCAST;
	}
	if (layout == DJB200_PARAMS_BROADCAST && n_params < 1)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "need at least one params block");
	if (layout == DJB200_PARAMS_PER_PAIR && n_params != n)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "PER_PAIR layout needs n_params == n");
	if (n == 0) return DJB200_OK;
	if (!a || !b) return fail(DJB200_ERR_INVALID_ARGUMENT, "direction arrays are NULL");
	if (op != OP_EVALP_IS && !out0) return fail(DJB200_ERR_INVALID_ARGUMENT, "output array is NULL");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;

	const bool uses_u = (op == OP_SAMPLE || op == OP_EVALP_IS);
	const size_t a_item = uses_u ? 8 : 12;
	const size_t out0_item = (op == OP_PDF) ? 4 : 12;

	MfLaunch L;
	memset(&L, 0, sizeof L);
	L.op = op;
	L.ndf = mf->ndf;
	L.shadow = mf->shadow;
	L.fresnel_kind = mf->fresnel.kind;
	memcpy(L.fv, mf->fresnel.v, sizeof L.fv);
	L.layout = layout;
	L.n_params = n_params;

	if (mem == DJB200_MEM_DEVICE) {
		cudaStream_t st = (cudaStream_t)stream;
		struct Scratch { // stream-ordered descriptor scratch, released on every exit path
			void *p = nullptr;
			cudaStream_t st;
			explicit Scratch(cudaStream_t s) : st(s) {}
			~Scratch() { if (p) cudaFreeAsync(p, st); }
		} d_params(st), d_spline(st);
		if (layout == DJB200_PARAMS_BROADCAST && !tab && n_params <= MF_INLINE_PARAMS) {
			L.params = nullptr; // inside the kernel arguments: no upload (a pageable H2D copy synchronises the stream first)
			L.params_host = params;
		} else if (layout == DJB200_PARAMS_BROADCAST) {
			CU(upload_small(params, sizeof(djb200_params) * (size_t)n_params, &d_params.p, st));
			L.params = d_params.p;
		} else {
			L.params = params; // bulk device array
		}
		if (L.fresnel_kind == DJB200_FRESNEL_SPLINE) {
			CU(upload_small(mf->fresnel.points, sizeof(float) * 3 * (size_t)mf->fresnel.n_points, &d_spline.p, st));
			L.spline_pts = (const float *)d_spline.p;
			L.spline_n = mf->fresnel.n_points;
		}
		L.a = a; L.b = b; L.n = n; L.out_stride = n;
		L.out0 = out0; L.out1 = out1; L.out2 = out2;
		cudaError_t e = launch(L, st);
		if (e != cudaSuccess) return cuda_fail(e, "microfacet kernel launch");
		return DJB200_OK;
	}

	// host memory: stage through the device in chunks; descriptors go to the arena's aux area (no per-call malloc).
	// A BROADCAST call with very many materials is tiled over blocks of materials, so that one staging slot (the pipeline's
	// smallest chunk x every material's output) stays inside the DJB200_CHUNK_MB budget.
	if (layout == DJB200_PARAMS_BROADCAST) {
		const size_t in_bytes = a_item + 12;
		const size_t out_bytes = (out0 ? out0_item : 0) + (out1 ? 12 : 0) + (out2 ? 4 : 0);
		const size_t per_chunk = staging_budget_bytes() / (size_t)MIN_CHUNK_PAIRS;
		int64_t max_mats = out_bytes && per_chunk > in_bytes ? (int64_t)((per_chunk - in_bytes) / out_bytes) : 1;
		if (max_mats < 1) max_mats = 1;
		if (n_params > max_mats) {
			for (int64_t m0 = 0; m0 < n_params; m0 += max_mats) {
				const int64_t mc = n_params - m0 < max_mats ? n_params - m0 : max_mats;
				djb200_status rc = microfacet_call(op, mf, params + m0, mc, layout, a, b, n,
				                                   out0 ? out0 + (size_t)m0 * (size_t)n * (out0_item / 4) : nullptr,
				                                   out1 ? out1 + (size_t)m0 * (size_t)n * 3 : nullptr,
				                                   out2 ? out2 + (size_t)m0 * (size_t)n : nullptr, mem, stream, tab);
				if (rc != DJB200_OK) return rc;
			}
			return DJB200_OK;
		}
	}
	void *d_params = nullptr;
	{
		StagingArena &A = t_arena;
		const size_t pbytes = layout == DJB200_PARAMS_BROADCAST ? sizeof(djb200_params) * (size_t)n_params : 0;
		const size_t poff = (pbytes + 255) & ~(size_t)255;
		const size_t sbytes = L.fresnel_kind == DJB200_FRESNEL_SPLINE ? sizeof(float) * 3 * (size_t)mf->fresnel.n_points : 0;
		CU(A.reserve(0));
		CU(cudaStreamSynchronize(A.st[0])); // aux may still be read by the previous call's last chunks
		CU(cudaStreamSynchronize(A.st[1]));
		CU(cudaStreamSynchronize(A.st[2]));
		CU(A.reserve_aux(poff + sbytes + 256));
		if (pbytes) {
			CU(A.upload_aux(params, 0, pbytes));
			d_params = A.aux;
		}
		if (sbytes) {
			CU(A.upload_aux(mf->fresnel.points, poff, sbytes));
			L.spline_pts = (const float *)((char *)A.aux + poff);
			L.spline_n = mf->fresnel.n_points;
		}
	}
	std::vector<BulkIn> ins = {{a, a_item}, {b, 12}};
	if (layout == DJB200_PARAMS_PER_PAIR) ins.push_back({params, sizeof(djb200_params)});
	std::vector<BulkOut> outs;
	int slot0 = -1, slot1 = -1, slot2 = -1;
	if (out0) { slot0 = (int)outs.size(); outs.push_back({out0, out0_item}); }
	if (out1) { slot1 = (int)outs.size(); outs.push_back({out1, 12}); }
	if (out2) { slot2 = (int)outs.size(); outs.push_back({out2, 4}); }
	const int64_t reps = (layout == DJB200_PARAMS_BROADCAST) ? n_params : 1;
	djb200_status rc = host_pipeline(n, ins, outs, reps,
		[&](const std::vector<void *> &din, const std::vector<void *> &dout, int64_t cn, cudaStream_t st) {
			MfLaunch C = L;
			C.a = (const float *)din[0];
			C.b = (const float *)din[1];
			C.params = (layout == DJB200_PARAMS_PER_PAIR) ? din[2] : d_params;
			C.n_params = (layout == DJB200_PARAMS_PER_PAIR) ? cn : n_params;
			C.n = cn;
			C.out_stride = cn;
			C.out0 = slot0 >= 0 ? (float *)dout[slot0] : nullptr;
			C.out1 = slot1 >= 0 ? (float *)dout[slot1] : nullptr;
			C.out2 = slot2 >= 0 ? (float *)dout[slot2] : nullptr;
			return launch(C, st);
		});
	return rc;
}

// LEAN-filtered shading: per-pair params from the renderer's texture fetches, fused in front of the query
// (op < 0: only the params construction).  Bulk arrays: E (n x 5), alpha (n x 3, optional), a, b, outputs.
static djb200_status lean_shading_call(int op, const djb200_microfacet *mf, const djb200_lean_shading *cfg,
                                       const float *alpha, const float *E, const float *a, const float *b, int64_t n,
                                       float *out0, float *out1, float *out2, int mem, void *stream)
{
	if (!cfg) return fail(DJB200_ERR_INVALID_ARGUMENT, "lean shading descriptor is NULL");
	if (!(cfg->dmap_scale >= 0.0f)) return fail(DJB200_ERR_INVALID_ARGUMENT, "Invalid scale"); // DJB_ASSERT, dj_brdf.h:2022
	if (!cfg->alpha_per_pair && !(cfg->alpha[0] > 0.0f && cfg->alpha[1] > 0.0f))
		return fail(DJB200_ERR_INVALID_ARGUMENT, "Invalid ellipse radii"); // DJB_ASSERT, dj_brdf.h:1453
	if (n < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "negative pair count");
	if (mem != DJB200_MEM_HOST && mem != DJB200_MEM_DEVICE)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown memory space %d", mem);
	if (op >= 0) {
		if (!mf) return fail(DJB200_ERR_INVALID_ARGUMENT, "microfacet descriptor is NULL");
		if (mf->ndf != DJB200_NDF_BECKMANN && mf->ndf != DJB200_NDF_GGX)
			return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown ndf %d", mf->ndf);
		if (mf->fresnel.kind < 0 || mf->fresnel.kind > DJB200_FRESNEL_SPLINE)
			return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown fresnel kind %d", mf->fresnel.kind);
		if (mf->fresnel.kind == DJB200_FRESNEL_SPLINE && (!mf->fresnel.points || mf->fresnel.n_points < 1))
			return fail(DJB200_ERR_INVALID_ARGUMENT, "spline fresnel needs points");
	}
	if (n == 0) return DJB200_OK;
	if (!E || (cfg->alpha_per_pair && !alpha)) return fail(DJB200_ERR_INVALID_ARGUMENT, "LEAN moment / roughness arrays are NULL");
	if (op >= 0 && (!a || !b)) return fail(DJB200_ERR_INVALID_ARGUMENT, "direction arrays are NULL");
	if (op != OP_EVALP_IS && !out0) return fail(DJB200_ERR_INVALID_ARGUMENT, "output array is NULL");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;

	MfLaunch L;
	memset(&L, 0, sizeof L);
	L.op = op;
	L.layout = PARAMS_LEAN_SHADING;
	L.lean_bias = cfg->bias;
	L.lean_dmap_scale = cfg->dmap_scale;
	L.lean_filtering = cfg->lean_filtering;
	memcpy(L.lean_alpha0, cfg->alpha, sizeof L.lean_alpha0);
	if (op >= 0) {
		L.ndf = mf->ndf;
		L.shadow = mf->shadow;
		L.fresnel_kind = mf->fresnel.kind;
		memcpy(L.fv, mf->fresnel.v, sizeof L.fv);
	}
	const bool uses_u = (op == OP_SAMPLE || op == OP_EVALP_IS);
	const size_t out0_item = op < 0 ? 48 : (op == OP_PDF ? 4 : 12);

	std::vector<BulkIn> ins = {{E, 20}};
	int slot_alpha = -1, slot_a = -1, slot_b = -1;
	if (cfg->alpha_per_pair) { slot_alpha = (int)ins.size(); ins.push_back({alpha, 12}); }
	if (op >= 0) {
		slot_a = (int)ins.size(); ins.push_back({a, uses_u ? (size_t)8 : (size_t)12});
		slot_b = (int)ins.size(); ins.push_back({b, 12});
	}
	std::vector<BulkOut> outs;
	int slot0 = -1, slot1 = -1, slot2 = -1;
	if (out0) { slot0 = (int)outs.size(); outs.push_back({out0, out0_item}); }
	if (out1) { slot1 = (int)outs.size(); outs.push_back({out1, 12}); }
	if (out2) { slot2 = (int)outs.size(); outs.push_back({out2, 4}); }

	void *d_spline = nullptr;
	const bool spline = op >= 0 && L.fresnel_kind == DJB200_FRESNEL_SPLINE;
	if (spline) { // small descriptor: stream-ordered scratch for device calls, a plain allocation for the staged host path
		const size_t bytes = sizeof(float) * 3 * (size_t)mf->fresnel.n_points;
		if (mem == DJB200_MEM_DEVICE) {
			CU(upload_small(mf->fresnel.points, bytes, &d_spline, (cudaStream_t)stream));
		} else {
			CU(cudaMalloc(&d_spline, bytes));
			cudaError_t e = cudaMemcpy(d_spline, mf->fresnel.points, bytes, cudaMemcpyHostToDevice);
			if (e != cudaSuccess) { cudaFree(d_spline); return cuda_fail(e, "fresnel spline upload"); }
		}
		L.spline_pts = (const float *)d_spline;
		L.spline_n = mf->fresnel.n_points;
	}
	auto body = [&](const std::vector<void *> &din, const std::vector<void *> &dout, int64_t cn, cudaStream_t st) {
		MfLaunch C = L;
		C.lean_E = (const float *)din[0];
		C.lean_alpha = slot_alpha >= 0 ? (const float *)din[slot_alpha] : nullptr;
		C.n = cn;
		C.n_params = cn;
		C.out_stride = cn;
		if (op < 0) return launch_lean_shading_params(C, (float *)dout[slot0], st);
		C.a = (const float *)din[slot_a];
		C.b = (const float *)din[slot_b];
		C.out0 = slot0 >= 0 ? (float *)dout[slot0] : nullptr;
		C.out1 = slot1 >= 0 ? (float *)dout[slot1] : nullptr;
		C.out2 = slot2 >= 0 ? (float *)dout[slot2] : nullptr;
		return launch_microfacet(C, st);
	};
	djb200_status rc;
	if (mem == DJB200_MEM_DEVICE) {
		std::vector<void *> din, dout;
		for (auto &i : ins) din.push_back(const_cast<void *>(i.host));
		for (auto &o : outs) dout.push_back(o.host);
		cudaError_t e = body(din, dout, n, (cudaStream_t)stream);
		if (d_spline) cudaFreeAsync(d_spline, (cudaStream_t)stream);
		rc = e == cudaSuccess ? DJB200_OK : cuda_fail(e, "lean shading launch");
	} else {
		rc = host_pipeline(n, ins, outs, 1, body);
		if (d_spline) cudaFree(d_spline);
	}
	return rc;
}

// generic "n items in, n items out" call used by the table / frame / LEAN entry points
template <class Body>
static djb200_status map_call(int64_t n, const std::vector<BulkIn> &ins, const std::vector<BulkOut> &outs, int mem,
                              void *stream, Body body)
{
	if (n < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "negative element count");
	if (mem != DJB200_MEM_HOST && mem != DJB200_MEM_DEVICE)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown memory space %d", mem);
	if (n == 0) return DJB200_OK;
	for (auto &i : ins) if (!i.host) return fail(DJB200_ERR_INVALID_ARGUMENT, "input array is NULL");
	for (auto &o : outs) if (!o.host) return fail(DJB200_ERR_INVALID_ARGUMENT, "output array is NULL");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	if (mem == DJB200_MEM_DEVICE) {
		std::vector<void *> din, dout;
		for (auto &i : ins) din.push_back(const_cast<void *>(i.host));
		for (auto &o : outs) dout.push_back(o.host);
		cudaError_t e = body(din, dout, n, (cudaStream_t)stream);
		if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
		return DJB200_OK;
	}
	return host_pipeline(n, ins, outs, 1, body);
}

} // namespace djb200

using namespace djb200;

// ====================================================================================================
extern "C" {

const char *djb200_last_error(void) { return t_error.c_str(); }
const char *djb200_version(void) { return "djb200 0.1 (sm_100a)"; }

djb200_status djb200_device_count(int *count)
{
	if (!count) return fail(DJB200_ERR_INVALID_ARGUMENT, "count is NULL");
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) {
		*count = 0;
		return cuda_fail(e, "cudaGetDeviceCount");
	}
	*count = n;
	return DJB200_OK;
}

djb200_status djb200_set_device(int device)
{
	CU(cudaSetDevice(device));
	return DJB200_OK;
}

uint64_t djb200_kernel_launch_count(void) { return g_kernel_launches.load(); }

djb200_status djb200_debug_force_generic(int on)
{
	g_force_generic.store(on ? 1 : 0);
	return DJB200_OK;
}

djb200_status djb200_set_precision(int mode)
{
	if (mode != DJB200_PRECISION_REFERENCE_BITS && mode != DJB200_PRECISION_1E5)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown precision mode %d", mode);
	g_fast_tier.store(mode == DJB200_PRECISION_1E5 ? 1 : 0);
	return DJB200_OK;
}

int djb200_get_precision(void) { return g_fast_tier.load() ? DJB200_PRECISION_1E5 : DJB200_PRECISION_REFERENCE_BITS; }

djb200_status djb200_debug_beckmann_compaction(int on)
{
	g_beck_compact.store(on ? 1 : 0);
	return DJB200_OK;
}

djb200_status djb200_release_cache(void)
{
	t_arena.release();
	return DJB200_OK;
}

djb200_status djb200_params_standard(djb200_params *out) { return djb200_params_elliptic(1.0f, 1.0f, 0.0f, out); }
djb200_status djb200_params_isotropic(float a, djb200_params *out) { return djb200_params_elliptic(a, a, 0.0f, out); }

djb200_status djb200_params_elliptic(float a1, float a2, float phi_a, djb200_params *out)
{
	if (!out) return fail(DJB200_ERR_INVALID_ARGUMENT, "out is NULL");
	// DJB_ASSERT(a1 > 0.0 && a2 > 0.0 && "Invalid ellipse radii"), dj_brdf.h:1453
	if (!(a1 > 0.0f && a2 > 0.0f)) return fail(DJB200_ERR_INVALID_ARGUMENT, "Invalid ellipse radii");
	elliptic_h(a1, a2, phi_a, out);
	return DJB200_OK;
}

djb200_status djb200_params_pdfparams(float ax, float ay, float rho, float tx_n, float ty_n, djb200_params *out)
{
	if (!out) return fail(DJB200_ERR_INVALID_ARGUMENT, "out is NULL");
	// dj_brdf.h:1466-1467
	if (!(ax > 0.0f && ay > 0.0f)) return fail(DJB200_ERR_INVALID_ARGUMENT, "Invalid scale parameters");
	if (!(std::fabs((double)rho) < 1.0)) return fail(DJB200_ERR_INVALID_ARGUMENT, "Invalid correlation parameter");
	pdfparams_h(ax, ay, rho, tx_n, ty_n, out);
	return DJB200_OK;
}

djb200_status djb200_microfacet_eval(const djb200_microfacet *mf, const djb200_params *params, int64_t n_params,
                                     int params_layout, const float *wi, const float *wo, int64_t n,
                                     float *out_rgb, int mem, void *stream)
{
	return microfacet_call(OP_EVAL, mf, params, n_params, params_layout, wi, wo, n, out_rgb, nullptr, nullptr, mem, stream);
}

djb200_status djb200_microfacet_evalp(const djb200_microfacet *mf, const djb200_params *params, int64_t n_params,
                                      int params_layout, const float *wi, const float *wo, int64_t n,
                                      float *out_rgb, int mem, void *stream)
{
	return microfacet_call(OP_EVALP, mf, params, n_params, params_layout, wi, wo, n, out_rgb, nullptr, nullptr, mem, stream);
}

djb200_status djb200_microfacet_pdf(const djb200_microfacet *mf, const djb200_params *params, int64_t n_params,
                                    int params_layout, const float *wi, const float *wo, int64_t n, float *out_pdf,
                                    int mem, void *stream)
{
	return microfacet_call(OP_PDF, mf, params, n_params, params_layout, wi, wo, n, out_pdf, nullptr, nullptr, mem, stream);
}

djb200_status djb200_microfacet_sample(const djb200_microfacet *mf, const djb200_params *params, int64_t n_params,
                                       int params_layout, const float *u, const float *wo, int64_t n, float *out_wi,
                                       int mem, void *stream)
{
	return microfacet_call(OP_SAMPLE, mf, params, n_params, params_layout, u, wo, n, out_wi, nullptr, nullptr, mem, stream);
}

djb200_status djb200_microfacet_evalp_is(const djb200_microfacet *mf, const djb200_params *params, int64_t n_params,
                                         int params_layout, const float *u, const float *wo, int64_t n,
                                         float *out_weight_rgb, float *out_wi, float *out_pdf, int mem, void *stream)
{
	if (!out_weight_rgb && !out_wi && !out_pdf) return fail(DJB200_ERR_INVALID_ARGUMENT, "all outputs are NULL");
	return microfacet_call(OP_EVALP_IS, mf, params, n_params, params_layout, u, wo, n, out_weight_rgb, out_wi, out_pdf,
	                       mem, stream);
}

djb200_status djb200_io_to_hd(const float *wi, const float *wo, int64_t n, float *h, float *d, int mem, void *stream)
{
	return map_call(n, {{wi, 12}, {wo, 12}}, {{h, 12}, {d, 12}}, mem, stream,
		[](const std::vector<void *> &i, const std::vector<void *> &o, int64_t cn, cudaStream_t st) {
			return launch_io_to_hd((const float *)i[0], (const float *)i[1], cn, (float *)o[0], (float *)o[1], st);
		});
}

djb200_status djb200_hd_to_io(const float *h, const float *d, int64_t n, float *wi, float *wo, int mem, void *stream)
{
	return map_call(n, {{h, 12}, {d, 12}}, {{wi, 12}, {wo, 12}}, mem, stream,
		[](const std::vector<void *> &i, const std::vector<void *> &o, int64_t cn, cudaStream_t st) {
			return launch_hd_to_io((const float *)i[0], (const float *)i[1], cn, (float *)o[0], (float *)o[1], st);
		});
}

// ---- djb::tabular as a BRDF ----------------------------------------------------------------------------
djb200_status djb200_tabular_create(const djb200_tabular_fit *fit, int32_t shadow, djb200_tabular **out)
{
	if (!fit || !out) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	if (fit->res <= 2 || !fit->p22 || !fit->sigma || !fit->qf || !fit->fresnel)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "incomplete fit (needs res > 2, p22, sigma, qf, fresnel)");
	if (fit->res > 8192) return fail(DJB200_ERR_UNSUPPORTED, "resolution %d too large", fit->res);
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	const size_t res = (size_t)fit->res;
	std::vector<float> h(7 * res, 0.0f);
	memcpy(h.data(), fit->p22, 4 * res);
	memcpy(h.data() + res, fit->sigma, 4 * res);
	memcpy(h.data() + 2 * res, fit->qf, 4 * res);
	memcpy(h.data() + 3 * res, fit->fresnel, 12 * res);
	if (fit->cdf) memcpy(h.data() + 6 * res, fit->cdf, 4 * res); // only radial_query(CDF_RADIAL) reads it
	float *d = nullptr;
	CU(cudaMalloc(&d, 4 * h.size()));
	cudaError_t e = cudaMemcpy(d, h.data(), 4 * h.size(), cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { cudaFree(d); return cuda_fail(e, "tabular upload"); }
	djb200_tabular *t = new djb200_tabular;
	t->tables = d; t->res = fit->res; t->shadow = shadow ? 1 : 0; t->azim_res = 0;
	cudaGetDevice(&t->device);
	*out = t;
	return DJB200_OK;
}

djb200_status djb200_tabular_anisotropic_create(const djb200_tabular_anisotropic_fit *fit, int32_t shadow, djb200_tabular **out)
{
	if (!fit || !out) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	if (fit->elev_res <= 1 || fit->azim_res <= 1 || !fit->p22 || !fit->sigma || !fit->fresnel)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "incomplete fit (needs resolutions > 1, p22, sigma, fresnel)");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	const size_t tab = (size_t)fit->elev_res * fit->azim_res, er = (size_t)fit->elev_res;
	if (fit->elev_res > 4096 || fit->azim_res > 4096) return fail(DJB200_ERR_UNSUPPORTED, "resolution too large");
	std::vector<float> h(2 * tab + 3 * er);
	memcpy(h.data(), fit->p22, 4 * tab);
	memcpy(h.data() + tab, fit->sigma, 4 * tab);
	memcpy(h.data() + 2 * tab, fit->fresnel, 12 * er);
	float *d = nullptr;
	CU(cudaMalloc(&d, 4 * aniso_table_floats(fit->elev_res, fit->azim_res)));
	cudaError_t e = cudaMemcpy(d, h.data(), 4 * h.size(), cudaMemcpyHostToDevice);
	// the marginal / conditional sampling tables are built on the device from the p22 block (dj_brdf.h:2266-2272)
	int counts[2] = {0, 0};
	if (e == cudaSuccess) e = build_aniso_sampling_tables(d, fit->elev_res, fit->azim_res, counts, 0);
	if (e != cudaSuccess) { cudaFree(d); return cuda_fail(e, "tabular_anisotropic upload / sampling tables"); }
	djb200_tabular *t = new djb200_tabular;
	t->tables = d; t->res = fit->elev_res; t->azim_res = fit->azim_res; t->shadow = shadow ? 1 : 0;
	t->n_qf1 = counts[0]; t->n_qf2 = counts[1];
	cudaGetDevice(&t->device);
	*out = t;
	return DJB200_OK;
}

djb200_status djb200_microfacet_component(const djb200_microfacet *mf, const djb200_params *params, int what, const float *a,
                                          const float *b, const float *c, int64_t n, float *out, int mem, void *stream)
{
	if (!mf) return fail(DJB200_ERR_INVALID_ARGUMENT, "microfacet descriptor is NULL");
	if (mf->ndf != DJB200_NDF_BECKMANN && mf->ndf != DJB200_NDF_GGX) return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown ndf %d", mf->ndf);
	if (mf->fresnel.kind < 0 || mf->fresnel.kind > DJB200_FRESNEL_SPLINE)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown fresnel kind %d", mf->fresnel.kind);
	if (what < DJB200_COMP_NDF || what > DJB200_COMP_FRESNEL) return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown component %d", what);
	const bool need_b = what == DJB200_COMP_GAF || what == DJB200_COMP_G1 || what == DJB200_COMP_VP22 || what == DJB200_COMP_VNDF;
	const bool need_c = what == DJB200_COMP_GAF;
	if (n > 0 && ((need_b && !b) || (need_c && !c))) return fail(DJB200_ERR_INVALID_ARGUMENT, "direction array is NULL");
	djb200_params P;
	if (params) P = *params; else elliptic_h(1.0f, 1.0f, 0.0f, &P);
	const djb200_microfacet M = *mf;
	const bool spline = what == DJB200_COMP_FRESNEL && M.fresnel.kind == DJB200_FRESNEL_SPLINE;
	if (spline && (!M.fresnel.points || M.fresnel.n_points < 1)) return fail(DJB200_ERR_INVALID_ARGUMENT, "spline fresnel needs points");
	void *d_spline = nullptr;
	if (spline && n > 0) {
		djb200_status rs = require_device();
		if (rs != DJB200_OK) return rs;
		const size_t bytes = sizeof(float) * 3 * (size_t)M.fresnel.n_points;
		CU(cudaMalloc(&d_spline, bytes));
		cudaError_t e = cudaMemcpy(d_spline, M.fresnel.points, bytes, cudaMemcpyHostToDevice);
		if (e != cudaSuccess) { cudaFree(d_spline); return cuda_fail(e, "fresnel spline upload"); }
	}
	std::vector<BulkIn> ins = {{a, 12}};
	int sb = -1, sc = -1;
	if (need_b) { sb = (int)ins.size(); ins.push_back({b, 12}); }
	if (need_c) { sc = (int)ins.size(); ins.push_back({c, 12}); }
	const size_t out_item = what == DJB200_COMP_FRESNEL ? 12 : 4;
	djb200_status rc = map_call(n, ins, {{out, out_item}}, mem, stream,
		[&](const std::vector<void *> &i, const std::vector<void *> &o, int64_t cn, cudaStream_t st) {
			return launch_microfacet_component(M.ndf, M.shadow, M.fresnel.kind, M.fresnel.v, (const float *)d_spline,
			                                   M.fresnel.n_points, &P, what, (const float *)i[0],
			                                   sb >= 0 ? (const float *)i[sb] : nullptr, sc >= 0 ? (const float *)i[sc] : nullptr, cn,
			                                   (float *)o[0], st);
		});
	if (d_spline) {
		if (mem == DJB200_MEM_DEVICE) cudaStreamSynchronize((cudaStream_t)stream);
		cudaFree(d_spline);
	}
	return rc;
}

djb200_status djb200_radial_query(int what, int ndf, const djb200_tabular *t, const float *x, int64_t n, float *out, int mem,
                                  void *stream)
{
	if (what < DJB200_RADIAL_P22 || what > DJB200_RADIAL_QF) return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown radial query %d", what);
	int family = ndf, res = 0;
	const float *tables = nullptr;
	if (t) {
		if (t->azim_res > 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "tabular_anisotropic is not a radial distribution");
		family = 2; res = t->res; tables = t->tables;
	} else if (ndf != DJB200_NDF_BECKMANN && ndf != DJB200_NDF_GGX) {
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown ndf %d", ndf);
	}
	return map_call(n, {{x, 4}}, {{out, 4}}, mem, stream,
		[=](const std::vector<void *> &i, const std::vector<void *> &o, int64_t cn, cudaStream_t st) {
			return launch_radial_query(family, what, tables, res, (const float *)i[0], cn, (float *)o[0], st);
		});
}

// the remaining public scalar members: one kernel, the argument plumbing of map_call
static djb200_status member_call(int family, int what, const djb200_tabular *t, const double *coef, int n_coef, const float *a,
                                 size_t a_item, const float *b, size_t b_item, const float *c, size_t c_item, size_t out_item, int64_t n,
                                 float *out, int mem, void *stream)
{
	if (n < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "negative count");
	if (n == 0) return DJB200_OK;
	if ((a_item && !a) || (b_item && !b) || (c_item && !c) || !out) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument array");
	std::vector<BulkIn> ins;
	int sa = -1, sb = -1, sc = -1;
	if (a_item) { sa = (int)ins.size(); ins.push_back({a, a_item}); }
	if (b_item) { sb = (int)ins.size(); ins.push_back({b, b_item}); }
	if (c_item) { sc = (int)ins.size(); ins.push_back({c, c_item}); }
	const float *tables = t ? t->tables : nullptr;
	const int er = t ? t->res : 0, ar = t ? t->azim_res : 0, nq = t ? t->n_qf1 : 0;
	return map_call(n, ins, {{out, out_item}}, mem, stream,
		[=](const std::vector<void *> &i, const std::vector<void *> &o, int64_t cn, cudaStream_t st) {
			return launch_member_query(family, what, tables, er, ar, nq, coef, n_coef, sa >= 0 ? (const float *)i[sa] : nullptr,
			                           sb >= 0 ? (const float *)i[sb] : nullptr, sc >= 0 ? (const float *)i[sc] : nullptr, cn,
			                           (float *)o[0], st);
		});
}

djb200_status djb200_quantile_query(int32_t ndf, int32_t what, const float *a, const float *b, const float *c, int64_t n, float *out,
                                    int mem, void *stream)
{
	if (ndf != DJB200_NDF_BECKMANN && ndf != DJB200_NDF_GGX) return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown ndf %d", ndf);
	if (what < DJB200_MEMBER_QF1 || what > DJB200_MEMBER_QF3_RADIAL) return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown quantile member %d", what);
	const bool two = what != DJB200_MEMBER_QF1, three = what == DJB200_MEMBER_QF2_RADIAL;
	return member_call(ndf, what, nullptr, nullptr, 0, a, 4, b, two ? 4 : 0, c, three ? 4 : 0, 4, n, out, mem, stream);
}

djb200_status djb200_tabular_anisotropic_query(const djb200_tabular *t, int32_t what, const float *a, const float *b, int64_t n,
                                               float *out, int mem, void *stream)
{
	if (!t) return fail(DJB200_ERR_INVALID_ARGUMENT, "tabular handle is NULL");
	if (t->azim_res <= 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "not a tabular_anisotropic handle");
	if (what < DJB200_MEMBER_PDF1 || what > DJB200_MEMBER_TQF2) return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown table member %d", what);
	const bool two = what == DJB200_MEMBER_PDF2 || what == DJB200_MEMBER_CDF2 || what == DJB200_MEMBER_TQF2;
	return member_call(2, what, t, nullptr, 0, a, 4, b, two ? 4 : 0, nullptr, 0, 4, n, out, mem, stream);
}

djb200_status djb200_sgd_member(const djb200_sgd_data *m, int32_t what, const float *a, const float *b, const float *c, int64_t n,
                                float *out, int mem, void *stream)
{
	if (!m) return fail(DJB200_ERR_INVALID_ARGUMENT, "sgd coefficients are NULL");
	if (what < DJB200_MEMBER_NDF || what > DJB200_MEMBER_FRESNEL) return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown sgd member %d", what);
	const bool gaf = what == DJB200_MEMBER_GAF;
	const size_t a_item = what == DJB200_MEMBER_FRESNEL ? 4 : 12;
	return member_call(3, what, nullptr, &m->ch[0][0], 33, a, a_item, b, gaf ? 12 : 0, c, gaf ? 12 : 0, 12, n, out, mem, stream);
}

djb200_status djb200_abc_member(const djb200_abc_data *m, int32_t what, const float *a, const float *b, const float *c, int64_t n,
                                float *out, int mem, void *stream)
{
	if (!m) return fail(DJB200_ERR_INVALID_ARGUMENT, "abc coefficients are NULL");
	if (what != DJB200_MEMBER_NDF && what != DJB200_MEMBER_GAF && what != DJB200_MEMBER_FRESNEL)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown abc member %d", what);
	double coef[9];
	for (int k = 0; k < 3; ++k) { coef[k] = m->kD[k]; coef[3 + k] = m->A[k]; }
	coef[6] = m->B; coef[7] = m->C; coef[8] = m->ior;
	const bool gaf = what == DJB200_MEMBER_GAF;
	const size_t a_item = what == DJB200_MEMBER_FRESNEL ? 4 : 12;
	return member_call(4, what, nullptr, coef, 9, a, a_item, b, gaf ? 12 : 0, c, gaf ? 12 : 0, gaf ? 4 : 12, n, out, mem, stream);
}

djb200_status djb200_tabular_anisotropic_sampling_tables(const djb200_tabular *t, float *pdf1, float *cdf1, float *qf1,
                                                         float *pdf2, float *cdf2, float *qf2, int32_t counts[2])
{
	if (!t) return fail(DJB200_ERR_INVALID_ARGUMENT, "tabular handle is NULL");
	if (t->azim_res <= 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "not a tabular_anisotropic handle");
	const int er = t->res, ar = t->azim_res;
	const size_t T = (size_t)er * ar;
	struct { float *dst; size_t off, n; } parts[6] = {
		{pdf1, aniso_off_pdf1(er, ar), (size_t)ar}, {cdf1, aniso_off_cdf1(er, ar), (size_t)ar}, {qf1, aniso_off_qf1(er, ar), (size_t)ar},
		{pdf2, aniso_off_pdf2(er, ar), T}, {cdf2, aniso_off_cdf2(er, ar), T}, {qf2, aniso_off_qf2(er, ar), T}};
	for (auto &p : parts)
		if (p.dst) CU(cudaMemcpy(p.dst, t->tables + p.off, 4 * p.n, cudaMemcpyDeviceToHost));
	if (counts) { counts[0] = t->n_qf1; counts[1] = t->n_qf2; }
	return DJB200_OK;
}

djb200_status djb200_tabular_destroy(djb200_tabular *t)
{
	if (!t) return DJB200_OK;
	cudaFree(t->tables);
	delete t;
	return DJB200_OK;
}

#define DJB200_TAB_NULLCHECK if (!t) return fail(DJB200_ERR_INVALID_ARGUMENT, "tabular handle is NULL")
djb200_status djb200_tabular_eval(const djb200_tabular *t, const djb200_params *params, int64_t n_params, int params_layout,
                                  const float *wi, const float *wo, int64_t n, float *out_rgb, int mem, void *stream)
{
	DJB200_TAB_NULLCHECK;
	return microfacet_call(OP_EVAL, nullptr, params, n_params, params_layout, wi, wo, n, out_rgb, nullptr, nullptr, mem, stream, t);
}
djb200_status djb200_tabular_evalp(const djb200_tabular *t, const djb200_params *params, int64_t n_params, int params_layout,
                                   const float *wi, const float *wo, int64_t n, float *out_rgb, int mem, void *stream)
{
	DJB200_TAB_NULLCHECK;
	return microfacet_call(OP_EVALP, nullptr, params, n_params, params_layout, wi, wo, n, out_rgb, nullptr, nullptr, mem, stream, t);
}
djb200_status djb200_tabular_pdf(const djb200_tabular *t, const djb200_params *params, int64_t n_params, int params_layout,
                                 const float *wi, const float *wo, int64_t n, float *out_pdf, int mem, void *stream)
{
	DJB200_TAB_NULLCHECK;
	return microfacet_call(OP_PDF, nullptr, params, n_params, params_layout, wi, wo, n, out_pdf, nullptr, nullptr, mem, stream, t);
}
djb200_status djb200_tabular_sample(const djb200_tabular *t, const djb200_params *params, int64_t n_params, int params_layout,
                                    const float *u, const float *wo, int64_t n, float *out_wi, int mem, void *stream)
{
	DJB200_TAB_NULLCHECK;
	return microfacet_call(OP_SAMPLE, nullptr, params, n_params, params_layout, u, wo, n, out_wi, nullptr, nullptr, mem, stream, t);
}
djb200_status djb200_tabular_evalp_is(const djb200_tabular *t, const djb200_params *params, int64_t n_params, int params_layout,
                                      const float *u, const float *wo, int64_t n, float *out_weight_rgb, float *out_wi,
                                      float *out_pdf, int mem, void *stream)
{
	DJB200_TAB_NULLCHECK;
	if (!out_weight_rgb && !out_wi && !out_pdf) return fail(DJB200_ERR_INVALID_ARGUMENT, "all outputs are NULL");
	return microfacet_call(OP_EVALP_IS, nullptr, params, n_params, params_layout, u, wo, n, out_weight_rgb, out_wi, out_pdf, mem,
	                       stream, t);
}
#undef DJB200_TAB_NULLCHECK

// ---- MERL ------------------------------------------------------------------------------------------
static const int64_t MERL_N = 90 * 90 * 180;

djb200_status djb200_merl_create(const double *samples, djb200_merl **out)
{
	if (!samples || !out) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	double *tmp = nullptr;
	float4 *cells = nullptr;
	CU(cudaMalloc(&tmp, sizeof(double) * 3 * MERL_N));
	cudaError_t e = cudaMalloc(&cells, sizeof(float4) * MERL_N);
	if (e != cudaSuccess) { cudaFree(tmp); return cuda_fail(e, "cudaMalloc(merl cells)"); }
	e = cudaMemcpy(tmp, samples, sizeof(double) * 3 * MERL_N, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = launch_merl_convert(tmp, cells, 0);
	if (e == cudaSuccess) e = cudaDeviceSynchronize();
	cudaFree(tmp);
	if (e != cudaSuccess) { cudaFree(cells); return cuda_fail(e, "merl upload"); }
	djb200_merl *m = new djb200_merl;
	m->cells = cells;
	cudaGetDevice(&m->device);
	*out = m;
	return DJB200_OK;
}

djb200_status djb200_merl_load(const char *filename, djb200_merl **out)
{
	if (!filename || !out) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	FILE *f = fopen(filename, "rb");
	if (!f) return fail(DJB200_ERR_IO, "djb_error: Failed to open %s", filename); // dj_brdf.h:970
	int32_t dims[3];
	if (fread(dims, 4, 3, f) != 3) { fclose(f); return fail(DJB200_ERR_IO, "djb_error: Failed to read MERL header"); }
	int64_t n = (int64_t)dims[0] * dims[1] * dims[2];
	if (n <= 0) { fclose(f); return fail(DJB200_ERR_IO, "djb_error: Failed to read MERL header"); } // :976
	if (n != MERL_N) { fclose(f); return fail(DJB200_ERR_UNSUPPORTED, "MERL header (%d,%d,%d) is not 90x90x180", dims[0], dims[1], dims[2]); }
	std::vector<double> buf((size_t)3 * n);
	size_t got = fread(buf.data(), sizeof(double), buf.size(), f);
	fclose(f);
	if (got != buf.size()) return fail(DJB200_ERR_IO, "djb_error: Reading %s failed", filename); // :982
	return djb200_merl_create(buf.data(), out);
}

djb200_status djb200_merl_destroy(djb200_merl *m)
{
	if (!m) return DJB200_OK;
	cudaFree(m->cells);
	delete m;
	return DJB200_OK;
}

djb200_status djb200_merl_eval(const djb200_merl *m, const float *wi, const float *wo, int64_t n, float *out_rgb,
                               int mem, void *stream)
{
	if (!m) return fail(DJB200_ERR_INVALID_ARGUMENT, "merl handle is NULL");
	const float4 *cells = m->cells;
	return map_call(n, {{wi, 12}, {wo, 12}}, {{out_rgb, 12}}, mem, stream,
		[cells](const std::vector<void *> &i, const std::vector<void *> &o, int64_t cn, cudaStream_t st) {
			return launch_merl_eval(cells, (const float *)i[0], (const float *)i[1], cn, (float *)o[0], st);
		});
}

djb200_status djb200_merl_index(const float *wi, const float *wo, int64_t n, int32_t *out_index, int mem, void *stream)
{
	return map_call(n, {{wi, 12}, {wo, 12}}, {{out_index, 4}}, mem, stream,
		[](const std::vector<void *> &i, const std::vector<void *> &o, int64_t cn, cudaStream_t st) {
			return launch_merl_index((const float *)i[0], (const float *)i[1], cn, (int32_t *)o[0], st);
		});
}

djb200_status djb200_debug_merl_filter_stats(const float *wi_dev, const float *wo_dev, int64_t n, uint64_t out_stats[5],
                                             void *stream)
{
	if (!wi_dev || !wo_dev || !out_stats || n < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "bad argument");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	unsigned long long *d = nullptr;
	CU(cudaMalloc(&d, 5 * sizeof(unsigned long long)));
	cudaError_t e = cudaMemsetAsync(d, 0, 5 * sizeof(unsigned long long), (cudaStream_t)stream);
	if (e == cudaSuccess) e = launch_merl_filter_stats(wi_dev, wo_dev, n, d, (cudaStream_t)stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(out_stats, d, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
	cudaFree(d);
	if (e != cudaSuccess) return cuda_fail(e, "merl filter stats");
	return DJB200_OK;
}

djb200_status djb200_debug_dmath(int fn, const double *x_dev, const double *y_dev, int64_t n, double *out_dev, void *stream)
{
	if (!x_dev || !out_dev || n < 0 || fn < 0 || fn > 9) return fail(DJB200_ERR_INVALID_ARGUMENT, "bad argument");
	if ((fn == 5 || fn == 8 || fn == 9) && !y_dev) return fail(DJB200_ERR_INVALID_ARGUMENT, "function %d takes two arguments", fn);
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	cudaError_t e = launch_debug_dmath(fn, x_dev, y_dev, n, out_dev, (cudaStream_t)stream);
	if (e != cudaSuccess) return cuda_fail(e, "debug dmath");
	return DJB200_OK;
}

// ---- UTIA ------------------------------------------------------------------------------------------
static const int64_t UTIA_N = 3 * 6 * 48 * 6 * 48;

djb200_status djb200_utia_create(const double *raw_samples, djb200_utia **out)
{
	if (!raw_samples || !out) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	double *tmp = nullptr;
	djb200::UtiaEntry *table = nullptr;
	CU(cudaMalloc(&tmp, sizeof(double) * UTIA_N));
	cudaError_t e = cudaMalloc(&table, sizeof(djb200::UtiaEntry) * (UTIA_N / 3));
	if (e != cudaSuccess) { cudaFree(tmp); return cuda_fail(e, "cudaMalloc(utia table)"); }
	e = cudaMemcpy(tmp, raw_samples, sizeof(double) * UTIA_N, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = launch_utia_convert(tmp, table, 0);
	if (e == cudaSuccess) e = cudaDeviceSynchronize();
	cudaFree(tmp);
	if (e != cudaSuccess) { cudaFree(table); return cuda_fail(e, "utia upload"); }
	djb200_utia *u = new djb200_utia;
	u->table = table;
	cudaGetDevice(&u->device);
	*out = u;
	return DJB200_OK;
}

djb200_status djb200_utia_load(const char *filename, djb200_utia **out)
{
	if (!filename || !out) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL argument");
	FILE *f = fopen(filename, "rb");
	if (!f) return fail(DJB200_ERR_IO, "djb_error: Failed to open %s", filename); // dj_brdf.h:1044
	std::vector<double> buf((size_t)UTIA_N);
	size_t got = fread(buf.data(), sizeof(double), buf.size(), f);
	fclose(f);
	if (got != buf.size()) return fail(DJB200_ERR_IO, "djb_error: Reading %s failed", filename); // :1058
	return djb200_utia_create(buf.data(), out);
}

djb200_status djb200_utia_destroy(djb200_utia *u)
{
	if (!u) return DJB200_OK;
	cudaFree(u->table);
	delete u;
	return DJB200_OK;
}

djb200_status djb200_utia_eval(const djb200_utia *u, const float *wi, const float *wo, int64_t n, float *out_rgb,
                               int mem, void *stream)
{
	if (!u) return fail(DJB200_ERR_INVALID_ARGUMENT, "utia handle is NULL");
	const djb200::UtiaEntry *table = u->table;
	return map_call(n, {{wi, 12}, {wo, 12}}, {{out_rgb, 12}}, mem, stream,
		[table](const std::vector<void *> &i, const std::vector<void *> &o, int64_t cn, cudaStream_t st) {
			return launch_utia_eval(table, (const float *)i[0], (const float *)i[1], cn, (float *)o[0], st);
		});
}

// ---- LEAN ------------------------------------------------------------------------------------------
int32_t djb200_leanmap_mip_levels(int32_t w, int32_t h, int32_t levels) { return w < 1 || h < 1 ? 0 : lean_mip_levels(w, h, levels); }

int64_t djb200_leanmap_mip_texels(int32_t w, int32_t h, int32_t levels)
{
	if (w < 1 || h < 1) return 0;
	int64_t total = 0;
	int a = w, b = h;
	for (int L = 0; L < lean_mip_levels(w, h, levels); ++L) {
		total += (int64_t)a * b;
		a = a > 1 ? a / 2 : 1;
		b = b > 1 ? b / 2 : 1;
	}
	return total;
}

djb200_status djb200_leanmap_to_half_mips(const float *leanmap, int32_t w, int32_t h, int32_t levels, uint16_t *out_rgba16f, int mem,
                                          void *stream)
{
	if (w < 0 || h < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "negative image size");
	if (mem != DJB200_MEM_HOST && mem != DJB200_MEM_DEVICE) return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown memory space %d", mem);
	const int64_t npix = (int64_t)w * h;
	if (npix == 0) return DJB200_OK;
	if (!leanmap || !out_rgba16f) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL image pointer");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	cudaStream_t st = (cudaStream_t)stream;
	const int64_t texels = djb200_leanmap_mip_texels(w, h, levels);
	const size_t qa = (size_t)(w > 1 ? w / 2 : 1) * (h > 1 ? h / 2 : 1), qb = (size_t)(w > 3 ? w / 4 : 1) * (h > 3 ? h / 4 : 1);
	float4 *scratch = nullptr;
	float *d_in = nullptr;
	uint16_t *d_out = nullptr;
	cudaError_t e = cudaMallocAsync((void **)&scratch, sizeof(float4) * (qa + qb), st);
	if (e == cudaSuccess && mem == DJB200_MEM_HOST) {
		e = cudaMallocAsync((void **)&d_in, sizeof(float) * 4 * (size_t)npix, st);
		if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_out, sizeof(uint16_t) * 4 * (size_t)texels, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, leanmap, sizeof(float) * 4 * (size_t)npix, cudaMemcpyHostToDevice, st);
	}
	if (e == cudaSuccess)
		e = launch_leanmap_half_mips(d_in ? d_in : leanmap, w, h, levels, d_out ? d_out : out_rgba16f, scratch, scratch + qa, st);
	if (e == cudaSuccess && mem == DJB200_MEM_HOST)
		e = cudaMemcpyAsync(out_rgba16f, d_out, sizeof(uint16_t) * 4 * (size_t)texels, cudaMemcpyDeviceToHost, st);
	if (scratch) cudaFreeAsync(scratch, st);
	if (d_in) cudaFreeAsync(d_in, st);
	if (d_out) cudaFreeAsync(d_out, st);
	if (e == cudaSuccess && mem == DJB200_MEM_HOST) e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) return cuda_fail(e, "leanmap half / mip conversion");
	return DJB200_OK;
}

djb200_status djb200_nmap_to_leanmap(const uint8_t *nmap, int32_t w, int32_t h, float base_roughness, float bias,
                                     float *lean1, float *lean2, int mem, void *stream)
{
	if (w < 0 || h < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "negative image size");
	if (mem != DJB200_MEM_HOST && mem != DJB200_MEM_DEVICE)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown memory space %d", mem);
	int64_t npix = (int64_t)w * h;
	if (npix == 0) return DJB200_OK;
	if (!nmap || !lean1 || !lean2) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL image pointer");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	if (mem == DJB200_MEM_DEVICE) {
		cudaError_t e = launch_nmap_to_leanmap(nmap, npix, base_roughness, bias, lean1, lean2, (cudaStream_t)stream);
		if (e != cudaSuccess) return cuda_fail(e, "lean kernel launch");
		return DJB200_OK;
	}
	// host images: planar data, so the whole image is staged in row bands that keep every plane
	// contiguous on the device (band-local planar layout), then copied plane by plane
	const int64_t band_rows_budget = ((int64_t)192 << 20) / ((int64_t)w * 35);
	int64_t band = band_rows_budget < 1 ? 1 : band_rows_budget;
	if (band > h) band = h;
	const int SLOTS = band < h ? 2 : 1;
	cudaStream_t st[2] = {nullptr, nullptr};
	uint8_t *d_in[2] = {nullptr, nullptr};
	float *d_o1[2] = {nullptr, nullptr}, *d_o2[2] = {nullptr, nullptr};
	cudaError_t e = cudaSuccess;
	for (int s = 0; s < SLOTS && e == cudaSuccess; ++s) {
		e = cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaMalloc(&d_in[s], (size_t)band * w * 3);
		if (e == cudaSuccess) e = cudaMalloc(&d_o1[s], (size_t)band * w * 16);
		if (e == cudaSuccess) e = cudaMalloc(&d_o2[s], (size_t)band * w * 16);
	}
	int c = 0;
	for (int64_t r0 = 0; r0 < h && e == cudaSuccess; r0 += band, ++c) {
		int s = c % SLOTS;
		int64_t rows = h - r0 < band ? h - r0 : band;
		int64_t bp = rows * w; // band plane size
		for (int ch = 0; ch < 3 && e == cudaSuccess; ++ch)
			e = cudaMemcpyAsync(d_in[s] + ch * bp, nmap + ch * npix + r0 * w, (size_t)bp, cudaMemcpyHostToDevice, st[s]);
		if (e == cudaSuccess) e = launch_nmap_to_leanmap(d_in[s], bp, base_roughness, bias, d_o1[s], d_o2[s], st[s]);
		for (int ch = 0; ch < 4 && e == cudaSuccess; ++ch) {
			e = cudaMemcpyAsync(lean1 + ch * npix + r0 * w, d_o1[s] + ch * bp, (size_t)bp * 4, cudaMemcpyDeviceToHost, st[s]);
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(lean2 + ch * npix + r0 * w, d_o2[s] + ch * bp, (size_t)bp * 4, cudaMemcpyDeviceToHost, st[s]);
		}
	}
	for (int s = 0; s < SLOTS; ++s) {
		if (st[s]) {
			cudaError_t e2 = cudaStreamSynchronize(st[s]);
			if (e == cudaSuccess) e = e2;
		}
		cudaFree(d_in[s]);
		cudaFree(d_o1[s]);
		cudaFree(d_o2[s]);
		if (st[s]) cudaStreamDestroy(st[s]);
	}
	if (e != cudaSuccess) return cuda_fail(e, "lean staging");
	return DJB200_OK;
}

djb200_status djb200_sgd_eval(const djb200_sgd_data *material, const float *wi, const float *wo, int64_t n,
                              float *out_rgb, int mem, void *stream)
{
	if (!material) return fail(DJB200_ERR_INVALID_ARGUMENT, "material is NULL");
	const djb200_sgd_data m = *material;
	return map_call(n, {{wi, 12}, {wo, 12}}, {{out_rgb, 12}}, mem, stream,
		[&m](const std::vector<void *> &i, const std::vector<void *> &o, int64_t cn, cudaStream_t st) {
			return launch_analytic_eval(DJB200_SOURCE_SGD, &m.ch[0][0], 33, (const float *)i[0], (const float *)i[1], cn,
			                            (float *)o[0], st);
		});
}

djb200_status djb200_abc_eval(const djb200_abc_data *material, const float *wi, const float *wo, int64_t n,
                              float *out_rgb, int mem, void *stream)
{
	if (!material) return fail(DJB200_ERR_INVALID_ARGUMENT, "material is NULL");
	const double m[9] = {material->kD[0], material->kD[1], material->kD[2], material->A[0], material->A[1],
	                     material->A[2], material->B, material->C, material->ior};
	return map_call(n, {{wi, 12}, {wo, 12}}, {{out_rgb, 12}}, mem, stream,
		[&m](const std::vector<void *> &i, const std::vector<void *> &o, int64_t cn, cudaStream_t st) {
			return launch_analytic_eval(DJB200_SOURCE_ABC, m, 9, (const float *)i[0], (const float *)i[1], cn,
			                            (float *)o[0], st);
		});
}

djb200_status djb200_dmap_to_nmap(const uint8_t *dmap, int32_t w, int32_t h, float scale, uint8_t *nmap, int mem, void *stream)
{
	if (w < 0 || h < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "negative image size");
	if (mem != DJB200_MEM_HOST && mem != DJB200_MEM_DEVICE)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown memory space %d", mem);
	const size_t npix = (size_t)w * h;
	if (npix == 0) return DJB200_OK;
	if (!dmap || !nmap) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL pointer");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	if (mem == DJB200_MEM_DEVICE) {
		cudaError_t e = launch_dmap2nmap(dmap, w, h, scale, nmap, (cudaStream_t)stream);
		return e == cudaSuccess ? DJB200_OK : cuda_fail(e, "dmap2nmap launch");
	}
	// host: the stencil needs whole rows and their neighbours; a map is 1 + 3 bytes per texel, staged in one piece
	uint8_t *d_in = nullptr, *d_out = nullptr;
	cudaError_t e = cudaMalloc(&d_in, npix);
	if (e == cudaSuccess) e = cudaMalloc(&d_out, 3 * npix);
	if (e == cudaSuccess) e = cudaMemcpy(d_in, dmap, npix, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = launch_dmap2nmap(d_in, w, h, scale, d_out, 0);
	if (e == cudaSuccess) e = cudaMemcpy(nmap, d_out, 3 * npix, cudaMemcpyDeviceToHost);
	cudaFree(d_in);
	cudaFree(d_out);
	return e == cudaSuccess ? DJB200_OK : cuda_fail(e, "dmap2nmap staging");
}

djb200_status djb200_lean_shading_params(const djb200_lean_shading *cfg, const float *alpha, const float *E, int64_t n,
                                         djb200_params *out, int mem, void *stream)
{
	return lean_shading_call(-1, nullptr, cfg, alpha, E, nullptr, nullptr, n, (float *)out, nullptr, nullptr, mem, stream);
}
djb200_status djb200_lean_shading_evalp(const djb200_microfacet *mf, const djb200_lean_shading *cfg, const float *alpha,
                                        const float *E, const float *wi, const float *wo, int64_t n, float *out_rgb, int mem,
                                        void *stream)
{
	return lean_shading_call(OP_EVALP, mf, cfg, alpha, E, wi, wo, n, out_rgb, nullptr, nullptr, mem, stream);
}
djb200_status djb200_lean_shading_pdf(const djb200_microfacet *mf, const djb200_lean_shading *cfg, const float *alpha,
                                      const float *E, const float *wi, const float *wo, int64_t n, float *out_pdf, int mem,
                                      void *stream)
{
	return lean_shading_call(OP_PDF, mf, cfg, alpha, E, wi, wo, n, out_pdf, nullptr, nullptr, mem, stream);
}
djb200_status djb200_lean_shading_evalp_is(const djb200_microfacet *mf, const djb200_lean_shading *cfg, const float *alpha,
                                           const float *E, const float *u, const float *wo, int64_t n,
                                           float *out_weight_rgb, float *out_wi, float *out_pdf, int mem, void *stream)
{
	return lean_shading_call(OP_EVALP_IS, mf, cfg, alpha, E, u, wo, n, out_weight_rgb, out_wi, out_pdf, mem, stream);
}

djb200_status djb200_lrep_to_params(const float *E, int64_t n, djb200_params *out, int mem, void *stream)
{
	return map_call(n, {{E, 20}}, {{out, 48}}, mem, stream,
		[](const std::vector<void *> &i, const std::vector<void *> &o, int64_t cn, cudaStream_t st) {
			return launch_lrep_to_params((const float *)i[0], cn, o[0], st);
		});
}

djb200_status djb200_params_to_lrep(const djb200_params *params, int64_t n, float *E, int mem, void *stream)
{
	return map_call(n, {{params, 48}}, {{E, 20}}, mem, stream,
		[](const std::vector<void *> &i, const std::vector<void *> &o, int64_t cn, cudaStream_t st) {
			return launch_params_to_lrep(i[0], cn, (float *)o[0], st);
		});
}

djb200_status djb200_leanmap_to_params(const float *lean1, const float *lean2, int32_t w, int32_t h, float bias,
                                       djb200_params *out, int mem, void *stream)
{
	if (w < 0 || h < 0) return fail(DJB200_ERR_INVALID_ARGUMENT, "negative image size");
	if (mem != DJB200_MEM_HOST && mem != DJB200_MEM_DEVICE)
		return fail(DJB200_ERR_INVALID_ARGUMENT, "unknown memory space %d", mem);
	int64_t npix = (int64_t)w * h;
	if (npix == 0) return DJB200_OK;
	if (!lean1 || !lean2 || !out) return fail(DJB200_ERR_INVALID_ARGUMENT, "NULL pointer");
	djb200_status rs = require_device();
	if (rs != DJB200_OK) return rs;
	if (mem == DJB200_MEM_DEVICE) {
		cudaError_t e = launch_leanmap_to_params(lean1, lean2, npix, bias, out, (cudaStream_t)stream);
		if (e != cudaSuccess) return cuda_fail(e, "leanmap_to_params launch");
		return DJB200_OK;
	}
	// host: whole maps staged at once (5 used planes of 4 B + 48 B out per texel)
	float *d1 = nullptr, *d2 = nullptr;
	void *dp = nullptr;
	cudaError_t e = cudaMalloc(&d1, (size_t)npix * 16);
	if (e == cudaSuccess) e = cudaMalloc(&d2, (size_t)npix * 16);
	if (e == cudaSuccess) e = cudaMalloc(&dp, (size_t)npix * 48);
	if (e == cudaSuccess) e = cudaMemcpy(d1, lean1, (size_t)npix * 16, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemcpy(d2, lean2, (size_t)npix * 16, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = launch_leanmap_to_params(d1, d2, npix, bias, dp, 0);
	if (e == cudaSuccess) e = cudaMemcpy(out, dp, (size_t)npix * 48, cudaMemcpyDeviceToHost);
	cudaFree(d1);
	cudaFree(d2);
	cudaFree(dp);
	if (e != cudaSuccess) return cuda_fail(e, "leanmap_to_params staging");
	return DJB200_OK;
}

} // extern "C"
