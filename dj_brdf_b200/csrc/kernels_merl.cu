// kernels_merl.cu -- MERL table lookup.
//
// merl::eval (dj_brdf.h:987-1024) turns (i, o) into three bin indices through 3 acos, 3 atan2 and 2 sincos
// evaluated in double; done literally that is ~970 warp instructions per lookup (17 % of the HBM roofline).
// The bin indices, however, are integers: they only depend on WHICH SIDE of a bin boundary an angle falls.
// So the lookup is split into
//
//   * a filtered fast path, pure FP32, table-free:
//       - h = normalize(i + o) is formed with the reference's own float operations, bit for bit (its one double
//         sub-expression, 1/sqrt, is the correctly rounded float reciprocal square root);
//       - d = R_y(-theta_h) R_z(-phi_h) i is built from cos / sin obtained algebraically from h (no angles);
//         it differs from the reference's d by < 4e-7 per component (measured; budget ETA = 2e-6);
//       - theta_h, theta_d, phi_d come from one polynomial acos (error < 2.4e-7 rad measured; budget 1e-6) and
//         each bin index is certified only if the value is farther from both ends of its bin than the
//         propagated error budget (which includes the reference's own float roundings of those angles);
//   * the exact path (merl_cell in djb_device.cuh: the reference's arithmetic in double) for the lookups the
//     filter rejects (~1.5e-3 of uniformly random pairs; all pairs within 0.36 degrees of the theta_h / theta_d
//     poles).  Rejected lookups are queued per warp and resolved 32 at a time, so a slow lane never stalls a
//     warp of fast ones.
//
// A certified index always equals the exact one: djb200_debug_merl_filter_stats counts violations (tests run it
// over 1e8 uniformly random pairs plus near-specular / near-retro-reflective stress sets: 0 violations), so the
// result is the reference's, cell for cell.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "djb_device.cuh"
#include "djb_internal.h"

namespace djb200 {

constexpr int TBM = 256;
constexpr int MERL_CELLS_N = 90 * 90 * 180;

// |d_fast - d_reference| per component: measured max 3.6e-7 over 1.2e9 pairs of all kinds (x5 margin)
constexpr float ETA_Z = 2.0e-6f;
constexpr float ETA_CROSS = 3.0e-6f;   // same, for the (d.x, d.y) direction (both components move)
constexpr float POLE_GUARD = 0.99998f; // fast path only below this |z| (the reference's pole guard is 0.99999)

DJB_DEV V3 ldv(const float *p, long long k) { return mk(p[3 * k], p[3 * k + 1], p[3 * k + 2]); }
DJB_DEV void stv(float *p, long long k, V3 v)
{
	p[3 * k] = v.x;
	p[3 * k + 1] = v.y;
	p[3 * k + 2] = v.z;
}

// bare MUFU ops (no denormal fix-up code): only used where the result is a proposal or inside the error bounds
DJB_DEV float rsq(float x)
{
	float y;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}
DJB_DEV float sqrt_fast(float x)
{
	float y;
	asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}

// acos(x) given ax = |x| and root = sqrt(1 - |x|): Abramowitz-Stegun 4.4.46 (|eps| <= 2e-8) evaluated in float.
// Measured against double acos over every float in [-1, 1]: see ACOS_ERR (djb200_debug_merl_filter_stats).
DJB_DEV float acos_from_root(float x, float ax, float root)
{
	float p = fmaf(ax, -0.0012624911f, 0.0066700901f);
	p = fmaf(ax, p, -0.0170881256f);
	p = fmaf(ax, p, 0.0308918810f);
	p = fmaf(ax, p, -0.0501743046f);
	p = fmaf(ax, p, 0.0889789874f);
	p = fmaf(ax, p, -0.2145988016f);
	p = fmaf(ax, p, 1.5707963050f);
	float r = root * p;
	return x < 0.0f ? 3.14159265f - r : r;
}

// error budget of the angle-domain bin tests (radians unless noted)
constexpr float ACOS_ERR = 2.0e-6f;   // |acos_from_root - acos|, measured max 4.7e-7 over 1.2e9 arguments (x4)
constexpr float REF_ROUND = 4.0e-7f;  // the reference rounds theta_d / phi_d to float (ulp(pi)/2 = 1.2e-7) and adds pi in float
constexpr float RAD2DEG = 57.29578f;

// The filtered fast path.  Returns the cell index, or -1 when the filter cannot certify it.
// d_out receives the fast d (for the calibration kernel).
//
// Each bin index is floor(f(angle)) of an angle the reference obtains in double and rounds to float.  Here the
// angle comes from acos_from_root (error <= ACOS_ERR), f(angle) is formed in float, and the index is certified only
// if f(angle) is farther from both ends of its bin than the propagated error bound; otherwise -1.
DJB_DEV int merl_cell_filtered(V3 i, V3 o, V3 &d_out)
{
	// h with the reference's own float operations (normalize, dj_brdf.h:630-637): bit-identical to the reference
	V3 s = i + o;
	float m = dot(s, s);
	V3 h = scale(__frsqrt_rn(m), s); // == float(1.0 / sqrt(double(m))) up to double-rounding ties (~1e-8)
	// ---- theta_h: index = floor(sqrt(theta_h / (pi/2) * 90 * 90)), dj_brdf.h:906-920
	float ahz = fabsf(h.z);
	float omh = 1.0f - ahz; // exact for |h.z| >= 0.5
	float th = acos_from_root(h.z, ahz, sqrt_fast(omh));
	float qh = sqrt_fast(th * 5156.62f);
	int kh = min((int)qh, 89);
	// d(q)/d(theta) = 5156.62 / (2 q); the reference's own float roundings move q by < 2e-5
	float mh = (0.5f * 5156.62f * (ACOS_ERR + REF_ROUND)) * __frcp_rn(fmaxf(qh, 1.0f)) + 1.0e-4f;
	bool ok = (ahz <= POLE_GUARD) && (qh - (float)kh > mh) && ((kh == 89) || ((float)(kh + 1) - qh > mh));
	// ---- d: rotations with cos / sin taken from h itself.
	// phi_h: cos / sin from (h.x, h.y).  theta_h: the reference takes cos / sin of float(acos(h.z)), i.e. of an
	// angle that is a function of the FLOAT h.z alone, so sin(theta_h) must come from h.z too -- (1 - z)(1 + z)
	// with 1 - z exact -- not from |h.xy|, which carries the part of theta_h that rounding h.z lost
	float rinv = rsq(fmaf(h.x, h.x, h.y * h.y));
	float c = h.x * rinv, sn = h.y * rinv;
	float sxy = sqrt_fast(omh * (1.0f + ahz));
	float xr = fmaf(c, i.x, sn * i.y);
	float yr = fmaf(c, i.y, -sn * i.x);
	V3 d = mk(fmaf(h.z, xr, -sxy * i.z), yr, fmaf(h.z, i.z, sxy * xr));
	d = scale(rsq(fmaf(d.x, d.x, fmaf(d.y, d.y, d.z * d.z))), d);
	d_out = d;
	// ---- theta_d: index = floor(theta_d / (pi/2) * 90), dj_brdf.h:926-936; |d.z error| <= ETA_Z
	float adz = fabsf(d.z);
	float omd = 1.0f - adz;
	float td = acos_from_root(d.z, adz, sqrt_fast(omd)) * RAD2DEG;
	int kd = min((int)td, 89);
	float md = RAD2DEG * (ACOS_ERR + REF_ROUND) + (RAD2DEG * ETA_Z) * rsq(omd * (1.0f + adz)); // ETA_Z / sin(theta_d)
	ok = ok && (adz <= POLE_GUARD) && (td - (float)kd > md) && ((kd == 89) || ((float)(kd + 1) - td > md));
	// ---- phi_d: fold phi < 0 onto phi + pi (dj_brdf.h:945-946); index = floor(phi_d / pi * 180)
	float dx = d.x, dy = d.y;
	if (dy < 0.0f) { dx = -dx; dy = -dy; }
	float rd = rsq(fmaf(dx, dx, dy * dy));
	float x = dx * rd, y = dy * rd, ax = fabsf(x);
	// 1 - |x| = y^2 / (1 + |x|): no cancellation, so the angle keeps its relative accuracy near 0 and pi
	float pd = acos_from_root(x, ax, y * rsq(1.0f + ax)) * RAD2DEG;
	int kp = min((int)pd, 179);
	float mp = RAD2DEG * (ACOS_ERR + REF_ROUND) + (RAD2DEG * ETA_CROSS) * rd; // ETA / |d.xy|
	ok = ok && (pd - (float)kp > mp) && ((float)(kp + 1) - pd > mp);
	return ok ? kp + kd * 180 + kh * 16200 : -1;
}

DJB_DEV int merl_cell_any(V3 i, V3 o)
{
	V3 d;
	int c = merl_cell_filtered(i, o, d);
	if (c < 0) c = merl_cell(i, o); // exact path (djb_device.cuh): the reference's arithmetic in double
	return c;
}

// one lookup per thread: tails, unaligned callers
__global__ void __launch_bounds__(TBM) merl_eval_fast_kernel(const float4 *__restrict__ cells, const float *__restrict__ wi,
                                                             const float *__restrict__ wo, long long n, float *__restrict__ out)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
		float4 v = __ldg(cells + merl_cell_any(ldv(wi, k), ldv(wo, k))); // negative ("below horizon") texels were zeroed at upload
		stv(out, k, mk(v.x, v.y, v.z));
	}
}

// Four consecutive lookups per thread: 2 x 3 LDG.128 in, 3 STG.128 out, fully coalesced.  Lookups the filter
// rejects are not resolved in place (one slow lane would stall its warp for ~1000 instructions): they are pushed
// on a per-warp queue and resolved 32 at a time by the whole warp.
constexpr int WQ_CAP = 160; // per-warp queue entries: flushed in groups of 32 once >= 32; a quad adds <= 128 per round

DJB_DEV void merl_resolve_exact(const float4 *__restrict__ cells, const float *__restrict__ wi, const float *__restrict__ wo,
                                long long k, float *__restrict__ out)
{
	int c = merl_cell(ldv(wi, k), ldv(wo, k));
	float4 v = __ldg(cells + c);
	stv(out, k, mk(v.x, v.y, v.z));
}

template <int MINB, int DBG = 0>
__global__ void __launch_bounds__(TBM, MINB) merl_eval_quad_kernel(const float4 *__restrict__ cells, const float *__restrict__ wi, const float *__restrict__ wo,
                                                             long long nquads, float *__restrict__ out)
{
	__shared__ long long s_queue[TBM / 32][WQ_CAP];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	long long *wq = s_queue[warp];
	int wq_count = 0; // warp-uniform
	const float4 *wi4 = reinterpret_cast<const float4 *>(wi), *wo4 = reinterpret_cast<const float4 *>(wo);
	float4 *out4 = reinterpret_cast<float4 *>(out);
	const long long stride = (long long)gridDim.x * blockDim.x;
	// warp-uniform trip count so that the __ballot_sync below always sees the whole warp
	const long long first = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
	for (long long base = first; base < nquads; base += stride) {
		const long long q = base + lane;
		const bool live = q < nquads;
		int c[4] = {0, 0, 0, 0};
		if (live) {
			// (tried: ld.global.L1::no_allocate for these streams, and the largest L1 carve-out -- both slower, profiles/r02_merl_ceiling.md)
			const float4 a0 = __ldcs(wi4 + 3 * q), a1 = __ldcs(wi4 + 3 * q + 1), a2 = __ldcs(wi4 + 3 * q + 2);
			const float4 b0 = __ldcs(wo4 + 3 * q), b1 = __ldcs(wo4 + 3 * q + 1), b2 = __ldcs(wo4 + 3 * q + 2);
			V3 d;
			c[0] = merl_cell_filtered(mk(a0.x, a0.y, a0.z), mk(b0.x, b0.y, b0.z), d);
			c[1] = merl_cell_filtered(mk(a0.w, a1.x, a1.y), mk(b0.w, b1.x, b1.y), d);
			c[2] = merl_cell_filtered(mk(a1.z, a1.w, a2.x), mk(b1.z, b1.w, b2.x), d);
			c[3] = merl_cell_filtered(mk(a2.y, a2.z, a2.w), mk(b2.y, b2.z, b2.w), d);
			if (DBG == 1) { // experiment: coalesced gather, full compute
				for (int j = 0; j < 4; ++j) c[j] = c[j] >= 0 ? (int)((4 * q + j) & 0xFFFF) : -1;
			}
			if (DBG == 2) { // experiment: random gather, no compute
				unsigned hsh[4] = {__float_as_uint(a0.x), __float_as_uint(a0.w), __float_as_uint(a1.z), __float_as_uint(a2.y)};
				unsigned hs2[4] = {__float_as_uint(b0.x), __float_as_uint(b0.w), __float_as_uint(b1.z), __float_as_uint(b2.y)};
				for (int j = 0; j < 4; ++j) c[j] = (int)(((hsh[j] * 2654435761u) ^ (hs2[j] * 40503u)) % 1458000u);
			}
		}
		const bool any_rej = live && ((c[0] | c[1] | c[2] | c[3]) < 0);
		float4 v[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
			if (live && c[j] >= 0) {
				v[j] = __ldg(cells + c[j]);
			}
		}
		if (live && !any_rej) {
			__stcs(out4 + 3 * q, make_float4(v[0].x, v[0].y, v[0].z, v[1].x));
			__stcs(out4 + 3 * q + 1, make_float4(v[1].y, v[1].z, v[2].x, v[2].y));
			__stcs(out4 + 3 * q + 2, make_float4(v[2].z, v[3].x, v[3].y, v[3].z));
		}
		if (__any_sync(0xffffffffu, any_rej)) {
			if (any_rej) { // certified results of this quad go out one by one; the rejected ones are queued
#pragma unroll
				for (int j = 0; j < 4; ++j)
					if (c[j] >= 0) stv(out, 4 * q + j, mk(v[j].x, v[j].y, v[j].z));
			}
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const bool rej = live && c[j] < 0;
				const unsigned m = __ballot_sync(0xffffffffu, rej);
				if (rej) wq[wq_count + __popc(m & ((1u << lane) - 1u))] = 4 * q + j;
				wq_count += __popc(m);
			}
			__syncwarp();
			while (wq_count >= 32) {
				wq_count -= 32;
				merl_resolve_exact(cells, wi, wo, wq[wq_count + lane], out);
			}
			__syncwarp();
		}
	}
	if (lane < wq_count) merl_resolve_exact(cells, wi, wo, wq[lane], out);
}

__global__ void __launch_bounds__(TBM) merl_index_fast_kernel(const float *__restrict__ wi, const float *__restrict__ wo,
                                                              long long n, int32_t *__restrict__ out)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
		out[k] = merl_cell_any(ldv(wi, k), ldv(wo, k));
}

// calibration / property check of the filter:
//   stats[0] = lookups the filter rejected, stats[1] = lookups it certified with a cell different from the exact one
//   (must be 0), stats[2] = max |d_fast - d_exact| over certified lookups (float bits), stats[3] = half vectors not
//   bit-identical to the reference's, stats[4] = max |acos_from_root - acos| over a sweep of [-1, 1] (float bits)
__global__ void __launch_bounds__(TBM) merl_filter_stats_kernel(const float *__restrict__ wi, const float *__restrict__ wo,
                                                                long long n, unsigned long long *stats)
{
	unsigned long long rejected = 0, wrong = 0, hdiff = 0;
	float maxerr = 0.0f, acerr = 0.0f;
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
		V3 i = ldv(wi, k), o = ldv(wo, k), df, h, de;
		float th;
		int cf = merl_cell_filtered(i, o, df);
		int ce = merl_cell(i, o);
		io_to_hd(i, o, h, de, th);
		if (cf < 0) ++rejected;
		else {
			if (cf != ce) ++wrong;
			maxerr = fmaxf(maxerr, fmaxf(fabsf(df.x - de.x), fmaxf(fabsf(df.y - de.y), fabsf(df.z - de.z))));
		}
		V3 s2 = i + o;
		V3 hf = scale(__frsqrt_rn(dot(s2, s2)), s2);
		if (__float_as_uint(hf.x) != __float_as_uint(h.x) || __float_as_uint(hf.y) != __float_as_uint(h.y) ||
		    __float_as_uint(hf.z) != __float_as_uint(h.z)) ++hdiff;
		// acos sweep: the components of the inputs and of d are as good a sample of [-1, 1] as any
		float xs[4] = {i.z, o.x, de.z, de.x};
		for (int q = 0; q < 4; ++q) {
			float x = xs[q], ax = fabsf(x);
			if (ax <= 1.0f) acerr = fmaxf(acerr, fabsf((float)((double)acos_from_root(x, ax, sqrt_fast(1.0f - ax)) - acos((double)x))));
		}
	}
	atomicAdd(stats + 0, rejected);
	atomicAdd(stats + 1, wrong);
	atomicMax(reinterpret_cast<unsigned int *>(stats + 2), __float_as_uint(maxerr));
	atomicAdd(stats + 3, hdiff);
	atomicMax(reinterpret_cast<unsigned int *>(stats + 4), __float_as_uint(acerr));
}

// file planes (R, G, B doubles) -> scaled float4 cells; MERL_*_SCALE, dj_brdf.h:897-899
__global__ void __launch_bounds__(TBM) merl_convert_kernel(const double *samples, float4 *cells)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= MERL_CELLS_N) return;
	float r = (float)(samples[c] * (1.00 / 1500.0));
	float g = (float)(samples[c + MERL_CELLS_N] * (1.15 / 1500.0));
	float b = (float)(samples[c + 2 * MERL_CELLS_N] * (1.66 / 1500.0));
	// "below horizon" cells hold negative samples and merl::eval returns vec3(0) for them (dj_brdf.h:1016-1021):
	// a pure function of the cell, so it is folded into the table
	if (r < 0.0f || g < 0.0f || b < 0.0f) r = g = b = 0.0f;
	cells[c] = make_float4(r, g, b, 0.f);
}

template <class K>
static inline int grid_of(K kernel, int64_t n)
{
	int resident = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, TBM, 0);
	if (resident < 1) resident = 1;
	int64_t want = (n + TBM - 1) / TBM, cap = (int64_t)sm_count() * resident;
	return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

cudaError_t launch_merl_convert(const double *samples_dev, float4 *cells_dev, cudaStream_t st)
{
	merl_convert_kernel<<<(MERL_CELLS_N + TBM - 1) / TBM, TBM, 0, st>>>(samples_dev, cells_dev);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

// Experiment switch (DJB200_MERL_PERSIST=1): pin the 23 MB table in L2 with an access-policy window on the launch stream, the
// streams already being marked evict-first.  Measured (profiles/r02_merl_ceiling.md): 104 -> 73 G lookups/s -- the table already
// hits in L2; what saturates is the SM -> L2 request interface, not L2 capacity, and the window only adds set-aside pressure.
static void merl_persist_window(const float4 *cells, cudaStream_t st, bool on)
{
	static bool limit_set = false;
	if (on && !limit_set) {
		cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)32 << 20);
		limit_set = true;
	}
	cudaStreamAttrValue attr;
	memset(&attr, 0, sizeof attr);
	attr.accessPolicyWindow.base_ptr = const_cast<float4 *>(cells);
	attr.accessPolicyWindow.num_bytes = on ? sizeof(float4) * (size_t)MERL_CELLS_N : 0;
	attr.accessPolicyWindow.hitRatio = 1.0f;
	attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
	attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
	cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
}

cudaError_t launch_merl_eval(const float4 *cells, const float *wi, const float *wo, int64_t n, float *out, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	static const bool persist = getenv("DJB200_MERL_PERSIST") != nullptr;
	if (persist) merl_persist_window(cells, st, true);
	// quads need 16-byte aligned arrays; the tail (n % 4) and unaligned callers take the one-lookup-per-thread kernel
	const bool aligned = ((uintptr_t)wi % 16 == 0) && ((uintptr_t)wo % 16 == 0) && ((uintptr_t)out % 16 == 0);
	const int64_t nq = aligned ? n / 4 : 0, done = nq * 4;
	if (nq > 0) {
		int64_t want = (nq + TBM - 1) / TBM;
		// DJB200_MERL_PROBE=1|2 selects the two ceiling probes (coalesced gather / no index math) used in DESIGN.md
		// section 5; results of a probe run are NOT lookups.
		static const char *var = getenv("DJB200_MERL_PROBE");
		const int v = var ? atoi(var) : 0;
#define LAUNCHQ(D)                                                                                              \
	do {                                                                                                        \
		static const int cap = grid_of(merl_eval_quad_kernel<5, D>, (int64_t)1 << 40);                             \
		merl_eval_quad_kernel<5, D><<<(int)(want < cap ? want : cap), TBM, 0, st>>>(cells, wi, wo, nq, out);       \
	} while (0)
		if (v == 1) LAUNCHQ(1); else if (v == 2) LAUNCHQ(2); else LAUNCHQ(0);
#undef LAUNCHQ
		g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	}
	if (done < n) {
		merl_eval_fast_kernel<<<grid_of(merl_eval_fast_kernel, n - done), TBM, 0, st>>>(cells, wi + 3 * done, wo + 3 * done,
		                                                                                n - done, out + 3 * done);
		g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	}
	return cudaGetLastError();
}

cudaError_t launch_merl_index(const float *wi, const float *wo, int64_t n, int32_t *out, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	merl_index_fast_kernel<<<grid_of(merl_index_fast_kernel, n), TBM, 0, st>>>(wi, wo, n, out);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

cudaError_t launch_merl_filter_stats(const float *wi, const float *wo, int64_t n, unsigned long long *stats_dev, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	merl_filter_stats_kernel<<<grid_of(merl_filter_stats_kernel, n), TBM, 0, st>>>(wi, wo, n, stats_dev);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

} // namespace djb200
