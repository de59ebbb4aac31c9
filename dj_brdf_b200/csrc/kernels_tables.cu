// kernels_tables.cu -- MERL / UTIA table lookups, the Rusinkiewicz frame, and the LEAN-map
// streaming kernels.
//
// MERL device layout: one float4 (r, g, b, 0) per cell, already multiplied by the channel scales.
// The reference computes float(double_sample * SCALE) at every lookup (dj_brdf.h:1012-1014), which
// is a pure function of the cell, so doing it once at upload time is bit-identical and turns the
// three 8-byte gathers from three planes 11.7 MB apart into one 16-byte gather.  The 23.3 MB table
// stays resident in the 126 MB L2; HBM traffic per lookup is the 24 B of directions + 12 B result.
#include "djb_device.cuh"
#include "djb_internal.h"

namespace djb200 {

constexpr int TB = 256;

static inline int grid_for(int64_t n, int per_thread = 1)
{
	int64_t want = (n + (int64_t)TB * per_thread - 1) / ((int64_t)TB * per_thread);
	int64_t cap = (int64_t)sm_count() * 8;
	return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

DJB_DEV V3 ld3(const float *p, long long k) { return mk(p[3 * k], p[3 * k + 1], p[3 * k + 2]); }
DJB_DEV void st3t(float *p, long long k, V3 v)
{
	p[3 * k] = v.x;
	p[3 * k + 1] = v.y;
	p[3 * k + 2] = v.z;
}

// ---- Rusinkiewicz frame ------------------------------------------------------------------------
__global__ void __launch_bounds__(TB) io_to_hd_kernel(const float *wi, const float *wo, long long n, float *h, float *d)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
		V3 hv, dv;
		float th;
		io_to_hd(ld3(wi, k), ld3(wo, k), hv, dv, th);
		st3t(h, k, hv);
		st3t(d, k, dv);
	}
}

__global__ void __launch_bounds__(TB) hd_to_io_kernel(const float *h, const float *d, long long n, float *wi, float *wo)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
		V3 iv, ov;
		hd_to_io(ld3(h, k), ld3(d, k), iv, ov);
		st3t(wi, k, iv);
		st3t(wo, k, ov);
	}
}

cudaError_t launch_io_to_hd(const float *wi, const float *wo, int64_t n, float *h, float *d, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	io_to_hd_kernel<<<grid_for(n), TB, 0, st>>>(wi, wo, n, h, d);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

cudaError_t launch_hd_to_io(const float *h, const float *d, int64_t n, float *wi, float *wo, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	hd_to_io_kernel<<<grid_for(n), TB, 0, st>>>(h, d, n, wi, wo);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

// ---- UTIA, dj_brdf.h:1039-1177 ---------------------------------------------------------------------
// utia::normalize (clamp at 0, scale by the float constant 1/140 in double) followed by the
// (float_t) cast utia::eval applies to every fetched sample (:1144, 1162-1177)
// Device layout (djb_device.cuh: UtiaEntry): the file's three channel planes interleaved, each cell together with its phi_v
// neighbour -- the 16 taps of a query are 8 256-bit loads (2.65 MB table, L2 resident) instead of 48 scattered 4-byte ones
__global__ void __launch_bounds__(TB) utia_convert_kernel(const double *raw, UtiaEntry *table)
{
	constexpr int PLANE = UT_CELLS / 3;
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= PLANE) return;
	const float k = 1.f / 140.f;
	const int c1 = (c % UT_NPV) == UT_NPV - 1 ? c - (UT_NPV - 1) : c + 1; // phi_v + 1, wrapped (dj_brdf.h:1118-1121)
	float ch[6];
	for (int s = 0; s < 6; ++s) {
		double v = raw[(s % 3) * PLANE + (s < 3 ? c : c1)];
		v = 0.0 > v ? 0.0 : v;
		ch[s] = (float)(v * (double)k);
	}
	UtiaEntry e;
	e.lo = make_float4(ch[0], ch[1], ch[2], 0.0f);
	e.hi = make_float4(ch[3], ch[4], ch[5], 0.0f);
	table[c] = e;
}

#ifndef DJB200_UTIA_MINB
#define DJB200_UTIA_MINB 6
#endif
__global__ void __launch_bounds__(TB, DJB200_UTIA_MINB) utia_eval_kernel(const UtiaEntry *__restrict__ tab, const float *__restrict__ wi,
                                                       const float *__restrict__ wo, long long n, float *__restrict__ out)
{
	__shared__ __align__(16) double s_dm[DMT_COUNT];
	dm_load_tables(s_dm);
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
		st3t(out, k, utia_eval1(tab, ld3(wi, k), ld3(wo, k), s_dm));
}

cudaError_t launch_utia_convert(const double *raw_dev, UtiaEntry *table_dev, cudaStream_t st)
{
	utia_convert_kernel<<<(UT_CELLS / 3 + TB - 1) / TB, TB, 0, st>>>(raw_dev, table_dev);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

cudaError_t launch_utia_eval(const UtiaEntry *table, const float *wi, const float *wo, int64_t n, float *out,
                             cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	utia_eval_kernel<<<grid_for(n), TB, 0, st>>>(table, wi, wo, n, out);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

// ---- LEAN maps -------------------------------------------------------------------------------------
// nmap2leanmap (utils/nmap2leanmap.cpp:18-54, _biased.cpp:23-63): pure float, no libm.  Planar
// uint8 in, two planar RGBA float images out: 3 B read + 32 B written per texel, nothing reused,
// so the kernel is a pure HBM stream.  Each thread converts 4 consecutive texels: one 32-bit load
// per input plane and one 128-bit store per output plane.
DJB_DEV void lean_texel(float r, float g, float b, float half_br2, float bias, float &e1, float &e2, float &e3,
                        float &e4, float &e5)
{
	float t1 = (r / 255.f) * 2.0f - 1.0f;
	float t2 = (g / 255.f) * 2.0f - 1.0f;
	float t3 = b / 255.f;
	float sx = -t1 / t3, sy = -t2 / t3;
	e1 = bias != 0.0f ? sx + bias : sx;
	e2 = bias != 0.0f ? sy + bias : sy;
	e3 = sx * sx + half_br2;
	e4 = sy * sy + half_br2;
	e5 = bias != 0.0f ? sx * sy + bias * bias : sx * sy;
}

__global__ void __launch_bounds__(TB) lean_kernel_vec4(const uchar4 *__restrict__ pr, const uchar4 *__restrict__ pg,
                                                       const uchar4 *__restrict__ pb, long long nquads, long long plane,
                                                       float base_roughness, float bias, float *__restrict__ l1,
                                                       float *__restrict__ l2)
{
	const float half_br2 = 0.5f * base_roughness * base_roughness;
	const float4 ones = make_float4(1.f, 1.f, 1.f, 1.f);
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += stride) {
		uchar4 r = __ldcs(pr + q), g = __ldcs(pg + q), b = __ldcs(pb + q);
		float4 E1, E2, E3, E4, E5;
		lean_texel(r.x, g.x, b.x, half_br2, bias, E1.x, E2.x, E3.x, E4.x, E5.x);
		lean_texel(r.y, g.y, b.y, half_br2, bias, E1.y, E2.y, E3.y, E4.y, E5.y);
		lean_texel(r.z, g.z, b.z, half_br2, bias, E1.z, E2.z, E3.z, E4.z, E5.z);
		lean_texel(r.w, g.w, b.w, half_br2, bias, E1.w, E2.w, E3.w, E4.w, E5.w);
		float4 *o1 = reinterpret_cast<float4 *>(l1) + q;
		float4 *o2 = reinterpret_cast<float4 *>(l2) + q;
		const long long pq = plane / 4;
		__stcs(o1, E1);
		__stcs(o1 + pq, E2);
		__stcs(o1 + 2 * pq, ones);
		__stcs(o1 + 3 * pq, ones);
		__stcs(o2, E3);
		__stcs(o2 + pq, E4);
		__stcs(o2 + 2 * pq, E5);
		__stcs(o2 + 3 * pq, ones);
	}
}

// scalar variant for planes whose size is not a multiple of 4 texels (alignment of planes 1..3)
__global__ void __launch_bounds__(TB) lean_kernel_scalar(const uint8_t *__restrict__ nmap, long long plane,
                                                         float base_roughness, float bias, float *__restrict__ l1,
                                                         float *__restrict__ l2)
{
	const float half_br2 = 0.5f * base_roughness * base_roughness;
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < plane; px += stride) {
		float e1, e2, e3, e4, e5;
		lean_texel(nmap[px], nmap[plane + px], nmap[2 * plane + px], half_br2, bias, e1, e2, e3, e4, e5);
		l1[px] = e1; l1[plane + px] = e2; l1[2 * plane + px] = 1.f; l1[3 * plane + px] = 1.f;
		l2[px] = e3; l2[plane + px] = e4; l2[2 * plane + px] = e5; l2[3 * plane + px] = 1.f;
	}
}

cudaError_t launch_nmap_to_leanmap(const uint8_t *nmap, int64_t npix, float base_roughness, float bias,
                                   float *lean1, float *lean2, cudaStream_t st)
{
	if (npix <= 0) return cudaSuccess;
	bool aligned = (npix % 4 == 0) && ((uintptr_t)nmap % 4 == 0) && ((uintptr_t)lean1 % 16 == 0) &&
	               ((uintptr_t)lean2 % 16 == 0);
	if (aligned) {
		const uchar4 *p = reinterpret_cast<const uchar4 *>(nmap);
		int64_t nq = npix / 4;
		lean_kernel_vec4<<<grid_for(nq), TB, 0, st>>>(p, p + nq, p + 2 * nq, nq, npix, base_roughness, bias, lean1, lean2);
	} else {
		lean_kernel_scalar<<<grid_for(npix), TB, 0, st>>>(nmap, npix, base_roughness, bias, lean1, lean2);
	}
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

// ---- LEAN maps -> half-float RGBA texture with a mip pyramid ---------------------------------------
// What happens to the two maps after nmap2leanmap in the reference's asset chain: save_exr converts the four float planes to
// interleaved half RGBA (utils/CImg.h:44940-44947, called at utils/nmap2leanmap.cpp:128-131: (half)(float), round to nearest
// even), and the renderer's texture unit filters the moments LINEARLY over a mip pyramid -- which is the whole point of the LEAN
// representation (mitsuba/dj_beckmannconductor.cpp:295-314 fetches the filtered texels).  Level 0 is that conversion; level L is
// the 2 x 2 box filter of level L - 1 carried in float32 (((a + b) + (c + d)) * 0.25, edge texels repeated for odd sizes), each
// level rounded to half once.  One kernel per level: HBM bound, 16 B in + 8 B out per level-0 texel.
#include <cuda_fp16.h>
__global__ void __launch_bounds__(TB) lean_half_level0_kernel(const float *__restrict__ planar, int64_t npix, uint2 *__restrict__ out)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < npix; k += stride) {
		const __half2 rg = __floats2half2_rn(planar[k], planar[npix + k]);
		const __half2 ba = __floats2half2_rn(planar[2 * npix + k], planar[3 * npix + k]);
		out[k] = make_uint2(*reinterpret_cast<const unsigned *>(&rg), *reinterpret_cast<const unsigned *>(&ba));
	}
}
// src: float4 RGBA texels of the previous level (level 1 reads the planar level-0 image instead: src_planar != NULL)
__global__ void __launch_bounds__(TB) lean_mip_level_kernel(const float4 *__restrict__ src, const float *__restrict__ src_planar, int sw,
                                                            int sh, int dw, int dh, float4 *__restrict__ dst, uint2 *__restrict__ out)
{
	const int64_t n = (int64_t)dw * dh, splane = (int64_t)sw * sh;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
		const int y = (int)(k / dw), x = (int)(k - (int64_t)y * dw);
		const int x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1), y0 = min(2 * y, sh - 1), y1 = min(2 * y + 1, sh - 1);
		const int64_t i00 = (int64_t)y0 * sw + x0, i10 = (int64_t)y0 * sw + x1, i01 = (int64_t)y1 * sw + x0, i11 = (int64_t)y1 * sw + x1;
		float4 a, b, c, d;
		if (src_planar) {
			auto ld = [&](int64_t i) { return make_float4(src_planar[i], src_planar[splane + i], src_planar[2 * splane + i], src_planar[3 * splane + i]); };
			a = ld(i00); b = ld(i10); c = ld(i01); d = ld(i11);
		} else {
			a = src[i00]; b = src[i10]; c = src[i01]; d = src[i11];
		}
		float4 r;
		r.x = ((a.x + b.x) + (c.x + d.x)) * 0.25f;
		r.y = ((a.y + b.y) + (c.y + d.y)) * 0.25f;
		r.z = ((a.z + b.z) + (c.z + d.z)) * 0.25f;
		r.w = ((a.w + b.w) + (c.w + d.w)) * 0.25f;
		dst[k] = r;
		const __half2 rg = __floats2half2_rn(r.x, r.y), ba = __floats2half2_rn(r.z, r.w);
		out[k] = make_uint2(*reinterpret_cast<const unsigned *>(&rg), *reinterpret_cast<const unsigned *>(&ba));
	}
}

int lean_mip_levels(int w, int h, int levels)
{
	int full = 1;
	for (int a = w, b = h; a > 1 || b > 1; a = a > 1 ? a / 2 : 1, b = b > 1 ? b / 2 : 1) ++full;
	return levels <= 0 || levels > full ? full : levels;
}

// out: all levels back to back, level L = [h_L][w_L] RGBA half texels; scratch: 2 float4 buffers of (w/2)(h/2) and (w/4)(h/4) texels
cudaError_t launch_leanmap_half_mips(const float *planar, int w, int h, int levels, uint16_t *out, float4 *scratch_a, float4 *scratch_b,
                                     cudaStream_t st)
{
	const int64_t npix = (int64_t)w * h;
	if (npix <= 0) return cudaSuccess;
	uint2 *o = reinterpret_cast<uint2 *>(out);
	lean_half_level0_kernel<<<grid_for(npix), TB, 0, st>>>(planar, npix, o);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	o += npix;
	int sw = w, sh = h;
	const float4 *src = nullptr;
	float4 *bufs[2] = {scratch_a, scratch_b};
	for (int L = 1; L < lean_mip_levels(w, h, levels); ++L) {
		const int dw = sw > 1 ? sw / 2 : 1, dh = sh > 1 ? sh / 2 : 1;
		float4 *dst = bufs[(L - 1) & 1];
		lean_mip_level_kernel<<<grid_for((int64_t)dw * dh), TB, 0, st>>>(src, L == 1 ? planar : nullptr, sw, sh, dw, dh, dst, o);
		g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
		o += (int64_t)dw * dh;
		src = dst;
		sw = dw;
		sh = dh;
	}
	return cudaGetLastError();
}

// ---- dmap2nmap, utils/dmap2nmap.cpp:13-44 --------------------------------------------------------------------------
// Central differences of an 8-bit displacement map (CImg atXY(): clamped at the borders) -> unit normal packed as planar
// 8-bit RGB.  1 B in (+ 4 neighbours, which the row / L2 locality makes free) and 3 B out per texel: HBM bound.
// That utility is compiled with <math.h> in scope, so its sqrt(float) is sqrtf; 1.0 / x and 0.5 * n + 0.5 are double
// expressions rounded once to float -- one IEEE reciprocal and one fused multiply-add give the same floats.
// (uint8_t)v for 0 <= v < 2^23 without the quarter-rate F2I: adding 2^23 with round-toward-zero leaves floor(v) in the low
// mantissa bits (the float grid at 2^23 is the integers)
DJB_DEV uint8_t trunc_u8(float v) { return (uint8_t)(__float_as_uint(__fadd_rz(v, 8388608.0f)) & 0xffu); }

// one texel from its four neighbours' heights (already divided by 255)
DJB_DEV void dmap_normal(float z_l, float z_r, float z_b, float z_t, float kx, float ky, uint8_t &r, uint8_t &g, uint8_t &b)
{
	const float sx = kx * (z_r - z_l), sy = ky * (z_t - z_b);
	const float nrm_sqr = 1.f + sx * sx + sy * sy;
	const float nrm_inv = __frcp_rn(__fsqrt_rn(nrm_sqr));
	const float nx = -sx * nrm_inv, ny = -sy * nrm_inv;
	r = trunc_u8(__fmaf_rn(0.5f, nx, 0.5f) * 255.f);
	g = trunc_u8(__fmaf_rn(0.5f, ny, 0.5f) * 255.f);
	b = trunc_u8(nrm_inv * 255.f);
}

// Row-walking variant for w % 4 == 0 and 4-byte aligned rows: a thread owns four consecutive texels of a column quad and
// walks down the rows (no index divisions); three 32-bit loads + two bytes bring in all 14 neighbours, and the
// reference's `(float)px / 255.f` -- 256 possible operands -- comes from a shared-memory table filled with that division.
__global__ void __launch_bounds__(TB) dmap2nmap_rows_kernel(const uint8_t *__restrict__ d, int w, int h, float scale,
                                                            uint8_t *__restrict__ out)
{
	__shared__ float s_z[256];
	for (int t = threadIdx.x; t < 256; t += blockDim.x) s_z[t] = (float)t / 255.f;
	__syncthreads();
	const int iq = blockIdx.x * blockDim.x + threadIdx.x;
	if (iq >= w / 4) return;
	const int i0 = 4 * iq;
	const float kx = (float)w * 0.5f * scale, ky = (float)h * 0.5f * scale;
	const size_t plane = (size_t)w * h;
	for (int j = blockIdx.y; j < h; j += gridDim.y) {
		const uint8_t *row = d + (size_t)j * w;
		const uchar4 c = *reinterpret_cast<const uchar4 *>(row + i0);
		const uchar4 t = *reinterpret_cast<const uchar4 *>(d + (size_t)(j > 0 ? j - 1 : 0) * w + i0);
		const uchar4 b = *reinterpret_cast<const uchar4 *>(d + (size_t)(j + 1 < h ? j + 1 : h - 1) * w + i0);
		const float zl = s_z[row[i0 > 0 ? i0 - 1 : 0]], zr = s_z[row[i0 + 4 < w ? i0 + 4 : w - 1]];
		const float z0 = s_z[c.x], z1 = s_z[c.y], z2 = s_z[c.z], z3 = s_z[c.w];
		uchar4 R, G, B;
		dmap_normal(zl, z1, s_z[b.x], s_z[t.x], kx, ky, R.x, G.x, B.x);
		dmap_normal(z0, z2, s_z[b.y], s_z[t.y], kx, ky, R.y, G.y, B.y);
		dmap_normal(z1, z3, s_z[b.z], s_z[t.z], kx, ky, R.z, G.z, B.z);
		dmap_normal(z2, zr, s_z[b.w], s_z[t.w], kx, ky, R.w, G.w, B.w);
		const size_t o = (size_t)j * w + i0;
		__stcs(reinterpret_cast<uchar4 *>(out + o), R);
		__stcs(reinterpret_cast<uchar4 *>(out + plane + o), G);
		__stcs(reinterpret_cast<uchar4 *>(out + 2 * plane + o), B);
	}
}

// any size / alignment: one texel per thread
__global__ void __launch_bounds__(TB) dmap2nmap_scalar_kernel(const uint8_t *__restrict__ d, int w, int h, float scale,
                                                              uint8_t *__restrict__ out)
{
	const float kx = (float)w * 0.5f * scale, ky = (float)h * 0.5f * scale;
	const long long plane = (long long)w * h, stride = (long long)gridDim.x * blockDim.x;
	for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < plane; px += stride) {
		const int i = (int)(px % w), j = (int)(px / w);
		const int il = i > 0 ? i - 1 : 0, ir = i + 1 < w ? i + 1 : w - 1;
		const int jt = j > 0 ? j - 1 : 0, jb = j + 1 < h ? j + 1 : h - 1;
		const size_t row = (size_t)j * w;
		uint8_t r, g, b;
		dmap_normal((float)d[il + row] / 255.f, (float)d[ir + row] / 255.f, (float)d[i + (size_t)jb * w] / 255.f,
		            (float)d[i + (size_t)jt * w] / 255.f, kx, ky, r, g, b);
		out[px] = r;
		out[plane + px] = g;
		out[2 * plane + px] = b;
	}
}

cudaError_t launch_dmap2nmap(const uint8_t *dmap, int w, int h, float scale, uint8_t *nmap, cudaStream_t st)
{
	const int64_t plane = (int64_t)w * h;
	if (plane <= 0) return cudaSuccess;
	const bool vec = (w % 4 == 0) && ((uintptr_t)nmap % 4 == 0) && ((uintptr_t)dmap % 4 == 0);
	if (vec) {
		const int wq = w / 4, gx = (wq + TB - 1) / TB;
		int gy = (sm_count() * 8 + gx - 1) / gx; // ~8 resident CTAs per SM in all
		if (gy > h) gy = h;
		if (gy > 65535) gy = 65535;
		dmap2nmap_rows_kernel<<<dim3(gx, gy), TB, 0, st>>>(dmap, w, h, scale, nmap);
	} else {
		dmap2nmap_scalar_kernel<<<grid_for(plane), TB, 0, st>>>(dmap, w, h, scale, nmap);
	}
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

// ---- LEAN algebra per texel --------------------------------------------------------------------------
DJB_DEV void store_params(void *out, long long k, const Params &p)
{
	float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<Params *>(out) + k);
	o[0] = make_float4(p.nx, p.ny, p.nz, p.a1);
	o[1] = make_float4(p.a2, p.phi_a, p.ax, p.ay);
	o[2] = make_float4(p.rho, p.srho, p.tx, p.ty);
}

__global__ void __launch_bounds__(TB) lrep_to_params_kernel(const float *__restrict__ E, long long n, void *out)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
		Params p;
		lrep_to_params(E[5 * k], E[5 * k + 1], E[5 * k + 2], E[5 * k + 3], E[5 * k + 4], p);
		store_params(out, k, p);
	}
}

// beckmann::params_to_lrep, dj_brdf.h:1965-1974
__global__ void __launch_bounds__(TB) params_to_lrep_kernel(const Params *__restrict__ P, long long n, float *__restrict__ E)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
		Params p = P[k];
		E[5 * k] = p.tx;
		E[5 * k + 1] = p.ty;
		E[5 * k + 2] = 0.5f * p.ax * p.ax + p.tx * p.tx;
		E[5 * k + 3] = 0.5f * p.ay * p.ay + p.ty * p.ty;
		E[5 * k + 4] = 0.5f * p.rho * p.ax * p.ay + p.tx * p.ty;
	}
}

// check_lean_maps (utils/nmap2leanmap.cpp:57-76) as a producer of per-texel params; with a bias the
// plugin removes it first (mitsuba/dj_beckmannconductor.cpp:295-314)
__global__ void __launch_bounds__(TB) leanmap_to_params_kernel(const float *__restrict__ l1, const float *__restrict__ l2,
                                                               long long plane, float bias, void *out)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < plane; px += stride) {
		float e1 = l1[px], e2 = l1[plane + px], e3 = l2[px], e4 = l2[plane + px], e5 = l2[2 * plane + px];
		if (bias != 0.0f) {
			e1 -= bias;
			e2 -= bias;
			e5 -= bias * bias;
		}
		Params p;
		lrep_to_params(e1, e2, e3, e4, e5, p);
		store_params(out, px, p);
	}
}

cudaError_t launch_lrep_to_params(const float *E, int64_t n, void *out_params, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	lrep_to_params_kernel<<<grid_for(n), TB, 0, st>>>(E, n, out_params);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

cudaError_t launch_params_to_lrep(const void *params, int64_t n, float *E, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	params_to_lrep_kernel<<<grid_for(n), TB, 0, st>>>(reinterpret_cast<const Params *>(params), n, E);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

cudaError_t launch_leanmap_to_params(const float *lean1, const float *lean2, int64_t npix, float bias,
                                     void *out_params, cudaStream_t st)
{
	if (npix <= 0) return cudaSuccess;
	leanmap_to_params_kernel<<<grid_for(npix), TB, 0, st>>>(lean1, lean2, npix, bias, out_params);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

} // namespace djb200
