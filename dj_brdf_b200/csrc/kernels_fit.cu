// kernels_fit.cu -- the "power iteration" fits on the device.
//
// Isotropic fit (djb::tabular, dj_brdf.h:2215-2236): ONE CTA PER MATERIAL, the whole pipeline of the
// reference's constructor in one launch -- kernel matrix, power iterations, normalisation, Smith
// projected-area table, Fresnel table, CDF / quantile tables, Beckmann and GGX roughness -- with
// every intermediate in shared memory (the matrix, cnt^2 doubles, lives in an L2-resident global
// workspace).  A batch of materials fills the GPU: 128 materials = 128 CTAs on 148 SMs.
//
// Parity: every sum keeps the reference's order (a float sum is not associative), so the tables are
// bit-comparable with the reference.  Work that does not depend on the summation index is hoisted
// (the 361 cos(phi) of the kernel-matrix integral, the 180 x 90 NDF grid of the sigma integral) and the
// independent terms of each quadrature are computed in parallel before one thread adds them in order.
//
// Anisotropic fit (djb::tabular_anisotropic, dj_brdf.h:2238-2273): the kernel matrix has (w h)^2
// entries (8010^2 doubles = 513 MB at 90 x 90), so each stage is a grid-wide kernel over the rows of
// ONE material; a row range [row0, row1) makes the same kernels the per-GPU shard of a fit that spans
// GPUs (the iterate is all-gathered between iterations by the host layer, capi_fit.cu).
#include "djb_fit.cuh"
#include "djb_internal.h"

namespace djb200 {

constexpr int FIT_THREADS = 256;
constexpr int SIG_NTHETA = 90, SIG_NPHI = 180; // compute_sigma quadrature, dj_brdf.h:2350-2351
constexpr int NORM_NTHETA = 128;               // normalize_p22 / fit_*_parameters quadrature
constexpr int MAX_PHI_STEPS = 400;             // the phi loop of compute_p22_smith runs 361 times

struct IsoFitArgs {
	const FitSourceDev *sources; // [n_materials]
	int res, shadow, iterations;
	double *K;        // workspace: n_materials x cnt x cnt, K[b * cnt + a] = km(b, a)   (column of row a contiguous in a)
	float4 *fres_ws;  // workspace: n_materials x cnt x (cnt + 2) Fresnel ratios (rx, ry, rz, valid)
	float *grid_ws;   // workspace of the split mode: n_materials x SIG_NPHI x SIG_NTHETA NDF grid values
	int hist;         // the iterates of the power iteration are kept in shared memory (iso_smem_plan)
	int phase;        // 0: the whole fit in one launch, one CTA per material.  1..6: ONE phase of the split mode (below)
	// outputs, n_materials x ...
	float *p22, *sigma, *cdf, *qf, *fresnel, *alpha, *residuals;
};

// dynamic shared memory layout (floats unless noted), res = cnt + 1:
//   double v0[cnt], v1[cnt]
//   double cphi_d[SIG_NPHI], cth_d[SIG_NTHETA]
//   float p22[res], sigma[res], cdf[res]
//   float row_theta[cnt], row_tan[cnt], row_cos[cnt], row_kji[cnt]
//   float cosphi[MAX_PHI_STEPS], terms[2 * NORM_NTHETA], sth[SIG_NTHETA], ui[SIG_NTHETA]
//   float scan[8 * cnt]
//   float grid[SIG_NPHI * SIG_NTHETA]      the NDF over the sigma quadrature nodes
//   int   fr_count[cnt], fr_offset[cnt + 1]  trip counts of the Fresnel loops
// then: double gsp_d[SIG_NPHI], gcp_d[SIG_NPHI] (sincos of the azimuths: the vec3(theta, phi) of the NDF grid), double ck_d[cnt],
// float sk_f[cnt] (per-view-angle constants of compute_sigma), and, when it fits (iso_smem_plan), double Ks[cnt * cnt], the
// kernel matrix (63 KB at res 90).  The kernel lays the doubles out first, then the floats / ints: every pointer is plain
// arithmetic on the shared base.
static size_t iso_smem_base(int res)
{
	size_t cnt = res - 1;
	return sizeof(double) * (2 * cnt + SIG_NPHI + SIG_NTHETA + 2 * SIG_NPHI + cnt) +
	       sizeof(float) * (3 * (size_t)res + 4 * cnt + MAX_PHI_STEPS + 2 * NORM_NTHETA + 2 * SIG_NTHETA + 8 * cnt + 8 +
	                        SIG_NPHI * SIG_NTHETA + cnt) +
	       sizeof(int) * (2 * cnt + 2);
}
struct IsoSmemPlan { size_t bytes; int k_in_smem, hist; };
static IsoSmemPlan iso_smem_plan(int res, int iterations)
{
	const size_t cnt = res - 1, limit = 227 * 1024, kmat = sizeof(double) * cnt * cnt;
	IsoSmemPlan p;
	p.bytes = iso_smem_base(res) + 16;
	p.k_in_smem = p.bytes + kmat <= limit;
	if (p.k_in_smem) p.bytes += kmat;
	// every iterate of the power iteration (the residual diagnostics are then computed after the loop, in parallel), at the end
	const size_t hist = sizeof(double) * cnt * ((size_t)(iterations > 0 ? iterations : 0) + 1);
	p.hist = p.k_in_smem && iterations > 0 && p.bytes + hist <= limit;
	if (p.hist) p.bytes += hist;
	return p;
}
static size_t iso_smem_bytes(int res) { return iso_smem_plan(res, 0).bytes; }

// SM clock at the phase boundaries of material 0's CTA (djb200_debug_fit_phase_clocks): rows | matrix | iterations | normalise |
// NDF grid | sigma | Fresnel ratios | Fresnel sums + cdf | qf + parameters
__device__ long long g_fit_phase_clock[10];
#define FIT_PHASE(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_fit_phase_clock[k] = clock64(); } while (0)

// SPLIT MODE (small batches: fewer materials than SMs / 3).  One CTA per material leaves most of the device idle when a call
// fits a handful of materials -- the reference's own example fits one at a time -- and the fit is 0.7 ms of dependent work inside that
// CTA.  The same kernel then runs once per phase with gridDim.y CTAs per material (blockIdx.y = part), passing the state
// through L2-resident global arrays (the outputs themselves + K / grid_ws / fres_ws):
//   1 MATRIX   rows in every CTA, the kernel-matrix entries split over the parts                -> K
//   2 ITER     one CTA: K -> shared memory, power iterations, normalize_p22                     -> p22
//   3 GRID     the 16 200 NDF grid values split over the parts                                   -> grid_ws
//   4 SIGMA    the 89 ordered sums split over the parts; inside a CTA seven producer warps compute the terms of the next
//              azimuth slab while lanes of warp 0 add the current slab's terms in the reference's order        -> sigma
//   5 FRESNEL  the 5 456 ratio evaluations split over the parts                                  -> fres_ws
//   6 FINISH   one CTA: Fresnel sums, cdf, qf, roughness parameters                              -> fresnel, cdf, qf, alpha
// Same operations on the same operands in the same order as the single launch: bit-identical (tests/test_gpu_fit.py).
enum { PH_ALL = 0, PH_MATRIX = 1, PH_ITER = 2, PH_GRID = 3, PH_SIGMA = 4, PH_FRESNEL = 5, PH_FINISH = 6 };
constexpr int SLAB_PITCH = 92; // floats per chain and slab: rows 16-byte aligned

template <bool K_SMEM>
__global__ void __launch_bounds__(FIT_THREADS) fit_tabular_kernel(IsoFitArgs A)
{
	extern __shared__ double smem_d[];
	const int res = A.res, cnt = res - 1, tid = threadIdx.x, nt = blockDim.x, mat = blockIdx.x;
	const int ph = A.phase, part = blockIdx.y, nparts = gridDim.y;
	const bool split = ph != PH_ALL;
	auto on = [&](int p) { return ph == PH_ALL || ph == p; };
	double *v0 = smem_d, *v1 = v0 + cnt, *cphi_d = v1 + cnt, *cth_d = cphi_d + SIG_NPHI;
	double *gsp_d = cth_d + SIG_NTHETA, *gcp_d = gsp_d + SIG_NPHI, *ck_d = gcp_d + SIG_NPHI;
	double *Ks = ck_d + cnt; // cnt * cnt doubles when K_SMEM
	float *s_p22 = reinterpret_cast<float *>(Ks + (K_SMEM ? cnt * cnt : 0));
	float *s_sigma = s_p22 + res, *s_cdf = s_sigma + res;
	float *row_theta = s_cdf + res, *row_tan = row_theta + cnt, *row_cos = row_tan + cnt, *row_kji = row_cos + cnt;
	float *cosphi = row_kji + cnt, *terms = cosphi + MAX_PHI_STEPS, *sth = terms + 2 * NORM_NTHETA, *ui = sth + SIG_NTHETA;
	float *scan = ui + SIG_NTHETA;
	float *grid = scan + 8 * cnt;
	int *fr_count = reinterpret_cast<int *>(grid + SIG_NPHI * SIG_NTHETA), *fr_offset = fr_count + cnt;
	float *sk_f = reinterpret_cast<float *>(fr_offset + cnt + 1);
	// (iterations + 1) x cnt doubles when A.hist: starts at the next 8-byte boundary after sk_f
	double *hist = reinterpret_cast<double *>(sk_f + cnt + ((3 * res + 8 * cnt + 4 * cnt + MAX_PHI_STEPS + 2 * NORM_NTHETA + 2 * SIG_NTHETA +
	                                                       SIG_NPHI * SIG_NTHETA + 2 * cnt + 1 + cnt) & 1));
	__shared__ int s_nphi;
	__shared__ double s_red[3][FIT_THREADS / 32];
	__shared__ float s_scale;

	const FitSourceDev src = A.sources[mat];
	const bool shadow = A.shadow != 0;
	double *Kg = A.K + (size_t)mat * cnt * cnt;
	double *K = Ks;
	if (!K_SMEM) K = Kg;
	double *Kw = split ? Kg : K; // the matrix phase of the split mode writes to global memory, whatever K_SMEM is
	float *grid_g = A.grid_ws + (size_t)mat * SIG_NPHI * SIG_NTHETA;
	float *o_p22 = A.p22 + (size_t)mat * res, *o_sigma = A.sigma + (size_t)mat * res;
	float4 *fres_ws = A.fres_ws + (size_t)mat * cnt * (cnt + 2);
	const double sqrt_half_pi = sqrt(DJB_PI * 0.5);
	const Params sp = standard_params();
	FIT_PHASE(0);
	TabIso tab;
	tab.p22 = s_p22; tab.sigma = s_sigma; tab.n = res; tab.T = g_dm_table_dev;

	// ---- compute_p22_smith, dj_brdf.h:2482-2522 ------------------------------------------------
	const float dphi_h = (float)(DJB_PI / 180.0);
	if (on(PH_MATRIX)) {
	if (tid == 0) { // for (phi = 0; phi < 2 pi; phi += dphi) with a float counter: 361 steps
		int n = 0;
		for (float phi = 0.0f; (double)phi < 2.0 * DJB_PI && n < MAX_PHI_STEPS; phi += dphi_h) cosphi[n++] = phi;
		s_nphi = n;
	}
	__syncthreads();
	const int nphi = s_nphi;
	for (int k = tid; k < nphi; k += nt) cosphi[k] = (float)cos((double)cosphi[k]);
	{
		const float dtheta = (float)(sqrt_half_pi / (double)(float)cnt);
		for (int i = tid; i < cnt; i += nt) {
			float t = (float)i / (float)cnt;
			float theta = (float)((double)t * sqrt_half_pi);
			float theta_o = theta * theta;
			float cos_o = (float)cos((double)theta_o), tan_o = (float)tan((double)theta_o);
			V3 dir = spherical(theta_o, 0.0f);
			float fr_i = intensity(source_eval(src, dir, dir));
			row_theta[i] = theta;
			row_tan[i] = tan_o;
			row_cos[i] = cos_o;
			row_kji[i] = (float)(((double)dtheta * pow((double)cos_o, 6.0)) * (8.0 * (double)fr_i));
		}
	}
	__syncthreads();
	FIT_PHASE(1);
	for (int e = part * nt + tid; e < cnt * cnt; e += nt * nparts) {
		int j = e / cnt, i = e - j * cnt; // consecutive threads walk i: coalesced writes of K[j * cnt + i]
		float tan_product = row_tan[j] * row_tan[i];
		float nint;
		if (tan_product <= 1.0f) {
			nint = (float)nphi; // every term is max(1, tan_product * cos) == 1: the float sum of nphi ones is exact
		} else {
			nint = 0.0f;
			for (int k = 0; k < nphi; ++k) nint += fmax_ref(1.0f, tan_product * cosphi[k]);
		}
		nint *= dphi_h;
		float entry = row_theta[j] * row_kji[i] * nint * row_tan[j] / (row_cos[j] * row_cos[j]);
		Kw[(size_t)j * cnt + i] = (double)entry; // out[i] = sum_j K(i, j) v[j]; stored transposed for coalesced reads
	}
	} // PH_MATRIX
	if (on(PH_ITER)) {
	if (split && K_SMEM) for (int e = tid; e < cnt * cnt; e += nt) Ks[e] = Kg[e];
	for (int a = tid; a < cnt; a += nt) v0[a] = 1.0;
	__syncthreads();
	FIT_PHASE(2);
	// matrix::eigenvector, dj_brdf.h:2467-2480: un-normalised power iterations, sums in index order
	double *vin = v0, *vout = v1;
	if (A.hist) {
		// every iterate stays in shared memory: the loop is the 89 dependent sums and one barrier per iteration; the diagnostics
		// (not in the reference, never fed back: ||v1/|v1| - v0/|v0||| = sqrt(2 - 2 cos)) follow, one warp per iteration
		for (int a = tid; a < cnt; a += nt) hist[a] = 1.0;
		__syncthreads();
		for (int it = 0; it < A.iterations; ++it) {
			const double *hin = hist + (size_t)it * cnt;
			double *hout = hist + (size_t)(it + 1) * cnt;
			for (int a = tid; a < cnt; a += nt) {
				double acc = 0.0;
#pragma unroll 8
				for (int b = 0; b < cnt; ++b) acc += K[(size_t)b * cnt + a] * hin[b]; // loads and products ahead of the ordered adds
				hout[a] = acc;
			}
			__syncthreads();
		}
		if (A.residuals) {
			const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
			for (int it = warp; it < A.iterations; it += nw) {
				const double *hin = hist + (size_t)it * cnt, *hout = hin + cnt;
				double n0 = 0.0, n1 = 0.0, d01 = 0.0;
				for (int a = lane; a < cnt; a += 32) { n0 += hin[a] * hin[a]; n1 += hout[a] * hout[a]; d01 += hin[a] * hout[a]; }
				for (int o = 16; o > 0; o >>= 1) {
					n0 += __shfl_xor_sync(0xffffffffu, n0, o);
					n1 += __shfl_xor_sync(0xffffffffu, n1, o);
					d01 += __shfl_xor_sync(0xffffffffu, d01, o);
				}
				if (lane == 0) {
					const double d = 2.0 - 2.0 * d01 / (sqrt(n0) * sqrt(n1));
					A.residuals[(size_t)mat * A.iterations + it] = (float)sqrt(d > 0.0 ? d : 0.0);
				}
			}
		}
		vin = hist + (size_t)A.iterations * cnt;
	} else
	for (int it = 0; it < A.iterations; ++it) {
		for (int a = tid; a < cnt; a += nt) {
			double acc = 0.0;
			for (int b = 0; b < cnt; ++b) acc += K[(size_t)b * cnt + a] * vin[b];
			vout[a] = acc;
		}
		__syncthreads();
		if (A.residuals) { // diagnostic only (not in the reference, never fed back): ||v1/|v1| - v0/|v0||| = sqrt(2 - 2 cos)
			double n0 = 0.0, n1 = 0.0, d01 = 0.0;
			for (int a = tid; a < cnt; a += nt) { n0 += vin[a] * vin[a]; n1 += vout[a] * vout[a]; d01 += vin[a] * vout[a]; }
			for (int o = 16; o > 0; o >>= 1) {
				n0 += __shfl_xor_sync(0xffffffffu, n0, o);
				n1 += __shfl_xor_sync(0xffffffffu, n1, o);
				d01 += __shfl_xor_sync(0xffffffffu, d01, o);
			}
			if ((tid & 31) == 0) { s_red[0][tid >> 5] = n0; s_red[1][tid >> 5] = n1; s_red[2][tid >> 5] = d01; }
			__syncthreads();
			if (tid == 0) {
				n0 = n1 = d01 = 0.0;
				for (int w = 0; w < nt / 32; ++w) { n0 += s_red[0][w]; n1 += s_red[1][w]; d01 += s_red[2][w]; }
				double d = 2.0 - 2.0 * d01 / (sqrt(n0) * sqrt(n1));
				A.residuals[(size_t)mat * A.iterations + it] = (float)sqrt(d > 0.0 ? d : 0.0);
			}
		}
		double *tmp = vin; vin = vout; vout = tmp;
		__syncthreads();
	}
	for (int a = tid; a < cnt; a += nt) s_p22[a] = (float)(1e-2 * vin[a]);
	if (tid == 0) s_p22[cnt] = 0.0f;
	__syncthreads();
	FIT_PHASE(3);

	// ---- normalize_p22, dj_brdf.h:2277-2304 ------------------------------------------------------
	for (int i = tid; i < NORM_NTHETA; i += nt) {
		float u = (float)i / (float)NORM_NTHETA;
		float theta_h = (float)((double)(u * u) * DJB_PI * 0.5);
		float r_h = (float)tan((double)theta_h), c_h = (float)cos((double)theta_h);
		terms[i] = (u * tab.p22_radial(r_h * r_h) * r_h) / (c_h * c_h);
	}
	__syncthreads();
	if (tid == 0) {
		float nint = 0.0f;
		for (int i = 0; i < NORM_NTHETA; ++i) nint += terms[i];
		const float dphi = (float)(2.0 * DJB_PI), dtheta = (float)(DJB_PI / (double)(float)NORM_NTHETA);
		nint *= dtheta * dphi;
		s_scale = (float)(1.0 / (double)nint);
	}
	__syncthreads();
	for (int a = tid; a < res; a += nt) s_p22[a] *= s_scale;
	__syncthreads();
	if (split) for (int a = tid; a < res; a += nt) o_p22[a] = s_p22[a];
	FIT_PHASE(4);
	} // PH_ITER
	if (split && ph >= PH_GRID) { // the later phases of the split mode start from the tables the earlier ones left in global memory
		for (int a = tid; a < res; a += nt) { s_p22[a] = o_p22[a]; s_sigma[a] = ph >= PH_FRESNEL ? o_sigma[a] : 0.0f; }
		__syncthreads();
	}

	// ---- compute_sigma, dj_brdf.h:2348-2386 -------------------------------------------------------
	// vec3(theta_h, phi_h) of the NDF grid (dj_brdf.h:589-595) is built from sincos of each angle: tabulated per angle (90 + 180
	// calls instead of 2 x 16 200); gst_f / gct_f borrow `terms`, which is idle between normalize_p22 and the parameter fits
	float *gst_f = terms, *gct_f = terms + SIG_NTHETA;
	if (on(PH_GRID) || on(PH_SIGMA)) {
	for (int j = tid; j < SIG_NPHI; j += nt) {
		float u_j = (float)j / (float)SIG_NPHI;
		const double phi_h = (double)(float)((double)u_j * 2.0 * DJB_PI);
		cphi_d[j] = cos(phi_h);
		sincos(phi_h, &gsp_d[j], &gcp_d[j]);
	}
	for (int j = tid; j < SIG_NTHETA; j += nt) {
		float u_i = (float)j / (float)SIG_NTHETA;
		float theta_h = (float)((double)(u_i * u_i) * DJB_PI * 0.5);
		cth_d[j] = cos((double)theta_h);
		sth[j] = (float)sin((double)theta_h);
		ui[j] = u_i;
		double st_, ct_;
		sincos((double)theta_h, &st_, &ct_);
		gst_f[j] = (float)st_;
		gct_f[j] = (float)ct_;
	}
	for (int i = tid; i < cnt; i += nt) { // the view angles
		float t = (float)i / (float)cnt;
		float theta_k = (float)((double)t * 0.5 * DJB_PI);
		ck_d[i] = (double)(float)cos((double)theta_k);
		sk_f[i] = (float)sin((double)theta_k);
	}
	__syncthreads();
	}
	if (on(PH_GRID)) {
	float *gw = split ? grid_g : grid;
	for (int e = part * nt + tid; e < SIG_NPHI * SIG_NTHETA; e += nt * nparts) { // ndf(vec3(theta_h, phi_h)) does not depend on the view angle
		int j2 = e / SIG_NTHETA, j1 = e - j2 * SIG_NTHETA;
		const double sd = (double)gst_f[j1];
		gw[e] = tab_ndf(tab, sp, mk((float)(sd * gcp_d[j2]), (float)(sd * gsp_d[j2]), gct_f[j1]));
	}
	__syncthreads();
	FIT_PHASE(5);
	} // PH_GRID
	if (on(PH_SIGMA)) {
	if (split) {
		for (int e = tid; e < SIG_NPHI * SIG_NTHETA; e += nt) grid[e] = grid_g[e];
		__syncthreads();
	}
	{
		// 89 ordered sums of 16 200 terms each, one thread per view angle; every operand of the inner loop is a shared-memory
		// broadcast.  Per term: two float <-> double conversions and three FP64 operations; a warp-wide FP64 / conversion
		// instruction occupies its scheduler's pipe for ~8 cycles whatever the number of active lanes, which is what bounds the
		// stage (36 cycles per term on each of the three schedulers in use).  Tried and dropped (profiles/r02_h_fit_phases.md):
		// (a) the whole CTA computing the 89 x 90 terms of one azimuth slab into shared memory, lane i adding its 90 terms in
		// order -- the same FP64 work plus the staging: 810 k cycles against 588 k; (b) tabulating the two azimuth-independent
		// operands (128 KB): 530 k, but the matrix then no longer fits in shared memory and the 50 iterations lose more.
		const float dtheta = (float)(DJB_PI / (double)(float)SIG_NTHETA);
		const float dphi = (float)(2.0 * DJB_PI / (double)(float)SIG_NPHI);
		if (split && K_SMEM) {
			// this part's chains c0 .. c1 - 1 (at most 32); two slab buffers in the idle matrix block
			const int cpp = (cnt + nparts - 1) / nparts, c0 = part * cpp, nc = (c0 + cpp < cnt ? c0 + cpp : cnt) - c0;
			float *slab = reinterpret_cast<float *>(Ks) + (((3 * cnt + 2 * SIG_NPHI + SIG_NPHI + SIG_NTHETA) & 1) ? 2 : 0); // 16-byte aligned
			const int nterm = nc * SIG_NTHETA, lane = tid & 31, warp = tid >> 5, nprod = nt - 32;
			auto produce = [&](int j2, float *buf) { // independent terms of azimuth slab j2, warps 1..7
				const double cp = cphi_d[j2];
				const float *g = grid + j2 * SIG_NTHETA;
				for (int e = tid - 32; e < nterm; e += nprod) {
					const int il = e / SIG_NTHETA, j1 = e - il * SIG_NTHETA;
					const float s_h = sth[j1];
					const float kh = (float)((double)(sk_f[c0 + il] * s_h) * cp + ck_d[c0 + il] * cth_d[j1]);
					buf[il * SLAB_PITCH + j1] = fmax_ref(0.0f, kh) * g[j1] * ui[j1] * s_h;
				}
			};
			if (warp != 0) produce(0, slab);
			__syncthreads();
			float nint = 0.0f;
			for (int j2 = 0; j2 < SIG_NPHI; ++j2) {
				float *cur = slab + (j2 & 1) * 32 * SLAB_PITCH, *nxt = slab + ((j2 + 1) & 1) * 32 * SLAB_PITCH;
				if (warp == 0) {
					if (lane < nc) {
						const float4 *row = reinterpret_cast<const float4 *>(cur + lane * SLAB_PITCH);
#pragma unroll
						for (int q = 0; q < SIG_NTHETA / 4; ++q) { // 22 x 4 terms, in index order
							const float4 t = row[q];
							nint += t.x; nint += t.y; nint += t.z; nint += t.w;
						}
						nint += cur[lane * SLAB_PITCH + 88];
						nint += cur[lane * SLAB_PITCH + 89];
					}
				} else if (j2 + 1 < SIG_NPHI) {
					produce(j2 + 1, nxt);
				}
				__syncthreads();
			}
			if (warp == 0 && lane < nc) {
				nint *= dtheta * dphi;
				const float sg = fmax_ref((float)ck_d[c0 + lane], nint);
				o_sigma[c0 + lane] = sg;
				if (c0 + lane == cnt - 1) o_sigma[cnt] = sg;
			}
		} else {
			for (int i = tid; i < cnt; i += nt) { // every operand of the inner loop is a shared-memory broadcast
				const float sk = sk_f[i];
				const double ckd = ck_d[i];
				float nint = 0.0f;
				for (int j2 = 0; j2 < SIG_NPHI; ++j2) {
					const double cp = cphi_d[j2];
					const float *g = grid + j2 * SIG_NTHETA;
#pragma unroll 10
					for (int j1 = 0; j1 < SIG_NTHETA; ++j1) {
						float s_h = sth[j1];
						float kh = (float)((double)(sk * s_h) * cp + ckd * cth_d[j1]);
						nint += fmax_ref(0.0f, kh) * g[j1] * ui[j1] * s_h;
					}
				}
				nint *= dtheta * dphi;
				s_sigma[i] = fmax_ref((float)ckd, nint);
			}
		}
	}
	__syncthreads();
	if (!split) {
		if (tid == 0) s_sigma[cnt] = s_sigma[cnt - 1];
		__syncthreads();
	}
	FIT_PHASE(6);
	} // PH_SIGMA

	// ---- compute_fresnel, dj_brdf.h:2583-2641 -------------------------------------------------------
	// The reference walks theta_h for every theta_d bin i; the (i, j) evaluations are independent, only the running
	// sums are ordered.  So: trip counts per bin, all ratios in parallel over the flattened (i, j) list, ordered sums.
	if (on(PH_FRESNEL) || on(PH_FINISH)) {
	for (int i = tid; i < cnt; i += nt) {
		float tt = (float)i / (float)cnt;
		float theta_d = (float)((double)tt * DJB_PI * 0.5);
		const double bound = DJB_PI * 0.5 - (double)theta_d;
		float theta_h = 0.0f;
		int J = 0;
		for (int j = 0; (double)theta_h < bound && j < cnt + 2; ++j) { // theta_h(cnt + 1) > pi/2 >= bound: at most cnt + 2 trips
			float t1 = (float)j / (float)cnt;
			theta_h = (float)((double)(t1 * t1) * DJB_PI * 0.5);
			++J;
		}
		fr_count[i] = J;
	}
	__syncthreads();
	if (tid == 0) {
		int acc = 0;
		for (int i = 0; i < cnt; ++i) { fr_offset[i] = acc; acc += fr_count[i]; }
		fr_offset[cnt] = acc;
	}
	__syncthreads();
	}
	if (on(PH_FRESNEL)) {
		const int total = fr_offset[cnt];
		const float phi_d = (float)(DJB_PI * 0.5), phi_h = 0.0f;
		for (int e = part * nt + tid; e < total; e += nt * nparts) {
			int lo = 0, hi = cnt - 1; // bin i with fr_offset[i] <= e < fr_offset[i + 1]
			while (lo < hi) {
				int mid = (lo + hi + 1) >> 1;
				if (fr_offset[mid] <= e) lo = mid; else hi = mid - 1;
			}
			const int i = lo, j = e - fr_offset[lo];
			float tt = (float)i / (float)cnt;
			float theta_d = (float)((double)tt * DJB_PI * 0.5);
			float t1 = (float)j / (float)cnt;
			float theta_h = (float)((double)(t1 * t1) * DJB_PI * 0.5);
			float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
			if (!((double)theta_h > DJB_PI * 0.5)) {
				V3 dir_i, dir_o;
				hd_to_io(spherical(theta_h, phi_h), spherical(theta_d, phi_d), dir_i, dir_o);
				dir_i = mk(0.f, 0.f, 1.f); // "hack to reproduce my EGSR fits", dj_brdf.h:2609
				V3 fr1 = source_eval(src, dir_i, dir_o);
				float fr2 = tab_eval_ideal(tab, sp, shadow, dir_i, dir_o); // ideal Fresnel: r == g == b
				if ((double)fr2 > 1e-4) r = make_float4(fr1.x / fr2, fr1.y / fr2, fr1.z / fr2, 1.0f);
			}
			fres_ws[e] = r;
		}
	}
	__syncthreads();
	FIT_PHASE(7);
	if (!on(PH_FINISH)) return;
	float *o_fres = A.fresnel + (size_t)mat * res * 3;
	for (int i = tid; i < cnt; i += nt) {
		V3 f = mk(0.f, 0.f, 0.f);
		int c = 0;
		const float4 *w = fres_ws + fr_offset[i];
		for (int j = 0; j < fr_count[i]; ++j) {
			float4 r = w[j];
			if (r.w != 0.0f) { f.x += r.x; f.y += r.y; f.z += r.z; ++c; }
		}
		V3 o;
		o.x = c == 0 ? 1.0f : fmin_ref(1.0f, f.x / (float)c);
		o.y = c == 0 ? 1.0f : fmin_ref(1.0f, f.y / (float)c);
		o.z = c == 0 ? 1.0f : fmin_ref(1.0f, f.z / (float)c);
		o_fres[3 * i] = o.x; o_fres[3 * i + 1] = o.y; o_fres[3 * i + 2] = o.z;
		if (i == cnt - 1) { o_fres[3 * cnt] = o.x; o_fres[3 * cnt + 1] = o.y; o_fres[3 * cnt + 2] = o.z; }
	}

	// ---- compute_cdf, dj_brdf.h:2705-2727 ------------------------------------------------------------
	for (int i = tid; i < cnt; i += nt) {
		float u = (float)i / (float)cnt;
		float theta_h = (float)((double)(u * u) * DJB_PI * 0.5);
		float c_h = (float)cos((double)theta_h), r_h = (float)tan((double)theta_h);
		scan[i] = (u * r_h * tab.p22_radial(r_h * r_h)) / (c_h * c_h);
	}
	__syncthreads();
	if (tid == 0) {
		const float dtheta = (float)(DJB_PI / (double)(float)cnt);
		float nint = 0.0f;
		for (int i = 0; i < cnt; ++i) {
			nint += scan[i];
			s_cdf[i] = (float)((double)(nint * dtheta) * (2.0 * DJB_PI));
		}
		s_cdf[cnt] = 1.0f;
	}
	__syncthreads();
	FIT_PHASE(8);

	// ---- compute_qf, dj_brdf.h:2731-2762: cdf_radial on the 8x finer grid in parallel, then the scan ---
	const int qres = cnt * 8;
	for (int j = tid; j < qres; j += nt) {
		float u = (float)j / (float)qres;
		float theta_h = (float)((double)u * DJB_PI * 0.5);
		float r = (float)tan((double)theta_h);
		float uu = (float)(atan((double)r) * (double)2.0f / (double)(float)DJB_PI); // cdf_radial, :2165-2170
		if (uu < 0.0f) uu = 0.0f;
		scan[j] = spline_f(s_cdf, res, (float)sqrt((double)uu));
	}
	__syncthreads();
	float *o_qf = A.qf + (size_t)mat * res;
	if (tid == 0) {
		int j = 0, n = 0;
		o_qf[n++] = 0.0f;
		for (int i = 1; i < cnt; ++i) {
			float c = (float)i / (float)cnt;
			for (; j < qres; ++j)
				if (scan[j] >= c) { o_qf[n++] = (float)j / (float)qres; break; }
		}
		if (n < res) o_qf[n++] = 1.0f;
		for (; n < res; ++n) o_qf[n] = 0.0f; // entries the reference never pushes
	}

	// ---- fit_beckmann_parameters / fit_ggx_parameters, dj_brdf.h:3133-3184 -----------------------------
	for (int i = tid; i < NORM_NTHETA; i += nt) {
		float u = (float)i / (float)NORM_NTHETA;
		float theta_h = (float)((double)(u * u) * DJB_PI * 0.5);
		float c_h = (float)cos((double)theta_h), r_h = (float)tan((double)theta_h);
		float r2 = r_h * r_h;
		float p = tab.p22_radial(r2);
		terms[i] = (u * r2 * r_h * p) / (c_h * c_h);
		terms[NORM_NTHETA + i] = (u * r2 * p) / (c_h * c_h);
	}
	__syncthreads();
	if (tid == 0) {
		const float dtheta = (float)(DJB_PI / (double)(float)NORM_NTHETA);
		float nb = 0.0f, ng = 0.0f;
		for (int i = 0; i < NORM_NTHETA; ++i) { nb += terms[i]; ng += terms[NORM_NTHETA + i]; }
		nb = (float)((double)nb * ((double)dtheta * DJB_PI));
		ng = (float)((double)ng * ((double)dtheta * 4.0));
		A.alpha[2 * mat] = (float)sqrt(2.0 * (double)nb);
		A.alpha[2 * mat + 1] = ng;
	}
	for (int a = tid; a < res; a += nt) {
		A.p22[(size_t)mat * res + a] = s_p22[a];
		A.sigma[(size_t)mat * res + a] = s_sigma[a];
		A.cdf[(size_t)mat * res + a] = s_cdf[a];
	}
	FIT_PHASE(9);
}

cudaError_t fit_phase_clocks(long long out[10])
{
	return cudaMemcpyFromSymbol(out, g_fit_phase_clock, sizeof(long long) * 10);
}

std::atomic<int> g_fit_parts{0};
// CTAs per material: 1 = the whole fit in one launch; 3..8 = the split mode (six launches), chosen when the batch would leave
// most SMs idle.  The split mode needs the matrix block in shared memory (the slab buffers of its sigma phase live there) and
// at most 32 chains per part.
int fit_tabular_parts(int n_materials, int res)
{
	const int forced = g_fit_parts.load(std::memory_order_relaxed); // djb200_debug_fit_parts: A/B tests
	const int cnt = res - 1;
	const size_t slab_bytes = 2 * 32 * SLAB_PITCH * sizeof(float) + 8;
	if (!iso_smem_plan(res, 0).k_in_smem || sizeof(double) * cnt * cnt < slab_bytes) return 1;
	int parts = forced > 0 ? forced : sm_count() / (n_materials > 0 ? n_materials : 1);
	if (parts > 8) parts = 8;
	if (parts < 3 || (cnt + parts - 1) / parts > 32) return 1;
	return parts;
}


cudaError_t launch_fit_tabular(const FitSourceDev *sources_dev, int n_materials, int res, int shadow, int iterations,
                               double *K_ws, float4 *fres_ws, float *grid_ws, float *p22, float *sigma, float *cdf, float *qf,
                               float *fresnel, float *alpha, float *residuals, cudaStream_t st)
{
	if (n_materials <= 0) return cudaSuccess;
	IsoFitArgs A;
	A.sources = sources_dev;
	A.res = res; A.shadow = shadow; A.iterations = iterations;
	A.K = K_ws; A.fres_ws = fres_ws; A.grid_ws = grid_ws; A.phase = PH_ALL;
	A.p22 = p22; A.sigma = sigma; A.cdf = cdf; A.qf = qf; A.fresnel = fresnel; A.alpha = alpha; A.residuals = residuals;
	const IsoSmemPlan plan = iso_smem_plan(res, iterations);
	A.hist = plan.hist;
	size_t smem = plan.bytes;
	const int parts = grid_ws ? fit_tabular_parts(n_materials, res) : 1;
	auto go = [&](auto kernel) {
		cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return e;
		if (parts == 1) {
			kernel<<<n_materials, FIT_THREADS, smem, st>>>(A);
			g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
			return cudaGetLastError();
		}
		for (int ph = PH_MATRIX; ph <= PH_FINISH; ++ph) {
			A.phase = ph;
			const bool one = ph == PH_ITER || ph == PH_FINISH;
			kernel<<<dim3(n_materials, one ? 1 : parts), FIT_THREADS, smem, st>>>(A);
			g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
		}
		return cudaGetLastError();
	};
	return plan.k_in_smem ? go(fit_tabular_kernel<true>) : go(fit_tabular_kernel<false>);
}

size_t fit_tabular_smem_bytes(int res) { return iso_smem_bytes(res); }

} // namespace djb200

// =====================================================================================================
// Anisotropic fit, djb::tabular_anisotropic (dj_brdf.h:2238-2273).  n = w * h unknowns (w = elevation_res - 1,
// h = azimuthal_res).  The reference materialises the n x n kernel matrix in fp64 (513 MB at 90 x 90) and
// multiplies by it four times; an entry is a product of a row factor and a "ReLU-dot" of a row vector with a
// column vector (dj_brdf.h:2548-2564), so the matvec below recomputes entries on the fly -- same float
// roundings, same summation order, no 513 MB of HBM traffic per iteration.
namespace djb200 {

struct AnisoGeom {
	int er, ar, w, h, n; // elevation_res, azimuthal_res, w = er - 1, h = ar, n = w * h
};

// row factors {zo, xo, yo, kji_tmp1} and column factors {tan_theta, slope1, slope2, cos_theta^2}: both live on
// the same (theta_i1, phi_i2) grid, dj_brdf.h:2536-2561
__global__ void __launch_bounds__(FIT_THREADS) aniso_pre_kernel(FitSourceDev src, AnisoGeom g, float4 *rowpre, float4 *colpre,
                                                                float *colrcp)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= g.n) return;
	const double sqrt_half_pi = sqrt(DJB_PI * 0.5);
	const float dtheta = (float)(sqrt_half_pi / (double)(float)g.w);
	const float dphi = (float)(2.0 * DJB_PI / (double)(float)g.h);
	int i2 = r / g.w, i1 = r - i2 * g.w;
	float t1 = (float)i1 / (float)g.w, t2 = (float)i2 / (float)g.h;
	float theta = (float)((double)t1 * 0.5 * DJB_PI), phi = (float)((double)t2 * 2.0 * DJB_PI);
	double st, ct, sp, cp;
	sincos((double)theta, &st, &ct);
	sincos((double)phi, &sp, &cp);
	float sin_theta = (float)st, zo = (float)ct;
	float xo = (float)((double)sin_theta * cp), yo = (float)((double)sin_theta * sp);
	V3 dir = spherical(theta, phi);
	float fr_i = intensity(source_eval(src, dir, dir));
	float kji1 = (float)((double)(dtheta * dphi) * (4.0 * (double)fr_i * pow((double)zo, 5.0)));
	rowpre[r] = make_float4(zo, xo, yo, kji1);
	float cos_theta = zo, tan_theta = (float)tan((double)theta);
	float s1 = (float)((double)(-tan_theta) * cp), s2 = (float)((double)(-tan_theta) * sp);
	colpre[r] = make_float4(tan_theta, s1, s2, cos_theta * cos_theta);
	colrcp[r] = __frcp_rn(cos_theta * cos_theta); // correctly rounded 1 / den: the matvec divides by den with two FMAs
}

// out[r] = sum_k K(r, k) v[k] for r in [row0, row1), k ascending (matrix::transform, dj_brdf.h:2456-2465).
// The double sum of a row is a chain of n dependent additions: it cannot be split without changing the result, so a row's
// chain (8010 x the latency of a double add at 90 x 90) is the floor of this kernel whatever the number of rows -- which is also
// why sharding the rows over GPUs cannot shorten it.  Everything else is made parallel around that chain.  A CTA owns 32 rows
// (lane = row).  Seven PRODUCER warps form the products K(r, k) v[k] of a block of MV_CHUNK columns -- independent work: entry,
// conversion, multiplication -- into shared memory; the ADDER warp walks the previous block and does nothing but the ordered
// additions (one shared-memory load and one add per term).  Blocks are double-buffered, one barrier per block.  The division by
// den = cos^2(theta_k) uses the column's correctly rounded reciprocal (quotient + one FMA correction = the IEEE quotient).
constexpr int MV_WARPS = 8, MV_THREADS = 32 * MV_WARPS, MV_PER_WARP = 8, MV_CHUNK = (MV_WARPS - 1) * MV_PER_WARP;
DJB_DEV float aniso_entry(const float4 rp, const float4 c, const float y)
{
	const float m_dot_o = rp.x - rp.y * c.y - rp.z * c.z;
	const float p = c.x * fmax_ref(0.0f, m_dot_o);
	// p / c.w given y = RN(1 / c.w): quotient estimate + one FMA correction is the IEEE quotient when the residual is exact,
	// i.e. for normal p; a tiny numerator is lifted by 2^64 first (exact) and the quotient lowered again -- branch-free
	const bool tiny = p < 1e-30f;
	const float ps = tiny ? p * 0x1p64f : p;
	float q = ps * y;
	q = __fmaf_rn(__fmaf_rn(-c.w, q, ps), y, q);
	q = tiny ? q * 0x1p-64f : q;
	return rp.w * q;
}
__global__ void __launch_bounds__(MV_THREADS) aniso_matvec_kernel(AnisoGeom g, const float4 *__restrict__ rowpre,
                                                                  const float4 *__restrict__ colpre,
                                                                  const float *__restrict__ colrcp, const double *__restrict__ v,
                                                                  double *__restrict__ out, int row0, int row1)
{
	__shared__ double s_prod[2][MV_CHUNK][32]; // [buffer][column of the block][row of the CTA]: 28 KB
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int r = row0 + blockIdx.x * 32 + lane;
	const bool live = r < row1;
	const float4 rp = live ? rowpre[r] : make_float4(0.f, 0.f, 0.f, 0.f);
	const int nblocks = (g.n + MV_CHUNK - 1) / MV_CHUNK;
	double acc = 0.0;
	for (int b = 0; b <= nblocks; ++b) {
		if (warp > 0 && b < nblocks) { // producers: block b
			const int k0 = b * MV_CHUNK + (warp - 1) * MV_PER_WARP;
#pragma unroll
			for (int j = 0; j < MV_PER_WARP; ++j) {
				const int k = k0 + j;
				double t = 0.0;
				if (k < g.n) t = (double)aniso_entry(rp, __ldg(colpre + k), __ldg(colrcp + k)) * __ldg(v + k);
				s_prod[b & 1][(warp - 1) * MV_PER_WARP + j][lane] = t;
			}
		} else if (warp == 0 && b > 0) { // adder: block b - 1, in column order
			const int cnt = g.n - (b - 1) * MV_CHUNK < MV_CHUNK ? g.n - (b - 1) * MV_CHUNK : MV_CHUNK;
			const double(*p)[32] = s_prod[(b - 1) & 1];
			if (cnt == MV_CHUNK) {
#pragma unroll
				for (int j = 0; j < MV_CHUNK; ++j) acc += p[j][lane];
			} else {
				for (int j = 0; j < cnt; ++j) acc += p[j][lane];
			}
		}
		__syncthreads();
	}
	if (warp == 0 && live) out[r] = acc;
}

__global__ void fill_ones_kernel(double *v, int n)
{
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) v[k] = 1.0;
}

// diagnostic: || v1/|v1| - v0/|v0| ||_2 (not in the reference; order of this reduction is irrelevant)
__global__ void __launch_bounds__(FIT_THREADS) aniso_residual_kernel(const double *v0, const double *v1, int n, float *out)
{
	__shared__ double red[3][FIT_THREADS];
	double a = 0, b = 0, c = 0;
	for (int k = threadIdx.x; k < n; k += blockDim.x) { a += v0[k] * v0[k]; b += v1[k] * v1[k]; c += v0[k] * v1[k]; }
	red[0][threadIdx.x] = a; red[1][threadIdx.x] = b; red[2][threadIdx.x] = c;
	__syncthreads();
	for (int s = blockDim.x / 2; s > 0; s >>= 1) {
		if (threadIdx.x < s)
			for (int q = 0; q < 3; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + s];
		__syncthreads();
	}
	if (threadIdx.x == 0) { // |a/|a| - b/|b||^2 = 2 - 2 a.b / (|a||b|)
		double d = 2.0 - 2.0 * red[2][0] / (sqrt(red[0][0]) * sqrt(red[1][0]));
		*out = (float)sqrt(d > 0.0 ? d : 0.0);
	}
}

// m_p22 from the iterate, with the zero column at theta = pi/2 (dj_brdf.h:2570-2578)
__global__ void aniso_p22_kernel(AnisoGeom g, const double *v, float *p22)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= g.er * g.ar) return;
	int j = e / g.er, i = e - j * g.er;
	p22[e] = i < g.w ? (float)v[j * g.w + i] : 0.0f;
}

constexpr int AN_NTHETA = 128, AN_NPHI = 256;   // normalize_p22, dj_brdf.h:2308-2309
constexpr int AP_NTHETA = 128, AP_NPHI = 512;   // fit_*_parameters, dj_brdf.h:3189-3190
constexpr int AS_NTHETA = 45, AS_NPHI = 90;     // compute_sigma, dj_brdf.h:2390-2391

// the 256 x 128 terms of normalize_p22 (dj_brdf.h:2314-2327), phi-major like the reference's loops
__global__ void aniso_norm_terms_kernel(AnisoGeom g, const float *p22, float *terms)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= AN_NPHI * AN_NTHETA) return;
	int j = e / AN_NTHETA, i = e - j * AN_NTHETA;
	const double sqrt_half_pi = sqrt(DJB_PI * 0.5);
	TabAniso t; t.p22 = p22; t.sigma = nullptr; t.w = g.er; t.h = g.ar; t.T = g_dm_table_dev;
	float u = (float)j / (float)AN_NPHI;
	float phi = (float)((double)u * 2.0 * DJB_PI);
	float ui = (float)i / (float)AN_NTHETA;
	float theta = (float)((double)ui * sqrt_half_pi);
	float theta_sqr = theta * theta;
	float c = (float)cos((double)theta_sqr);
	float pdf = t.p22_theta_phi(theta_sqr, phi);
	float weight = (float)(((double)theta * tan((double)theta_sqr)) / (double)(c * c));
	terms[e] = weight * pdf;
}

// The sum of p[0 .. n) in index order (a float sum is not associative: the reference's order is the result), computed by a whole
// warp: the lanes fetch the next 1024 terms with coalesced 16-byte loads while the current 1024, staged in shared memory, are
// added in order (every lane redundantly: one broadcast 16-byte load per four additions) -- the chain of dependent additions is
// the only serial part left (one thread walking global memory paid a cache latency per term).  n must be a multiple of 4 and
// p 16-byte aligned; `stage`: 1024 floats of shared memory owned by this warp.
constexpr int OS_CHUNK = 1024;
DJB_DEV float ordered_sum_warp(const float *__restrict__ p, int n, float *stage)
{
	const int lane = threadIdx.x & 31;
	const float4 *p4 = reinterpret_cast<const float4 *>(p);
	float4 *s4 = reinterpret_cast<float4 *>(stage);
	const int n4 = n >> 2;
	float4 nxt[OS_CHUNK / 128];
	auto fetch = [&](int base4) {
#pragma unroll
		for (int q = 0; q < OS_CHUNK / 128; ++q) {
			const int i = base4 + q * 32 + lane;
			nxt[q] = i < n4 ? p4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
		}
	};
	float acc = 0.0f;
	fetch(0);
	for (int base4 = 0; base4 < n4; base4 += OS_CHUNK / 4) {
		__syncwarp();
#pragma unroll
		for (int q = 0; q < OS_CHUNK / 128; ++q) s4[q * 32 + lane] = nxt[q];
		__syncwarp();
		fetch(base4 + OS_CHUNK / 4); // in flight while this chunk is added
		const int c4 = n4 - base4 < OS_CHUNK / 4 ? n4 - base4 : OS_CHUNK / 4;
		if (c4 == OS_CHUNK / 4) {
#pragma unroll 16
			for (int j = 0; j < OS_CHUNK / 4; ++j) {
				const float4 x = s4[j];
				acc += x.x; acc += x.y; acc += x.z; acc += x.w;
			}
		} else {
			for (int j = 0; j < c4; ++j) {
				const float4 x = s4[j];
				acc += x.x; acc += x.y; acc += x.z; acc += x.w;
			}
		}
	}
	return acc;
}

// one warp adds the terms in the reference's order and derives the normalisation constant
__global__ void aniso_norm_sum_kernel(const float *terms, float *scale_out)
{
	__shared__ __align__(16) float s_stage[OS_CHUNK];
	const double sqrt_half_pi = sqrt(DJB_PI * 0.5);
	float k = ordered_sum_warp(terms, AN_NPHI * AN_NTHETA, s_stage);
	if (threadIdx.x != 0) return;
	float dtheta = (float)(sqrt(0.5 * DJB_PI) / (double)(float)AN_NTHETA);
	float dphi = (float)(2.0 * DJB_PI / (double)(float)AN_NPHI);
	(void)sqrt_half_pi;
	k = (float)((double)k * (2.0 * (double)dtheta * (double)dphi));
	*scale_out = (float)(1.0 / (double)k);
}

__global__ void aniso_scale_kernel(float *p22, int n, const float *scale)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < n) p22[e] *= *scale;
}

// compute_sigma tables that do not depend on the view direction (dj_brdf.h:2397-2421):
//   nd[j2 * 45 + j1] = ndf(vec3(theta_sqr, phi)),  cth[j1] = cos(theta_sqr) (double), sth[j1], th[j1],
//   cdp[i2 * 90 + j2] = cos(phi_j2 - phi_k(i2)) (double, the difference formed in float)
__global__ void aniso_sigma_pre_kernel(AnisoGeom g, const float *p22, float *nd, double *cth, float *sth, float *th, double *cdp)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	const double sqrt_half_pi = sqrt(DJB_PI * 0.5);
	if (e < AS_NPHI * AS_NTHETA) {
		int j2 = e / AS_NTHETA, j1 = e - j2 * AS_NTHETA;
		TabAniso t; t.p22 = p22; t.sigma = nullptr; t.w = g.er; t.h = g.ar; t.T = g_dm_table_dev;
		float phi = (float)((double)((float)j2 / (float)AS_NPHI) * 2.0 * DJB_PI);
		float theta = (float)((double)((float)j1 / (float)AS_NTHETA) * sqrt_half_pi);
		nd[e] = tab_ndf(t, standard_params(), spherical(theta * theta, phi));
	}
	if (e < AS_NTHETA) {
		float theta = (float)((double)((float)e / (float)AS_NTHETA) * sqrt_half_pi);
		float theta_sqr = theta * theta;
		cth[e] = cos((double)theta_sqr);
		sth[e] = (float)sin((double)theta_sqr);
		th[e] = theta;
	}
	if (e < g.h * AS_NPHI) {
		int i2 = e / AS_NPHI, j2 = e - i2 * AS_NPHI;
		float phi_k = (float)((double)((float)i2 / (float)g.h) * 2.0 * DJB_PI);
		float phi = (float)((double)((float)j2 / (float)AS_NPHI) * 2.0 * DJB_PI);
		cdp[e] = cos((double)(phi - phi_k));
	}
}

// sigma_rows[r] for r = i2 * w + i1 in [row0, row1): the 90 x 45 quadrature in the reference's order
__global__ void __launch_bounds__(64) aniso_sigma_kernel(AnisoGeom g, const float *__restrict__ nd, const double *__restrict__ cth,
                                                         const float *__restrict__ sth, const float *__restrict__ th,
                                                         const double *__restrict__ cdp, float *__restrict__ sigma_rows,
                                                         int row0, int row1)
{
	__shared__ double s_cth[AS_NTHETA];
	__shared__ float s_sth[AS_NTHETA], s_w[AS_NTHETA];
	for (int k = threadIdx.x; k < AS_NTHETA; k += blockDim.x) {
		s_cth[k] = cth[k];
		s_sth[k] = sth[k];
		s_w[k] = th[k] * sth[k]; // weight = theta * sin_theta
	}
	__syncthreads();
	int r = row0 + blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= row1) return;
	const double sqrt_half_pi = sqrt(DJB_PI * 0.5);
	const float dtheta = (float)(sqrt_half_pi / (double)(float)AS_NTHETA);
	const float dphi = (float)(2.0 * DJB_PI / (double)(float)AS_NPHI);
	int i2 = r / g.w, i1 = r - i2 * g.w;
	float theta_k = (float)((double)((float)i1 / (float)g.w) * 0.5 * DJB_PI);
	float cos_theta_k = (float)cos((double)theta_k);
	const double sin_k = sin((double)theta_k), cos_k = (double)cos_theta_k;
	float nint = 0.0f;
	for (int j2 = 0; j2 < AS_NPHI; ++j2) {
		const double cd = __ldg(cdp + i2 * AS_NPHI + j2);
		const float *ndrow = nd + j2 * AS_NTHETA;
#pragma unroll 5
		for (int j1 = 0; j1 < AS_NTHETA; ++j1) {
			float m_dot_k = (float)(sin_k * (double)s_sth[j1] * cd + cos_k * s_cth[j1]);
			float masking = fmax_ref(0.0f, m_dot_k) * __ldg(ndrow + j1);
			nint += s_w[j1] * masking;
		}
	}
	nint = (float)((double)nint * (2.0 * (double)dtheta * (double)dphi));
	sigma_rows[r] = fmax_ref(cos_theta_k, nint);
}

// sigma table (er x ar, with the duplicated last elevation, dj_brdf.h:2426) from the per-row values
__global__ void aniso_sigma_table_kernel(AnisoGeom g, const float *sigma_rows, float *sigma)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= g.er * g.ar) return;
	int j = e / g.er, i = e - j * g.er;
	sigma[e] = sigma_rows[j * g.w + (i < g.w ? i : g.w - 1)];
}

// compute_fresnel, dj_brdf.h:2643-2701.  The reference walks theta_h for every theta_d bin i; the (i, j) evaluations are independent,
// only the running sums are ordered.  ws[i * (cnt + 2) + j] = (ratio rgb, 1) when trip j of bin i runs and passes the 1e-4 gate,
// zeros otherwise.  Trip j runs iff theta_h of trip j - 1 is below the bin's bound (theta_h grows with j: the trips are a prefix).
__global__ void __launch_bounds__(128) aniso_fresnel_ratio_kernel(FitSourceDev src, AnisoGeom g, int shadow, const float *p22,
                                                                  const float *sigma, float4 *ws)
{
	TabAniso t; t.p22 = p22; t.sigma = sigma; t.w = g.er; t.h = g.ar; t.T = g_dm_table_dev;
	const int cnt = g.er - 1, stride = cnt + 2;
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= cnt * stride) return;
	const int i = e / stride, j = e - i * stride;
	const Params sp = standard_params();
	const float phi_d = (float)(DJB_PI * 0.5), phi_h = 0.0f;
	const float theta_d = (float)((double)((float)i / (float)cnt) * DJB_PI * 0.5);
	const double bound = DJB_PI * 0.5 - (double)theta_d;
	float prev = 0.0f;
	if (j > 0) {
		const float t0 = (float)(j - 1) / (float)cnt;
		prev = (float)((double)(t0 * t0) * DJB_PI * 0.5);
	}
	float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
	const float t1 = (float)j / (float)cnt;
	const float theta_h = (float)((double)(t1 * t1) * DJB_PI * 0.5);
	if ((double)prev < bound && !((double)theta_h > DJB_PI * 0.5)) {
		V3 dir_i, dir_o;
		hd_to_io(spherical(theta_h, phi_h), spherical(theta_d, phi_d), dir_i, dir_o);
		dir_i = mk(0.f, 0.f, 1.f); // "hack to reproduce my EGSR fits", dj_brdf.h:2669
		const V3 fr1 = source_eval(src, dir_i, dir_o);
		const float fr2 = tab_eval_ideal(t, sp, shadow != 0, dir_i, dir_o); // ideal Fresnel: r == g == b
		if ((double)fr2 > 1e-4) r = make_float4(fr1.x / fr2, fr1.y / fr2, fr1.z / fr2, 1.0f);
	}
	ws[e] = r;
}
__global__ void __launch_bounds__(FIT_THREADS) aniso_fresnel_sum_kernel(AnisoGeom g, const float4 *ws, float *fresnel)
{
	const int cnt = g.er - 1, stride = cnt + 2;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
		V3 f = mk(0.f, 0.f, 0.f);
		int c = 0;
		for (int j = 0; j < stride; ++j) {
			const float4 r = ws[i * stride + j];
			if (r.w != 0.0f) { f.x += r.x; f.y += r.y; f.z += r.z; ++c; }
		}
		V3 o;
		o.x = c == 0 ? 1.0f : fmin_ref(1.0f, f.x / (float)c);
		o.y = c == 0 ? 1.0f : fmin_ref(1.0f, f.y / (float)c);
		o.z = c == 0 ? 1.0f : fmin_ref(1.0f, f.z / (float)c);
		fresnel[3 * i] = o.x; fresnel[3 * i + 1] = o.y; fresnel[3 * i + 2] = o.z;
		if (i == cnt - 1) { fresnel[3 * cnt] = o.x; fresnel[3 * cnt + 1] = o.y; fresnel[3 * cnt + 2] = o.z; }
	}
}

// the 512 x 128 terms of fit_beckmann_parameters / fit_ggx_parameters (dj_brdf.h:3196-3226, 3260-3286):
// seven distinct sequences -- tmp2*e1, tmp2*e2 (shared), tmp2*e3, tmp2*e4, tmp2*e5 (beckmann), tmp2*|e1|, tmp2*|e2| (ggx)
__global__ void aniso_param_terms_kernel(AnisoGeom g, const float *p22, float *terms)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	const int N = AP_NPHI * AP_NTHETA;
	if (e >= N) return;
	int j = e / AP_NTHETA, i = e - j * AP_NTHETA;
	const double sqrt_half_pi = sqrt(DJB_PI * 0.5);
	TabAniso t; t.p22 = p22; t.sigma = nullptr; t.w = g.er; t.h = g.ar; t.T = g_dm_table_dev;
	float phi = (float)((double)((float)j / (float)AP_NPHI) * 2.0 * DJB_PI);
	float cos_phi = (float)cos((double)phi), sin_phi = (float)sin((double)phi);
	float cos_phi_sqr = cos_phi * cos_phi, sin_phi_sqr = sin_phi * sin_phi;
	float theta = (float)((double)((float)i / (float)AP_NTHETA) * sqrt_half_pi);
	float theta_sqr = theta * theta;
	float pv = t.p22_theta_phi(theta_sqr, phi);
	float tan_theta = (float)tan((double)theta_sqr), cos_theta = (float)cos((double)theta_sqr);
	float tan_theta_sqr = tan_theta * tan_theta, cos_theta_sqr = cos_theta * cos_theta;
	float tmp2 = theta * pv * tan_theta / cos_theta_sqr;
	float e1 = -tan_theta * cos_phi, e2 = -tan_theta * sin_phi;
	float e3 = tan_theta_sqr * cos_phi_sqr, e4 = tan_theta_sqr * sin_phi_sqr;
	float e5 = tan_theta_sqr * cos_phi * sin_phi;
	terms[0 * N + e] = tmp2 * e1;
	terms[1 * N + e] = tmp2 * e2;
	terms[2 * N + e] = tmp2 * e3;
	terms[3 * N + e] = tmp2 * e4;
	terms[4 * N + e] = tmp2 * e5;
	terms[5 * N + e] = tmp2 * fabsf(e1);
	terms[6 * N + e] = tmp2 * fabsf(e2);
}

// seven warps add one sequence each, in order; thread 0 then forms both parameter sets
__global__ void __launch_bounds__(7 * 32) aniso_param_sums_kernel(const float *terms, float *beckmann5, float *ggx5)
{
	__shared__ float s[7];
	__shared__ __align__(16) float s_stage[7][OS_CHUNK];
	const int N = AP_NPHI * AP_NTHETA;
	{
		const int seq = threadIdx.x >> 5;
		const float acc = ordered_sum_warp(terms + (size_t)seq * N, N, s_stage[seq]);
		const double sqrt_half_pi = sqrt(DJB_PI * 0.5);
		float dtheta = (float)(sqrt_half_pi / (double)(float)AP_NTHETA);
		float dphi = (float)(2.0 * DJB_PI / (double)(float)AP_NPHI);
		if ((threadIdx.x & 31) == 0) s[seq] = (float)((double)acc * (2.0 * (double)dtheta * (double)dphi));
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		float mux = s[0], muy = s[1];
		float ax = (float)sqrt((double)(2.0f * (s[2] - mux * mux)));
		float ay = (float)sqrt((double)(2.0f * (s[3] - muy * muy)));
		float rho = (float)(2.0 * (double)(s[4] - mux * muy) / (double)(ax * ay));
		beckmann5[0] = ax; beckmann5[1] = ay; beckmann5[2] = rho; beckmann5[3] = mux; beckmann5[4] = muy;
		float gx = (float)sqrt((double)(s[5] * s[5] - mux * mux));
		float gy = (float)sqrt((double)(s[6] * s[6] - muy * muy));
		ggx5[0] = gx; ggx5[1] = gy; ggx5[2] = 0.0f; ggx5[3] = mux; ggx5[4] = muy;
	}
}

// ---- host-side launchers (one material; [row0, row1) is this GPU's shard of the rows) ---------------------
static inline void count_launch() { g_kernel_launches.fetch_add(1, std::memory_order_relaxed); }
static inline int blocks_for(int n, int t) { return (n + t - 1) / t; }

cudaError_t aniso_launch_pre(const FitSourceDev &src, int er, int ar, float4 *rowpre, float4 *colpre, float *colrcp, double *v_ones,
                             cudaStream_t st)
{
	AnisoGeom g = {er, ar, er - 1, ar, (er - 1) * ar};
	aniso_pre_kernel<<<blocks_for(g.n, FIT_THREADS), FIT_THREADS, 0, st>>>(src, g, rowpre, colpre, colrcp);
	count_launch();
	if (v_ones) {
		fill_ones_kernel<<<blocks_for(g.n, 256), 256, 0, st>>>(v_ones, g.n);
		count_launch();
	}
	return cudaGetLastError();
}

cudaError_t aniso_launch_matvec(int er, int ar, const float4 *rowpre, const float4 *colpre, const float *colrcp, const double *v_in,
                                double *v_out, int row0, int row1, cudaStream_t st)
{
	AnisoGeom g = {er, ar, er - 1, ar, (er - 1) * ar};
	if (row1 <= row0) return cudaSuccess;
	aniso_matvec_kernel<<<blocks_for(row1 - row0, 32), MV_THREADS, 0, st>>>(g, rowpre, colpre, colrcp, v_in, v_out, row0, row1);
	count_launch();
	return cudaGetLastError();
}

cudaError_t aniso_launch_residual(int n, const double *v0, const double *v1, float *out, cudaStream_t st)
{
	aniso_residual_kernel<<<1, FIT_THREADS, 0, st>>>(v0, v1, n, out);
	count_launch();
	return cudaGetLastError();
}

// p22 table from the iterate + normalize_p22; `terms` holds >= 7 * 512 * 128 floats, scale_tmp 1 float
cudaError_t aniso_launch_p22(int er, int ar, const double *v, float *p22, float *terms, float *scale_tmp, cudaStream_t st)
{
	AnisoGeom g = {er, ar, er - 1, ar, (er - 1) * ar};
	aniso_p22_kernel<<<blocks_for(er * ar, 256), 256, 0, st>>>(g, v, p22);
	aniso_norm_terms_kernel<<<blocks_for(AN_NPHI * AN_NTHETA, 256), 256, 0, st>>>(g, p22, terms);
	aniso_norm_sum_kernel<<<1, 32, 0, st>>>(terms, scale_tmp);
	aniso_scale_kernel<<<blocks_for(er * ar, 256), 256, 0, st>>>(p22, er * ar, scale_tmp);
	for (int k = 0; k < 4; ++k) count_launch();
	return cudaGetLastError();
}

size_t aniso_sigma_pre_floats(int ar) { return (size_t)AS_NPHI * AS_NTHETA + 2 * AS_NTHETA; }
size_t aniso_sigma_pre_doubles(int ar) { return (size_t)AS_NTHETA + (size_t)ar * AS_NPHI; }

cudaError_t aniso_launch_sigma(int er, int ar, const float *p22, float *pre_f, double *pre_d, float *sigma_rows, int row0,
                               int row1, cudaStream_t st)
{
	AnisoGeom g = {er, ar, er - 1, ar, (er - 1) * ar};
	float *nd = pre_f, *sth = nd + AS_NPHI * AS_NTHETA, *th = sth + AS_NTHETA;
	double *cth = pre_d, *cdp = cth + AS_NTHETA;
	int pre_n = AS_NPHI * AS_NTHETA;
	if (g.h * AS_NPHI > pre_n) pre_n = g.h * AS_NPHI;
	aniso_sigma_pre_kernel<<<blocks_for(pre_n, 256), 256, 0, st>>>(g, p22, nd, cth, sth, th, cdp);
	count_launch();
	if (row1 > row0) {
		aniso_sigma_kernel<<<blocks_for(row1 - row0, 64), 64, 0, st>>>(g, nd, cth, sth, th, cdp, sigma_rows, row0, row1);
		count_launch();
	}
	return cudaGetLastError();
}

cudaError_t aniso_launch_finish(const FitSourceDev &src, int er, int ar, int shadow, const float *p22, const float *sigma_rows,
                                float *sigma, float *fresnel, float *terms, float *beckmann5, float *ggx5, cudaStream_t st)
{
	AnisoGeom g = {er, ar, er - 1, ar, (er - 1) * ar};
	aniso_sigma_table_kernel<<<blocks_for(er * ar, 256), 256, 0, st>>>(g, sigma_rows, sigma);
	// the Fresnel ratios borrow the head of `terms` (er x (er + 1) float4 <= 7 x 512 x 128 floats for er <= 330); the stream orders
	// their use before aniso_param_terms_kernel overwrites it
	float4 *fres_ws = reinterpret_cast<float4 *>(terms);
	aniso_fresnel_ratio_kernel<<<blocks_for((er - 1) * (er + 1), 128), 128, 0, st>>>(src, g, shadow, p22, sigma, fres_ws);
	aniso_fresnel_sum_kernel<<<blocks_for(er - 1, FIT_THREADS), FIT_THREADS, 0, st>>>(g, fres_ws, fresnel);
	aniso_param_terms_kernel<<<blocks_for(AP_NPHI * AP_NTHETA, 256), 256, 0, st>>>(g, p22, terms);
	aniso_param_sums_kernel<<<1, 7 * 32, 0, st>>>(terms, beckmann5, ggx5);
	for (int k = 0; k < 5; ++k) count_launch();
	return cudaGetLastError();
}

} // namespace djb200
