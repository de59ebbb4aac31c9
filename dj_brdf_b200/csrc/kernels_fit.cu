// kernels_fit.cu -- the "power iteration" fits on the device.
//
// Isotropic fit (djb::tabular, dj_brdf.h:2215-2236): ONE CTA PER MATERIAL, the whole pipeline of the
// reference's constructor in one launch -- kernel matrix, power iterations, normalisation, Smith
// projected-area table, Fresnel table, CDF / quantile tables, Beckmann and GGX roughness -- with
// every intermediate in shared memory (the matrix, cnt^2 doubles, lives in an L2-resident global
// workspace).  A batch of materials fills the GPU: 128 materials = 128 CTAs on 148 SMs.
//
// Parity: every sum keeps the reference's order (a float sum is not associative), so the tables are
// bit-comparable with the CPU oracle.  Work that does not depend on the summation index is hoisted
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
	float *ndf_grid;  // workspace: n_materials x SIG_NPHI x SIG_NTHETA
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
static size_t iso_smem_bytes(int res)
{
	size_t cnt = res - 1;
	return sizeof(double) * (2 * cnt + SIG_NPHI + SIG_NTHETA) +
	       sizeof(float) * (3 * (size_t)res + 4 * cnt + MAX_PHI_STEPS + 2 * NORM_NTHETA + 2 * SIG_NTHETA + 8 * cnt + 8);
}

__global__ void __launch_bounds__(FIT_THREADS) fit_tabular_kernel(IsoFitArgs A)
{
	extern __shared__ double smem_d[];
	const int res = A.res, cnt = res - 1, tid = threadIdx.x, nt = blockDim.x, mat = blockIdx.x;
	double *v0 = smem_d, *v1 = v0 + cnt, *cphi_d = v1 + cnt, *cth_d = cphi_d + SIG_NPHI;
	float *s_p22 = reinterpret_cast<float *>(cth_d + SIG_NTHETA);
	float *s_sigma = s_p22 + res, *s_cdf = s_sigma + res;
	float *row_theta = s_cdf + res, *row_tan = row_theta + cnt, *row_cos = row_tan + cnt, *row_kji = row_cos + cnt;
	float *cosphi = row_kji + cnt, *terms = cosphi + MAX_PHI_STEPS, *sth = terms + 2 * NORM_NTHETA, *ui = sth + SIG_NTHETA;
	float *scan = ui + SIG_NTHETA;
	__shared__ int s_nphi;
	__shared__ float s_scale;

	const FitSourceDev src = A.sources[mat];
	const bool shadow = A.shadow != 0;
	double *K = A.K + (size_t)mat * cnt * cnt;
	float *grid = A.ndf_grid + (size_t)mat * SIG_NPHI * SIG_NTHETA;
	const double sqrt_half_pi = sqrt(DJB_PI * 0.5);
	const Params sp = standard_params();

	// ---- compute_p22_smith, dj_brdf.h:2482-2522 ------------------------------------------------
	const float dphi_h = (float)(DJB_PI / 180.0);
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
	for (int e = tid; e < cnt * cnt; e += nt) {
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
		K[(size_t)j * cnt + i] = (double)entry; // out[i] = sum_j K(i, j) v[j]; stored transposed for coalesced reads
	}
	for (int a = tid; a < cnt; a += nt) v0[a] = 1.0;
	__syncthreads();
	// matrix::eigenvector, dj_brdf.h:2467-2480: un-normalised power iterations, sums in index order
	double *vin = v0, *vout = v1;
	for (int it = 0; it < A.iterations; ++it) {
		for (int a = tid; a < cnt; a += nt) {
			double acc = 0.0;
			for (int b = 0; b < cnt; ++b) acc += K[(size_t)b * cnt + a] * vin[b];
			vout[a] = acc;
		}
		__syncthreads();
		if (A.residuals && tid == 0) { // diagnostic only (not in the reference, never fed back)
			double n0 = 0.0, n1 = 0.0, d = 0.0;
			for (int a = 0; a < cnt; ++a) { n0 += vin[a] * vin[a]; n1 += vout[a] * vout[a]; }
			n0 = sqrt(n0); n1 = sqrt(n1);
			for (int a = 0; a < cnt; ++a) { double x = vout[a] / n1 - vin[a] / n0; d += x * x; }
			A.residuals[(size_t)mat * A.iterations + it] = (float)sqrt(d);
		}
		double *tmp = vin; vin = vout; vout = tmp;
		__syncthreads();
	}
	for (int a = tid; a < cnt; a += nt) s_p22[a] = (float)(1e-2 * vin[a]);
	if (tid == 0) s_p22[cnt] = 0.0f;
	__syncthreads();
	TabIso tab;
	tab.p22 = s_p22; tab.sigma = s_sigma; tab.n = res;

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

	// ---- compute_sigma, dj_brdf.h:2348-2386 -------------------------------------------------------
	for (int j = tid; j < SIG_NPHI; j += nt) {
		float u_j = (float)j / (float)SIG_NPHI;
		cphi_d[j] = cos((double)(float)((double)u_j * 2.0 * DJB_PI));
	}
	for (int j = tid; j < SIG_NTHETA; j += nt) {
		float u_i = (float)j / (float)SIG_NTHETA;
		float theta_h = (float)((double)(u_i * u_i) * DJB_PI * 0.5);
		cth_d[j] = cos((double)theta_h);
		sth[j] = (float)sin((double)theta_h);
		ui[j] = u_i;
	}
	for (int e = tid; e < SIG_NPHI * SIG_NTHETA; e += nt) { // ndf(vec3(theta_h, phi_h)) does not depend on the view angle
		int j2 = e / SIG_NTHETA, j1 = e - j2 * SIG_NTHETA;
		float u_j = (float)j2 / (float)SIG_NPHI, u_i = (float)j1 / (float)SIG_NTHETA;
		float phi_h = (float)((double)u_j * 2.0 * DJB_PI);
		float theta_h = (float)((double)(u_i * u_i) * DJB_PI * 0.5);
		grid[e] = tab_ndf(tab, sp, spherical(theta_h, phi_h));
	}
	__syncthreads();
	{
		const float dtheta = (float)(DJB_PI / (double)(float)SIG_NTHETA);
		const float dphi = (float)(2.0 * DJB_PI / (double)(float)SIG_NPHI);
		for (int i = tid; i < cnt; i += nt) {
			float t = (float)i / (float)cnt;
			float theta_k = (float)((double)t * 0.5 * DJB_PI);
			float ck = (float)cos((double)theta_k), sk = (float)sin((double)theta_k);
			const double ckd = (double)ck;
			float nint = 0.0f;
			for (int j2 = 0; j2 < SIG_NPHI; ++j2) {
				const double cp = cphi_d[j2];
				const float *g = grid + j2 * SIG_NTHETA;
#pragma unroll 6
				for (int j1 = 0; j1 < SIG_NTHETA; ++j1) {
					float s_h = sth[j1];
					float kh = (float)((double)(sk * s_h) * cp + ckd * cth_d[j1]);
					nint += fmax_ref(0.0f, kh) * g[j1] * ui[j1] * s_h;
				}
			}
			nint *= dtheta * dphi;
			s_sigma[i] = fmax_ref(ck, nint);
		}
	}
	__syncthreads();
	if (tid == 0) s_sigma[cnt] = s_sigma[cnt - 1];
	__syncthreads();

	// ---- compute_fresnel, dj_brdf.h:2583-2641 -------------------------------------------------------
	float *o_fres = A.fresnel + (size_t)mat * res * 3;
	for (int i = tid; i < cnt; i += nt) {
		V3 f = fresnel_bin(tab, src, shadow, i, cnt);
		o_fres[3 * i] = f.x; o_fres[3 * i + 1] = f.y; o_fres[3 * i + 2] = f.z;
		if (i == cnt - 1) { o_fres[3 * cnt] = f.x; o_fres[3 * cnt + 1] = f.y; o_fres[3 * cnt + 2] = f.z; }
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
}

cudaError_t launch_fit_tabular(const FitSourceDev *sources_dev, int n_materials, int res, int shadow, int iterations,
                               double *K_ws, float *grid_ws, float *p22, float *sigma, float *cdf, float *qf,
                               float *fresnel, float *alpha, float *residuals, cudaStream_t st)
{
	if (n_materials <= 0) return cudaSuccess;
	IsoFitArgs A;
	A.sources = sources_dev;
	A.res = res; A.shadow = shadow; A.iterations = iterations;
	A.K = K_ws; A.ndf_grid = grid_ws;
	A.p22 = p22; A.sigma = sigma; A.cdf = cdf; A.qf = qf; A.fresnel = fresnel; A.alpha = alpha; A.residuals = residuals;
	size_t smem = iso_smem_bytes(res);
	cudaError_t e = cudaFuncSetAttribute(fit_tabular_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	fit_tabular_kernel<<<n_materials, FIT_THREADS, smem, st>>>(A);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

size_t fit_tabular_smem_bytes(int res) { return iso_smem_bytes(res); }

} // namespace djb200
