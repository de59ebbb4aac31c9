// djb_internal.h -- declarations shared by the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/djb200.h"

// opaque handles of include/djb200.h: device-resident, immutable after creation
struct djb200_merl {
	float4 *cells; // one (r, g, b, 0) per cell, already multiplied by the MERL channel scales
	int device;
};
struct djb200_tabular {
	float *tables; // radial: p22[res] | sigma[res] | qf[res] | fresnel[res][3] | cdf[res]; anisotropic: see aniso_table_floats()
	int res, shadow, device;
	int azim_res; // 0: radial tables (djb::tabular); > 0: djb::tabular_anisotropic with res = elevation resolution
	int n_qf1, n_qf2; // anisotropic: entries the quantile-table searches produced (dj_brdf.h:2904-2935, 3004-3037)
};
namespace djb200 {
// per (theta_i, phi_i, theta_v, phi_v) cell: the three channels of the cell (lo) and of its phi_v neighbour, wrapped (hi)
#ifndef DJB200_UTIA_ENTRY_DEFINED
#define DJB200_UTIA_ENTRY_DEFINED
struct __align__(32) UtiaEntry { float4 lo, hi; };
#endif
} // namespace djb200
struct djb200_utia {
	djb200::UtiaEntry *table; // utia::normalize()d samples cast to float, one entry per (theta_i, phi_i, theta_v, phi_v) cell
	int device;
};

namespace djb200 {

// device layout of a djb::tabular_anisotropic handle, T = er * ar floats per 2-D table:
//   p22[T] | sigma[T] | fresnel[er][3] | qf1[ar] | qf2[T] | pdf1[ar] | cdf1[ar] | pdf2[T] | cdf2[T]
__host__ __device__ inline size_t aniso_off_qf1(int er, int ar) { return 2 * (size_t)er * ar + 3 * (size_t)er; }
__host__ __device__ inline size_t aniso_off_qf2(int er, int ar) { return aniso_off_qf1(er, ar) + ar; }
__host__ __device__ inline size_t aniso_off_pdf1(int er, int ar) { return aniso_off_qf2(er, ar) + (size_t)er * ar; }
__host__ __device__ inline size_t aniso_off_cdf1(int er, int ar) { return aniso_off_pdf1(er, ar) + ar; }
__host__ __device__ inline size_t aniso_off_pdf2(int er, int ar) { return aniso_off_cdf1(er, ar) + ar; }
__host__ __device__ inline size_t aniso_off_cdf2(int er, int ar) { return aniso_off_pdf2(er, ar) + (size_t)er * ar; }
__host__ __device__ inline size_t aniso_table_floats(int er, int ar) { return aniso_off_cdf2(er, ar) + (size_t)er * ar; }

// error plumbing of the C-ABI layer (capi.cu)
djb200_status fail(djb200_status s, const char *fmt, ...);
djb200_status cuda_fail(cudaError_t e, const char *what);
djb200_status require_device();

constexpr int MF_INLINE_PARAMS = 16; // params blocks (48 B each) a launch can carry in its kernel arguments
constexpr int PARAMS_LEAN_SHADING = 2; // third params layout, only reachable through djb200_lean_shading_*
enum MfOp { OP_EVAL = 0, OP_EVALP = 1, OP_PDF = 2, OP_SAMPLE = 3, OP_EVALP_IS = 4 };

// One launch of a microfacet query over device-resident arrays.
struct MfLaunch {
	int op, ndf, shadow, fresnel_kind;
	float fv[6];
	const float *spline_pts; // device
	int spline_n;
	const void *params;      // device, djb200_params blocks
	const void *params_host; // BROADCAST only: the same blocks in HOST memory; up to MF_INLINE_PARAMS of them travel inside the
	                         // kernel arguments instead (no descriptor upload, no stream synchronisation): `params` may be NULL then
	int64_t n_params;
	int layout;              // djb200_params_layout, or PARAMS_LEAN_SHADING (internal)
	// PARAMS_LEAN_SHADING: params are built per pair from the renderer's texture fetches (kernels_mf.cu)
	const float *lean_E;     // device, n x 5 LEAN moments
	const float *lean_alpha; // device, n x 3 (alpha1, alpha2, alphaAngle), or NULL: lean_alpha0 for every pair
	float lean_alpha0[3];
	float lean_bias, lean_dmap_scale;
	int lean_filtering;
	const float *a;          // wi (eval/evalp/pdf) or u (sample/evalp_is)
	const float *b;          // wo
	int64_t n;               // pairs in this launch
	int64_t out_stride;      // elements between consecutive params blocks in the outputs (BROADCAST)
	float *out0, *out1, *out2;
};

extern std::atomic<uint64_t> g_kernel_launches;
extern std::atomic<int> g_force_generic; // kernels_mf.cu: 1 = never take the lean FP32 kernels
extern std::atomic<int> g_fast_tier;     // kernels_mf.cu: 1 = eval / evalp / pdf run the 1e-5 tier (default), 0 = the reference's bits
extern std::atomic<int> g_beck_compact;  // kernels_mf.cu: 0 = Beckmann BROADCAST queries stay on the uncompacted lean kernel
int sm_count();

cudaError_t launch_microfacet(const MfLaunch &L, cudaStream_t st);
// only the per-pair params construction of PARAMS_LEAN_SHADING; params_out: device, n x 12 floats
cudaError_t launch_lean_shading_params(const MfLaunch &L, float *params_out, cudaStream_t st);
// djb::tabular as a BRDF (kernels_tabular.cu); tables: device, p22[res] | sigma[res] | qf[res] | fresnel[res][3]
cudaError_t launch_tabular_query(const float *tables, int res, const MfLaunch &L, cudaStream_t st);
// djb::tabular_anisotropic as a BRDF (eval / evalp / pdf); tables: device, p22[er * ar] | sigma[er * ar] | fresnel[er][3]
cudaError_t launch_tabular_aniso_query(const float *tables, int elev_res, int azim_res, int n_qf1, const MfLaunch &L,
                                       cudaStream_t st);
// djb::radial's scalar queries; family: djb200_ndf or 2 = tabular (radial tables: p22 | sigma | qf | fresnel[3] | cdf)
cudaError_t launch_radial_query(int family, int what, const float *tables, int res, const float *x, int64_t n, float *out,
                                cudaStream_t st);
// the remaining public scalar members (kernels_tabular.cu); family: 0 beckmann, 1 ggx, 2 tabular_anisotropic, 3 sgd, 4 abc
cudaError_t launch_member_query(int family, int what, const float *tables, int er, int ar, int n_qf1, const double *coef, int n_coef,
                                const float *a, const float *b, const float *c, int64_t n, float *out, cudaStream_t st);
// builds qf1 | qf2 | pdf1 | cdf1 | pdf2 | cdf2 inside `tables` from its p22 block (dj_brdf.h:2848-3103); counts_host[2]
// receives the fill counts of qf1 / qf2 (the call synchronises the stream)
cudaError_t build_aniso_sampling_tables(float *tables, int elev_res, int azim_res, int counts_host[2], cudaStream_t st);

// tables / frames / LEAN (kernels_tables.cu)
cudaError_t launch_io_to_hd(const float *wi, const float *wo, int64_t n, float *h, float *d, cudaStream_t st);
cudaError_t launch_hd_to_io(const float *h, const float *d, int64_t n, float *wi, float *wo, cudaStream_t st);
cudaError_t launch_merl_convert(const double *samples_dev, float4 *cells_dev, cudaStream_t st);
cudaError_t launch_merl_eval(const float4 *cells, const float *wi, const float *wo, int64_t n, float *out,
                             cudaStream_t st);
cudaError_t launch_merl_index(const float *wi, const float *wo, int64_t n, int32_t *out, cudaStream_t st);
cudaError_t launch_merl_filter_stats(const float *wi, const float *wo, int64_t n, unsigned long long *stats_dev,
                                     cudaStream_t st);
cudaError_t launch_debug_dmath(int fn, const double *x, const double *y, int64_t n, double *out, cudaStream_t st);
cudaError_t launch_utia_convert(const double *raw_dev, UtiaEntry *table_dev, cudaStream_t st);
cudaError_t launch_utia_eval(const UtiaEntry *table, const float *wi, const float *wo, int64_t n, float *out,
                             cudaStream_t st);
cudaError_t launch_nmap_to_leanmap(const uint8_t *nmap, int64_t npix, float base_roughness, float bias,
                                   float *lean1, float *lean2, cudaStream_t st);
int lean_mip_levels(int w, int h, int levels); // levels <= 0: the full chain down to 1 x 1
cudaError_t launch_leanmap_half_mips(const float *planar, int w, int h, int levels, uint16_t *out, float4 *scratch_a, float4 *scratch_b,
                                     cudaStream_t st);
cudaError_t launch_dmap2nmap(const uint8_t *dmap, int w, int h, float scale, uint8_t *nmap, cudaStream_t st);
cudaError_t launch_lrep_to_params(const float *E, int64_t n, void *out_params, cudaStream_t st);
cudaError_t launch_params_to_lrep(const void *params, int64_t n, float *E, cudaStream_t st);
cudaError_t launch_leanmap_to_params(const float *lean1, const float *lean2, int64_t npix, float bias,
                                     void *out_params, cudaStream_t st);

// djb::sgd / djb::abc (kernels_analytic.cu); kind = DJB200_SOURCE_SGD / _ABC, coef = host pointer to the material's doubles
cudaError_t launch_analytic_eval(int kind, const double *coef, int n_coef, const float *wi, const float *wo, int64_t n,
                                 float *out, cudaStream_t st);

// djb::microfacet's component queries (kernels_analytic.cu); params_host: one 48-byte block; spline_pts: device
cudaError_t launch_microfacet_component(int ndf, int shadow, int fresnel_kind, const float fv[6], const float *spline_pts,
                                        int spline_n, const void *params_host, int what, const float *a, const float *b,
                                        const float *c, int64_t n, float *out, cudaStream_t st);

// fits (kernels_fit.cu)
struct FitSourceDev;
size_t fit_tabular_smem_bytes(int res);
cudaError_t fit_phase_clocks(long long out[10]); // SM clock at the phase boundaries of material 0 of the last isotropic fit
// grid_ws: n_materials x 16 200 floats, needed when fit_tabular_parts(n_materials, res) > 1 (small batches run one launch per
// phase with several CTAs per material); NULL forces the single launch
int fit_tabular_parts(int n_materials, int res);
extern std::atomic<int> g_fit_parts; // 0: automatic; 1: always the single launch; 3..8: that many CTAs per material where possible
cudaError_t launch_fit_tabular(const FitSourceDev *sources_dev, int n_materials, int res, int shadow, int iterations,
                               double *K_ws, float4 *fres_ws, float *grid_ws, float *p22, float *sigma, float *cdf, float *qf,
                               float *fresnel, float *alpha, float *residuals, cudaStream_t st);

// anisotropic fit stages; [row0, row1) = this GPU's shard of the n = (er - 1) * ar rows
cudaError_t aniso_launch_pre(const FitSourceDev &src, int er, int ar, float4 *rowpre, float4 *colpre, float *colrcp, double *v_ones,
                             cudaStream_t st);
cudaError_t aniso_launch_matvec(int er, int ar, const float4 *rowpre, const float4 *colpre, const float *colrcp, const double *v_in,
                                double *v_out, int row0, int row1, cudaStream_t st);
cudaError_t aniso_launch_residual(int n, const double *v0, const double *v1, float *out, cudaStream_t st);
cudaError_t aniso_launch_p22(int er, int ar, const double *v, float *p22, float *terms, float *scale_tmp, cudaStream_t st);
size_t aniso_sigma_pre_floats(int ar);
size_t aniso_sigma_pre_doubles(int ar);
cudaError_t aniso_launch_sigma(int er, int ar, const float *p22, float *pre_f, double *pre_d, float *sigma_rows, int row0,
                               int row1, cudaStream_t st);
cudaError_t aniso_launch_finish(const FitSourceDev &src, int er, int ar, int shadow, const float *p22,
                                const float *sigma_rows, float *sigma, float *fresnel, float *terms, float *beckmann5,
                                float *ggx5, cudaStream_t st);

} // namespace djb200
