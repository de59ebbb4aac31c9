/* djb200.h -- C-ABI of the B200-native BRDF engine (libdjb200.so).
 *
 * This is the drop-in boundary for the hot paths of jdupuy/dj_brdf: every entry point replaces
 * one call a host program makes into the reference's single header `dj_brdf.h` (cited per
 * function as dj_brdf.h:LINE) or into utils/nmap2leanmap*.cpp.  Plain C, plain pointers and
 * sizes, no C++ or torch types.  The C++ facade in include/djb200_facade.hpp re-creates the
 * reference's `djb::` class surface on top of these functions; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - Every function returns a djb200_status; djb200_last_error() gives a thread-local message.
 *     Nothing throws across this boundary (the reference throws djb::exc, dj_brdf.h:54-59).
 *   - Direction arrays are packed float[3] (x,y,z), `n` of them; i = towards the light,
 *     o = towards the viewer (dj_brdf.h:23-26).
 *   - `mem` says where the *bulk* arrays (directions, uniforms, images, outputs) live:
 *     DJB200_MEM_HOST   -- host pointers; the library stages them through the GPU in chunks
 *                          (pinned host memory is copied asynchronously without staging),
 *     DJB200_MEM_DEVICE -- device pointers on the current CUDA device; the call only enqueues
 *                          work on `stream` (a cudaStream_t, NULL = default stream).
 *     Small descriptors (djb200_params blocks, Fresnel data) are always host pointers.
 *   - The library works on the calling thread's current CUDA device.  All entry points are
 *     re-entrant; concurrent calls may share handles (handles are immutable after creation).
 *   - There is no CPU fallback: without a CUDA device every compute call fails with
 *     DJB200_ERR_NO_DEVICE.
 */
#ifndef DJB200_H
#define DJB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define DJB200_API
#else
#define DJB200_API __attribute__((visibility("default")))
#endif

typedef enum djb200_status {
	DJB200_OK = 0,
	DJB200_ERR_INVALID_ARGUMENT = 1,
	DJB200_ERR_NO_DEVICE = 2,
	DJB200_ERR_CUDA = 3,
	DJB200_ERR_OUT_OF_MEMORY = 4,
	DJB200_ERR_IO = 5,
	DJB200_ERR_UNSUPPORTED = 6
} djb200_status;

typedef enum djb200_mem { DJB200_MEM_HOST = 0, DJB200_MEM_DEVICE = 1 } djb200_mem;

/* NDF families: djb::beckmann (dj_brdf.h:327-371), djb::ggx (dj_brdf.h:374-391) */
typedef enum djb200_ndf { DJB200_NDF_BECKMANN = 0, DJB200_NDF_GGX = 1 } djb200_ndf;

/* djb::fresnel::{ideal,schlick,unpolarized,sgd,spline} (dj_brdf.h:149-207, 1292-1344) */
typedef enum djb200_fresnel_kind {
	DJB200_FRESNEL_IDEAL = 0,
	DJB200_FRESNEL_SCHLICK = 1,     /* v[0..2] = f0 */
	DJB200_FRESNEL_UNPOLARIZED = 2, /* v[0..2] = ior */
	DJB200_FRESNEL_SGD = 3,         /* v[0..2] = f0, v[3..5] = f1 */
	DJB200_FRESNEL_SPLINE = 4       /* points: n_points x rgb, uniform in theta_d (dj_brdf.h:1338-1344) */
} djb200_fresnel_kind;

typedef struct djb200_fresnel {
	int32_t kind;
	float v[6];
	const float *points; /* host pointer, n_points * 3 floats (SPLINE only) */
	int32_t n_points;
} djb200_fresnel;

/* Layout-identical to djb::microfacet::params (dj_brdf.h:238-242; 12 floats, 48 bytes), so a
 * reference `const void *user_param` can be passed through unchanged. */
typedef struct djb200_params {
	float n[3];                      /* mean normal */
	float a1, a2, phi_a;             /* ellipse radii and orientation */
	float ax, ay;                    /* slope-space scales */
	float rho, sqrt_one_minus_rho2;  /* slope correlation */
	float tx_n, ty_n;                /* slope-space location */
} djb200_params;

/* How `params` pairs up with the direction arrays:
 *   BROADCAST: n_params blocks, every pair is evaluated under every block; output element
 *              (m, k) is stored at index m * n + k  (material-major, each material contiguous),
 *   PER_PAIR:  n_params == n, pair k uses block k (what a renderer does when roughness comes
 *              from textures, mitsuba/dj_brdf.cpp:353-357); params is then a bulk array (`mem`). */
typedef enum djb200_params_layout { DJB200_PARAMS_BROADCAST = 0, DJB200_PARAMS_PER_PAIR = 1 } djb200_params_layout;

/* djb::microfacet construction state: family, Fresnel term, shadowing flag (dj_brdf.h:285-297) */
typedef struct djb200_microfacet {
	int32_t ndf;    /* djb200_ndf */
	int32_t shadow; /* m_shadow, dj_brdf.h:297 */
	djb200_fresnel fresnel;
} djb200_microfacet;

/* ---- library / device ------------------------------------------------------------------ */
DJB200_API const char *djb200_last_error(void);
DJB200_API const char *djb200_version(void);
DJB200_API djb200_status djb200_device_count(int *count);
DJB200_API djb200_status djb200_set_device(int device);
/* counts kernels launched by this library in the calling process (for bench.py `gpu_launches`) */
DJB200_API uint64_t djb200_kernel_launch_count(void);
/* DJB200_MEM_HOST calls stage through per-thread device buffers that are kept between calls; this frees the
 * calling thread's buffers (they are also freed when the thread exits) */
DJB200_API djb200_status djb200_release_cache(void);

/* Precision of microfacet eval / evalp / pdf / sample (ideal and Schlick Fresnel terms; process-wide, read at every launch).
 *   DJB200_PRECISION_1E5 (default):
 *     eval / evalp / pdf within 1e-5 relative of the reference header's float results, identical zero / NaN pattern (the
 *     tolerance BASELINE.json's north_star states): MUFU reciprocals / exp2 and fused multiply-adds instead of correctly
 *     rounded divisions, square roots and a float-float exponential -- about 2x the throughput.  Measured worst relative
 *     difference against the reference's floats over 3.2e8 results per query: 2.0e-6.
 *     sample: the same algorithm, trip for trip, with MUFU log2 / exp2 / reciprocals; the (0, 0, 1) pattern is identical.
 *     GGX directions are within 1e-5 per component of the reference's (measured maximum 5.3e-6 over 3.2e8 samples).  Beckmann:
 *     median difference 6e-8; the quantile search of dj_brdf.h:1897-1952 stops at |CDF(b) - u| < 1e-5, and where the two
 *     evaluations straddle that threshold the search makes one trip more or less than the reference's, so a sample can move by
 *     the reference's own convergence tolerance: 2.2e-4 of the samples differ by more than 1e-5, 4.4e-6 by more than 1e-4.
 *   DJB200_PRECISION_REFERENCE_BITS: every query reproduces the reference's rounded floats (bit-identical on >= 99.99 % of
 *     results; the remainder are double-rounding ties of a device libm call).  Also selected by DJB200_PRECISION=bits in the
 *     environment.  evalp_is, Beckmann eval / pdf with per-pair params (PER_PAIR, LEAN texels), the table BRDFs, the fits and the maps always run at this
 *     level. */
enum { DJB200_PRECISION_REFERENCE_BITS = 0, DJB200_PRECISION_1E5 = 1 };
DJB200_API djb200_status djb200_set_precision(int mode);
DJB200_API int djb200_get_precision(void);

/* Debug switch for A/B tests: 1 = microfacet eval / evalp / pdf always run the mirrored-rounding kernel
 * (double sub-expressions literally as in the reference), 0 (default) = the lean FP32 kernels where they apply.
 * Both give the reference's rounded results; tests/test_gpu_parity.py compares them at full size. */
DJB200_API djb200_status djb200_debug_force_generic(int on);
/* Debug switch for A/B tests: 0 = Beckmann eval / evalp / pdf over several materials (BROADCAST) run on the plain lean
 * kernel instead of the warp-compacting one (csrc/kernels_mf.cu, mf_beck_compact_kernel); 1 (default) = compaction on.
 * The two produce the same floats (same functions on the same operands, scheduled on other lanes). */
DJB200_API djb200_status djb200_debug_beckmann_compaction(int on);
/* Profiling aid: SM clock values at the ten phase boundaries of material 0's CTA in the last isotropic fit launched by this
 * process (start | rows | matrix | iterations | normalise | NDF grid | sigma | Fresnel ratios | Fresnel sums + cdf | end). */
DJB200_API djb200_status djb200_debug_fit_phase_clocks(int64_t out_clocks[10]);
/* A/B switch of the isotropic fit's launch shape: 0 (default) = automatic -- batches that would leave most SMs idle run one
 * launch per phase with up to 8 CTAs per material, larger ones the whole fit in one launch with one CTA per material;
 * 1 = always the single launch; 3..8 = that many CTAs per material wherever the resolution allows.  Same results either way. */
DJB200_API djb200_status djb200_debug_fit_parts(int parts);
/* Test aid: the double functions of csrc/djb_dmath.cuh evaluated ON THE DEVICE over DEVICE arrays of doubles (the host build of the
 * same source is checked against libm by tests/cpp/dmath_check.cpp; this entry shows that the device build -- MUFU seeds, fused
 * operations -- gives the same accuracy).  fn: 0 exp_t, 1 log_t, 2 sqrt_d, 3 acos_d, 4 atan_t, 5 atan2_t(x, y), 6 sin, 7 cos (sincos_d),
 * 8 x / y (div_core), 9 pow_pos_t(x, y).  y_dev may be NULL for the one-argument functions. */
DJB200_API djb200_status djb200_debug_dmath(int fn, const double *x_dev, const double *y_dev, int64_t n, double *out_dev, void *stream);

/* ---- params factories (host side, dj_brdf.h:1355-1474) ---------------------------------- */
DJB200_API djb200_status djb200_params_standard(djb200_params *out);                           /* :1412 */
DJB200_API djb200_status djb200_params_isotropic(float a, djb200_params *out);                 /* :1417 */
DJB200_API djb200_status djb200_params_elliptic(float a1, float a2, float phi_a, djb200_params *out); /* :1422 */
DJB200_API djb200_status djb200_params_pdfparams(float ax, float ay, float rho, float tx_n, float ty_n,
                                                 djb200_params *out);                          /* :1428 */

/* ---- batched microfacet queries ---------------------------------------------------------- */
/* f_r: microfacet::eval, dj_brdf.h:1551-1555.  out_rgb: n_out x 3 floats. */
DJB200_API djb200_status djb200_microfacet_eval(const djb200_microfacet *mf, const djb200_params *params,
                                                int64_t n_params, int params_layout, const float *wi,
                                                const float *wo, int64_t n, float *out_rgb, int mem,
                                                void *stream);
/* f_r cos: microfacet::evalp, dj_brdf.h:1529-1547 */
DJB200_API djb200_status djb200_microfacet_evalp(const djb200_microfacet *mf, const djb200_params *params,
                                                 int64_t n_params, int params_layout, const float *wi,
                                                 const float *wo, int64_t n, float *out_rgb, int mem,
                                                 void *stream);
/* microfacet::pdf, dj_brdf.h:1713-1730.  out_pdf: n_out floats. */
DJB200_API djb200_status djb200_microfacet_pdf(const djb200_microfacet *mf, const djb200_params *params,
                                               int64_t n_params, int params_layout, const float *wi,
                                               const float *wo, int64_t n, float *out_pdf, int mem,
                                               void *stream);
/* microfacet::sample, dj_brdf.h:1669-1709.  u: n x 2 uniforms; out_wi: n_out x 3. */
DJB200_API djb200_status djb200_microfacet_sample(const djb200_microfacet *mf, const djb200_params *params,
                                                  int64_t n_params, int params_layout, const float *u,
                                                  const float *wo, int64_t n, float *out_wi, int mem,
                                                  void *stream);
/* microfacet::evalp_is, dj_brdf.h:1734-1765.  Any of the three outputs may be NULL.
 * out_wi is written as (0,0,0) where the reference leaves *i untouched (G <= 0). */
DJB200_API djb200_status djb200_microfacet_evalp_is(const djb200_microfacet *mf, const djb200_params *params,
                                                    int64_t n_params, int params_layout, const float *u,
                                                    const float *wo, int64_t n, float *out_weight_rgb,
                                                    float *out_wi, float *out_pdf, int mem, void *stream);

/* The public component queries of djb::microfacet (dj_brdf.h:258-272; implementations :1559-1665) under ONE params block
 * (NULL = params::standard()).  a / b / c are n x 3 arrays whose meaning depends on `what` (unused ones may be NULL):
 *   NDF (a = h) | GAF (a = h, b = i, c = o) | G1 (a = h, b = k) | SIGMA (a = k) | P22 (a = (x, y, -)) |
 *   VP22 (a = (x, y, -), b = k) | VNDF (a = h, b = k) | FRESNEL (a = (cos_theta_d, -, -)).
 * out: n floats; FRESNEL: n x 3 (rgb). */
typedef enum djb200_component {
	DJB200_COMP_NDF = 0, DJB200_COMP_GAF = 1, DJB200_COMP_G1 = 2, DJB200_COMP_SIGMA = 3, DJB200_COMP_P22 = 4,
	DJB200_COMP_VP22 = 5, DJB200_COMP_VNDF = 6, DJB200_COMP_FRESNEL = 7
} djb200_component;
DJB200_API djb200_status djb200_microfacet_component(const djb200_microfacet *mf, const djb200_params *params, int what,
                                                     const float *a, const float *b, const float *c, int64_t n, float *out,
                                                     int mem, void *stream);

/* ---- Rusinkiewicz frame (brdf::io_to_hd / hd_to_io, dj_brdf.h:771-793) ------------------ */
DJB200_API djb200_status djb200_io_to_hd(const float *wi, const float *wo, int64_t n, float *h, float *d,
                                         int mem, void *stream);
DJB200_API djb200_status djb200_hd_to_io(const float *h, const float *d, int64_t n, float *wi, float *wo,
                                         int mem, void *stream);

/* ---- MERL (djb::merl, dj_brdf.h:126-133, 893-1024) -------------------------------------- */
typedef struct djb200_merl djb200_merl; /* device-resident table, immutable */
/* samples: 3 planes (R, G, B) of 90*90*180 doubles exactly as in a MERL .binary file (host pointer) */
DJB200_API djb200_status djb200_merl_create(const double *samples, djb200_merl **out);
/* merl::merl(const char*), dj_brdf.h:963-983 */
DJB200_API djb200_status djb200_merl_load(const char *filename, djb200_merl **out);
DJB200_API djb200_status djb200_merl_destroy(djb200_merl *m);
/* merl::eval, dj_brdf.h:987-1024 */
DJB200_API djb200_status djb200_merl_eval(const djb200_merl *m, const float *wi, const float *wo, int64_t n,
                                          float *out_rgb, int mem, void *stream);
/* the cell index merl::eval forms at dj_brdf.h:997-1002 (phi_d + 180 theta_d + 16200 theta_h) */
DJB200_API djb200_status djb200_merl_index(const float *wi, const float *wo, int64_t n, int32_t *out_index,
                                           int mem, void *stream);

/* Property check of the filtered lookup (csrc/kernels_merl.cu) on DEVICE arrays: out_stats[0] = lookups the FP32
 * filter handed to the exact double path, [1] = lookups it certified with a cell different from the exact one
 * (must be 0), [2] = float bits of max |d_fast - d_exact|, [3] = half vectors not bit-identical to the reference's,
 * [4] = float bits of max |acos_poly - acos| over the sampled arguments. */
DJB200_API djb200_status djb200_debug_merl_filter_stats(const float *wi_dev, const float *wo_dev, int64_t n,
                                                        uint64_t out_stats[5], void *stream);

/* ---- UTIA (djb::utia, dj_brdf.h:136-146, 1029-1177) ------------------------------------- */
typedef struct djb200_utia djb200_utia;
/* raw_samples: 3*6*48*6*48 doubles as stored in a UTIA .bin file, before utia::normalize() */
DJB200_API djb200_status djb200_utia_create(const double *raw_samples, djb200_utia **out);
DJB200_API djb200_status djb200_utia_load(const char *filename, djb200_utia **out);
DJB200_API djb200_status djb200_utia_destroy(djb200_utia *u);
DJB200_API djb200_status djb200_utia_eval(const djb200_utia *u, const float *wi, const float *wo, int64_t n,
                                          float *out_rgb, int mem, void *stream);

/* ---- SGD and ABC analytic BRDFs (djb::sgd, djb::abc; dj_brdf.h:481-535, 3416-3499, 3608-3668) ------ *
 * One material = a block of double coefficients (the reference keeps them in static tables of 100 MERL fits each,
 * dj_brdf.h:3312-3413, 3505-3606, and looks them up by name in the constructor, :3436-3451, :3617-3631). */
enum { /* columns of djb200_sgd_data.ch[c]: the fields of sgd::data (dj_brdf.h:482-497) for colour channel c */
	DJB200_SGD_RHO_D = 0, DJB200_SGD_RHO_S, DJB200_SGD_ALPHA, DJB200_SGD_P, DJB200_SGD_F0, DJB200_SGD_F1,
	DJB200_SGD_KAP, DJB200_SGD_LAMBDA, DJB200_SGD_C, DJB200_SGD_K, DJB200_SGD_THETA0, DJB200_SGD_NCOEF
};
typedef struct djb200_sgd_data { double ch[3][DJB200_SGD_NCOEF]; } djb200_sgd_data;
typedef struct djb200_abc_data { double kD[3], A[3], B, C, ior; } djb200_abc_data; /* abc::data, dj_brdf.h:515-522 */
/* name lookup of the constructors; an unknown name fails with DJB200_ERR_INVALID_ARGUMENT and the reference's message
 * ("djb_error: No SGD parameters for ...", dj_brdf.h:3449).  SGD materials answer to two names (name, otherName). */
DJB200_API djb200_status djb200_sgd_preset(const char *name, djb200_sgd_data *out);
DJB200_API djb200_status djb200_abc_preset(const char *name, djb200_abc_data *out);
DJB200_API int32_t djb200_preset_count(void);                    /* 100 */
DJB200_API const char *djb200_sgd_preset_name(int32_t index);    /* NULL when out of range */
DJB200_API const char *djb200_abc_preset_name(int32_t index);
/* sgd::eval, dj_brdf.h:3454-3469; abc::eval, dj_brdf.h:3633-3647 (f_r; f_r cos is the base class's eval * i.z, :803) */
DJB200_API djb200_status djb200_sgd_eval(const djb200_sgd_data *material, const float *wi, const float *wo, int64_t n,
                                         float *out_rgb, int mem, void *stream);
DJB200_API djb200_status djb200_abc_eval(const djb200_abc_data *material, const float *wi, const float *wo, int64_t n,
                                         float *out_rgb, int mem, void *stream);

/* ---- LEAN / LEADR ------------------------------------------------------------------------ */
/* nmap2leanmap, utils/nmap2leanmap.cpp:18-54 (bias = 0) and nmap2leanmap_biased.cpp:23-63
 * (bias = 25).  nmap: planar uint8 [3][h][w]; lean1/lean2: planar float [4][h][w] (CImg layout). */
DJB200_API djb200_status djb200_nmap_to_leanmap(const uint8_t *nmap, int32_t w, int32_t h, float base_roughness,
                                                float bias, float *lean1, float *lean2, int mem, void *stream);

/* The two LEAN maps as the renderer consumes them: interleaved half-float RGBA with a mip pyramid.  Level 0 is what save_exr does
 * with the planar float image (utils/CImg.h:44940-44947, called at utils/nmap2leanmap.cpp:128-131: (half)(float), round to nearest
 * even); level L is the 2 x 2 box filter of level L - 1 -- linear filtering of the moments, which is what the renderer's texture
 * unit does before mitsuba/dj_beckmannconductor.cpp:295-314 reads them -- carried in float32 (((a + b) + (c + d)) * 0.25, edge
 * texels repeated at odd sizes) and rounded to half once per level.  leanmap: planar float32 [4][h][w]; out: all levels back to
 * back, level L = [h_L][w_L][4] halves with (w_L, h_L) = max(1, floor(w / 2^L)), ...; levels <= 0: the whole chain to 1 x 1. */
DJB200_API int32_t djb200_leanmap_mip_levels(int32_t w, int32_t h, int32_t levels);
DJB200_API int64_t djb200_leanmap_mip_texels(int32_t w, int32_t h, int32_t levels);
DJB200_API djb200_status djb200_leanmap_to_half_mips(const float *leanmap, int32_t w, int32_t h, int32_t levels,
                                                     uint16_t *out_rgba16f, int mem, void *stream);
/* dmap2nmap, utils/dmap2nmap.cpp:13-44: 8-bit displacement map [h][w] -> planar 8-bit normal map [3][h][w] (central
 * differences clamped at the borders, slopes scaled by (size / 2) * scale; the tool's default scale is 0.01, :69) */
DJB200_API djb200_status djb200_dmap_to_nmap(const uint8_t *dmap, int32_t w, int32_t h, float scale, uint8_t *nmap,
                                             int mem, void *stream);
/* beckmann::lrep_to_params, dj_brdf.h:1976-1990.  E: n x 5 moments (E1..E5). */
DJB200_API djb200_status djb200_lrep_to_params(const float *E, int64_t n, djb200_params *out, int mem,
                                               void *stream);
/* beckmann::params_to_lrep, dj_brdf.h:1965-1974 */
DJB200_API djb200_status djb200_params_to_lrep(const djb200_params *params, int64_t n, float *E, int mem,
                                               void *stream);
/* check_lean_maps, utils/nmap2leanmap.cpp:57-76: per-texel lrep_to_params over the two planar maps
 * (bias is subtracted from E1, E2 and bias^2 from E5 first, mitsuba/dj_beckmannconductor.cpp:295-314) */
DJB200_API djb200_status djb200_leanmap_to_params(const float *lean1, const float *lean2, int32_t w, int32_t h,
                                                  float bias, djb200_params *out, int mem, void *stream);

/* LEAN-filtered shading, fused: what mitsuba/dj_beckmannconductor.cpp does on the CPU before EVERY query (:283-314 in
 * eval, :338-366 in pdf, :379-400 in sample) -- base roughness ellipse -> params -> lrep; LEAN texel (E1..E5 as fetched
 * from leanmap1.rg / leanmap2.rgb, still biased) -> lrep, scaled by dmapscale; lrep sum -> params -- done per pair on the
 * device in front of evalp / pdf / evalp_is, so that neither the 48-byte params blocks nor the intermediate lreps ever
 * touch memory.  Per pair the kernel reads 20 B of moments (+ 12 B of roughness when it is textured) next to the
 * directions. */
typedef struct djb200_lean_shading {
	float bias;             /* BIAS of utils/nmap2leanmap_biased.cpp:11 (25 in the plugin, :300): E1 -= b, E2 -= b, E5 -= b*b */
	float dmap_scale;       /* lrep1 *= dmapscale (:312); >= 0 */
	int32_t lean_filtering; /* 1: LEAN filtering (:307), 0: naive MIP mapping, second moments rebuilt from the means (:309) */
	int32_t alpha_per_pair; /* 0: every pair uses alpha[3] below; 1: the `alpha` bulk array holds n x (alpha1, alpha2, alphaAngle) */
	float alpha[3];         /* (alpha1, alpha2, alphaAngle in radians) of params::elliptic (:290-294) */
} djb200_lean_shading;
/* only the parameter construction: out = n params blocks (what the plugin hands to evalp as user_param) */
DJB200_API djb200_status djb200_lean_shading_params(const djb200_lean_shading *cfg, const float *alpha, const float *E,
                                                    int64_t n, djb200_params *out, int mem, void *stream);
/* m_brdf->evalp(i, o, &params) of dj_beckmannconductor.cpp:316-319 with the construction fused in */
DJB200_API djb200_status djb200_lean_shading_evalp(const djb200_microfacet *mf, const djb200_lean_shading *cfg,
                                                   const float *alpha, const float *E, const float *wi, const float *wo,
                                                   int64_t n, float *out_rgb, int mem, void *stream);
/* m_brdf->pdf(i, o, &params), dj_beckmannconductor.cpp:362-364 */
DJB200_API djb200_status djb200_lean_shading_pdf(const djb200_microfacet *mf, const djb200_lean_shading *cfg,
                                                 const float *alpha, const float *E, const float *wi, const float *wo,
                                                 int64_t n, float *out_pdf, int mem, void *stream);
/* m_brdf->evalp_is(u1, u2, o, &i, &pdf, &params), dj_beckmannconductor.cpp:402-410; outputs as djb200_microfacet_evalp_is */
DJB200_API djb200_status djb200_lean_shading_evalp_is(const djb200_microfacet *mf, const djb200_lean_shading *cfg,
                                                      const float *alpha, const float *E, const float *u, const float *wo,
                                                      int64_t n, float *out_weight_rgb, float *out_wi, float *out_pdf,
                                                      int mem, void *stream);

/* ---- fits ("power iterations") ----------------------------------------------------------- */
/* What a fit reads from its source BRDF: a MERL/UTIA table on the device, or an analytic
 * microfacet BRDF.  (The reference takes any `const brdf&`, dj_brdf.h:401, 441.) */
typedef enum djb200_source_kind {
	DJB200_SOURCE_MERL = 0,
	DJB200_SOURCE_UTIA = 1,
	DJB200_SOURCE_MICROFACET = 2,
	DJB200_SOURCE_SGD = 3, /* what mitsuba/dj_sgd.cpp:29-30 fits */
	DJB200_SOURCE_ABC = 4  /* mitsuba/dj_abc.cpp:30-32 */
} djb200_source_kind;

typedef struct djb200_source {
	int32_t kind;
	const djb200_merl *merl;
	const djb200_utia *utia;
	djb200_microfacet microfacet; /* evaluated with params::standard(), as the reference's NULL user_param */
	const djb200_sgd_data *sgd;   /* host pointers, copied by the call */
	const djb200_abc_data *abc;
} djb200_source;

/* Result of djb::tabular::tabular + fit_*_parameters for one material (dj_brdf.h:2215-2236,
 * 3133-3184).  All arrays have `res` entries (fresnel: res x rgb). */
typedef struct djb200_tabular_fit {
	int32_t res;
	float *p22;     /* get_p22v()   */
	float *sigma;   /* get_sigmav() */
	float *cdf;     /* get_cdfv()   */
	float *qf;      /* get_qfv()    */
	float *fresnel; /* get_fresnel() spline points */
	float alpha_beckmann, alpha_ggx;
	float *residuals; /* optional, `iterations` floats: ||v_k/|v_k| - v_{k-1}/|v_{k-1}|||_2 (diagnostic,
	                     not in the reference, never fed back) */
} djb200_tabular_fit;

/* Batched isotropic fit: n_sources materials, host result structs with host arrays.
 * iterations = 4 reproduces the reference (km.eigenvector(4), dj_brdf.h:2518). */
DJB200_API djb200_status djb200_fit_tabular(const djb200_source *sources, int32_t n_sources, int32_t res,
                                            int32_t shadow, int32_t iterations, djb200_tabular_fit *results,
                                            void *stream);

/* The same batched fit with every result in ONE packed array (no per-material host objects, one device -> host copy; `mem` =
 * DJB200_MEM_DEVICE leaves the results on the device, stream ordered): the form a pipeline that fits hundreds of materials per
 * call wants (what examples/merl_params.cpp:55-60 does per file, batched).  Layout of `out`, n = n_sources:
 *   p22[n][res] | sigma[n][res] | cdf[n][res] | qf[n][res] | fresnel[n][res][3] | alpha[n][2] = (beckmann, ggx)
 * i.e. djb200_fit_tabular_packed_floats(n, res) = n (7 res + 2) floats; `residuals` (optional): [n][iterations]. */
DJB200_API int64_t djb200_fit_tabular_packed_floats(int32_t n_sources, int32_t res);
DJB200_API djb200_status djb200_fit_tabular_packed(const djb200_source *sources, int32_t n_sources, int32_t res,
                                                   int32_t shadow, int32_t iterations, float *out, float *residuals,
                                                   int mem, void *stream);

/* ---- djb::tabular as a BRDF (dj_brdf.h:394-425) ---------------------------------------------- *
 * The fitted tables evaluated / sampled like any other microfacet BRDF: the queries of dj_brdf.h:1529-1765 with
 * tabular::p22_radial / sigma_std_radial / qf_radial (:2151-2176), the fitted Fresnel spline, and -- because
 * tabular::supports_smith_vndf_sampling() is false (:413) -- normal-map sampling (:1806-1816) with its pdf
 * D cos(theta_h) / (4 i.h).  `params` as in the microfacet queries (the tabulated distribution is the
 * standard-space one; params stretch / shear / offset it). */
typedef struct djb200_tabular djb200_tabular;
DJB200_API djb200_status djb200_tabular_create(const djb200_tabular_fit *fit, int32_t shadow, djb200_tabular **out);
DJB200_API djb200_status djb200_tabular_destroy(djb200_tabular *t);
DJB200_API djb200_status djb200_tabular_eval(const djb200_tabular *t, const djb200_params *params, int64_t n_params,
                                             int params_layout, const float *wi, const float *wo, int64_t n,
                                             float *out_rgb, int mem, void *stream);
DJB200_API djb200_status djb200_tabular_evalp(const djb200_tabular *t, const djb200_params *params, int64_t n_params,
                                              int params_layout, const float *wi, const float *wo, int64_t n,
                                              float *out_rgb, int mem, void *stream);
DJB200_API djb200_status djb200_tabular_pdf(const djb200_tabular *t, const djb200_params *params, int64_t n_params,
                                            int params_layout, const float *wi, const float *wo, int64_t n,
                                            float *out_pdf, int mem, void *stream);
DJB200_API djb200_status djb200_tabular_sample(const djb200_tabular *t, const djb200_params *params, int64_t n_params,
                                               int params_layout, const float *u, const float *wo, int64_t n,
                                               float *out_wi, int mem, void *stream);
DJB200_API djb200_status djb200_tabular_evalp_is(const djb200_tabular *t, const djb200_params *params, int64_t n_params,
                                                 int params_layout, const float *u, const float *wo, int64_t n,
                                                 float *out_weight_rgb, float *out_wi, float *out_pdf, int mem,
                                                 void *stream);

/* The public scalar queries of djb::radial (dj_brdf.h:307-310) -- what tests/plot_qf.cpp and tests/plot_cdf.cpp of the
 * reference tabulate: x[k] is r^2 (P22), cos(theta_k) (SIGMA_STD), r (CDF) or u (QF).  `t` = a radial djb200_tabular handle
 * (then `ndf` is ignored; dj_brdf.h:2151-2176; CDF needs the fit's cdf table at djb200_tabular_create), or NULL for the
 * analytic family `ndf` (dj_brdf.h:1866-1889, 2056-2076). */
typedef enum djb200_radial_what {
	DJB200_RADIAL_P22 = 0, DJB200_RADIAL_SIGMA_STD = 1, DJB200_RADIAL_CDF = 2, DJB200_RADIAL_QF = 3
} djb200_radial_what;
DJB200_API djb200_status djb200_radial_query(int what, int ndf, const djb200_tabular *t, const float *x, int64_t n, float *out,
                                             int mem, void *stream);

/* Result of djb::tabular_anisotropic (eval tables) + fit_*_parameters (dj_brdf.h:2238-2273,
 * 3186-3307): p22/sigma are elev_res x azim_res, fresnel elev_res x rgb,
 * beckmann/ggx = (ax, ay, rho, tx_n, ty_n). */
typedef struct djb200_tabular_anisotropic_fit {
	int32_t elev_res, azim_res;
	float *p22;
	float *sigma;
	float *fresnel;
	float beckmann[5];
	float ggx[5];
	float *residuals;
} djb200_tabular_anisotropic_fit;

DJB200_API djb200_status djb200_fit_tabular_anisotropic(const djb200_source *sources, int32_t n_sources,
                                                        int32_t elev_res, int32_t azim_res, int32_t shadow,
                                                        int32_t iterations,
                                                        djb200_tabular_anisotropic_fit *results, void *stream);
/* djb::tabular_anisotropic as an evaluable / samplable BRDF: a djb200_tabular handle on the elevation x azimuth tables.
 * Every djb200_tabular_* query works on it: p22_std / sigma_std of dj_brdf.h:2178-2211, pdf = D cos / (4 i.h), and
 * normal-map sampling through the azimuth-marginal / elevation-conditional quantile tables (dj_brdf.h:2766-2837), which
 * this call builds on the device from the p22 table exactly as the constructor does (compute_pdf1 / cdf1 / qf1 / pdf2 /
 * cdf2 / qf2, dj_brdf.h:2266-2272, 2848-3103). */
DJB200_API djb200_status djb200_tabular_anisotropic_create(const djb200_tabular_anisotropic_fit *fit, int32_t shadow,
                                                           djb200_tabular **out);
/* The sampling tables behind tabular_anisotropic::pdf1 / cdf1 / qf1 / pdf2 / cdf2 / qf2 (dj_brdf.h:2766-2824), copied to
 * host arrays (any may be NULL): 1-D tables azim_res floats, 2-D tables elev_res * azim_res floats (azimuth-major rows of
 * elev_res).  counts = entries the reference's push_back loops produce for qf1 / qf2 (normally the full sizes). */
DJB200_API djb200_status djb200_tabular_anisotropic_sampling_tables(const djb200_tabular *t, float *pdf1, float *cdf1,
                                                                    float *qf1, float *pdf2, float *cdf2, float *qf2,
                                                                    int32_t counts[2]);

/* ---- the remaining public scalar members of the reference's classes ------------------------------------------------------ *
 * One call evaluates one member for n arguments; unused argument arrays are NULL.  Scalars in, one float out per argument,
 * except the sgd / abc members that return a vec3 (out: n x 3).
 *   djb200_quantile_query(ndf, what, ...)   beckmann / ggx (dj_brdf.h:366-369, 384-389):
 *       DJB200_MEMBER_QF1 (a = u) | DJB200_MEMBER_QF2_RADIAL (a = u, b = cos_theta_k, c = sin_theta_k) | DJB200_MEMBER_QF3_RADIAL (a = u, b = qf2)
 *   djb200_tabular_anisotropic_query(t, what, ...)   tabular_anisotropic (dj_brdf.h:450-455, 2766-2824):
 *       DJB200_MEMBER_PDF1 / CDF1 (a = phi) | DJB200_MEMBER_TQF1 (a = u1) | DJB200_MEMBER_PDF2 / CDF2 (a = theta, b = phi) | DJB200_MEMBER_TQF2 (a = u, b = phi)
 *   djb200_sgd_member / djb200_abc_member (dj_brdf.h:506-509, 531-533, 3471-3499, 3649-3668), vec3 arguments as n x 3:
 *       DJB200_MEMBER_NDF (a = h) | DJB200_MEMBER_GAF (a = h, b = i, c = o; abc returns a scalar: out n x 1) | DJB200_MEMBER_G1 (sgd: a = k)
 *       | DJB200_MEMBER_FRESNEL (a = cos_theta_d, n x 1)
 * (microfacet::qf2 / qf3 and radial's own qf2_radial / qf3_radial throw "Not Implemented" in the reference, dj_brdf.h:1783-1791,
 * 1848-1860: the facade does the same, no device entry.) */
enum {
	DJB200_MEMBER_QF1 = 0, DJB200_MEMBER_QF2_RADIAL = 1, DJB200_MEMBER_QF3_RADIAL = 2,
	DJB200_MEMBER_PDF1 = 10, DJB200_MEMBER_CDF1 = 11, DJB200_MEMBER_TQF1 = 12, DJB200_MEMBER_PDF2 = 13, DJB200_MEMBER_CDF2 = 14,
	DJB200_MEMBER_TQF2 = 15,
	DJB200_MEMBER_NDF = 20, DJB200_MEMBER_GAF = 21, DJB200_MEMBER_G1 = 22, DJB200_MEMBER_FRESNEL = 23
};
DJB200_API djb200_status djb200_quantile_query(int32_t ndf, int32_t what, const float *a, const float *b, const float *c, int64_t n,
                                               float *out, int mem, void *stream);
DJB200_API djb200_status djb200_tabular_anisotropic_query(const djb200_tabular *t, int32_t what, const float *a, const float *b,
                                                          int64_t n, float *out, int mem, void *stream);
DJB200_API djb200_status djb200_sgd_member(const djb200_sgd_data *m, int32_t what, const float *a, const float *b, const float *c,
                                           int64_t n, float *out, int mem, void *stream);
DJB200_API djb200_status djb200_abc_member(const djb200_abc_data *m, int32_t what, const float *a, const float *b, const float *c,
                                           int64_t n, float *out, int mem, void *stream);

/* ---- anisotropic fit, stage by stage ------------------------------------------------------- *
 * The same fit as djb200_fit_tabular_anisotropic(), split at the two places where a fit whose
 * n = (elev_res - 1) * azim_res matrix rows are sharded over several GPUs has to exchange data: the
 * iterate after every power iteration (matrix::transform, dj_brdf.h:2456-2465) and the projected-area
 * values (compute_sigma, dj_brdf.h:2388-2432).  Each rank owns the rows [row0, row1); the host layer
 * all-gathers `v_out` / `sigma_rows` between calls (NCCL through torch.distributed in
 * dj_brdf_b200/fit_sharded.py).  All pointers below are DEVICE pointers. */
typedef struct djb200_aniso_fit djb200_aniso_fit;
DJB200_API djb200_status djb200_aniso_fit_create(const djb200_source *source, int32_t elev_res, int32_t azim_res,
                                                 int32_t shadow, void *stream, djb200_aniso_fit **out);
DJB200_API djb200_status djb200_aniso_fit_destroy(djb200_aniso_fit *f);
/* number of unknowns n (length of the iterate) */
DJB200_API int64_t djb200_aniso_fit_size(const djb200_aniso_fit *f);
/* v_out[row0:row1] = (K v_in)[row0:row1]; v_in = NULL means the all-ones start vector (dj_brdf.h:2473-2474) */
DJB200_API djb200_status djb200_aniso_fit_matvec(djb200_aniso_fit *f, const double *v_in, double *v_out, int64_t row0,
                                                 int64_t row1, void *stream);
/* slope pdf table from the final iterate + normalize_p22 (dj_brdf.h:2570-2578, 2306-2338) */
DJB200_API djb200_status djb200_aniso_fit_set_iterate(djb200_aniso_fit *f, const double *v, void *stream);
/* sigma_rows[row0:row1]: projected area for view direction r = i2 * (elev_res - 1) + i1 */
DJB200_API djb200_status djb200_aniso_fit_sigma(djb200_aniso_fit *f, float *sigma_rows, int64_t row0, int64_t row1,
                                                void *stream);
/* sigma table, Fresnel table and the Beckmann / GGX parameter fits from the complete sigma_rows[n] */
DJB200_API djb200_status djb200_aniso_fit_finish(djb200_aniso_fit *f, const float *sigma_rows, void *stream);
/* copies the tables to host arrays of `result` (residuals are not touched) */
DJB200_API djb200_status djb200_aniso_fit_download(djb200_aniso_fit *f, djb200_tabular_anisotropic_fit *result,
                                                   void *stream);

/* ---- one fit spanning GPUs: the exchange step inside the library --------------------------- *
 * One process per GPU.  Rank 0 calls djb200_comm_unique_id and hands the DJB200_COMM_ID_BYTES bytes to the other ranks by
 * whatever the host program uses (torch.distributed, MPI, a file); every rank then creates its communicator on its current
 * device.  NCCL (libnccl.so.2) is loaded at run time: a process without it gets DJB200_ERR_UNSUPPORTED here and nothing else
 * in the library needs it. */
#define DJB200_COMM_ID_BYTES 128
typedef struct djb200_comm djb200_comm;
DJB200_API djb200_status djb200_comm_unique_id(void *out_id /* DJB200_COMM_ID_BYTES */);
DJB200_API djb200_status djb200_comm_create(const void *unique_id, int32_t world, int32_t rank, djb200_comm **out);
DJB200_API djb200_status djb200_comm_destroy(djb200_comm *c);
/* matrix::eigenvector(iterations) (dj_brdf.h:2467-2480) and every stage after it (:2238-2273) on a created fit, the n rows of the
 * operator split in contiguous blocks over the ranks of `comm` (NULL: this GPU alone): per iteration each rank computes its rows
 * and the iterate is all-gathered in place over NCCL (n doubles; 64 KB at 90 x 90), the projected-area rows likewise.  Every
 * rank ends with the complete result (djb200_aniso_fit_download).  residuals (optional, host): `iterations` floats.
 * timing_ms (optional, host): [0] = device time of the whole run, [1] = the part spent in the exchanges. */
DJB200_API djb200_status djb200_aniso_fit_run(djb200_aniso_fit *f, djb200_comm *comm, int32_t iterations, float *residuals,
                                              float *timing_ms, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DJB200_H */
