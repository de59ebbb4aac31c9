/* oracle/djb_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C11) of the reference's numeric hot paths, written from the
 * behaviour of /root/reference/dj_brdf.h and utils/nmap2leanmap*.cpp.  Every function cites
 * the reference lines it restates.  All float/double rounding points of the *pinned* reference
 * build (SURVEY.md section 0 finding 2 and appendix A) are made explicit with casts:
 * inside `namespace djb` the unqualified libm calls bind to the C double functions, double
 * literals promote their sub-expression to double, and nothing is contracted into FMAs.
 *
 * Pinning: tests/test_oracle_vs_reference.py checks this port bit-for-bit against
 * oracle/_ref/libdjbref.so (the unmodified reference compiled in place) whenever that library is
 * present, and against the committed vectors in tests/golden/ (generated from the reference by
 * tests/golden/make_golden.py) everywhere else.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (dj_brdf_b200/, include/) never does.
 */
#ifndef DJB_ORACLE_H
#define DJB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* layout-identical to djb::microfacet::params (dj_brdf.h:238-242), 48 bytes */
typedef struct orc_params {
	float n[3];          /* mean normal */
	float a1, a2, phi_a; /* ellipse */
	float ax, ay;        /* scales */
	float rho, srho;     /* correlation, sqrt(1 - rho^2) */
	float tx, ty;        /* location */
} orc_params;

enum { ORC_NDF_BECKMANN = 0, ORC_NDF_GGX = 1 };
enum { ORC_F_IDEAL = 0, ORC_F_SCHLICK = 1, ORC_F_UNPOLARIZED = 2, ORC_F_SGD = 3, ORC_F_SPLINE = 4 };

typedef struct orc_fresnel {
	int kind;
	float v[6];       /* schlick: f0[3]; unpolarized: ior[3]; sgd: f0[3], f1[3] */
	const float *pts; /* spline: npts * 3 floats */
	int npts;
} orc_fresnel;

/* params factories (dj_brdf.h:1355-1474) */
void orc_params_elliptic(float a1, float a2, float phi_a, orc_params *out);
void orc_params_pdfparams(float ax, float ay, float rho, float tx, float ty, orc_params *out);

/* microfacet batch queries (dj_brdf.h:1529-1765); params==NULL => params::standard() */
void orc_microfacet_eval(int ndf, const orc_fresnel *F, int shadow, const orc_params *P,
                         const float *wi, const float *wo, int64_t n, float *out3, int nthreads);
void orc_microfacet_evalp(int ndf, const orc_fresnel *F, int shadow, const orc_params *P,
                          const float *wi, const float *wo, int64_t n, float *out3, int nthreads);
void orc_microfacet_pdf(int ndf, const orc_fresnel *F, int shadow, const orc_params *P,
                        const float *wi, const float *wo, int64_t n, float *out1, int nthreads);
void orc_microfacet_sample(int ndf, const orc_fresnel *F, int shadow, const orc_params *P,
                           const float *u2, const float *wo, int64_t n, float *out3, int nthreads);
void orc_microfacet_evalp_is(int ndf, const orc_fresnel *F, int shadow, const orc_params *P,
                             const float *u2, const float *wo, int64_t n,
                             float *out_w3, float *out_i3, float *out_pdf, int nthreads);

/* djb::microfacet's public component queries (dj_brdf.h:258-272): what = 0 ndf, 1 gaf, 2 g1, 3 sigma, 4 p22, 5 vp22, 6 vndf,
 * 7 fresnel (rgb); argument arrays as in include/djb200.h's djb200_microfacet_component */
void orc_microfacet_component(int ndf, const orc_fresnel *F, int shadow, const orc_params *P, int what, const float *a,
                              const float *b, const float *c, int64_t n, float *out);

/* Rusinkiewicz transforms (dj_brdf.h:771-793) */
void orc_io_to_hd(const float *wi, const float *wo, int64_t n, float *h3, float *d3);
void orc_hd_to_io(const float *h3, const float *d3, int64_t n, float *wi, float *wo);

/* MERL (dj_brdf.h:906-1024): table = 3 planes of 90*90*180 doubles as in the .binary file */
void orc_merl_index(const float *wi, const float *wo, int64_t n, int32_t *idx, int nthreads);
void orc_merl_eval(const double *table, const float *wi, const float *wo, int64_t n, float *out3,
                   int nthreads);

/* UTIA (dj_brdf.h:1039-1177): table = 3*6*48*6*48 doubles already normalised by orc_utia_normalize */
void orc_utia_normalize(double *table);
void orc_utia_eval(const double *table, const float *wi, const float *wo, int64_t n, float *out3,
                   int nthreads);

/* djb::sgd / djb::abc (dj_brdf.h:481-535, 3416-3499, 3608-3668): one material's coefficients, field-major as in the
 * reference's static tables */
typedef struct orc_sgd {
	double rhoD[3], rhoS[3], alpha[3], p[3], f0[3], f1[3], kap[3], lambda[3], c[3], k[3], theta0[3];
} orc_sgd;
typedef struct orc_abc { double kD[3], A[3], B, C, ior; } orc_abc;
void orc_sgd_eval(const orc_sgd *m, const float *wi, const float *wo, int64_t n, float *out3, int nthreads);
void orc_abc_eval(const orc_abc *m, const float *wi, const float *wo, int64_t n, float *out3, int nthreads);

/* LEAN (dj_brdf.h:1965-1990; utils/nmap2leanmap.cpp:18-54; nmap2leanmap_biased.cpp:23-63) */
void orc_lrep_to_params(const float *E5, int64_t n, orc_params *out);
void orc_params_to_lrep(const orc_params *p, int64_t n, float *E5);
/* per-shading-point params of the LEAN-filtering plugin (mitsuba/dj_beckmannconductor.cpp:283-314) */
void orc_lean_shading_params(float bias, float dmap_scale, int lean_filtering, int alpha_per_pair, const float *alpha,
                             const float *E5, int64_t n, orc_params *out);
void orc_nmap2leanmap(const uint8_t *nmap_planar_rgb, int w, int h, float base_roughness, float bias,
                      float *lean1_planar_rgba, float *lean2_planar_rgba);

/* djb::erf (dj_brdf.h:667-688) */
void orc_erf(const float *x, int64_t n, float *out);

/* djb::radial's public scalar queries (dj_brdf.h:307-310); family 0 beckmann, 1 ggx, 2 tabular (tables of length res) */
void orc_radial_query(int family, int what, const float *p22, const float *sigma, const float *qf, const float *cdf, int res,
                      const float *x, int64_t n, float *out);

/* dmap2nmap (utils/dmap2nmap.cpp:13-44): uint8 [h][w] displacement map -> planar uint8 [3][h][w] normal map */
void orc_dmap2nmap(const uint8_t *dmap, int w, int h, float scale, uint8_t *nmap_planar_rgb);

/* ---- fits: djb_oracle_fit.c ------------------------------------------------------------- */
/* generic BRDF handle used as the fit input */
enum { ORC_SRC_MICROFACET = 0, ORC_SRC_MERL = 1, ORC_SRC_UTIA = 2, ORC_SRC_SGD = 3, ORC_SRC_ABC = 4 };
typedef struct orc_source {
	int kind;
	/* microfacet */
	int ndf;
	orc_fresnel F;
	int shadow;
	orc_params P; /* user_param is NULL in the reference's fit calls => P must be standard() */
	/* merl / utia */
	const double *table;
	/* sgd / abc */
	const orc_sgd *sgd;
	const orc_abc *abc;
} orc_source;
void orc_source_eval(const orc_source *src, const float *wi, const float *wo, float *out3);

/* isotropic fit: tabular::tabular + fit_*_parameters (dj_brdf.h:2215-2236, 3133-3184).
 * iterations = 4 reproduces the reference (dj_brdf.h:2518).
 * outputs: p22[res], sigma[res], cdf[res], qf[res], fresnel[3*res], alpha[2] = beckmann, ggx */
void orc_fit_tabular(const orc_source *src, int res, int shadow, int iterations,
                     float *p22, float *sigma, float *cdf, float *qf, float *fresnel3, float *alpha2);

/* djb::tabular as a BRDF: eval (op 0), evalp (1), pdf (2), sample (3), evalp_is (4) on fitted tables
 * (dj_brdf.h:1529-1765 with tabular::p22_radial / sigma_std_radial / qf_radial :2151-2176 and normal-map sampling
 * :1806-1816); F = the fitted Fresnel spline (or any other term), P = NULL for params::standard() */
void orc_tabular_query(int op, const float *p22, const float *sigma, const float *qf, int res,
                       const orc_fresnel *F, int shadow, const orc_params *P, const float *a, const float *b,
                       int64_t n, float *o0, float *o1, float *o2, int nthreads);

/* anisotropic fit: tabular_anisotropic (dj_brdf.h:2238-2273, 3186-3307), eval tables only.
 * outputs: p22[er*ar], sigma[er*ar], fresnel[3*er], beckmann5/ggx5 = ax, ay, rho, tx, ty */
void orc_fit_tabular_anisotropic(const orc_source *src, int elev_res, int azim_res, int shadow,
                                 int iterations, float *p22, float *sigma, float *fresnel3,
                                 float *beckmann5, float *ggx5, int nthreads);

/* djb::tabular_anisotropic as an evaluable BRDF: op 0 eval, 1 evalp, 2 pdf (dj_brdf.h:1529-1555, 1724-1725 with
 * tabular_anisotropic::p22_std / sigma_std :2178-2211) */
void orc_tabular_aniso_query(int op, const float *p22, const float *sigma, int elev_res, int azim_res,
                             const orc_fresnel *F, int shadow, const orc_params *P, const float *wi,
                             const float *wo, int64_t n, float *o0, int nthreads);

/* tabular_anisotropic's sampling tables (dj_brdf.h:2848-3103) from its normalised p22 table; 1-D outputs hold azim_res
 * floats, 2-D ones elev_res * azim_res; counts[0] / counts[1] = entries the reference's push_back loops produce for
 * qf1 / qf2 (azim_res and elev_res * azim_res unless a search runs out) */
void orc_aniso_sampling_tables(const float *p22, int elev_res, int azim_res, float *pdf1, float *cdf1, float *qf1,
                               float *pdf2, float *cdf2, float *qf2, int *counts);
/* djb::tabular_anisotropic sample (op 3) / evalp_is (op 4): normal-map sampling through qf1 / qf2 (:2826-2837) */
void orc_tabular_aniso_sample_query(int op, const float *p22, const float *sigma, const float *qf1, int n_qf1,
                                    const float *qf2, int elev_res, int azim_res, const orc_fresnel *F, int shadow,
                                    const orc_params *P, const float *u, const float *wo, int64_t n, float *o0,
                                    float *o1, float *o2, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
