/* oracle/djb_oracle_fit.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Included by djb_oracle.c.
 *
 * CPU restatement of the reference's "power iteration" fits:
 *   djb::tabular              dj_brdf.h:2215-2236 (ctor), 2482-2522, 2277-2304, 2348-2386, 2583-2641,
 *                             2705-2762, 3133-3184
 *   djb::tabular_anisotropic  dj_brdf.h:2238-2273 (ctor, eval tables only), 2525-2579, 2306-2338,
 *                             2388-2432, 2643-2701, 3186-3307
 * with the float / double rounding points of the pinned reference build made explicit (F / D, see
 * djb_oracle.c) and the reference's summation orders kept, so every output is bit-comparable.
 *
 * One extension: `iterations` (the reference hard-codes km.eigenvector(4), :2518, :2568).
 */

/* fit input: brdf.eval(i, o) with the reference's NULL user_param (=> params::standard()) */
static v3 source_eval(const orc_source *s, v3 i, v3 o)
{
	switch (s->kind) {
	case ORC_SRC_MERL: return merl_eval1(s->table, i, o);
	case ORC_SRC_UTIA: return utia_eval1(s->table, i, o);
	case ORC_SRC_SGD: return sgd_eval1(s->sgd, i, o);
	case ORC_SRC_ABC: return abc_eval1(s->abc, i, o);
	default: return mf_eval(s->ndf, &s->F, s->shadow, &s->P, i, o);
	}
}

ORC_API void orc_source_eval(const orc_source *src, const float *wi, const float *wo, float *out3)
{
	v3_st(out3, 0, source_eval(src, v3_ld(wi, 0), v3_ld(wo, 0)));
}

static inline float v3_intensity(v3 v) /* vec3::intensity, :69 */
{
	return (0.2126f * v.x + 0.7152f * v.y) + 0.0722f * v.z;
}

/* matrix::eigenvector, :2467-2480: v <- ones; `iterations` x (v <- K v), sums in index order, all double.
 * K is stored so that out[a] = sum_b K[a * n + b] * v[b]  (km(b, a) of the reference, :2448-2449). */
static void power_iterations(const double *K, int n, int iterations, double *v, double *tmp)
{
	for (int a = 0; a < n; ++a) v[a] = 1.0;
	for (int it = 0; it < iterations; ++it) {
		for (int a = 0; a < n; ++a) {
			double acc = 0.0;
			const double *row = K + (size_t)a * n;
			for (int b = 0; b < n; ++b) acc += row[b] * v[b];
			tmp[a] = acc;
		}
		memcpy(v, tmp, sizeof(double) * n);
	}
}

typedef struct { const double *K; int n; const double *v; double *out; } matvec_ctx;
static void matvec_range(void *vctx, int64_t s, int64_t e)
{
	matvec_ctx *c = (matvec_ctx *)vctx;
	for (int64_t a = s; a < e; ++a) {
		double acc = 0.0;
		const double *row = c->K + (size_t)a * c->n;
		for (int b = 0; b < c->n; ++b) acc += row[b] * c->v[b];
		c->out[a] = acc;
	}
}

/* ============================================================================================
 * isotropic */
ORC_API void orc_fit_tabular(const orc_source *src, int res, int shadow, int iterations,
                             float *p22, float *sigma, float *cdf, float *qf, float *fresnel3, float *alpha2)
{
	const int cnt = res - 1;
	orc_params std_params;
	orc_params_elliptic(1.0f, 1.0f, 0.0f, &std_params); /* params::standard() */
	const double sqrt_half_pi = sqrt(ORC_PI * 0.5);

	/* ---- compute_p22_smith, :2482-2522 ---- */
	{
		float dtheta = F(sqrt_half_pi / D((float)cnt));
		double *K = (double *)malloc(sizeof(double) * (size_t)cnt * cnt);
		double *v = (double *)malloc(sizeof(double) * cnt), *tmp = (double *)malloc(sizeof(double) * cnt);
		const float dphi_h = F(ORC_PI / 180.0);
		/* the phi loop does not depend on (i, j): for (phi = 0; phi < 2 pi; phi += dphi) in float */
		float cosphi[400];
		int nphi = 0;
		for (float phi_h = 0.0f; D(phi_h) < 2.0 * ORC_PI; phi_h += dphi_h) cosphi[nphi++] = F(cos(D(phi_h)));
		for (int i = 0; i < cnt; ++i) {
			float t = (float)i / (float)cnt;
			float theta = F(D(t) * sqrt_half_pi);
			float theta_o = theta * theta;
			float cos_theta_o = F(cos(D(theta_o)));
			float tan_theta_o = F(tan(D(theta_o)));
			v3 dir = v3_spherical(theta_o, 0.0f);
			float fr_i = v3_intensity(source_eval(src, dir, dir));
			float kji_tmp = F((D(dtheta) * pow(D(cos_theta_o), D(6.0f))) * (8.0 * D(fr_i)));
			for (int j = 0; j < cnt; ++j) {
				float tj = (float)j / (float)cnt;
				float thj = F(D(tj) * sqrt_half_pi);
				float theta_h = thj * thj;
				float cos_theta_h = F(cos(D(theta_h)));
				float tan_theta_h = F(tan(D(theta_h)));
				float tan_product = tan_theta_h * tan_theta_o;
				float nint = 0.0f;
				for (int k = 0; k < nphi; ++k) nint += f_max(1.0f, tan_product * cosphi[k]);
				nint *= dphi_h;
				K[(size_t)i * cnt + j] = D(thj * kji_tmp * nint * tan_theta_h / (cos_theta_h * cos_theta_h));
			}
		}
		power_iterations(K, cnt, iterations, v, tmp);
		for (int i = 0; i < cnt; ++i) p22[i] = F(1e-2 * v[i]);
		p22[cnt] = 0.0f;
		free(K); free(v); free(tmp);
	}
	orc__set_tabular(p22, res, sigma, res);

	/* ---- normalize_p22, :2277-2304 ---- */
	{
		const int ntheta = 128;
		const float dphi = F(2.0 * ORC_PI);
		const float dtheta = F(ORC_PI / D((float)ntheta));
		float nint = 0.0f;
		for (int i = 0; i < ntheta; ++i) {
			float u = (float)i / (float)ntheta;
			float theta_h = F(D(u * u) * ORC_PI * 0.5);
			float r_h = F(tan(D(theta_h)));
			float cos_theta_h = F(cos(D(theta_h)));
			float p22_r = p22_radial(ORC_NDF_TABULAR, r_h * r_h);
			nint += (u * p22_r * r_h) / (cos_theta_h * cos_theta_h);
		}
		nint *= dtheta * dphi;
		nint = F(1.0 / D(nint));
		for (int i = 0; i < res; ++i) p22[i] *= nint;
	}

	/* ---- compute_sigma, :2348-2386 ---- */
	{
		const int ntheta = 90, nphi = 180;
		float dtheta = F(ORC_PI / D((float)ntheta));
		float dphi = F(2.0 * ORC_PI / D((float)nphi));
		for (int i = 0; i < cnt; ++i) {
			float t = (float)i / (float)cnt;
			float theta_k = F(D(t) * 0.5 * ORC_PI);
			float cos_theta_k = F(cos(D(theta_k)));
			float sin_theta_k = F(sin(D(theta_k)));
			float nint = 0.0f;
			for (int j2 = 0; j2 < nphi; ++j2) {
				float u_j = (float)j2 / (float)nphi;
				float phi_h = F(D(u_j) * 2.0 * ORC_PI);
				for (int j1 = 0; j1 < ntheta; ++j1) {
					float u_i = (float)j1 / (float)ntheta;
					float theta_h = F(D(u_i * u_i) * ORC_PI * 0.5);
					float sin_theta_h = F(sin(D(theta_h)));
					float kh = F(D(sin_theta_k * sin_theta_h) * cos(D(phi_h)) + D(cos_theta_k) * cos(D(theta_h)));
					nint += f_max(0.0f, kh) * mf_ndf(ORC_NDF_TABULAR, &std_params, v3_spherical(theta_h, phi_h))
					        * u_i * sin_theta_h;
				}
			}
			nint *= dtheta * dphi;
			sigma[i] = f_max(cos_theta_k, nint);
		}
		sigma[cnt] = sigma[cnt - 1];
	}

	/* ---- compute_fresnel, :2583-2641 ---- */
	{
		const float phi_d = F(ORC_PI * 0.5), phi_h = 0.0f;
		for (int i = 0; i < cnt; ++i) {
			float t = (float)i / (float)cnt;
			float theta_d = F(D(t) * ORC_PI * 0.5);
			v3 f = v3_make(0, 0, 0);
			int count[3] = {0, 0, 0};
			float theta_h = 0.0f;
			for (int j = 0; D(theta_h) < ORC_PI * 0.5 - D(theta_d); ++j) {
				float t1 = (float)j / (float)cnt;
				theta_h = F(D(t1 * t1) * ORC_PI * 0.5);
				if (D(theta_h) > ORC_PI * 0.5) continue;
				v3 dir_h = v3_spherical(theta_h, phi_h), dir_d = v3_spherical(theta_d, phi_d), dir_i, dir_o;
				hd_to_io(dir_h, dir_d, &dir_i, &dir_o);
				dir_i = v3_make(0, 0, 1);
				v3 fr1 = source_eval(src, dir_i, dir_o);
				v3 fr2 = mf_eval(ORC_NDF_TABULAR, NULL, shadow, &std_params, dir_i, dir_o);
				if (D(fr2.x) > 1e-4) { f.x += fr1.x / fr2.x; ++count[0]; }
				if (D(fr2.y) > 1e-4) { f.y += fr1.y / fr2.y; ++count[1]; }
				if (D(fr2.z) > 1e-4) { f.z += fr1.z / fr2.z; ++count[2]; }
			}
			fresnel3[3 * i + 0] = count[0] == 0 ? 1.0f : f_min(1.0f, f.x / (float)count[0]);
			fresnel3[3 * i + 1] = count[1] == 0 ? 1.0f : f_min(1.0f, f.y / (float)count[1]);
			fresnel3[3 * i + 2] = count[2] == 0 ? 1.0f : f_min(1.0f, f.z / (float)count[2]);
		}
		for (int c = 0; c < 3; ++c) fresnel3[3 * (res - 1) + c] = fresnel3[3 * (res - 2) + c];
	}

	/* ---- compute_cdf, :2705-2727 ---- */
	{
		float dtheta = F(ORC_PI / D((float)cnt));
		float nint = 0.0f;
		for (int i = 0; i < cnt; ++i) {
			float u = (float)i / (float)cnt;
			float theta_h = F(D(u * u) * ORC_PI * 0.5);
			float cos_theta_h = F(cos(D(theta_h)));
			float r_h = F(tan(D(theta_h)));
			float p22_r = p22_radial(ORC_NDF_TABULAR, r_h * r_h);
			nint += (u * r_h * p22_r) / (cos_theta_h * cos_theta_h);
			cdf[i] = F(D(nint * dtheta) * (2.0 * ORC_PI));
		}
		cdf[cnt] = 1.0f;
	}

	/* ---- compute_qf, :2731-2762 (entries the reference never pushes stay 0) ---- */
	{
		int qres = cnt * 8, j = 0, n = 0;
		for (int i = 0; i < res; ++i) qf[i] = 0.0f;
		qf[n++] = 0.0f;
		for (int i = 1; i < cnt; ++i) {
			float c = (float)i / (float)cnt;
			for (; j < qres; ++j) {
				float u = (float)j / (float)qres;
				float theta_h = F(D(u) * ORC_PI * 0.5);
				/* cdf_radial(tan(theta_h)), :2165-2170 */
				float r = F(tan(D(theta_h)));
				float uu = F(atan(D(r)) * D(2.0f) / D(F(ORC_PI)));
				if (uu < 0.0f) uu = 0.0f;
				float q = orc__spline_eval_f(cdf, res, F(sqrt(D(uu))));
				if (q >= c) { qf[n++] = u; break; }
			}
		}
		if (n < res) qf[n++] = 1.0f;
	}

	/* ---- fit_beckmann_parameters / fit_ggx_parameters, :3133-3184 ---- */
	{
		const int ntheta = 128;
		float dtheta = F(ORC_PI / D((float)ntheta));
		float nb = 0.0f, ng = 0.0f;
		for (int i = 0; i < ntheta; ++i) {
			float u = (float)i / (float)ntheta;
			float theta_h = F(D(u * u) * ORC_PI * 0.5);
			float cos_theta_h = F(cos(D(theta_h)));
			float r_h = F(tan(D(theta_h)));
			float r2 = r_h * r_h;
			float p22_r = p22_radial(ORC_NDF_TABULAR, r2);
			nb += (u * r2 * r_h * p22_r) / (cos_theta_h * cos_theta_h);
			ng += (u * r2 * p22_r) / (cos_theta_h * cos_theta_h);
		}
		nb = F(D(nb) * (D(dtheta) * ORC_PI));
		alpha2[0] = F(sqrt(2.0 * D(nb)));
		ng = F(D(ng) * (D(dtheta) * 4.0));
		alpha2[1] = ng;
	}
	orc__set_tabular(NULL, 0, NULL, 0);
}

/* ============================================================================================
 * anisotropic (eval tables + parameter fits) */
typedef struct {
	const orc_source *src;
	int w, h;
	float dtheta, dphi;
	double *K;
} aniso_build_ctx;

static void aniso_build_range(void *vctx, int64_t s, int64_t e)
{
	aniso_build_ctx *c = (aniso_build_ctx *)vctx;
	const int w = c->w, h = c->h, n = w * h;
	/* per-column factors, shared by every row: tan(theta) * max(0, m.o) / cos^2 needs (slope1, slope2, tan/den) */
	float *tanth = (float *)malloc(sizeof(float) * n), *s1 = (float *)malloc(sizeof(float) * n),
	      *s2 = (float *)malloc(sizeof(float) * n), *den = (float *)malloc(sizeof(float) * n);
	for (int j2 = 0; j2 < h; ++j2)
		for (int j1 = 0; j1 < w; ++j1) {
			float t1 = (float)j1 / (float)w, t2 = (float)j2 / (float)h;
			float theta = F(D(t1) * 0.5 * ORC_PI), phi = F(D(t2) * 2.0 * ORC_PI);
			float cos_theta = F(cos(D(theta))), tan_theta = F(tan(D(theta)));
			int k = j2 * w + j1;
			tanth[k] = tan_theta;
			s1[k] = F(D(-tan_theta) * cos(D(phi)));
			s2[k] = F(D(-tan_theta) * sin(D(phi)));
			den[k] = cos_theta * cos_theta;
		}
	for (int64_t row = s; row < e; ++row) {
		int i2 = (int)(row / w), i1 = (int)(row % w);
		float t1 = (float)i1 / (float)w, t2 = (float)i2 / (float)h;
		float theta = F(D(t1) * 0.5 * ORC_PI), phi = F(D(t2) * 2.0 * ORC_PI);
		float sin_theta = F(sin(D(theta)));
		float zo = F(cos(D(theta)));
		float xo = F(D(sin_theta) * cos(D(phi)));
		float yo = F(D(sin_theta) * sin(D(phi)));
		v3 dir = v3_spherical(theta, phi);
		float fr_i = v3_intensity(source_eval(c->src, dir, dir));
		float kji1 = F(D(c->dtheta * c->dphi) * (4.0 * D(fr_i) * pow(D(zo), D(5.0f))));
		double *out = c->K + (size_t)row * n;
		for (int k = 0; k < n; ++k) {
			float m_dot_o = zo - xo * s1[k] - yo * s2[k];
			float kji2 = tanth[k] * f_max(0.0f, m_dot_o) / den[k];
			out[k] = D(kji1 * kji2);
		}
	}
	free(tanth); free(s1); free(s2); free(den);
}

typedef struct {
	int w, h; /* w = elevation_res - 1 */
	float dtheta, dphi;
	const orc_params *P;
	const float *p22;
	float *sigma;
	int er, ar;
	const float *sigma_tab;
} aniso_sigma_ctx;

static void aniso_sigma_range(void *vctx, int64_t s, int64_t e)
{
	aniso_sigma_ctx *c = (aniso_sigma_ctx *)vctx;
	const int ntheta = 45, nphi = 90;
	const double sqrt_half_pi = sqrt(ORC_PI * 0.5);
	orc__set_tabular(c->p22, c->er, c->sigma_tab, c->ar); /* thread-local */
	/* ndf(vec3(theta_sqr, phi)) does not depend on k */
	float *nd = (float *)malloc(sizeof(float) * ntheta * nphi);
	for (int j2 = 0; j2 < nphi; ++j2) {
		float t = (float)j2 / (float)nphi;
		float phi = F(D(t) * 2.0 * ORC_PI);
		for (int j1 = 0; j1 < ntheta; ++j1) {
			float tt = (float)j1 / (float)ntheta;
			float theta = F(D(tt) * sqrt_half_pi);
			nd[j2 * ntheta + j1] = mf_ndf(ORC_NDF_TABULAR_ANISO, c->P, v3_spherical(theta * theta, phi));
		}
	}
	for (int64_t idx = s; idx < e; ++idx) {
		int i2 = (int)(idx / c->w), i1 = (int)(idx % c->w);
		float t2 = (float)i2 / (float)c->h;
		float phi_k = F(D(t2) * 2.0 * ORC_PI);
		float t1 = (float)i1 / (float)c->w;
		float theta_k = F(D(t1) * 0.5 * ORC_PI);
		float cos_theta_k = F(cos(D(theta_k)));
		float nint = 0.0f;
		for (int j2 = 0; j2 < nphi; ++j2) {
			float t = (float)j2 / (float)nphi;
			float phi = F(D(t) * 2.0 * ORC_PI);
			for (int j1 = 0; j1 < ntheta; ++j1) {
				float tt = (float)j1 / (float)ntheta;
				float theta = F(D(tt) * sqrt_half_pi);
				float theta_sqr = theta * theta;
				float sin_theta = F(sin(D(theta_sqr)));
				float m_dot_k = F(sin(D(theta_k)) * D(sin_theta) * cos(D(phi - phi_k))
				                  + D(cos_theta_k) * cos(D(theta_sqr)));
				float weight = theta * sin_theta;
				float masking = f_max(0.0f, m_dot_k) * nd[j2 * ntheta + j1];
				nint += weight * masking;
			}
		}
		nint = F(D(nint) * (2.0 * D(c->dtheta) * D(c->dphi)));
		c->sigma[i2 * c->er + i1] = f_max(cos_theta_k, nint);
	}
	free(nd);
	orc__set_tabular(NULL, 0, NULL, 0);
}

ORC_API void orc_fit_tabular_anisotropic(const orc_source *src, int elev_res, int azim_res, int shadow,
                                         int iterations, float *p22, float *sigma, float *fresnel3,
                                         float *beckmann5, float *ggx5, int nthreads)
{
	const int w = elev_res - 1, h = azim_res, n = w * h;
	const double sqrt_half_pi = sqrt(ORC_PI * 0.5);
	orc_params std_params;
	orc_params_elliptic(1.0f, 1.0f, 0.0f, &std_params);

	/* ---- compute_p22_smith, :2525-2579 ---- */
	{
		aniso_build_ctx bc;
		bc.src = src; bc.w = w; bc.h = h;
		bc.dtheta = F(sqrt_half_pi / D((float)w));
		bc.dphi = F(2.0 * ORC_PI / D((float)h));
		bc.K = (double *)malloc(sizeof(double) * (size_t)n * n);
		orc_parallel_ranges(n, nthreads, aniso_build_range, &bc);
		double *v = (double *)malloc(sizeof(double) * n), *tmp = (double *)malloc(sizeof(double) * n);
		for (int a = 0; a < n; ++a) v[a] = 1.0;
		for (int it = 0; it < iterations; ++it) {
			matvec_ctx mc = {bc.K, n, v, tmp};
			orc_parallel_ranges(n, nthreads, matvec_range, &mc);
			memcpy(v, tmp, sizeof(double) * n);
		}
		for (int j = 0; j < h; ++j) {
			for (int i = 0; i < w; ++i) p22[j * elev_res + i] = F(v[j * w + i]);
			p22[j * elev_res + w] = 0.0f;
		}
		free(bc.K); free(v); free(tmp);
	}
	orc__set_tabular(p22, elev_res, sigma, azim_res);

	/* ---- normalize_p22, :2306-2338 ---- */
	{
		const int ntheta = 128, nphi = 256;
		float dtheta = F(sqrt(0.5 * ORC_PI) / D((float)ntheta));
		float dphi = F(2.0 * ORC_PI / D((float)nphi));
		float k = 0.0f;
		for (int j = 0; j < nphi; ++j) {
			float u = (float)j / (float)nphi;
			float phi = F(D(u) * 2.0 * ORC_PI);
			for (int i = 0; i < ntheta; ++i) {
				float ui = (float)i / (float)ntheta;
				float theta = F(D(ui) * sqrt_half_pi);
				float theta_sqr = theta * theta;
				float c = F(cos(D(theta_sqr)));
				float pdf = orc__aniso_p22_theta_phi(theta_sqr, phi);
				float weight = F((D(theta) * tan(D(theta_sqr))) / D(c * c));
				k += weight * pdf;
			}
		}
		k = F(D(k) * (2.0 * D(dtheta) * D(dphi)));
		k = F(1.0 / D(k));
		for (int i = 0; i < elev_res * azim_res; ++i) p22[i] *= k;
	}

	/* ---- compute_sigma, :2388-2432 ---- */
	{
		aniso_sigma_ctx sc;
		sc.w = w; sc.h = h; sc.er = elev_res; sc.ar = azim_res;
		sc.dtheta = F(sqrt_half_pi / D((float)45));
		sc.dphi = F(2.0 * ORC_PI / D((float)90));
		sc.P = &std_params; sc.p22 = p22; sc.sigma = sigma; sc.sigma_tab = sigma;
		orc_parallel_ranges(n, nthreads, aniso_sigma_range, &sc);
		/* a range that ran inline on this thread (nthreads <= 1, or few rows) cleared this thread's table pointers */
		orc__set_tabular(p22, elev_res, sigma, azim_res);
		for (int i2 = 0; i2 < h; ++i2) sigma[i2 * elev_res + w] = sigma[i2 * elev_res + w - 1];
	}

	/* ---- compute_fresnel, :2643-2701 ---- */
	{
		const int res = elev_res, cnt = res - 1;
		const float phi_d = F(ORC_PI * 0.5), phi_h = 0.0f;
		for (int i = 0; i < cnt; ++i) {
			float t = (float)i / (float)cnt;
			float theta_d = F(D(t) * ORC_PI * 0.5);
			v3 f = v3_make(0, 0, 0);
			int count[3] = {0, 0, 0};
			float theta_h = 0.0f;
			for (int j = 0; D(theta_h) < ORC_PI * 0.5 - D(theta_d); ++j) {
				float t1 = (float)j / (float)cnt;
				theta_h = F(D(t1 * t1) * ORC_PI * 0.5);
				if (D(theta_h) > ORC_PI * 0.5) continue;
				v3 dir_h = v3_spherical(theta_h, phi_h), dir_d = v3_spherical(theta_d, phi_d), dir_i, dir_o;
				hd_to_io(dir_h, dir_d, &dir_i, &dir_o);
				dir_i = v3_make(0, 0, 1);
				v3 fr1 = source_eval(src, dir_i, dir_o);
				v3 fr2 = mf_eval(ORC_NDF_TABULAR_ANISO, NULL, shadow, &std_params, dir_i, dir_o);
				if (D(fr2.x) > 1e-4) { f.x += fr1.x / fr2.x; ++count[0]; }
				if (D(fr2.y) > 1e-4) { f.y += fr1.y / fr2.y; ++count[1]; }
				if (D(fr2.z) > 1e-4) { f.z += fr1.z / fr2.z; ++count[2]; }
			}
			fresnel3[3 * i + 0] = count[0] == 0 ? 1.0f : f_min(1.0f, f.x / (float)count[0]);
			fresnel3[3 * i + 1] = count[1] == 0 ? 1.0f : f_min(1.0f, f.y / (float)count[1]);
			fresnel3[3 * i + 2] = count[2] == 0 ? 1.0f : f_min(1.0f, f.z / (float)count[2]);
		}
		for (int c = 0; c < 3; ++c) fresnel3[3 * (res - 1) + c] = fresnel3[3 * (res - 2) + c];
	}

	/* ---- fit_beckmann_parameters / fit_ggx_parameters, :3186-3307 ---- */
	{
		const int ntheta = 128, nphi = 512;
		float dtheta = F(sqrt_half_pi / D((float)ntheta));
		float dphi = F(2.0 * ORC_PI / D((float)nphi));
		float nb[5] = {0, 0, 0, 0, 0}, ng[5] = {0, 0, 0, 0, 0};
		for (int j = 0; j < nphi; ++j) {
			float t = (float)j / (float)nphi;
			float phi = F(D(t) * 2.0 * ORC_PI);
			float cos_phi = F(cos(D(phi))), sin_phi = F(sin(D(phi)));
			float cos_phi_sqr = cos_phi * cos_phi, sin_phi_sqr = sin_phi * sin_phi;
			for (int i = 0; i < ntheta; ++i) {
				float t1 = (float)i / (float)ntheta;
				float theta = F(D(t1) * sqrt_half_pi);
				float theta_sqr = theta * theta;
				float pv = orc__aniso_p22_theta_phi(theta_sqr, phi);
				float tan_theta = F(tan(D(theta_sqr)));
				float cos_theta = F(cos(D(theta_sqr)));
				float tan_theta_sqr = tan_theta * tan_theta;
				float cos_theta_sqr = cos_theta * cos_theta;
				float tmp2 = theta * pv * tan_theta / cos_theta_sqr;
				float e1 = -tan_theta * cos_phi, e2 = -tan_theta * sin_phi;
				float e3 = tan_theta_sqr * cos_phi_sqr, e4 = tan_theta_sqr * sin_phi_sqr;
				float e5 = tan_theta_sqr * cos_phi * sin_phi;
				nb[0] += tmp2 * e1; nb[1] += tmp2 * e2; nb[2] += tmp2 * e3; nb[3] += tmp2 * e4; nb[4] += tmp2 * e5;
				float g3 = F(fabs(D(e1))), g4 = F(fabs(D(e2)));
				ng[0] += tmp2 * e1; ng[1] += tmp2 * e2; ng[2] += tmp2 * g3; ng[3] += tmp2 * g4; ng[4] += tmp2 * 0.0f;
			}
		}
		for (int i = 0; i < 5; ++i) {
			nb[i] = F(D(nb[i]) * (2.0 * D(dtheta) * D(dphi)));
			ng[i] = F(D(ng[i]) * (2.0 * D(dtheta) * D(dphi)));
		}
		{
			float mux = nb[0], muy = nb[1];
			float ax = F(sqrt(D(2.0f * (nb[2] - mux * mux))));
			float ay = F(sqrt(D(2.0f * (nb[3] - muy * muy))));
			float rho = F(2.0 * D(nb[4] - mux * muy) / D(ax * ay));
			beckmann5[0] = ax; beckmann5[1] = ay; beckmann5[2] = rho; beckmann5[3] = mux; beckmann5[4] = muy;
		}
		{
			float mux = ng[0], muy = ng[1];
			float ax = F(sqrt(D(ng[2] * ng[2] - mux * mux)));
			float ay = F(sqrt(D(ng[3] * ng[3] - muy * muy)));
			ggx5[0] = ax; ggx5[1] = ay; ggx5[2] = 0.0f; ggx5[3] = mux; ggx5[4] = muy;
		}
	}
	orc__set_tabular(NULL, 0, NULL, 0);
}

/* ---------------------------------------------------------------------------------------------
 * tabular_anisotropic's sampling tables: the marginal density of the azimuth, the conditional density of the elevation,
 * their running integrals and the inverted tables (compute_pdf1 / cdf1 / qf1 / pdf2 / cdf2 / qf2 and the two
 * normalisations, dj_brdf.h:2848-3103), from the final (normalised) p22 table.
 * All six outputs have room for azim_res (1-D) or elev_res * azim_res (2-D) floats and are zero-filled first; the
 * reference builds qf1 / qf2 with push_back inside a search loop that can run out without pushing, so their fill
 * counts are returned in counts[0] (qf1) and counts[1] (qf2; rows are packed back to back exactly as push_back leaves
 * them). */
static float spline_eval_repeat_f(const float *pts, int n, float u) /* spline::eval with uwrap_repeat, :1183-1218 */
{
	double ip;
	float frac = F(modf(D(u * (float)n - u), &ip));
	int i1 = (int)ip, i2 = (int)ip + 1;
	i1 = wrap_repeat(i1, n);
	i2 = wrap_repeat(i2, n);
	float p1 = pts[i1], p2 = pts[i2];
	return p1 + frac * (p2 - p1);
}
static float aniso_lookup1(const float *tab, int n, float phi) /* pdf1 / cdf1, :2766-2778 */
{
	return spline_eval_repeat_f(tab, n, F(D(phi) * 0.5 / ORC_PI));
}
static float aniso_lookup2(const float *tab, int w, int h, float theta, float phi, float beyond) /* pdf2 / cdf2, :2786-2812 */
{
	if (D(theta) >= 0.5 * ORC_PI) return beyond;
	float u1 = F(D(theta) * 2.0 / ORC_PI), u2 = F(D(phi) * 0.5 / ORC_PI);
	return orc__spline_eval2d_f(tab, w, h, u1, u2);
}
/* one term of the three elevation quadratures: (f * tan(theta)) / (cos_theta * cos_theta), accumulated in float */
static float quad_step(float nint, float f, float theta)
{
	float c = F(cos(D(theta)));
	return F(D(nint) + (D(f) * tan(D(theta))) / D(c * c));
}

ORC_API void orc_aniso_sampling_tables(const float *p22, int elev_res, int azim_res, float *pdf1, float *cdf1,
                                       float *qf1, float *pdf2, float *cdf2, float *qf2, int *counts)
{
	const int er = elev_res, ar = azim_res;
	memset(pdf1, 0, sizeof(float) * ar); memset(cdf1, 0, sizeof(float) * ar); memset(qf1, 0, sizeof(float) * ar);
	memset(pdf2, 0, sizeof(float) * er * ar); memset(cdf2, 0, sizeof(float) * er * ar);
	memset(qf2, 0, sizeof(float) * er * ar);
	orc__set_tabular(p22, er, NULL, ar);
	/* compute_pdf1, :2848-2874 */
	{
		const int ntheta = 256;
		float dtheta = F(0.5 * ORC_PI / D((float)ntheta));
		for (int i = 0; i < ar; ++i) {
			float phi = F(D((float)i / (float)ar) * 2.0 * ORC_PI), nint = 0.0f;
			for (int j = 0; j < ntheta; ++j) {
				float theta = F(D((float)j / (float)ntheta) * 0.5 * ORC_PI);
				nint = quad_step(nint, orc__aniso_p22_theta_phi(theta, phi), theta);
			}
			pdf1[i] = nint * dtheta;
		}
	}
	/* normalize_pdf1, :3029-3051 */
	{
		const int cnt = 512;
		float dphi = F(2.0 * ORC_PI / D((float)cnt)), nint = 0.0f;
		for (int i = 0; i < cnt; ++i) nint += aniso_lookup1(pdf1, ar, F(D((float)i / (float)cnt) * 2.0 * ORC_PI));
		nint *= dphi;
		float k = F(1.0 / D(nint));
		for (int i = 0; i < ar; ++i) pdf1[i] *= k;
	}
	/* compute_cdf1, :2878-2900 */
	{
		int cnt = ar - 1;
		float dphi = F(2.0 * ORC_PI / D((float)cnt)), nint = 0.0f;
		cdf1[0] = 0.0f;
		for (int i = 1; i < cnt; ++i) {
			nint += aniso_lookup1(pdf1, ar, F(D((float)i / (float)cnt) * 2.0 * ORC_PI));
			cdf1[i] = nint * dphi;
		}
		cdf1[cnt] = 1.0f;
	}
	/* compute_qf1, :2904-2935 */
	int n1 = 0;
	{
		int cnt = ar - 1, res = cnt * 8, j = 0;
		qf1[n1++] = 0.0f;
		for (int i = 1; i < cnt; ++i) {
			float cdf = (float)i / (float)cnt;
			for (; j < res; ++j) {
				float u = (float)j / (float)res;
				if (aniso_lookup1(cdf1, ar, F(D(u) * 2.0 * ORC_PI)) >= cdf) { qf1[n1++] = u; break; }
			}
		}
		qf1[n1++] = 1.0f;
	}
	/* compute_pdf2, :2944-2969 */
	{
		int ntheta = er - 1;
		for (int i = 0; i < ar; ++i) {
			float phi = F(D((float)i / (float)ar) * 2.0 * ORC_PI);
			for (int j = 0; j < ntheta; ++j) {
				float theta = F(D((float)j / (float)ntheta) * 0.5 * ORC_PI);
				pdf2[i * er + j] = orc__aniso_p22_theta_phi(theta, phi) / aniso_lookup1(pdf1, ar, phi);
			}
			pdf2[i * er + ntheta] = 0.0f;
		}
	}
	/* normalize_pdf2, :3055-3088: every constant is computed on the unscaled table, then the rows are scaled */
	{
		const int ntheta = 256;
		float dtheta = F(0.5 * ORC_PI / D((float)ntheta));
		float *k = (float *)malloc(sizeof(float) * ar);
		for (int j = 0; j < ar; ++j) {
			float phi = F(D((float)j / (float)ar) * 2.0 * ORC_PI), nint = 0.0f;
			for (int i = 0; i < ntheta; ++i) {
				float theta = F(D((float)i / (float)ntheta) * 0.5 * ORC_PI);
				nint = quad_step(nint, aniso_lookup2(pdf2, er, ar, theta, phi, 0.0f), theta);
			}
			nint *= dtheta;
			k[j] = F(1.0 / D(nint));
		}
		for (int j = 0; j < ar; ++j)
			for (int i = 0; i < er; ++i) pdf2[i + er * j] *= k[j];
		free(k);
	}
	/* compute_cdf2, :2973-3000 */
	{
		int ntheta = er - 1;
		float dtheta = F(0.5 * ORC_PI / D((float)ntheta));
		for (int i = 0; i < ar; ++i) {
			float phi = F(D((float)i / (float)ar) * 2.0 * ORC_PI), nint = 0.0f;
			for (int j = 0; j < ntheta; ++j) {
				float theta = F(D((float)j / (float)ntheta) * 0.5 * ORC_PI);
				nint = quad_step(nint, aniso_lookup2(pdf2, er, ar, theta, phi, 0.0f), theta);
				cdf2[i * er + j] = nint * dtheta;
			}
			cdf2[i * er + ntheta] = 1.0f;
		}
	}
	/* compute_qf2, :3004-3037 */
	int n2 = 0;
	{
		int ntheta = er - 1, res = ntheta * 8;
		for (int kk = 0; kk < ar; ++kk) {
			float phi = F(D((float)kk / (float)ar) * 2.0 * ORC_PI);
			int j = 0;
			qf2[n2++] = 0.0f;
			for (int i = 1; i < ntheta; ++i) {
				float cdf = (float)i / (float)ntheta;
				for (; j < res; ++j) {
					float u = (float)j / (float)res;
					float theta = F(D(u) * 0.5 * ORC_PI);
					if (aniso_lookup2(cdf2, er, ar, theta, phi, 1.0f) >= cdf) { qf2[n2++] = u; break; }
				}
			}
			qf2[n2++] = 1.0f;
		}
	}
	if (counts) { counts[0] = n1; counts[1] = n2; }
	orc__set_tabular(NULL, 0, NULL, 0);
}
