// kernels_tabular.cu -- djb::tabular as an evaluable / samplable BRDF on the device:
// the microfacet queries of dj_brdf.h:1529-1765 on the radial tables a fit produced (tabular::p22_radial,
// sigma_std_radial, qf_radial, dj_brdf.h:2151-2176), with the fitted Fresnel spline.  tabular does not support Smith
// VNDF sampling (dj_brdf.h:413): sample() draws normals from the tabulated slope quantile function
// (radial::sample_vp22_std_nmap, dj_brdf.h:1806-1816) and pdf() is the matching D cos / (4 i.h).
//
// Mirrored-rounding tier (djb_device.cuh / djb_fit.cuh): double sub-expressions where the reference has them.
#include "djb_fit.cuh"
#include "djb_internal.h"

namespace djb200 {

constexpr int TQ_THREADS = 256;
constexpr int TQ_MAX_SMEM_PARAMS = 64;

struct TabQueryArgs {
	const float *tables; // p22[res] | sigma[res] | qf[res] | fresnel[res][3]
	int res, shadow;
	const Params *params;
	int n_params;
	const float *a, *b;
	long long n, out_stride;
	float *out0, *out1, *out2;
};

template <class TAB>
struct TabBrdfT {
	TAB t;
	const float *qf; // radial quantile table (isotropic only)
	const float *qf1, *qf2; // azimuth / elevation quantile tables (anisotropic only)
	int n_qf1;
	FresnelDev fr;
	bool shadow;
};
typedef TabBrdfT<TabIso> TabBrdf;

template <class TB>
DJB_DEV float tabq_gaf(const TB &B, const Params &p, V3 i, V3 o)
{
	float g1o = tab_g1(B.t, p, o);
	if (B.shadow) {
		float g1i = tab_g1(B.t, p, i);
		float t = g1i * g1o;
		return t > 0.0f ? t / (g1i + g1o - t) : 0.0f;
	}
	return g1o;
}

// microfacet::evalp, dj_brdf.h:1529-1547
template <class TB>
DJB_DEV V3 tabq_evalp(const TB &B, const Params &p, V3 i, V3 o)
{
	V3 h = normalize(i + o);
	float G = tabq_gaf(B, p, i, o);
	if (G > 0.0f) {
		float cd = sat_ref(dot(o, h));
		V3 Fr = fresnel_eval<FK_SPLINE>(B.fr, cd);
		float Dn = tab_ndf(B.t, p, h);
		return scale(__fdiv_rn(Dn * G, 4.0f * o.z), Fr); // == (float)((double)(Dn * G) / (4.0 * (double)o.z)): one IEEE op
	}
	return mk(0.f, 0.f, 0.f);
}

// microfacet::pdf without Smith VNDF sampling, dj_brdf.h:1724-1725
template <class TB>
DJB_DEV float tabq_pdf(const TB &B, const Params &p, V3 i, V3 o)
{
	V3 h = normalize(i + o);
	float G = tabq_gaf(B, p, i, o);
	if (G > 0.0f) return __fdiv_rn(h.z * tab_ndf(B.t, p, h), 4.0f * dot(i, h)); // == the quotient in double, rounded once more
	return 0.0f;
}

// spline::eval<float_t> with uwrap_repeat, dj_brdf.h:1183-1218
DJB_DEV float spline_repeat_f(const float *pts, int n, float u)
{
	const float x = u * (float)n - u;
	const float ip = truncf(x), frac = x - ip; // == (float)modf((double)x, &ip): exact in both precisions
	int i1 = (int)ip, i2 = (int)ip + 1;
	i1 = wrap_repeat(i1, n);
	i2 = wrap_repeat(i2, n);
	const float p1 = pts[i1], p2 = pts[i2];
	return p1 + frac * (p2 - p1);
}

// (float)tan((double)x) as sin / cos of djb_dmath.cuh (<= 3 ulp of double before the rounding to float)
DJB_DEV float tan_f(float x)
{
	double sn, cs;
	sincos_d((double)x, &sn, &cs);
	const double a = fabs(cs);
	if (!(a >= 1e-30 && fabs((double)x) <= 1e5)) return (float)tan((double)x);
	return (float)div_core(sn, cs);
}
// standard slopes of a normal drawn from the tabulated distribution
DJB_DEV void tabq_std_slopes(const TabBrdfT<TabIso> &B, float u1, float u2, float &txm, float &tym)
{
	// radial::sample_vp22_std_nmap, dj_brdf.h:1806-1816
	float phi_h = (float)((double)u1 * DJB_PI * 2.0);
	float q = spline_f(B.qf, B.t.n, u2);                        // tabular::qf_radial, :2172-2176
	float r_h = tan_f(q * (float)DJB_PI / 2.0f);
	double sp, cp;
	sincos_d((double)phi_h, &sp, &cp);
	txm = (float)((double)r_h * cp);
	tym = (float)((double)r_h * sp);
}
DJB_DEV void tabq_std_slopes(const TabBrdfT<TabAniso> &B, float u1, float u2, float &txm, float &tym)
{
	// tabular_anisotropic::sample_vp22_std_nmap with qf1 / qf2, dj_brdf.h:2826-2837, 2780-2784, 2814-2824
	const float phi = (float)((double)spline_f(B.qf1, B.n_qf1, u1) * 2.0 * DJB_PI);
	const float uphi = (float)((double)phi / (2.0 * DJB_PI));
	const float theta = (float)((double)spline2d_f(B.qf2, B.t.w, B.t.h, u2, uphi) * 0.5 * DJB_PI);
	const float tan_theta = tan_f(theta);
	double sp, cp;
	sincos_d((double)phi, &sp, &cp);
	txm = (float)((double)(-tan_theta) * cp);
	tym = (float)((double)(-tan_theta) * sp);
}

// microfacet::sample (dj_brdf.h:1669-1709) through sample_vp22_std_nmap
template <class TB>
DJB_DEV V3 tabq_sample(const TB &B, const Params &p, float u1, float u2, V3 o)
{
	u1 = sat_ref(u1) * 0.99998f + 0.00001f;
	u2 = sat_ref(u2) * 0.99998f + 0.00001f;
	float a = o.x * p.ax + o.y * p.ay * p.rho;
	float b = o.y * p.ay * p.srho;
	float c = o.z - o.x * p.tx - o.y * p.ty;
	V3 os = normalize(mk(a, b, c));
	if (os.z > 0.0f) {
		float txm, tym;
		tabq_std_slopes(B, u1, u2, txm, tym);
		float txh = p.ax * txm + p.tx;
		float chol = p.rho * txm + p.srho * tym;
		float tyh = p.ay * chol + p.ty;
		V3 h = normalize(mk(-txh, -tyh, 1.0f));
		float k = (float)(2.0 * (double)dot(o, h));
		return scale(k, h) - o;
	}
	return mk(0.f, 0.f, 1.f);
}

// microfacet::evalp_is without Smith VNDF sampling, dj_brdf.h:1734-1765
template <class TB>
DJB_DEV V3 tabq_evalp_is(const TB &B, const Params &p, float u1, float u2, V3 o, V3 &i_out, float &pdf_out)
{
	V3 i = tabq_sample(B, p, u1, u2, o);
	V3 h = normalize(i + o);
	float G = tabq_gaf(B, p, i, o);
	pdf_out = 0.0f;
	i_out = mk(0.f, 0.f, 0.f);
	if (G > 0.0f) {
		float cd = sat_ref(dot(o, h));
		i_out = i;
		float pdf = __fdiv_rn(h.z * tab_ndf(B.t, p, h), 4.0f * cd);
		pdf_out = pdf;
		return scale(rcp_via_double(pdf), tabq_evalp(B, p, i, o));
	}
	return mk(0.f, 0.f, 0.f);
}

DJB_DEV void tq_st3(float *p, long long k, V3 v)
{
	p[3 * k] = v.x;
	p[3 * k + 1] = v.y;
	p[3 * k + 2] = v.z;
}

template <int OP, bool PERPAIR>
__global__ void __launch_bounds__(TQ_THREADS) tabular_query_kernel(TabQueryArgs A)
{
	extern __shared__ float s_tab[]; // 6 * res floats
	__shared__ Params s_params[PERPAIR ? 1 : TQ_MAX_SMEM_PARAMS];
	__shared__ __align__(16) double s_dm[DMT_COUNT]; // djb_dmath.cuh: table of the double atan
	for (int t = threadIdx.x; t < DMT_COUNT; t += blockDim.x) s_dm[t] = g_dm_table_dev[t];
	for (int t = threadIdx.x; t < 6 * A.res; t += blockDim.x) s_tab[t] = A.tables[t];
	if (!PERPAIR) {
		const float *src = reinterpret_cast<const float *>(A.params);
		float *dst = reinterpret_cast<float *>(s_params);
		for (int t = threadIdx.x; t < A.n_params * 12; t += blockDim.x) dst[t] = src[t];
	}
	__syncthreads();
	TabBrdf B;
	B.t.p22 = s_tab; B.t.sigma = s_tab + A.res; B.t.n = A.res; B.t.T = s_dm;
	B.qf = s_tab + 2 * A.res;
	B.fr.pts = s_tab + 3 * A.res; B.fr.npts = A.res;
	B.shadow = A.shadow != 0;
	constexpr bool uses_u = (OP == OP_SAMPLE || OP == OP_EVALP_IS);
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < A.n; k += stride) {
		V3 va;
		if (uses_u) {
			const float2 u = reinterpret_cast<const float2 *>(A.a)[k];
			va = mk(u.x, u.y, 0.f);
		} else {
			va = mk(A.a[3 * k], A.a[3 * k + 1], A.a[3 * k + 2]);
		}
		const V3 o = mk(A.b[3 * k], A.b[3 * k + 1], A.b[3 * k + 2]);
		auto one = [&](const Params &p, long long slot) {
			if (OP == OP_EVAL) tq_st3(A.out0, slot, scale(rcp_via_double(va.z), tabq_evalp(B, p, va, o)));
			else if (OP == OP_EVALP) tq_st3(A.out0, slot, tabq_evalp(B, p, va, o));
			else if (OP == OP_PDF) A.out0[slot] = tabq_pdf(B, p, va, o);
			else if (OP == OP_SAMPLE) tq_st3(A.out0, slot, tabq_sample(B, p, va.x, va.y, o));
			else {
				V3 iv;
				float pdf;
				V3 w = tabq_evalp_is(B, p, va.x, va.y, o, iv, pdf);
				if (A.out0) tq_st3(A.out0, slot, w);
				if (A.out1) tq_st3(A.out1, slot, iv);
				if (A.out2) A.out2[slot] = pdf;
			}
		};
		if (PERPAIR) {
			const float4 *pp = reinterpret_cast<const float4 *>(A.params + k);
			const float4 q0 = pp[0], q1 = pp[1], q2 = pp[2];
			Params p;
			p.nx = q0.x; p.ny = q0.y; p.nz = q0.z; p.a1 = q0.w;
			p.a2 = q1.x; p.phi_a = q1.y; p.ax = q1.z; p.ay = q1.w;
			p.rho = q2.x; p.srho = q2.y; p.tx = q2.z; p.ty = q2.w;
			one(p, k);
		} else {
			for (int m = 0; m < A.n_params; ++m) one(s_params[m], (long long)m * A.out_stride + k);
		}
	}
}

template <int OP, bool PERPAIR>
static cudaError_t launch_tq(const TabQueryArgs &A, cudaStream_t st)
{
	const size_t smem = sizeof(float) * 6 * (size_t)A.res;
	if (smem > 40 * 1024) { // the attribute belongs to the current device's context: set whenever it is needed (cheap)
		cudaError_t e = cudaFuncSetAttribute(tabular_query_kernel<OP, PERPAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		if (e != cudaSuccess) return e;
	}
	long long want = (A.n + TQ_THREADS - 1) / TQ_THREADS, cap = (long long)sm_count() * 4;
	tabular_query_kernel<OP, PERPAIR><<<(int)(want < cap ? want : cap), TQ_THREADS, smem, st>>>(A);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

// djb::tabular_anisotropic (dj_brdf.h:428-478) as an evaluable BRDF: eval / evalp / pdf on the elevation x azimuth
// tables (2 x 32 KB at 90 x 90: read through the read-only cache, not staged)
template <int OP, bool PERPAIR>
__global__ void __launch_bounds__(TQ_THREADS) tabular_aniso_query_kernel(TabQueryArgs A, int azim_res, int n_qf1)
{
	__shared__ Params s_params[PERPAIR ? 1 : TQ_MAX_SMEM_PARAMS];
	__shared__ __align__(16) double s_dm[DMT_COUNT]; // djb_dmath.cuh: table of the double atan / atan2
	if (!PERPAIR) {
		const float *src = reinterpret_cast<const float *>(A.params);
		float *dst = reinterpret_cast<float *>(s_params);
		for (int t = threadIdx.x; t < A.n_params * 12; t += blockDim.x) dst[t] = src[t];
	}
	dm_load_tables(s_dm);
	TabBrdfT<TabAniso> B;
	const int tab = A.res * azim_res;
	B.t.p22 = A.tables; B.t.sigma = A.tables + tab; B.t.w = A.res; B.t.h = azim_res; B.t.T = s_dm;
	B.qf = nullptr;
	B.fr.pts = A.tables + 2 * tab; B.fr.npts = A.res;
	B.qf1 = A.tables + aniso_off_qf1(A.res, azim_res); B.n_qf1 = n_qf1;
	B.qf2 = A.tables + aniso_off_qf2(A.res, azim_res);
	B.shadow = A.shadow != 0;
	constexpr bool uses_u = (OP == OP_SAMPLE || OP == OP_EVALP_IS);
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < A.n; k += stride) {
		V3 i;
		if (uses_u) {
			const float2 u = reinterpret_cast<const float2 *>(A.a)[k];
			i = mk(u.x, u.y, 0.f);
		} else {
			i = mk(A.a[3 * k], A.a[3 * k + 1], A.a[3 * k + 2]);
		}
		const V3 o = mk(A.b[3 * k], A.b[3 * k + 1], A.b[3 * k + 2]);
		auto one = [&](const Params &p, long long slot) {
			if (OP == OP_EVAL) tq_st3(A.out0, slot, scale(rcp_via_double(i.z), tabq_evalp(B, p, i, o)));
			else if (OP == OP_EVALP) tq_st3(A.out0, slot, tabq_evalp(B, p, i, o));
			else if (OP == OP_PDF) A.out0[slot] = tabq_pdf(B, p, i, o);
			else if (OP == OP_SAMPLE) tq_st3(A.out0, slot, tabq_sample(B, p, i.x, i.y, o));
			else {
				V3 iv;
				float pdf;
				V3 w = tabq_evalp_is(B, p, i.x, i.y, o, iv, pdf);
				if (A.out0) tq_st3(A.out0, slot, w);
				if (A.out1) tq_st3(A.out1, slot, iv);
				if (A.out2) A.out2[slot] = pdf;
			}
		};
		if (PERPAIR) {
			const float4 *pp = reinterpret_cast<const float4 *>(A.params + k);
			const float4 q0 = pp[0], q1 = pp[1], q2 = pp[2];
			Params p;
			p.nx = q0.x; p.ny = q0.y; p.nz = q0.z; p.a1 = q0.w;
			p.a2 = q1.x; p.phi_a = q1.y; p.ax = q1.z; p.ay = q1.w;
			p.rho = q2.x; p.srho = q2.y; p.tx = q2.z; p.ty = q2.w;
			one(p, k);
		} else {
			for (int m = 0; m < A.n_params; ++m) one(s_params[m], (long long)m * A.out_stride + k);
		}
	}
}

// ---- the public scalar queries of djb::radial (dj_brdf.h:307-310): p22_radial, sigma_std_radial, cdf_radial, qf_radial ---
// what tests/plot_qf.cpp / plot_cdf.cpp of the reference call; one thread per argument.
// family: NDF_BECKMANN, NDF_GGX, or 2 = tabular (tables: p22 | sigma | qf | fresnel[3] | cdf, `res` entries each)
__global__ void __launch_bounds__(TQ_THREADS) radial_query_kernel(int family, int what, const float *__restrict__ tables, int res,
                                                                  const float *__restrict__ x, long long n, float *__restrict__ out)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
		const float v = x[k];
		float r = 0.0f;
		if (family == 2) {
			TabIso t; t.p22 = tables; t.sigma = tables + res; t.n = res; t.T = g_dm_table_dev;
			if (what == 0) r = t.p22_radial(v);
			else if (what == 1) r = spline_f(t.sigma, res, (float)(2.0 * acos((double)v) / (double)(float)DJB_PI)); // :2158-2162
			else if (what == 2) { // tabular::cdf_radial, :2164-2169
				float u = (float)(atan((double)v) * (double)2.0f / (double)(float)DJB_PI);
				if (u < 0.0f) u = 0.0f;
				r = spline_f(tables + 6 * res, res, (float)sqrt((double)u));
			} else { // tabular::qf_radial, :2171-2176
				const float q = spline_f(tables + 2 * res, res, v);
				r = (float)tan((double)(q * (float)DJB_PI / 2.0f));
			}
		} else if (family == NDF_GGX) {
			if (what == 0) r = p22_radial<NDF_GGX>(v);
			else if (what == 1) r = sigma_std_radial<NDF_GGX>(v);
			else if (what == 2) { const float t = v * v; r = (float)((double)t / (1.0 + (double)t)); }     // :2067-2071
			else r = (float)sqrt((double)v / (1.0 - (double)v));                                           // :2073-2076
		} else {
			if (what == 0) r = p22_radial<NDF_BECKMANN>(v);
			else if (what == 1) r = sigma_std_radial<NDF_BECKMANN>(v);
			else if (what == 2) r = (float)(1.0 - exp((double)(-v * v)));                                 // :1881-1884
			else r = (float)sqrt(-log(1.0 - (double)v));                                                   // :1886-1889
		}
		out[k] = r;
	}
}

cudaError_t launch_radial_query(int family, int what, const float *tables, int res, const float *x, int64_t n, float *out,
                                cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	long long want = (n + TQ_THREADS - 1) / TQ_THREADS, cap = (long long)sm_count() * 8;
	radial_query_kernel<<<(int)(want < cap ? want : cap), TQ_THREADS, 0, st>>>(family, what, tables, res, x, n, out);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

// ---- tabular_anisotropic's sampling tables (dj_brdf.h:2848-3103) ----------------------------------------------------
// A chain of eight small table passes, each depending on the previous one: marginal azimuth density (256-point elevation
// quadrature per azimuth), its normalisation, running integral and inverse; conditional elevation density, its per-azimuth
// normalisation, running integrals and inverses.  ~130 k table lookups in all: one CTA of 1024 threads walks the chain,
// every phase computes its independent terms in parallel (global scratch) and then adds them in the reference's order --
// float accumulation is not associative, so each ordered sum is done by a single thread (one per azimuth where the
// sums are per azimuth).
struct AnisoBuild {
	float *tab;       // handle tables (layout: aniso_table_floats)
	int er, ar;
	double *terms;    // ar * 256
	float *q;         // ar * 8 * (er - 1): running-integral lookups of the inversion searches (also 512 + 8 * ar for the 1-D passes)
	float *rows;      // ar * er: unpacked qf2 rows
	int *rowcnt;      // ar
	int *counts;      // 2
};

DJB_DEV float ab_angle(int i, int n, double scale) { return (float)((double)((float)i / (float)n) * scale); }
DJB_DEV float ab_lookup1(const float *t, int n, float phi) { return spline_repeat_f(t, n, (float)((double)phi * 0.5 / DJB_PI)); }
DJB_DEV float ab_lookup2(const float *t, int w, int h, float theta, float phi, float beyond)
{
	if ((double)theta >= 0.5 * DJB_PI) return beyond;
	return spline2d_f(t, w, h, (float)((double)theta * 2.0 / DJB_PI), (float)((double)phi * 0.5 / DJB_PI));
}
DJB_DEV double ab_term(float f, float theta) // (f * tan(theta)) / (cos_theta * cos_theta)
{
	const float c = (float)cos((double)theta);
	return ((double)f * tan((double)theta)) / (double)(c * c);
}

__global__ void __launch_bounds__(1024) aniso_sampling_tables_kernel(AnisoBuild A)
{
	const int er = A.er, ar = A.ar, T = er * ar, tid = threadIdx.x, nt = blockDim.x;
	const float *p22 = A.tab;
	float *qf1 = A.tab + aniso_off_qf1(er, ar), *qf2 = A.tab + aniso_off_qf2(er, ar);
	float *pdf1 = A.tab + aniso_off_pdf1(er, ar), *cdf1 = A.tab + aniso_off_cdf1(er, ar);
	float *pdf2 = A.tab + aniso_off_pdf2(er, ar), *cdf2 = A.tab + aniso_off_cdf2(er, ar);
	TabAniso tp; tp.p22 = p22; tp.sigma = nullptr; tp.w = er; tp.h = ar; tp.T = g_dm_table_dev;
	const double TWO_PI = 2.0 * DJB_PI, HALF_PI = 0.5 * DJB_PI;
	__shared__ float s_k;
	__shared__ int s_total;

	for (int t = tid; t < ar; t += nt) { qf1[t] = 0.f; pdf1[t] = 0.f; cdf1[t] = 0.f; }
	for (int t = tid; t < T; t += nt) { qf2[t] = 0.f; pdf2[t] = 0.f; cdf2[t] = 0.f; }
	// compute_pdf1, :2848-2874
	for (int t = tid; t < ar * 256; t += nt) {
		const int i = t / 256, j = t % 256;
		const float phi = ab_angle(i, ar, TWO_PI), theta = ab_angle(j, 256, HALF_PI);
		A.terms[t] = ab_term(tp.p22_theta_phi(theta, phi), theta);
	}
	__syncthreads();
	for (int i = tid; i < ar; i += nt) {
		const float dtheta = (float)(HALF_PI / (double)256.0f);
		float nint = 0.0f;
		for (int j = 0; j < 256; ++j) nint = (float)((double)nint + A.terms[i * 256 + j]);
		pdf1[i] = nint * dtheta;
	}
	__syncthreads();
	// normalize_pdf1, :3029-3051
	for (int t = tid; t < 512; t += nt) A.q[t] = ab_lookup1(pdf1, ar, ab_angle(t, 512, TWO_PI));
	__syncthreads();
	if (tid == 0) {
		float nint = 0.0f;
		for (int i = 0; i < 512; ++i) nint += A.q[i];
		nint *= (float)(TWO_PI / (double)512.0f);
		s_k = rcp_via_double(nint);
	}
	__syncthreads();
	for (int t = tid; t < ar; t += nt) pdf1[t] *= s_k;
	__syncthreads();
	// compute_cdf1, :2878-2900
	{
		const int cnt = ar - 1;
		for (int t = tid; t < cnt; t += nt) A.q[t] = ab_lookup1(pdf1, ar, ab_angle(t, cnt, TWO_PI));
		__syncthreads();
		if (tid == 0) {
			const float dphi = (float)(TWO_PI / (double)(float)cnt);
			float nint = 0.0f;
			cdf1[0] = 0.0f;
			for (int i = 1; i < cnt; ++i) {
				nint += A.q[i];
				cdf1[i] = nint * dphi;
			}
			cdf1[cnt] = 1.0f;
		}
		__syncthreads();
		// compute_qf1, :2904-2935: first j (never moving backwards) whose running integral reaches i / cnt
		const int res = cnt * 8;
		for (int t = tid; t < res; t += nt) A.q[t] = ab_lookup1(cdf1, ar, ab_angle(t, res, TWO_PI));
		__syncthreads();
		if (tid == 0) {
			int n1 = 0, j = 0;
			qf1[n1++] = 0.0f;
			for (int i = 1; i < cnt; ++i) {
				const float cdf = (float)i / (float)cnt;
				for (; j < res; ++j)
					if (A.q[j] >= cdf) { qf1[n1++] = (float)j / (float)res; break; }
			}
			qf1[n1++] = 1.0f;
			A.counts[0] = n1;
		}
		__syncthreads();
	}
	// compute_pdf2, :2944-2969
	{
		const int ntheta = er - 1;
		for (int t = tid; t < T; t += nt) {
			const int i = t / er, j = t % er;
			const float phi = ab_angle(i, ar, TWO_PI);
			pdf2[t] = j < ntheta ? tp.p22_theta_phi(ab_angle(j, ntheta, HALF_PI), phi) / ab_lookup1(pdf1, ar, phi) : 0.0f;
		}
		__syncthreads();
		// normalize_pdf2, :3055-3088 (all constants on the unscaled table, then the rows are scaled)
		for (int t = tid; t < ar * 256; t += nt) {
			const int j = t / 256, i = t % 256;
			const float phi = ab_angle(j, ar, TWO_PI), theta = ab_angle(i, 256, HALF_PI);
			A.terms[t] = ab_term(ab_lookup2(pdf2, er, ar, theta, phi, 0.0f), theta);
		}
		__syncthreads();
		for (int j = tid; j < ar; j += nt) {
			float nint = 0.0f;
			for (int i = 0; i < 256; ++i) nint = (float)((double)nint + A.terms[j * 256 + i]);
			nint *= (float)(HALF_PI / (double)256.0f);
			A.q[j] = rcp_via_double(nint);
		}
		__syncthreads();
		for (int t = tid; t < T; t += nt) pdf2[t] *= A.q[t / er];
		__syncthreads();
		// compute_cdf2, :2973-3000
		for (int t = tid; t < ar * ntheta; t += nt) {
			const int i = t / ntheta, j = t % ntheta;
			const float phi = ab_angle(i, ar, TWO_PI), theta = ab_angle(j, ntheta, HALF_PI);
			A.terms[t] = ab_term(ab_lookup2(pdf2, er, ar, theta, phi, 0.0f), theta);
		}
		__syncthreads();
		for (int i = tid; i < ar; i += nt) {
			const float dtheta = (float)(HALF_PI / (double)(float)ntheta);
			float nint = 0.0f;
			for (int j = 0; j < ntheta; ++j) {
				nint = (float)((double)nint + A.terms[i * ntheta + j]);
				cdf2[i * er + j] = nint * dtheta;
			}
			cdf2[i * er + ntheta] = 1.0f;
		}
		__syncthreads();
		// compute_qf2, :3004-3037
		const int res = ntheta * 8;
		for (int t = tid; t < ar * res; t += nt) {
			const int k = t / res, j = t % res;
			A.q[t] = ab_lookup2(cdf2, er, ar, ab_angle(j, res, HALF_PI), ab_angle(k, ar, TWO_PI), 1.0f);
		}
		__syncthreads();
		for (int k = tid; k < ar; k += nt) {
			float *row = A.rows + k * er;
			int n = 0, j = 0;
			row[n++] = 0.0f;
			for (int i = 1; i < ntheta; ++i) {
				const float cdf = (float)i / (float)ntheta;
				for (; j < res; ++j)
					if (A.q[k * res + j] >= cdf) { row[n++] = (float)j / (float)res; break; }
			}
			row[n++] = 1.0f;
			A.rowcnt[k] = n;
		}
		__syncthreads();
		// the reference appends the rows to one vector: pack them back to back
		if (tid == 0) {
			int off = 0;
			for (int k = 0; k < ar; ++k) {
				const int n = A.rowcnt[k];
				A.rowcnt[k] = off;
				off += n;
			}
			A.counts[1] = off;
			s_total = off;
		}
		__syncthreads();
		for (int t = tid; t < T; t += nt) {
			const int k = t / er, e = t % er;
			const int begin = A.rowcnt[k], end = k + 1 < ar ? A.rowcnt[k + 1] : s_total;
			if (e < end - begin) qf2[begin + e] = A.rows[t];
		}
	}
}

cudaError_t build_aniso_sampling_tables(float *tables, int elev_res, int azim_res, int counts_host[2], cudaStream_t st)
{
	const size_t er = (size_t)elev_res, ar = (size_t)azim_res;
	const size_t n_terms = ar * 256 > ar * er ? ar * 256 : ar * er;
	size_t n_q = ar * 8 * (er - 1);
	if (n_q < 512 + 8 * ar) n_q = 512 + 8 * ar;
	const size_t bytes = 8 * n_terms + 4 * n_q + 4 * ar * er + 4 * ar + 8;
	char *ws = nullptr;
	cudaError_t e = cudaMalloc(&ws, bytes);
	if (e != cudaSuccess) return e;
	AnisoBuild A;
	A.tab = tables; A.er = elev_res; A.ar = azim_res;
	A.terms = reinterpret_cast<double *>(ws);
	A.q = reinterpret_cast<float *>(ws + 8 * n_terms);
	A.rows = A.q + n_q;
	A.rowcnt = reinterpret_cast<int *>(A.rows + ar * er);
	A.counts = A.rowcnt + ar;
	aniso_sampling_tables_kernel<<<1, 1024, 0, st>>>(A);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	e = cudaGetLastError();
	if (e == cudaSuccess) e = cudaMemcpyAsync(counts_host, A.counts, 8, cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	cudaFree(ws);
	return e;
}

// tables: device, laid out as aniso_table_floats() describes
cudaError_t launch_tabular_aniso_query(const float *tables, int elev_res, int azim_res, int n_qf1, const MfLaunch &L,
                                       cudaStream_t st)
{
	if (L.n <= 0) return cudaSuccess;
	TabQueryArgs A;
	A.tables = tables; A.res = elev_res; A.shadow = L.shadow;
	A.a = L.a; A.b = L.b; A.n = L.n; A.out_stride = L.out_stride;
	A.out1 = L.out1; A.out2 = L.out2;
	const int per = (L.op == OP_PDF) ? 1 : 3;
	long long want = (L.n + TQ_THREADS - 1) / TQ_THREADS, cap = (long long)sm_count() * 4;
	const int grid = (int)(want < cap ? want : cap);
#define TA_DISPATCH(PP)                                                                                             \
	switch (L.op) {                                                                                                 \
	case OP_EVAL: tabular_aniso_query_kernel<OP_EVAL, PP><<<grid, TQ_THREADS, 0, st>>>(A, azim_res, n_qf1); break;          \
	case OP_EVALP: tabular_aniso_query_kernel<OP_EVALP, PP><<<grid, TQ_THREADS, 0, st>>>(A, azim_res, n_qf1); break;        \
	case OP_PDF: tabular_aniso_query_kernel<OP_PDF, PP><<<grid, TQ_THREADS, 0, st>>>(A, azim_res, n_qf1); break;            \
	case OP_SAMPLE: tabular_aniso_query_kernel<OP_SAMPLE, PP><<<grid, TQ_THREADS, 0, st>>>(A, azim_res, n_qf1); break;      \
	default: tabular_aniso_query_kernel<OP_EVALP_IS, PP><<<grid, TQ_THREADS, 0, st>>>(A, azim_res, n_qf1);                  \
	}                                                                                                               \
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	if (L.layout == DJB200_PARAMS_PER_PAIR) {
		A.params = reinterpret_cast<const Params *>(L.params);
		A.n_params = 1;
		A.out0 = L.out0; A.out1 = L.out1; A.out2 = L.out2;
		TA_DISPATCH(true)
		return cudaGetLastError();
	}
	for (int64_t m0 = 0; m0 < L.n_params; m0 += TQ_MAX_SMEM_PARAMS) {
		int64_t mc = L.n_params - m0 < TQ_MAX_SMEM_PARAMS ? L.n_params - m0 : TQ_MAX_SMEM_PARAMS;
		A.params = reinterpret_cast<const Params *>(L.params) + m0;
		A.n_params = (int)mc;
		A.out0 = L.out0 ? L.out0 + m0 * L.out_stride * per : nullptr;
		A.out1 = L.out1 ? L.out1 + m0 * L.out_stride * 3 : nullptr;
		A.out2 = L.out2 ? L.out2 + m0 * L.out_stride : nullptr;
		TA_DISPATCH(false)
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) return e;
	}
#undef TA_DISPATCH
	return cudaSuccess;
}

// L: same launch description as the microfacet queries; tables: device, p22 | sigma | qf | fresnel
cudaError_t launch_tabular_query(const float *tables, int res, const MfLaunch &L, cudaStream_t st)
{
	if (L.n <= 0) return cudaSuccess;
	TabQueryArgs A;
	A.tables = tables; A.res = res; A.shadow = L.shadow;
	A.a = L.a; A.b = L.b; A.n = L.n; A.out_stride = L.out_stride;
	const int per = (L.op == OP_PDF) ? 1 : 3;
#define TQ_DISPATCH(PP)                                                                   \
	switch (L.op) {                                                                       \
	case OP_EVAL: e = launch_tq<OP_EVAL, PP>(A, st); break;                                \
	case OP_EVALP: e = launch_tq<OP_EVALP, PP>(A, st); break;                              \
	case OP_PDF: e = launch_tq<OP_PDF, PP>(A, st); break;                                  \
	case OP_SAMPLE: e = launch_tq<OP_SAMPLE, PP>(A, st); break;                            \
	case OP_EVALP_IS: e = launch_tq<OP_EVALP_IS, PP>(A, st); break;                        \
	default: e = cudaErrorInvalidValue;                                                   \
	}
	cudaError_t e = cudaSuccess;
	if (L.layout == DJB200_PARAMS_PER_PAIR) {
		A.params = reinterpret_cast<const Params *>(L.params);
		A.n_params = 1;
		A.out0 = L.out0; A.out1 = L.out1; A.out2 = L.out2;
		TQ_DISPATCH(true)
		return e;
	}
	for (int64_t m0 = 0; m0 < L.n_params && e == cudaSuccess; m0 += TQ_MAX_SMEM_PARAMS) {
		int64_t mc = L.n_params - m0 < TQ_MAX_SMEM_PARAMS ? L.n_params - m0 : TQ_MAX_SMEM_PARAMS;
		A.params = reinterpret_cast<const Params *>(L.params) + m0;
		A.n_params = (int)mc;
		int64_t off = m0 * L.out_stride;
		A.out0 = L.out0 ? L.out0 + off * per : nullptr;
		A.out1 = L.out1 ? L.out1 + off * 3 : nullptr;
		A.out2 = L.out2 ? L.out2 + off : nullptr;
		TQ_DISPATCH(false)
	}
#undef TQ_DISPATCH
	return e;
}

// ---- the remaining public scalar members of the reference's classes (dj_brdf.h:366-369, 384-389, 450-455, 506-509, 531-533) ------
// beckmann / ggx ::qf1, qf2_radial, qf3_radial; tabular_anisotropic ::pdf1 / cdf1 / qf1 / pdf2 / cdf2 / qf2; sgd ::ndf / gaf / g1 /
// fresnel; abc ::ndf / gaf / fresnel.  One query per thread, mirrored-rounding tier (these are the very functions sample / eval
// are built from).  Rare calls: correctness and the reference's interface, not throughput.
struct MemberArgs {
	int family, what;         // family: 0 beckmann, 1 ggx, 2 tabular_anisotropic, 3 sgd, 4 abc
	const float *tables;      // tabular_anisotropic handle tables
	int er, ar, n_qf1;
	double coef[33];          // sgd / abc coefficients
	const float *a, *b, *c;
	long long n;
	float *out;
};
DJB_DEV float ggx_qf1_ref(float u) // ggx::qf1, dj_brdf.h:2078-2087
{
	if ((double)u < 0.5) {
		u = (float)((0.5 - (double)u) * 2.0);
		return -u * inv_sqrt((float)(1.0 - (double)(u * u)));
	}
	u = (float)(((double)u - 0.5) * 2.0);
	return u * inv_sqrt((float)(1.0 - (double)(u * u)));
}
__global__ void __launch_bounds__(TQ_THREADS) member_query_kernel(MemberArgs A)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < A.n; k += stride) {
		if (A.family <= 1) { // a = u, b = cos_theta_k or qf2, c = sin_theta_k
			const float u = A.a[k];
			float r;
			if (A.what == DJB200_MEMBER_QF1) r = A.family == NDF_GGX ? ggx_qf1_ref(u) : erfinv_giles((float)(2.0 * (double)u - 1.0));
			else if (A.what == DJB200_MEMBER_QF2_RADIAL) r = A.family == NDF_GGX ? ggx_qf2(u, A.b[k], A.c[k]) : beckmann_qf2(u, A.b[k], A.c[k]);
			else r = A.family == NDF_GGX ? ggx_qf3(u, A.b[k]) : erfinv_giles((float)(2.0 * (double)u - 1.0)); // beckmann::qf3_radial = qf1(u)
			A.out[k] = r;
		} else if (A.family == 2) { // dj_brdf.h:2766-2824
			const int er = A.er, ar = A.ar;
			const float x = A.a[k], y = A.b ? A.b[k] : 0.0f;
			float r;
			switch (A.what) {
			case DJB200_MEMBER_PDF1: r = ab_lookup1(A.tables + aniso_off_pdf1(er, ar), ar, x); break;
			case DJB200_MEMBER_CDF1: r = ab_lookup1(A.tables + aniso_off_cdf1(er, ar), ar, x); break;
			case DJB200_MEMBER_TQF1: r = (float)((double)spline_f(A.tables + aniso_off_qf1(er, ar), A.n_qf1, x) * 2.0 * DJB_PI); break;
			case DJB200_MEMBER_PDF2: r = ab_lookup2(A.tables + aniso_off_pdf2(er, ar), er, ar, x, y, 0.0f); break;
			case DJB200_MEMBER_CDF2: r = ab_lookup2(A.tables + aniso_off_cdf2(er, ar), er, ar, x, y, 1.0f); break;
			default: r = (float)((double)spline2d_f(A.tables + aniso_off_qf2(er, ar), er, ar, x, (float)((double)y / (2.0 * DJB_PI))) * 0.5 * DJB_PI);
			}
			A.out[k] = r;
		} else if (A.family == 3) { // sgd, dj_brdf.h:3471-3499: a = h (ndf) / unused (gaf) / k (g1) / cos (fresnel); b = i, c = o
			V3 r;
			if (A.what == DJB200_MEMBER_NDF) {
				const double ch = (double)A.a[3 * k + 2], c2 = ch * ch, t2 = (1.0 - c2) / c2, inv_pi = 1.0 / DJB_PI;
				float v[3];
				for (int c = 0; c < 3; ++c) {
					const double *mc = A.coef + 11 * c;
					const double ax = mc[2] + t2 / mc[2];
					v[c] = (float)((mc[6] * exp(-ax) * inv_pi) / (pow(ax, mc[3]) * c2 * c2));
				}
				r = mk(v[0], v[1], v[2]);
			} else if (A.what == DJB200_MEMBER_G1 || A.what == DJB200_MEMBER_GAF) {
				float g[3] = {1.f, 1.f, 1.f};
				const int terms = A.what == DJB200_MEMBER_GAF ? 2 : 1;
				for (int t = 0; t < terms; ++t) { // gaf(h, i, o) = g1(i) * g1(o), vec3 products in float
					const float *dir = A.what == DJB200_MEMBER_GAF ? (t == 0 ? A.b : A.c) : A.a;
					const double ak = acos((double)dir[3 * k + 2]);
					for (int c = 0; c < 3; ++c) {
						const float g1 = (float)sgd_g1_ch(ak, A.coef + 11 * c);
						g[c] = t == 0 ? g1 : g[c] * g1;
					}
				}
				r = mk(g[0], g[1], g[2]);
			} else {
				FresnelDev fr;
				fr.pts = nullptr; fr.npts = 0;
				for (int c = 0; c < 3; ++c) { fr.v[c] = (float)A.coef[11 * c + 4]; fr.v[3 + c] = (float)A.coef[11 * c + 5]; }
				r = fresnel_eval<FK_SGD>(fr, A.a[k]);
			}
			A.out[3 * k] = r.x; A.out[3 * k + 1] = r.y; A.out[3 * k + 2] = r.z;
		} else { // abc, dj_brdf.h:3649-3668: coef = kD[3] A[3] B C ior
			if (A.what == DJB200_MEMBER_GAF) {
				const V3 h = mk(A.a[3 * k], A.a[3 * k + 1], A.a[3 * k + 2]), i = mk(A.b[3 * k], A.b[3 * k + 1], A.b[3 * k + 2]);
				const V3 o = mk(A.c[3 * k], A.c[3 * k + 1], A.c[3 * k + 2]);
				const float g1_i = fmin_ref(1.0f, 2.0f * (h.z * i.z / dot(h, i))), g1_o = fmin_ref(1.0f, 2.0f * (h.z * o.z / dot(h, o)));
				A.out[k] = fmin_ref(g1_i, g1_o);
				continue;
			}
			V3 r;
			if (A.what == DJB200_MEMBER_NDF) {
				const double den = pow(1.0 + A.coef[6] * (1.0 - (double)A.a[3 * k + 2]), A.coef[7]);
				r = mk((float)(A.coef[3] / den), (float)(A.coef[4] / den), (float)(A.coef[5] / den));
			} else {
				const float f = unpolarized_channel(A.a[k], (float)A.coef[8]);
				r = mk(f, f, f);
			}
			A.out[3 * k] = r.x; A.out[3 * k + 1] = r.y; A.out[3 * k + 2] = r.z;
		}
	}
}

cudaError_t launch_member_query(int family, int what, const float *tables, int er, int ar, int n_qf1, const double *coef, int n_coef,
                                const float *a, const float *b, const float *c, int64_t n, float *out, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	MemberArgs A;
	A.family = family; A.what = what; A.tables = tables; A.er = er; A.ar = ar; A.n_qf1 = n_qf1;
	for (int k = 0; k < 33; ++k) A.coef[k] = coef && k < n_coef ? coef[k] : 0.0;
	A.a = a; A.b = b; A.c = c; A.n = n; A.out = out;
	long long want = (n + TQ_THREADS - 1) / TQ_THREADS, cap = (long long)sm_count() * 8;
	member_query_kernel<<<(int)(want < cap ? want : cap), TQ_THREADS, 0, st>>>(A);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

} // namespace djb200
