// kernels_analytic.cu -- the analytic data-driven BRDFs of the reference: djb::sgd (shifted gamma distribution) and
// djb::abc, one thread per (wi, wo) pair (dj_brdf.h:3416-3499, 3608-3668).
//
// The coefficients of one material (33 or 9 doubles) ride in the kernel parameter block, i.e. in constant memory:
// every lane reads the same address, so they cost one broadcast each and no shared-memory staging is needed.
// HBM traffic: 24 B in + 12 B out per pair.  The math is the reference's: acos / exp / pow in double per channel.
#include <cstring>

#include "djb_device.cuh"
#include "djb_internal.h"

namespace djb200 {

namespace {
constexpr int TB = 256;

struct AnalyticCoef { double v[33]; };

inline int grid_for(int64_t n)
{
	int64_t want = (n + TB - 1) / TB;
	int64_t cap = (int64_t)sm_count() * 8;
	return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

template <int KIND>
__global__ void __launch_bounds__(TB) analytic_eval_kernel(const __grid_constant__ AnalyticCoef m, const float *__restrict__ wi,
                                                           const float *__restrict__ wo, long long n,
                                                           float *__restrict__ out)
{
	__shared__ __align__(16) double s_dm[DMT_COUNT]; // djb_dmath.cuh: tables of the double exp / log
	dm_load_tables(s_dm);
	const bool plain = KIND == DJB200_SOURCE_SGD && sgd_material_plain(m.v);
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
		const V3 i = mk(wi[3 * k], wi[3 * k + 1], wi[3 * k + 2]);
		const V3 o = mk(wo[3 * k], wo[3 * k + 1], wo[3 * k + 2]);
		const V3 r = KIND == DJB200_SOURCE_SGD ? sgd_eval1(m.v, i, o, s_dm, plain) : abc_eval1(m.v, i, o, s_dm);
		out[3 * k] = r.x;
		out[3 * k + 1] = r.y;
		out[3 * k + 2] = r.z;
	}
}
// djb200_debug_dmath: the double functions of djb_dmath.cuh as the kernels above compile them
__global__ void __launch_bounds__(TB) dmath_kernel(int fn, const double *__restrict__ x, const double *__restrict__ y, long long n,
                                                   double *__restrict__ out)
{
	__shared__ __align__(16) double s_dm[DMT_COUNT];
	dm_load_tables(s_dm);
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
		const double a = x[k], b = y ? y[k] : 0.0;
		double r = 0.0, sn, cs;
		switch (fn) {
		case 0: r = exp_t(a, s_dm); break;
		case 1: r = log_t(a, s_dm); break;
		case 2: r = sqrt_d(a); break;
		case 3: r = acos_d(a); break;
		case 4: r = atan_t(a, s_dm); break;
		case 5: r = atan2_t(a, b, s_dm); break;
		case 6: sincos_d(a, &sn, &cs); r = sn; break;
		case 7: sincos_d(a, &sn, &cs); r = cs; break;
		case 8: r = (fabs(b) >= 1e-30 && fabs(b) <= 1e30) ? div_core(a, b) : a / b; break;
		default: r = pow_pos_t(a, b, s_dm); break;
		}
		out[k] = r;
	}
}
// the public component queries of djb::microfacet (dj_brdf.h:258-272, 1559-1665): ndf(h), gaf(h, i, o), g1(h, k), sigma(k),
// p22(x, y), vp22(x, y, k), vndf(h, k), fresnel(cos) under one params block -- mirrored-rounding tier, one query per thread
struct ComponentArgs {
	int what, shadow, fresnel_kind;
	FresnelDev fr;
	Params p;
	const float *a, *b, *c;
	long long n;
	float *out;
};

template <int NDF>
__global__ void __launch_bounds__(TB) component_kernel(ComponentArgs A)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < A.n; k += stride) {
		const V3 va = mk(A.a[3 * k], A.a[3 * k + 1], A.a[3 * k + 2]);
		V3 vb = mk(0.f, 0.f, 1.f), vc = mk(0.f, 0.f, 1.f);
		if (A.b) vb = mk(A.b[3 * k], A.b[3 * k + 1], A.b[3 * k + 2]);
		if (A.c) vc = mk(A.c[3 * k], A.c[3 * k + 1], A.c[3 * k + 2]);
		float r = 0.0f;
		switch (A.what) {
		case DJB200_COMP_NDF: r = mf_ndf<NDF>(A.p, va); break;
		case DJB200_COMP_GAF: r = mf_gaf<NDF>(A.p, A.shadow != 0, vb, vc); break; // (h, i, o): h is not used, :1644-1665
		case DJB200_COMP_G1: r = mf_g1<NDF>(A.p, vb); break;                       // (h, k): h is not used, :1633-1642
		case DJB200_COMP_SIGMA: r = mf_sigma<NDF>(A.p, va); break;
		case DJB200_COMP_P22: r = mf_p22<NDF>(A.p, va.x, va.y); break;
		case DJB200_COMP_VP22: { // :1589-1598
			const V3 h = normalize(mk(-va.x, -va.y, 1.0f));
			const float jacobian = h.z * h.z * h.z;
			r = jacobian * mf_vndf<NDF>(A.p, h, vb);
		} break;
		case DJB200_COMP_VNDF: r = mf_vndf<NDF>(A.p, va, vb); break;
		default: { // DJB200_COMP_FRESNEL: rgb
			const V3 f = fresnel_rt(A.fresnel_kind, A.fr, va.x);
			A.out[3 * k] = f.x;
			A.out[3 * k + 1] = f.y;
			A.out[3 * k + 2] = f.z;
			continue;
		}
		}
		A.out[k] = r;
	}
}
} // namespace

cudaError_t launch_microfacet_component(int ndf, int shadow, int fresnel_kind, const float fv[6], const float *spline_pts,
                                        int spline_n, const void *params_host, int what, const float *a, const float *b,
                                        const float *c, int64_t n, float *out, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	ComponentArgs A;
	A.what = what; A.shadow = shadow; A.fresnel_kind = fresnel_kind;
	for (int k = 0; k < 6; ++k) A.fr.v[k] = fv[k];
	A.fr.pts = spline_pts; A.fr.npts = spline_n;
	memcpy(&A.p, params_host, sizeof(Params));
	A.a = a; A.b = b; A.c = c; A.n = n; A.out = out;
	if (ndf == NDF_GGX) component_kernel<NDF_GGX><<<grid_for(n), TB, 0, st>>>(A);
	else if (ndf == NDF_BECKMANN) component_kernel<NDF_BECKMANN><<<grid_for(n), TB, 0, st>>>(A);
	else return cudaErrorInvalidValue;
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

cudaError_t launch_debug_dmath(int fn, const double *x, const double *y, int64_t n, double *out, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	dmath_kernel<<<grid_for(n), TB, 0, st>>>(fn, x, y, n, out);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

cudaError_t launch_analytic_eval(int kind, const double *coef, int n_coef, const float *wi, const float *wo, int64_t n,
                                 float *out, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	AnalyticCoef m;
	for (int k = 0; k < 33; ++k) m.v[k] = k < n_coef ? coef[k] : 0.0;
	if (kind == DJB200_SOURCE_SGD) analytic_eval_kernel<DJB200_SOURCE_SGD><<<grid_for(n), TB, 0, st>>>(m, wi, wo, n, out);
	else if (kind == DJB200_SOURCE_ABC) analytic_eval_kernel<DJB200_SOURCE_ABC><<<grid_for(n), TB, 0, st>>>(m, wi, wo, n, out);
	else return cudaErrorInvalidValue;
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

} // namespace djb200
