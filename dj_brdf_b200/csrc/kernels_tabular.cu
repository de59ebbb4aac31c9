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
		return scale((float)((double)(Dn * G) / (4.0 * (double)o.z)), Fr);
	}
	return mk(0.f, 0.f, 0.f);
}

// microfacet::pdf without Smith VNDF sampling, dj_brdf.h:1724-1725
template <class TB>
DJB_DEV float tabq_pdf(const TB &B, const Params &p, V3 i, V3 o)
{
	V3 h = normalize(i + o);
	float G = tabq_gaf(B, p, i, o);
	if (G > 0.0f) return (float)((double)(h.z * tab_ndf(B.t, p, h)) / (4.0 * (double)dot(i, h)));
	return 0.0f;
}

// microfacet::sample (dj_brdf.h:1669-1709) with radial::sample_vp22_std_nmap (:1806-1816)
DJB_DEV V3 tabq_sample(const TabBrdf &B, const Params &p, float u1, float u2, V3 o)
{
	u1 = sat_ref(u1) * 0.99998f + 0.00001f;
	u2 = sat_ref(u2) * 0.99998f + 0.00001f;
	float a = o.x * p.ax + o.y * p.ay * p.rho;
	float b = o.y * p.ay * p.srho;
	float c = o.z - o.x * p.tx - o.y * p.ty;
	V3 os = normalize(mk(a, b, c));
	if (os.z > 0.0f) {
		float phi_h = (float)((double)u1 * DJB_PI * 2.0);
		float q = spline_f(B.qf, B.t.n, u2);                        // tabular::qf_radial, :2172-2176
		float r_h = (float)tan((double)(q * (float)DJB_PI / 2.0f));
		double sp, cp;
		sincos((double)phi_h, &sp, &cp);
		float txm = (float)((double)r_h * cp), tym = (float)((double)r_h * sp);
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
DJB_DEV V3 tabq_evalp_is(const TabBrdf &B, const Params &p, float u1, float u2, V3 o, V3 &i_out, float &pdf_out)
{
	V3 i = tabq_sample(B, p, u1, u2, o);
	V3 h = normalize(i + o);
	float G = tabq_gaf(B, p, i, o);
	pdf_out = 0.0f;
	i_out = mk(0.f, 0.f, 0.f);
	if (G > 0.0f) {
		float cd = sat_ref(dot(o, h));
		i_out = i;
		float pdf = (float)((double)(h.z * tab_ndf(B.t, p, h)) / (4.0 * (double)cd));
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
	for (int t = threadIdx.x; t < 6 * A.res; t += blockDim.x) s_tab[t] = A.tables[t];
	if (!PERPAIR) {
		const float *src = reinterpret_cast<const float *>(A.params);
		float *dst = reinterpret_cast<float *>(s_params);
		for (int t = threadIdx.x; t < A.n_params * 12; t += blockDim.x) dst[t] = src[t];
	}
	__syncthreads();
	TabBrdf B;
	B.t.p22 = s_tab; B.t.sigma = s_tab + A.res; B.t.n = A.res;
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
	static bool attr_set = false;
	if (!attr_set && smem > 40 * 1024) {
		cudaError_t e = cudaFuncSetAttribute(tabular_query_kernel<OP, PERPAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		if (e != cudaSuccess) return e;
		attr_set = true;
	}
	long long want = (A.n + TQ_THREADS - 1) / TQ_THREADS, cap = (long long)sm_count() * 4;
	tabular_query_kernel<OP, PERPAIR><<<(int)(want < cap ? want : cap), TQ_THREADS, smem, st>>>(A);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

// djb::tabular_anisotropic (dj_brdf.h:428-478) as an evaluable BRDF: eval / evalp / pdf on the elevation x azimuth
// tables (2 x 32 KB at 90 x 90: read through the read-only cache, not staged)
template <int OP, bool PERPAIR>
__global__ void __launch_bounds__(TQ_THREADS) tabular_aniso_query_kernel(TabQueryArgs A, int azim_res)
{
	__shared__ Params s_params[PERPAIR ? 1 : TQ_MAX_SMEM_PARAMS];
	if (!PERPAIR) {
		const float *src = reinterpret_cast<const float *>(A.params);
		float *dst = reinterpret_cast<float *>(s_params);
		for (int t = threadIdx.x; t < A.n_params * 12; t += blockDim.x) dst[t] = src[t];
		__syncthreads();
	}
	TabBrdfT<TabAniso> B;
	const int tab = A.res * azim_res;
	B.t.p22 = A.tables; B.t.sigma = A.tables + tab; B.t.w = A.res; B.t.h = azim_res;
	B.qf = nullptr;
	B.fr.pts = A.tables + 2 * tab; B.fr.npts = A.res;
	B.shadow = A.shadow != 0;
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < A.n; k += stride) {
		const V3 i = mk(A.a[3 * k], A.a[3 * k + 1], A.a[3 * k + 2]);
		const V3 o = mk(A.b[3 * k], A.b[3 * k + 1], A.b[3 * k + 2]);
		auto one = [&](const Params &p, long long slot) {
			if (OP == OP_EVAL) tq_st3(A.out0, slot, scale(rcp_via_double(i.z), tabq_evalp(B, p, i, o)));
			else if (OP == OP_EVALP) tq_st3(A.out0, slot, tabq_evalp(B, p, i, o));
			else A.out0[slot] = tabq_pdf(B, p, i, o);
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

// tables: device, p22[er * ar] | sigma[er * ar] | fresnel[er][3]
cudaError_t launch_tabular_aniso_query(const float *tables, int elev_res, int azim_res, const MfLaunch &L, cudaStream_t st)
{
	if (L.n <= 0) return cudaSuccess;
	if (L.op != OP_EVAL && L.op != OP_EVALP && L.op != OP_PDF) return cudaErrorNotSupported;
	TabQueryArgs A;
	A.tables = tables; A.res = elev_res; A.shadow = L.shadow;
	A.a = L.a; A.b = L.b; A.n = L.n; A.out_stride = L.out_stride;
	A.out1 = A.out2 = nullptr;
	const int per = (L.op == OP_PDF) ? 1 : 3;
	long long want = (L.n + TQ_THREADS - 1) / TQ_THREADS, cap = (long long)sm_count() * 4;
	const int grid = (int)(want < cap ? want : cap);
#define TA_DISPATCH(PP)                                                                                             \
	switch (L.op) {                                                                                                 \
	case OP_EVAL: tabular_aniso_query_kernel<OP_EVAL, PP><<<grid, TQ_THREADS, 0, st>>>(A, azim_res); break;          \
	case OP_EVALP: tabular_aniso_query_kernel<OP_EVALP, PP><<<grid, TQ_THREADS, 0, st>>>(A, azim_res); break;        \
	default: tabular_aniso_query_kernel<OP_PDF, PP><<<grid, TQ_THREADS, 0, st>>>(A, azim_res);                       \
	}                                                                                                               \
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	if (L.layout == DJB200_PARAMS_PER_PAIR) {
		A.params = reinterpret_cast<const Params *>(L.params);
		A.n_params = 1;
		A.out0 = L.out0;
		TA_DISPATCH(true)
		return cudaGetLastError();
	}
	for (int64_t m0 = 0; m0 < L.n_params; m0 += TQ_MAX_SMEM_PARAMS) {
		int64_t mc = L.n_params - m0 < TQ_MAX_SMEM_PARAMS ? L.n_params - m0 : TQ_MAX_SMEM_PARAMS;
		A.params = reinterpret_cast<const Params *>(L.params) + m0;
		A.n_params = (int)mc;
		A.out0 = L.out0 + m0 * L.out_stride * per;
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

} // namespace djb200
