// kernels_analytic.cu -- the analytic data-driven BRDFs of the reference: djb::sgd (shifted gamma distribution) and
// djb::abc, one thread per (wi, wo) pair (dj_brdf.h:3416-3499, 3608-3668).
//
// The coefficients of one material (33 or 9 doubles) ride in the kernel parameter block, i.e. in constant memory:
// every lane reads the same address, so they cost one broadcast each and no shared-memory staging is needed.
// HBM traffic: 24 B in + 12 B out per pair.  The math is the reference's: acos / exp / pow in double per channel.
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
__global__ void __launch_bounds__(TB) analytic_eval_kernel(const AnalyticCoef m, const float *__restrict__ wi,
                                                           const float *__restrict__ wo, long long n,
                                                           float *__restrict__ out)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
		const V3 i = mk(wi[3 * k], wi[3 * k + 1], wi[3 * k + 2]);
		const V3 o = mk(wo[3 * k], wo[3 * k + 1], wo[3 * k + 2]);
		const V3 r = KIND == DJB200_SOURCE_SGD ? sgd_eval1(m.v, i, o) : abc_eval1(m.v, i, o);
		out[3 * k] = r.x;
		out[3 * k + 1] = r.y;
		out[3 * k + 2] = r.z;
	}
}
} // namespace

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
