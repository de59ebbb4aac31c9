// tests/cpp/dmath_check.cpp -- dj_brdf_b200/csrc/djb_dmath.cuh compiled for the host (the same source the kernels compile) against
// libm: largest error in units of the last place over dense sweeps of the ranges the analytic BRDF kernels use.
//   dmath_check            prints "exp max_ulp <e>  log max_ulp <l>" and exits 0 when both are <= 4
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "djb_dmath.cuh"

static double ulps(double got, double want)
{
	if (got == want) return 0.0;
	if (std::isnan(got) && std::isnan(want)) return 0.0;
	int e;
	frexp(want, &e);
	return fabs(got - want) / ldexp(1.0, e - 53);
}

int main()
{
	using namespace djb200;
	uint64_t st = 0x9E3779B97F4A7C15ull;
	auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
	double worst_e = 0, worst_l = 0;
	const int N = 20000000;
	for (int k = 0; k < N; ++k) {
		const double u = rnd();
		// exp: uniform in [-700, 700], plus a dense band near 0
		const double x = (k & 1) ? (u * 2.0 - 1.0) * 700.0 : (u * 2.0 - 1.0) * 2.0;
		const double ue = ulps(exp_d(x), exp(x));
		if (ue > worst_e) worst_e = ue;
		// log: log-uniform over [1e-300, 1e300], plus a dense band around 1
		const double y = (k & 1) ? pow(10.0, (u * 2.0 - 1.0) * 300.0) : 0.5 + u;
		const double ul = ulps(log_d(y), log(y));
		if (ul > worst_l) worst_l = ul;
	}
	const double edge[] = {0.0, -0.0, 1e-320, -1.0, INFINITY, -INFINITY, NAN, 700.0, -700.0, 709.0, -745.2, 1.0};
	int bad = 0;
	for (double x : edge) {
		const double a = exp_d(x), b = exp(x), c = log_d(x), d = log(x);
		if (!((a == b) || (std::isnan(a) && std::isnan(b)) || ulps(a, b) <= 4)) ++bad;
		if (!((c == d) || (std::isnan(c) && std::isnan(d)) || ulps(c, d) <= 4)) ++bad;
	}
	printf("exp max_ulp %.3f  log max_ulp %.3f  edge_mismatches %d\n", worst_e, worst_l, bad);
	return worst_e <= 4.0 && worst_l <= 4.0 && bad == 0 ? 0 : 1;
}
