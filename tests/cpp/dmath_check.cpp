// tests/cpp/dmath_check.cpp -- dj_brdf_b200/csrc/djb_dmath.cuh compiled for the host (the same source the kernels compile) against
// libm: largest error in units of the last place over dense sweeps of the ranges the analytic BRDF kernels use.
//   dmath_check            prints "exp max_ulp <e>  log max_ulp <l>" and exits 0 when both are <= 4
// and the table-driven set (exp_t / log_t / pow_pos_t / sqrt_d / acos_d / atan_t / atan2_t) the same way; for the two one-float
// coordinate maps of the table BRDFs, u = (float)(2 acos((double)c) / pi) (dj_brdf.h:2158-2162) and u = (float)sqrt(2 atan(r) / pi)
// (:2151-2156), EVERY float argument is compared with libm's result: the number of floats whose rounded coordinate differs is
// printed (each such float is a place where the device and the reference may differ in the last bit of u).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <atomic>
#include <thread>
#include <vector>
#include "djb_dmath.cuh"

static double ulps(double got, double want)
{
	if (got == want) return 0.0;
	if (std::isnan(got) && std::isnan(want)) return 0.0;
	int e;
	frexp(want, &e);
	return fabs(got - want) / ldexp(1.0, e - 53);
}

int main(int argc, char **argv)
{
	// `dmath_check full`: every float in the two coordinate maps (~40 s on 8 cores; profiles/r02_n_dmath_exhaustive.txt holds that
	// run); the default strides through them 16 apart from an odd offset
	const uint64_t step = argc > 1 && !strcmp(argv[1], "full") ? 1 : 16;
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
	int rc = worst_e <= 4.0 && worst_l <= 4.0 && bad == 0 ? 0 : 1;

	// ---- table-driven set ----
	const double *T = dm_table_host();
	double w_et = 0, w_lt = 0, w_pw = 0, w_sq = 0, w_ac = 0, w_at = 0, w_a2 = 0, w_sc = 0;
	for (int k = 0; k < N; ++k) {
		const double u = rnd(), v = rnd();
		const double x = (k & 1) ? (u * 2.0 - 1.0) * 700.0 : (u * 2.0 - 1.0) * 2.0;
		w_et = fmax(w_et, ulps(exp_t(x, T), exp(x)));
		const double y = (k & 1) ? pow(10.0, (u * 2.0 - 1.0) * 300.0) : 0.5 + u;
		const double ly = log(y);
		w_lt = fmax(w_lt, fabs(log_t(y, T) - ly) / (ldexp(1.0, -52) * fmax(1.0, fabs(ly)))); // absolute criterion
		// pow as the kernels use it: base in (0, 4), exponent up to 500, relative error in units of 2^-52 (1 + |y log x|)
		const double pb = 1e-3 + 4.0 * u, pe = (v * 2.0 - 1.0) * ((k & 2) ? 500.0 : 3.0), pw = pow(pb, pe);
		if (pw > 1e-300 && pw < 1e300)
			w_pw = fmax(w_pw, fabs(pow_pos_t(pb, pe, T) - pw) / (pw * ldexp(1.0, -52) * (1.0 + fabs(pe * log(pb)))));
		const double sz = (k & 1) ? pow(10.0, (u * 2.0 - 1.0) * 29.0) : u;
		w_sq = fmax(w_sq, ulps(sqrt_d(sz), sqrt(sz)));
		const double ca = (k & 1) ? u * 2.0 - 1.0 : 1.0 - u * u * u * u; // dense near 1
		w_ac = fmax(w_ac, ulps(acos_d(ca), acos(ca)));
		const double ta = (k & 1) ? pow(10.0, (u * 2.0 - 1.0) * 25.0) * (v < 0.5 ? -1.0 : 1.0) : (u * 2.0 - 1.0) * 4.0;
		w_at = fmax(w_at, ulps(atan_t(ta, T), atan(ta)));
		const double ay = (u * 2.0 - 1.0) * ((k & 4) ? 1.0 : 1e-3), bx = (v * 2.0 - 1.0) * ((k & 8) ? 1.0 : 1e-3);
		w_a2 = fmax(w_a2, ulps(atan2_t(ay, bx, T), atan2(ay, bx)));
		const double sa = (k & 1) ? (u * 2.0 - 1.0) * 1e5 : (u * 2.0 - 1.0) * 7.0;
		double ss, cc;
		sincos_d(sa, &ss, &cc);
		w_sc = fmax(w_sc, fmax(ulps(ss, sin(sa)), ulps(cc, cos(sa))));
	}
	const double e2[][2] = {{0.0, 1.0}, {-0.0, 1.0}, {0.0, -1.0}, {-0.0, -1.0}, {1.0, 0.0}, {-1.0, 0.0}, {1.0, -0.0}, {-1.0, -0.0},
	                        {0.0, 0.0}, {1.0, 1.0}, {-1.0, -1.0}, {INFINITY, 1.0}, {1.0, INFINITY}, {NAN, 1.0}, {1e-40, 1e-39}, {3.0, -4.0}};
	int bad2 = 0;
	for (auto &q : e2) {
		const double a = atan2_t(q[0], q[1], T), b = atan2(q[0], q[1]);
		if (!((a == b && std::signbit(a) == std::signbit(b)) || (std::isnan(a) && std::isnan(b)) || ulps(a, b) <= 2)) ++bad2;
	}
	const double e1[] = {0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1.0000001, NAN, 1e-300, INFINITY, -INFINITY, 1e31};
	for (double x : e1) {
		const double a = acos_d(x), b = acos(x), c = atan_t(x, T), d = atan(x);
		if (!((a == b) || (std::isnan(a) && std::isnan(b)) || ulps(a, b) <= 2)) ++bad2;
		if (!((c == d && std::signbit(c) == std::signbit(d)) || (std::isnan(c) && std::isnan(d)) || ulps(c, d) <= 2)) ++bad2;
	}
	printf("exp_t max_ulp %.3f  log_t max_abs %.3f  pow_pos_t %.3f  sqrt_d %.3f  acos_d %.3f  atan_t %.3f  atan2_t %.3f  sincos_d %.3f  edge_mismatches %d\n",
	       w_et, w_lt, w_pw, w_sq, w_ac, w_at, w_a2, w_sc, bad2);
	if (!(w_et <= 2.0 && w_lt <= 2.0 && w_pw <= 8.0 && w_sq <= 1.0 && w_ac <= 2.0 && w_at <= 2.0 && w_a2 <= 2.0 && w_sc <= 2.0 && bad2 == 0)) rc = 1;

	// ---- the two coordinate maps, every float ----
	// name, first and last bit pattern (non-negative floats), also the negated arguments?
	struct Map { const char *name; uint32_t lo, hi; bool both_signs; float (*want)(float); float (*got)(float); };
	static const double pi_f = (double)(float)M_PI;
	static const float r2d = (float)(180.0 / M_PI);
	const Map maps[] = {
		{"acos_coord", 0u, 0x3f800000u, true, [](float c) { return (float)(2.0 * acos((double)c) / pi_f); }, [](float c) { return acos_coord(c); }},
		{"acos_coord_pi", 0u, 0x3f800000u, true, [](float c) { return (float)(2.0 * acos((double)c) / M_PI); }, [](float c) { return acos_coord_pi(c); }},
		{"atan_coord", 0u, 0x7f7fffffu, false, [](float r) { return (float)sqrt(2.0 * atan((double)r) / pi_f); },
		 [](float r) { return atan_coord(r, dm_table_host()); }},
		{"(float)acos", 0u, 0x3f800000u, true, [](float c) { return (float)acos((double)c); }, [](float c) { return (float)acos_d((double)c); }},
		{"(float)atan(sqrt)", 0u, 0x7f7fffffu, false, [](float q) { return (float)atan(sqrt((double)q)); },
		 [](float q) { return (float)atan_t(sqrt_d((double)q), dm_table_host()); }},
		{"utia theta (degrees)", 0u, 0x3f800000u, true, [](float c) { return (float)((double)r2d * acos((double)c)); },
		 [](float c) { return (float)((double)r2d * acos_d((double)c)); }},
		{"(float)sin", 0u, 0x47c35000u, true, [](float x) { return (float)sin((double)x); },
		 [](float x) { double s, c; sincos_d((double)x, &s, &c); return (float)s; }},
		{"(float)cos", 0u, 0x47c35000u, true, [](float x) { return (float)cos((double)x); },
		 [](float x) { double s, c; sincos_d((double)x, &s, &c); return (float)c; }},
		{"pow5_one_minus", 0u, 0x3f800000u, false, [](float c) { return (float)pow(1.0 - (double)c, 5.0); }, [](float c) { return pow5_one_minus(c); }},
	};
	unsigned nt = std::thread::hardware_concurrency();
	nt = nt < 1 ? 1 : (nt > 64 ? 64 : nt);
	for (const Map &mp : maps) {
		std::atomic<long long> flips{0}, count{0};
		std::vector<std::thread> pool;
		for (unsigned t = 0; t < nt; ++t)
			pool.emplace_back([&, t]() {
				long long f = 0, n = 0;
				for (uint64_t bits = mp.lo + t * step + (step > 1 ? 5 : 0); bits <= mp.hi; bits += nt * step)
					for (int sgn = 0; sgn < (mp.both_signs ? 2 : 1); ++sgn) {
						float c;
						const uint32_t bb = (uint32_t)bits | (sgn ? 0x80000000u : 0u);
						memcpy(&c, &bb, 4);
						const float w = mp.want(c), g = mp.got(c);
						f += memcmp(&w, &g, 4) != 0 && !(w != w && g != g);
						++n;
					}
				flips += f;
				count += n;
			});
		for (auto &th : pool) th.join();
		printf("map %-22s %lld of %lld floats differ\n", mp.name, flips.load(), count.load());
		if (flips.load() > 64) rc = 1;
	}
	return rc;
}
