// glibcf_check.cpp -- pins dj_brdf_b200/csrc/djb_glibcf.h (the restatement of glibc's logf / expf / powf that the device
// sampling path runs) to the platform's libm, bit for bit.
//
//   g++ -O2 -ffp-contract=off -pthread tests/cpp/glibcf_check.cpp -o glibcf_check && ./glibcf_check [stride]
//
// stride = 1 walks EVERY float of the domains the sampling path uses (dj_brdf.h:695, 1917, 1935):
//   logf:  every positive normal float                                   (2.1e9 arguments)
//   expf:  every float with |x| < 88, both signs                         (2.2e9 arguments)
//   powf:  every float x in [1e-6, 1] under 8 exponents y in [0.49, 1]   (the range of the reference's `fit`), plus a
//          broad pseudo-random sweep of normal x and |y| in [1e-3, 64]
// and takes a few minutes on 8 cores; the test suite runs it with a large stride (seconds).  Prints one line per function:
// arguments checked, mismatches.  Exit code 0 iff there is no mismatch.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../dj_brdf_b200/csrc/djb_glibcf.h"

using namespace djb200;

static inline uint64_t mix(uint64_t z)
{
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

template <class F>
static void parallel(uint64_t lo, uint64_t hi, uint64_t stride, F body)
{
	unsigned nt = std::thread::hardware_concurrency();
	if (nt == 0) nt = 4;
	std::vector<std::thread> th;
	const uint64_t n = (hi - lo + stride - 1) / stride;
	for (unsigned t = 0; t < nt; ++t)
		th.emplace_back([=]() {
			const uint64_t a = n * t / nt, b = n * (t + 1) / nt;
			for (uint64_t j = a; j < b; ++j) body(lo + j * stride);
		});
	for (auto &x : th) x.join();
}

int main(int argc, char **argv)
{
	const uint64_t stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1021;
	const GlfTablePtr T = {g_glf_table};
	const GlfHot H = glf_hot();
	int bad_total = 0;

	{ // logf over the positive normal floats
		std::atomic<uint64_t> n{0}, bad{0};
		parallel(0x00800000u, 0x7f800000u, stride, [&](uint64_t b) {
			const float x = glf_as_f32((uint32_t)b);
			if (!glf_logf_ok(x)) { bad++; return; }
			const float want = logf(x), got = glf_logf(T, H, x);
			n.fetch_add(1, std::memory_order_relaxed);
			if (glf_as_u32(want) != glf_as_u32(got)) {
				if (bad++ < 5) printf("  logf(%a): libm %a, restated %a\n", x, want, got);
			}
		});
		printf("logf: %llu arguments, %llu mismatches\n", (unsigned long long)n, (unsigned long long)bad);
		bad_total += bad != 0;
	}
	{ // expf over |x| < 88
		std::atomic<uint64_t> n{0}, bad{0};
		for (uint32_t sign = 0; sign < 2; ++sign)
			parallel(0u, 0x42b00000u, stride, [&](uint64_t b) {
				const float x = glf_as_f32((uint32_t)b | (sign << 31));
				if (!glf_expf_ok(x)) { bad++; return; }
				const float want = expf(x), got = glf_expf(T, H, x);
				n.fetch_add(1, std::memory_order_relaxed);
				if (glf_as_u32(want) != glf_as_u32(got)) {
					if (bad++ < 5) printf("  expf(%a): libm %a, restated %a\n", x, want, got);
				}
			});
		printf("expf: %llu arguments, %llu mismatches\n", (unsigned long long)n, (unsigned long long)bad);
		bad_total += bad != 0;
	}
	{ // powf: the sampling path's domain, then a broad sweep
		std::atomic<uint64_t> n{0}, bad{0}, skipped{0};
		const uint32_t lo = glf_as_u32(1e-6f), hi = glf_as_u32(1.0f) + 1;
		parallel(lo, hi, stride, [&](uint64_t b) {
			const float x = glf_as_f32((uint32_t)b);
			for (int j = 0; j < 8; ++j) {
				const float y = 0.49f + 0.51f * (float)((mix(b * 8 + j) >> 40) * (1.0 / 16777216.0));
				bool ok;
				const float got = glf_powf(T, x, y, ok);
				if (!glf_powf_ok(x, y) || !ok) { skipped++; continue; }
				const float want = powf(x, y);
				n.fetch_add(1, std::memory_order_relaxed);
				if (glf_as_u32(want) != glf_as_u32(got)) {
					if (bad++ < 5) printf("  powf(%a, %a): libm %a, restated %a\n", x, y, want, got);
				}
			}
		});
		const uint64_t broad = 4000000000ull / stride + 100000;
		parallel(0, broad, 1, [&](uint64_t j) {
			const uint64_t h = mix(j * 2 + 1), g = mix(j * 2 + 2);
			const float x = glf_as_f32(0x00800000u + (uint32_t)(h % (0x7f800000u - 0x00800000u)));
			float y = ldexpf(1.0f + (float)((g >> 40) * (1.0 / 16777216.0)), (int)((g >> 8) % 17) - 10);
			if (g & 1) y = -y;
			bool ok;
			const float got = glf_powf(T, x, y, ok);
			if (!glf_powf_ok(x, y) || !ok) { skipped++; return; }
			const float want = powf(x, y);
			n.fetch_add(1, std::memory_order_relaxed);
			if (glf_as_u32(want) != glf_as_u32(got)) {
				if (bad++ < 5) printf("  powf(%a, %a): libm %a, restated %a\n", x, y, want, got);
			}
		});
		printf("powf: %llu arguments, %llu mismatches (%llu outside the restated branch)\n", (unsigned long long)n,
		       (unsigned long long)bad, (unsigned long long)skipped);
		bad_total += bad != 0;
		if (skipped * 2 > n) { printf("powf: too many skipped arguments\n"); bad_total++; }
	}
	return bad_total ? 1 : 0;
}
