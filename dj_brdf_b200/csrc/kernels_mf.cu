// kernels_mf.cu -- batched microfacet eval / evalp / pdf / sample / evalp_is (SURVEY.md rows E1-E10,
// P1, S1-S5).  One thread owns one (wi, wo) pair and walks the params blocks staged in shared
// memory, so a pair is read from HBM once however many materials it is evaluated under; the
// half vector and the reciprocals that do not depend on the material are hoisted.
//
// HBM traffic per (pair, material): (24 + 12 M) / M bytes for eval -- 13.5 B at M = 16.
// The kernels are bound by instruction issue, not by HBM: matching the reference to the bit
// costs ~190 (GGX eval) to ~1650 (Beckmann sample) warp instructions per result.  Three kernel families:
//   mf_lean_kernel          lean FP32 tier (djb_lean.cuh), every query, params BROADCAST / PER_PAIR / built per pair
//                           from LEAN texels (PSRC_LEAN: mitsuba/dj_beckmannconductor.cpp:283-314 fused in front);
//   mf_beck_compact_kernel  Beckmann eval / evalp / pdf over several materials with warp-level work compaction;
//   mf_broadcast_kernel / mf_perpair_kernel   mirrored-rounding tier (djb_device.cuh): the other Fresnel terms, A/B tests.
#include <cstdlib>
#include <cstring>

#include "djb_internal.h"
#include "djb_lean.cuh"

namespace djb200 {

std::atomic<int> g_force_generic{getenv("DJB200_MF_GENERIC") != nullptr ? 1 : 0};
// 0 (or DJB200_MF_NOCOMPACT=1): Beckmann BROADCAST eval / evalp / pdf stay on mf_lean_kernel (A/B tests)
std::atomic<int> g_beck_compact{getenv("DJB200_MF_NOCOMPACT") != nullptr ? 0 : 1};
// 1 (default): eval / evalp / pdf of the lean kernels run the 1e-5 tier (djb_lean.cuh, "The 1e-5 tier"); 0 (djb200_set_precision(
// DJB200_PRECISION_REFERENCE_BITS), or DJB200_PRECISION=bits in the environment): every query reproduces the reference's floats
static int initial_fast_tier()
{
	const char *e = getenv("DJB200_PRECISION");
	return e && (!strcmp(e, "bits") || !strcmp(e, "0")) ? 0 : 1;
}
std::atomic<int> g_fast_tier{initial_fast_tier()};

constexpr int MF_THREADS = 256;
constexpr int MF_MAX_SMEM_PARAMS = 256; // 12 KB of shared memory
constexpr int MF_MAX_SMEM_SPLINE = 256; // points, 3 KB

struct MfKernelArgs {
	int shadow, fresnel_kind;
	FresnelDev fr;
	const Params *params; // device blocks, or NULL: the blocks are in inline_params
	int n_params;
	Params inline_params[MF_INLINE_PARAMS];
	const float *a, *b;
	long long n, out_stride;
	float *out0, *out1, *out2;
	// PSRC_LEAN: per-pair params from LEAN texels + base roughness (mitsuba/dj_beckmannconductor.cpp:283-314)
	const float *lean_E, *lean_alpha;
	float lean_alpha0[3];
	LeanShadingCfg lean_cfg;
	float *params_out; // optional: the constructed blocks (djb200_lean_shading_params)
};

// where a thread's params block comes from
enum { PSRC_BROADCAST = 0, PSRC_PER_PAIR = 1, PSRC_LEAN = 2 };

template <int PSRC>
DJB_DEV Params pair_params(const MfKernelArgs &A, long long k)
{
	Params p;
	if (PSRC == PSRC_PER_PAIR) {
		const float4 *pp = reinterpret_cast<const float4 *>(A.params + k); // 48 B blocks: 16-B aligned
		const float4 q0 = pp[0], q1 = pp[1], q2 = pp[2];
		p.nx = q0.x; p.ny = q0.y; p.nz = q0.z; p.a1 = q0.w;
		p.a2 = q1.x; p.phi_a = q1.y; p.ax = q1.z; p.ay = q1.w;
		p.rho = q2.x; p.srho = q2.y; p.tx = q2.z; p.ty = q2.w;
	} else {
		float a1 = A.lean_alpha0[0], a2 = A.lean_alpha0[1], phi = A.lean_alpha0[2];
		if (A.lean_alpha) {
			a1 = A.lean_alpha[3 * k];
			a2 = A.lean_alpha[3 * k + 1];
			phi = A.lean_alpha[3 * k + 2];
		}
		const float *E = A.lean_E + 5 * k;
		lean_shading_params(A.lean_cfg, a1, a2, phi, E[0], E[1], E[2], E[3], E[4], p);
	}
	return p;
}

// evalp with a run-time Fresnel kind (uniform across the grid, so the switch never diverges)
template <int NDF>
DJB_DEV V3 evalp_rt(const Params &p, int fk, const FresnelDev &f, bool shadow, V3 i, V3 o, V3 h)
{
	float G = mf_gaf<NDF>(p, shadow, i, o);
	if (G > 0.0f) {
		float cd = sat_ref(dot(o, h));
		V3 Fr = fresnel_rt(fk, f, cd);
		float Dn = mf_ndf<NDF>(p, h);
		return scale(__fdiv_rn(Dn * G, 4.0f * o.z), Fr);
	}
	return mk(0.f, 0.f, 0.f);
}

template <int NDF>
DJB_DEV V3 evalp_is_rt(const Params &p, int fk, const FresnelDev &f, bool shadow, float u1, float u2, V3 o,
                       V3 &i_out, float &pdf_out)
{
	V3 i = mf_sample<NDF>(p, u1, u2, o);
	V3 h = normalize(i + o);
	float G = mf_gaf<NDF>(p, shadow, i, o);
	pdf_out = 0.0f;
	i_out = mk(0.f, 0.f, 0.f);
	if (G > 0.0f) {
		float cd = sat_ref(dot(o, h));
		i_out = i;
		V3 Fr = fresnel_rt(fk, f, cd);
		float g1 = mf_g1<NDF>(p, o);
		pdf_out = __fdiv_rn(mf_vndf<NDF>(p, h, o), 4.0f * cd);
		return scale(G / g1, Fr);
	}
	return mk(0.f, 0.f, 0.f);
}

DJB_DEV void st3(float *p, long long k, V3 v)
{
	p[3 * k] = v.x;
	p[3 * k + 1] = v.y;
	p[3 * k + 2] = v.z;
}

// one (pair, params) query; `h` and `inv_iz` are material independent and precomputed by the caller
template <int NDF, int OP>
DJB_DEV void mf_query(const MfKernelArgs &A, const FresnelDev &fr, const Params &p, long long slot, V3 va, V3 o,
                      V3 h, float inv_iz)
{
	const bool shadow = A.shadow != 0;
	if (OP == OP_EVAL) {
		V3 e = evalp_rt<NDF>(p, A.fresnel_kind, fr, shadow, va, o, h);
		st3(A.out0, slot, scale(inv_iz, e)); // evalp / i.z, dj_brdf.h:1554
	} else if (OP == OP_EVALP) {
		st3(A.out0, slot, evalp_rt<NDF>(p, A.fresnel_kind, fr, shadow, va, o, h));
	} else if (OP == OP_PDF) {
		A.out0[slot] = mf_pdf<NDF>(p, shadow, va, o, h);
	} else if (OP == OP_SAMPLE) {
		st3(A.out0, slot, mf_sample<NDF>(p, va.x, va.y, o));
	} else {
		V3 iv;
		float pdf;
		V3 w = evalp_is_rt<NDF>(p, A.fresnel_kind, fr, shadow, va.x, va.y, o, iv, pdf);
		if (A.out0) st3(A.out0, slot, w);
		if (A.out1) st3(A.out1, slot, iv);
		if (A.out2) A.out2[slot] = pdf;
	}
}

template <int NDF, int OP>
DJB_DEV void load_pair(const MfKernelArgs &A, long long k, V3 &va, V3 &o, V3 &h, float &inv_iz)
{
	constexpr bool uses_u = (OP == OP_SAMPLE || OP == OP_EVALP_IS);
	if (uses_u) {
		float2 u = reinterpret_cast<const float2 *>(A.a)[k];
		va = mk(u.x, u.y, 0.f);
	} else {
		va = mk(A.a[3 * k], A.a[3 * k + 1], A.a[3 * k + 2]);
	}
	o = mk(A.b[3 * k], A.b[3 * k + 1], A.b[3 * k + 2]);
	h = mk(0.f, 0.f, 0.f);
	inv_iz = 0.f;
	if (!uses_u) {
		h = normalize(va + o);
		if (OP == OP_EVAL) inv_iz = rcp_via_double(va.z);
	}
}

// BROADCAST layout: every pair under every params block; output (m, k) at m * out_stride + k.
template <int NDF, int OP>
__global__ void __launch_bounds__(MF_THREADS) mf_broadcast_kernel(const __grid_constant__ MfKernelArgs A)
{
	__shared__ Params s_params[MF_MAX_SMEM_PARAMS];
	__shared__ float s_spline[3 * MF_MAX_SMEM_SPLINE];

	// stage the per-material parameters (48 B each) and the Fresnel spline once per CTA
	{
		const float *src = reinterpret_cast<const float *>(A.params ? A.params : A.inline_params);
		float *dst = reinterpret_cast<float *>(s_params);
		for (int t = threadIdx.x; t < A.n_params * 12; t += blockDim.x) dst[t] = src[t];
	}
	FresnelDev fr = A.fr;
	if (A.fresnel_kind == FK_SPLINE && fr.npts <= MF_MAX_SMEM_SPLINE) {
		for (int t = threadIdx.x; t < fr.npts * 3; t += blockDim.x) s_spline[t] = A.fr.pts[t];
		fr.pts = s_spline;
	}
	__syncthreads();

	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < A.n; k += stride) {
		V3 va, o, h;
		float inv_iz;
		load_pair<NDF, OP>(A, k, va, o, h, inv_iz);
		for (int m = 0; m < A.n_params; ++m)
			mf_query<NDF, OP>(A, fr, s_params[m], (long long)m * A.out_stride + k, va, o, h, inv_iz);
	}
}

// The lean FP32 path (djb_lean.cuh) for an ideal or Schlick Fresnel term, every query, both params layouts.
// Same results as the mirrored kernels below (tests compare the two at full size), about half the instructions.
// resident CTAs per SM the register allocation aims for: the Beckmann sampling kernel is a long dependent chain per thread
// (stall_wait 39 % at 4 CTAs / 54 registers), so it is held to 51 registers = 5 CTAs
#ifndef DJB200_BSAMPLE_MINB
#define DJB200_BSAMPLE_MINB 5
#endif
#ifndef DJB200_FSAMPLE_MINB
#define DJB200_FSAMPLE_MINB 4
#endif
#ifndef DJB200_LEANSRC_MINB
#define DJB200_LEANSRC_MINB 4
#endif
constexpr int lean_min_blocks(int ndf, int op, int psrc, bool fast)
{
	// PSRC_LEAN: the params construction in front of the query (double sincos / atan / square roots, djb_dmath.cuh) would otherwise
	// take 70 - 76 registers
	if (psrc == 2 /* PSRC_LEAN */) return DJB200_LEANSRC_MINB;
	return (ndf == NDF_BECKMANN && op == OP_SAMPLE && psrc == 0 /* PSRC_BROADCAST */) ? (fast ? DJB200_FSAMPLE_MINB : DJB200_BSAMPLE_MINB) : 1;
}

template <int NDF, int FK, int OP, int PSRC, bool FAST>
__global__ void __launch_bounds__(MF_THREADS, lean_min_blocks(NDF, OP, PSRC, FAST)) mf_lean_kernel(const __grid_constant__ MfKernelArgs A)
{
	constexpr bool uses_u = (OP == OP_SAMPLE || OP == OP_EVALP_IS);
	constexpr bool PERPAIR = PSRC != PSRC_BROADCAST;
	__shared__ ParamsX s_params[PERPAIR ? 1 : MF_MAX_SMEM_PARAMS];
	__shared__ float2 s_exp2[64]; // 2^(j/64) as float-float, for the Beckmann exponentials
	__shared__ __align__(16) uint64_t s_glf_words[GLF_TABLE_WORDS]; // glibc's logf / powf / expf tables (Beckmann sampling path)
	uint32_t glf_addr = (uint32_t)__cvta_generic_to_shared(s_glf_words);
	asm volatile("" : "+r"(glf_addr)); // opaque: kept in a register instead of being re-derived (4 uniform instructions) at every lookup
	if (!PERPAIR)
		for (int t = threadIdx.x; t < A.n_params; t += blockDim.x)
			s_params[t] = extend_params(A.params ? A.params[t] : A.inline_params[t]);
	if (NDF == NDF_BECKMANN && threadIdx.x < 64) s_exp2[threadIdx.x] = g_exp2_64[threadIdx.x];
	if (NDF == NDF_BECKMANN && uses_u && threadIdx.x < GLF_TABLE_WORDS) s_glf_words[threadIdx.x] = g_glf_table[threadIdx.x];
	__syncthreads();
	GlfCtx s_glf;
	s_glf.T.s = glf_addr;
	if (NDF == NDF_BECKMANN && uses_u) s_glf.H = glf_hot(s_glf.T); // three constants held in registers for the whole kernel
	const FresnelDev fr = A.fr;
	const bool shadow = A.shadow != 0;
	// 1e-5 tier of pdf: every material a centred lobe => the shadowing gate needs no sigma(i) (fast_pdf_try)
	bool centred = FAST && OP == OP_PDF && !PERPAIR && shadow;
	if (FAST && OP == OP_PDF && !PERPAIR)
		for (int t = 0; t < A.n_params; ++t) centred = centred && params_centred(s_params[t].p);
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
		PairX c;
		if (!uses_u) c = make_pair<uses_u ? OP_EVAL : OP>(va, o);
		// sampling: the clamped uniforms and everything that follows from u2 alone, once per pair
		float u1c = 0.0f;
		SampleU2 su2;
		if (uses_u) {
			u1c = sample_clamp_u(va.x);
			su2 = lean_sample_u2<NDF>(s_glf, sample_clamp_u(va.y));
		}
		auto one = [&](const ParamsX &mx, long long slot) {
			if (OP == OP_PDF) {
				A.out0[slot] = FAST ? fast_pdf<NDF>(s_exp2, mx, shadow, c, centred) : lean_pdf<NDF>(s_exp2, mx, shadow, c);
			} else if (OP == OP_SAMPLE) {
				st3(A.out0, slot, FAST ? fast_sample<NDF>(s_exp2, s_glf, mx.p, u1c, su2, o) : lean_sample<NDF>(s_exp2, s_glf, mx.p, u1c, su2, o));
			} else if (OP == OP_EVALP_IS) {
				V3 iv;
				float pdf;
				const V3 w = lean_evalp_is<NDF, FK>(s_exp2, s_glf, mx, fr, shadow, u1c, su2, o, iv, pdf);
				if (A.out0) st3(A.out0, slot, w);
				if (A.out1) st3(A.out1, slot, iv);
				if (A.out2) A.out2[slot] = pdf;
			} else {
				constexpr int EOP = OP == OP_EVALP ? OP_EVALP : OP_EVAL;
				st3(A.out0, slot, FAST ? fast_evalp<NDF, FK, EOP>(s_exp2, mx, fr, shadow, c) : lean_evalp<NDF, FK, EOP>(s_exp2, mx, fr, shadow, c));
			}
		};
		if (PERPAIR) {
			one(extend_params(pair_params<PERPAIR ? PSRC : PSRC_PER_PAIR>(A, k)), k);
		} else {
			for (int m = 0; m < A.n_params; ++m) one(s_params[m], (long long)m * A.out_stride + k);
		}
	}
}

// Beckmann eval / evalp / pdf, BROADCAST layout, with the expensive work COMPACTED across the warp.
// For a material whose lobe is narrow, D underflows to zero for most random pairs, and the result is then +0 without the
// exponential and without the two projected areas (sqrt, reciprocal, exp and erf each).  In mf_lean_kernel the lanes that
// do need them run with the others idle -- measured 18 of 32 lanes active per instruction on the benchmark's 16 materials.
// Here every lane does only the cheap test for its own pair under every material (the squared standard-space slope radius
// r2; r2 > 103.5 <=> D == 0), lanes that pass push a work item (source lane, material, r2) into a per-warp queue in shared
// memory, and whenever 32 items are waiting the whole warp evaluates them, one item per lane, reading the source lane's
// pair from shared memory.  Results are the same floats: the same functions run on the same operands, on another lane.
// Tried and dropped (round 2): screening all 16 materials into two per-lane bit masks first (no ballots inside the material
// loop), one warp scan to scatter 16-bit items, r2 recomputed by the finishing lane -- bit-identical, but slower (1e-5 tier:
// eval 17.0 -> 18.7 ms, pdf 17.3 -> 17.7 ms; exact tier 24.9 -> 29.2 ms): interleaving the finishing batches with the
// screening of the next materials hides their latencies, separated phases do not.
struct PairS { float4 a, b, c, d; }; // i.xyz o.x | o.yz h.xy | h.z den rcp_den inv_iz | o.h den_ok c4 rcp_c4

#ifndef DJB200_COMPACT_MINB
#define DJB200_COMPACT_MINB 1
#endif
template <int FK, int OP, bool FAST>
__global__ void __launch_bounds__(MF_THREADS, DJB200_COMPACT_MINB) mf_beck_compact_kernel(MfKernelArgs A)
{
	constexpr int NDF = NDF_BECKMANN;
	constexpr int WARPS = MF_THREADS / 32;
	__shared__ ParamsX s_params[MF_MAX_SMEM_PARAMS];
	__shared__ float2 s_exp2[64];
	__shared__ PairS s_pair[MF_THREADS];
	__shared__ uint2 s_q[WARPS][64], s_qs[WARPS][64];
	for (int t = threadIdx.x; t < A.n_params; t += blockDim.x)
		s_params[t] = extend_params(A.params ? A.params[t] : A.inline_params[t]);
	if (threadIdx.x < 64) s_exp2[threadIdx.x] = g_exp2_64[threadIdx.x];
	for (int t = threadIdx.x; t < WARPS * 64; t += blockDim.x) (&s_q[0][0])[t] = (&s_qs[0][0])[t] = make_uint2(0u, 0u);
	__syncthreads();
	const FresnelDev fr = A.fr;
	const bool shadow = A.shadow != 0;
	bool centred = FAST && OP == OP_PDF && shadow; // 1e-5 tier of pdf over centred lobes: no sigma(i) for the gate (fast_pdf_try)
	if (FAST && OP == OP_PDF)
		for (int t = 0; t < A.n_params; ++t) centred = centred && params_centred(s_params[t].p);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
	uint2 *q = s_q[warp], *qs = s_qs[warp];
	PairS *pairs = s_pair + warp * 32;

	// one queued item, run by whichever lane picked it up: D from r2, then (unless D == 0) shadowing, Fresnel, quotient.
	// Returns true when the 1e-5 tier declines the item (the reference's G is ill-conditioned for it): nothing was written, the
	// caller moves the item to the slow queue, where every item takes the exact functions
	auto finish = [&](uint2 item, long long kb) -> bool {
		const int src = item.x & 31, m = item.x >> 8;
		const PairS s = pairs[src];
		PairX c;
		c.i = mk(s.a.x, s.a.y, s.a.z);
		c.o = mk(s.a.w, s.b.x, s.b.y);
		c.h = mk(s.b.z, s.b.w, s.c.x);
		c.den = s.c.y; c.rcp_den = s.c.z; c.inv_iz = s.c.w;
		c.kh = s.d.x; c.cd = sat_ref(c.kh); c.den_ok = s.d.y == 1.0f || s.d.y == 3.0f; c.both_up = s.d.y >= 2.0f; c.c4 = s.d.z; c.rcp_c4 = s.d.w;
		const ParamsX &mx = s_params[m];
		const long long slot = (long long)m * A.out_stride + kb + src;
		const float r2 = __uint_as_float(item.y);
		if (FAST && !(item.x & 0xC0u) && r2 <= FAST_BECK_R2_MAX) { // the 1e-5 tier; the underflow tail (the slow queue) stays exact
			bool ill;
			if (OP == OP_PDF) {
				const float v = fast_pdf_try<NDF>(mx, shadow, c, r2, ill, centred);
				if (!ill) A.out0[slot] = v;
			} else {
				const V3 v = fast_evalp_try<NDF, FK, OP>(mx, fr, shadow, c, r2, ill);
				if (!ill) st3(A.out0, slot, v);
			}
			return ill;
		}
		// item.y: r2, or the NaN-free marker "not facing" (D == 0 with a non-positive denominator: the rare literal case)
		const float Dn = (item.x & 0x80u) ? 0.0f : lean_ndf_from_r2<NDF>(s_exp2, mx, c, r2);
		if (OP == OP_PDF) A.out0[slot] = lean_skip(Dn, c) ? 0.0f : lean_pdf_tail<NDF>(s_exp2, mx.p, shadow, c, Dn);
		else st3(A.out0, slot, lean_skip(Dn, c) ? lean_zero<OP>(c) : lean_evalp_tail<NDF, FK, OP>(s_exp2, mx.p, fr, shadow, c, Dn));
		return false;
	};

	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long kb = (long long)blockIdx.x * blockDim.x + warp * 32; kb < A.n; kb += stride) { // warp-uniform
		const long long k = kb + lane;
		const bool valid = k < A.n;
		V3 va = mk(0.f, 0.f, 1.f), o = mk(0.f, 0.f, 1.f);
		if (valid) {
			va = mk(A.a[3 * k], A.a[3 * k + 1], A.a[3 * k + 2]);
			o = mk(A.b[3 * k], A.b[3 * k + 1], A.b[3 * k + 2]);
		}
		const PairX c = make_pair<OP>(va, o);
		PairS s;
		s.a = make_float4(c.i.x, c.i.y, c.i.z, c.o.x);
		s.b = make_float4(c.o.y, c.o.z, c.h.x, c.h.y);
		s.c = make_float4(c.h.z, c.den, c.rcp_den, c.inv_iz);
		s.d = make_float4(c.kh, (c.den_ok ? 1.0f : 0.0f) + (c.both_up ? 2.0f : 0.0f), c.c4, c.rcp_c4); // two flags in one slot
		__syncwarp(); // the previous round's items have all been finished: the pair slots may be overwritten
		pairs[lane] = s;
		__syncwarp();
		// D == 0 makes the result +0 when the denominator is positive (lean_skip); otherwise the item is queued.
		// Two queues: items whose exponential lands in the gradual-underflow range (r2 > 78: quotients near or below 2^-120 take
		// the guarded IEEE division and the hand-rounded subnormal path, ~10x the instructions) wait in their own queue,
		// so that those long paths also run with many lanes instead of the one or two a mixed batch would have.
		const bool zero_ok = c.den > 0.0f;
		int qn = 0, qsn = 0; // items waiting (regular / slow), warp-uniform
		for (int m = 0; m <= A.n_params; ++m) {
			const bool last = m == A.n_params; // one extra trip that only drains the queues (single call site of `finish`)
			if (!last) {
				const float r2 = lean_ndf_r2(s_params[m], c);
				const bool d_zero = !c.facing || r2 > 103.5f; // D == 0 exactly (lean_ndf / beck_p22_lean)
				const bool skip = d_zero && zero_ok;
				if (valid && skip) {
					const long long slot = (long long)m * A.out_stride + k;
					if (OP == OP_PDF) A.out0[slot] = 0.0f;
					else st3(A.out0, slot, lean_zero<OP>(c));
				}
				const bool need = valid && !skip;
				const bool slow = need && (d_zero || r2 > 78.0f);
				const unsigned mask_f = __ballot_sync(FULL, need && !slow), mask_s = __ballot_sync(FULL, slow);
				const uint2 item = make_uint2((unsigned)lane | (d_zero ? 0x80u : 0u) | ((unsigned)m << 8), __float_as_uint(r2));
				if (slow) qs[qsn + __popc(mask_s & lt)] = item;
				else if (need) q[qn + __popc(mask_f & lt)] = item;
				qn += __popc(mask_f);
				qsn += __popc(mask_s);
			}
			for (;;) { // full batches; on the last trip whatever is left (the pair slots are reused afterwards)
				uint2 *Q;
				int cnt;
				// a full slow batch goes first: a regular batch may hand up to 32 declined items over to the slow queue (64 slots)
				const bool fast = qsn < 32 && (qn >= 32 || (last && qn > 0));
				if (fast) { Q = q; cnt = qn; }
				else if (qsn >= 32 || (last && qsn > 0)) { Q = qs; cnt = qsn; }
				else break;
				__syncwarp();
				const int nb = cnt < 32 ? cnt : 32;
				const uint2 item = Q[lane];
				const uint2 rest = Q[32 + lane]; // only the first cnt - 32 are meaningful
				bool redo = false;
				if (lane < nb) redo = finish(item, kb);
				__syncwarp();
				cnt -= nb;
				if (lane < cnt) Q[lane] = rest;
				if (fast) qn = cnt; else qsn = cnt;
				if (FAST && fast) { // declined items: to the slow queue, marked "exact tier" (0x40)
					const unsigned mask_r = __ballot_sync(FULL, redo);
					if (redo) qs[qsn + __popc(mask_r & lt)] = make_uint2(item.x | 0x40u, item.y);
					qsn += __popc(mask_r);
				}
				__syncwarp();
			}
		}
	}
}

template <int FK, int OP, bool FAST>
static void launch_beck_compact(const MfKernelArgs &A, long long want, cudaStream_t st)
{
	static int resident = 0;
	if (!resident) {
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, mf_beck_compact_kernel<FK, OP, FAST>, MF_THREADS, 0);
		if (resident < 1) resident = 1;
	}
	const long long cap = (long long)sm_count() * resident;
	mf_beck_compact_kernel<FK, OP, FAST><<<(int)(want < cap ? want : cap), MF_THREADS, 0, st>>>(A);
}

// PER_PAIR layout: pair k under params block k (roughness from textures at every shading point).
template <int NDF, int OP, int PSRC>
__global__ void __launch_bounds__(MF_THREADS) mf_perpair_kernel(MfKernelArgs A)
{
	__shared__ float s_spline[3 * MF_MAX_SMEM_SPLINE];
	FresnelDev fr = A.fr;
	if (A.fresnel_kind == FK_SPLINE && fr.npts <= MF_MAX_SMEM_SPLINE) {
		for (int t = threadIdx.x; t < fr.npts * 3; t += blockDim.x) s_spline[t] = A.fr.pts[t];
		fr.pts = s_spline;
		__syncthreads();
	}
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < A.n; k += stride) {
		V3 va, o, h;
		float inv_iz;
		load_pair<NDF, OP>(A, k, va, o, h, inv_iz);
		mf_query<NDF, OP>(A, fr, pair_params<PSRC>(A, k), k, va, o, h, inv_iz);
	}
}

// the per-pair construction alone (djb200_lean_shading_params): what the plugin computes before every query
__global__ void __launch_bounds__(MF_THREADS) lean_shading_params_kernel(MfKernelArgs A)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < A.n; k += stride) {
		const Params p = pair_params<PSRC_LEAN>(A, k);
		float4 *o = reinterpret_cast<float4 *>(A.params_out + 12 * k);
		o[0] = make_float4(p.nx, p.ny, p.nz, p.a1);
		o[1] = make_float4(p.a2, p.phi_a, p.ax, p.ay);
		o[2] = make_float4(p.rho, p.srho, p.tx, p.ty);
	}
}

template <int NDF, int FK, int OP, int PSRC, bool FAST>
static int lean_grid_cap()
{
	static int resident = 0;
	if (!resident) {
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, mf_lean_kernel<NDF, FK, OP, PSRC, FAST>, MF_THREADS, 0);
		if (resident < 1) resident = 1;
	}
	return sm_count() * resident;
}

template <int NDF, int FK, int OP, int PSRC, bool FAST>
static void launch_lean_tier(const MfKernelArgs &A, long long want, cudaStream_t st)
{
	const long long cap = lean_grid_cap<NDF, FK, OP, PSRC, FAST>();
	mf_lean_kernel<NDF, FK, OP, PSRC, FAST><<<(int)(want < cap ? want : cap), MF_THREADS, 0, st>>>(A);
}
// eval / evalp / pdf / sample have the two tiers; evalp_is only the exact one
template <int NDF, int FK, int OP, int PSRC>
static void launch_lean(const MfKernelArgs &A, long long want, cudaStream_t st)
{
	// Beckmann eval / evalp / pdf with one params block per pair (PER_PAIR, LEAN texels) stay on the exact tier.  These kernels do
	// not compact their work, so a lane the fast tier declines (underflow tail, ill-conditioned G of an off-centre lobe) runs the
	// exact functions after its warp has run the fast ones.  LEAN-texel lobes are off-centre by construction -- measured on
	// bench.py's lean_shading leg: 2.48 ms against 2.01 fused, 3.2 against 2.4 through PER_PAIR blocks -- and for centred lobes
	// the two tiers take the same time there (0.72 vs 0.74 ms per 2e7 pairs: the params traffic dominates)
	constexpr bool exact_only = NDF == NDF_BECKMANN && PSRC != PSRC_BROADCAST && OP != OP_SAMPLE;
	constexpr bool has_fast = (OP == OP_EVAL || OP == OP_EVALP || OP == OP_PDF || OP == OP_SAMPLE) && !exact_only;
	if constexpr (has_fast) {
		if (g_fast_tier.load(std::memory_order_relaxed) != 0) {
			launch_lean_tier<NDF, FK, OP, PSRC, true>(A, want, st);
			return;
		}
	}
	launch_lean_tier<NDF, FK, OP, PSRC, false>(A, want, st);
}

template <int NDF, int OP>
static cudaError_t launch_T(const MfLaunch &L, cudaStream_t st)
{
	MfKernelArgs A;
	A.shadow = L.shadow;
	A.fresnel_kind = L.fresnel_kind;
	for (int k = 0; k < 6; ++k) A.fr.v[k] = L.fv[k];
	A.fr.pts = L.spline_pts;
	A.fr.npts = L.spline_n;
	A.a = L.a;
	A.b = L.b;
	A.n = L.n;
	A.out_stride = L.out_stride;
	A.lean_E = L.lean_E;
	A.lean_alpha = L.lean_alpha;
	for (int k = 0; k < 3; ++k) A.lean_alpha0[k] = L.lean_alpha0[k];
	A.lean_cfg.bias = L.lean_bias;
	A.lean_cfg.dmap_scale = L.lean_dmap_scale;
	A.lean_cfg.lean_filtering = L.lean_filtering;
	A.params_out = nullptr;
	if (L.n <= 0) return cudaSuccess;

	// persistent grid-stride launch: a whole number of waves (SM count x resident CTAs per SM)
	const long long want = (L.n + MF_THREADS - 1) / MF_THREADS;
	static int resident_bc = 0, resident_pp = 0;
	if (!resident_bc) {
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident_bc, mf_broadcast_kernel<NDF, OP>, MF_THREADS, 0);
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident_pp, mf_perpair_kernel<NDF, OP, PSRC_LEAN>, MF_THREADS, 0);
		if (resident_bc < 1) resident_bc = 1;
		if (resident_pp < 1) resident_pp = 1;
	}
	const long long cap = (long long)sm_count() * (L.layout != DJB200_PARAMS_BROADCAST ? resident_pp : resident_bc);
	const int grid = (int)(want < cap ? want : cap);
	// djb200_debug_force_generic(1) (or DJB200_MF_GENERIC=1) runs the mirrored-rounding kernels everywhere: A/B tests
	const bool force_generic = g_force_generic.load(std::memory_order_relaxed) != 0;
	// the lean kernels cover the ideal and Schlick Fresnel terms (sampling does not evaluate the Fresnel term at all)
	const bool lean = !force_generic && (OP == OP_SAMPLE || L.fresnel_kind == FK_IDEAL || L.fresnel_kind == FK_SCHLICK);
	const bool schlick = L.fresnel_kind == FK_SCHLICK && OP != OP_SAMPLE && OP != OP_PDF;

	if (L.layout == DJB200_PARAMS_PER_PAIR) {
		A.params = reinterpret_cast<const Params *>(L.params);
		A.n_params = 1;
		A.out0 = L.out0; A.out1 = L.out1; A.out2 = L.out2;
		if (lean && schlick) launch_lean<NDF, FK_SCHLICK, OP, PSRC_PER_PAIR>(A, want, st);
		else if (lean) launch_lean<NDF, FK_IDEAL, OP, PSRC_PER_PAIR>(A, want, st);
		else mf_perpair_kernel<NDF, OP, PSRC_PER_PAIR><<<grid, MF_THREADS, 0, st>>>(A);
		g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
		return cudaGetLastError();
	}
	if (L.layout == PARAMS_LEAN_SHADING) { // params built per pair from LEAN texels, fused in front of the query
		A.params = nullptr;
		A.n_params = 1;
		A.out0 = L.out0; A.out1 = L.out1; A.out2 = L.out2;
		if (lean && schlick) launch_lean<NDF, FK_SCHLICK, OP, PSRC_LEAN>(A, want, st);
		else if (lean) launch_lean<NDF, FK_IDEAL, OP, PSRC_LEAN>(A, want, st);
		else mf_perpair_kernel<NDF, OP, PSRC_LEAN><<<grid, MF_THREADS, 0, st>>>(A);
		g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
		return cudaGetLastError();
	}
	// BROADCAST: at most MF_MAX_SMEM_PARAMS blocks per launch
	const int per = (OP == OP_PDF) ? 1 : 3;
	for (int64_t m0 = 0; m0 < L.n_params; m0 += MF_MAX_SMEM_PARAMS) {
		int64_t mc = L.n_params - m0 < MF_MAX_SMEM_PARAMS ? L.n_params - m0 : MF_MAX_SMEM_PARAMS;
		if (L.params) {
			A.params = reinterpret_cast<const Params *>(L.params) + m0;
		} else { // a small set, carried by value in the kernel arguments
			A.params = nullptr;
			memcpy(A.inline_params, L.params_host, sizeof(Params) * (size_t)mc);
		}
		A.n_params = (int)mc;
		int64_t off = m0 * L.out_stride;
		A.out0 = L.out0 ? L.out0 + off * per : nullptr;
		A.out1 = L.out1 ? L.out1 + off * 3 : nullptr;
		A.out2 = L.out2 ? L.out2 + off : nullptr;
		// Beckmann eval / evalp / pdf over several materials: the warp-compacting kernel
		constexpr bool can_compact = NDF == NDF_BECKMANN && (OP == OP_EVAL || OP == OP_EVALP || OP == OP_PDF);
		bool compacted = false;
		if constexpr (can_compact) {
			if (lean && A.n_params >= 2 && g_beck_compact.load(std::memory_order_relaxed) != 0) {
				const bool fast = g_fast_tier.load(std::memory_order_relaxed) != 0;
				if (schlick && fast) launch_beck_compact<FK_SCHLICK, OP, true>(A, want, st);
				else if (schlick) launch_beck_compact<FK_SCHLICK, OP, false>(A, want, st);
				else if (fast) launch_beck_compact<FK_IDEAL, OP, true>(A, want, st);
				else launch_beck_compact<FK_IDEAL, OP, false>(A, want, st);
				compacted = true;
			}
		}
		if (!compacted) {
			if (lean && schlick) launch_lean<NDF, FK_SCHLICK, OP, PSRC_BROADCAST>(A, want, st);
			else if (lean) launch_lean<NDF, FK_IDEAL, OP, PSRC_BROADCAST>(A, want, st);
			else mf_broadcast_kernel<NDF, OP><<<grid, MF_THREADS, 0, st>>>(A);
		}
		g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) return e;
	}
	return cudaSuccess;
}

template <int NDF>
static cudaError_t launch_N(const MfLaunch &L, cudaStream_t st)
{
	switch (L.op) {
	case OP_EVAL: return launch_T<NDF, OP_EVAL>(L, st);
	case OP_EVALP: return launch_T<NDF, OP_EVALP>(L, st);
	case OP_PDF: return launch_T<NDF, OP_PDF>(L, st);
	case OP_SAMPLE: return launch_T<NDF, OP_SAMPLE>(L, st);
	case OP_EVALP_IS: return launch_T<NDF, OP_EVALP_IS>(L, st);
	}
	return cudaErrorInvalidValue;
}

cudaError_t launch_lean_shading_params(const MfLaunch &L, float *params_out, cudaStream_t st)
{
	if (L.n <= 0) return cudaSuccess;
	MfKernelArgs A;
	memset(&A, 0, sizeof A);
	A.n = L.n;
	A.lean_E = L.lean_E;
	A.lean_alpha = L.lean_alpha;
	for (int k = 0; k < 3; ++k) A.lean_alpha0[k] = L.lean_alpha0[k];
	A.lean_cfg.bias = L.lean_bias;
	A.lean_cfg.dmap_scale = L.lean_dmap_scale;
	A.lean_cfg.lean_filtering = L.lean_filtering;
	A.params_out = params_out;
	const long long want = (L.n + MF_THREADS - 1) / MF_THREADS, cap = (long long)sm_count() * 8;
	lean_shading_params_kernel<<<(int)(want < cap ? want : cap), MF_THREADS, 0, st>>>(A);
	g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
	return cudaGetLastError();
}

cudaError_t launch_microfacet(const MfLaunch &L, cudaStream_t st)
{
	if (L.ndf == NDF_GGX) return launch_N<NDF_GGX>(L, st);
	if (L.ndf == NDF_BECKMANN) return launch_N<NDF_BECKMANN>(L, st);
	return cudaErrorInvalidValue;
}

} // namespace djb200
