/*
 * reheatfunq_b200 — FP64 device math of the heat-flow anomaly posterior.
 *
 * Everything here is __host__ __device__ so that the numerics can be unit
 * tested on the build container's CPU (tests/ compile this header into a
 * test-only helper); the PRODUCT only ever calls it from the sm_100a kernels
 * in rhfq_kernels.cu.  There is no CPU execution path in the product.
 *
 * What is computed (reference file:line in /root/reference):
 *   g(a) = lgamma(v a) - n lgamma(a) + C a       include/anomaly/posterior/integrand.hpp:111-157
 *   a_max(z) by Newton                           include/anomaly/posterior/integrand.hpp:201-284
 *   I(z) = int_{amin}^inf exp(..) da             include/anomaly/posterior/integrand.hpp:381-480
 *   large-z (y -> 0) expansion                   include/anomaly/posterior/large_z.hpp:82-538
 *   max over (a,z), y_max                        include/anomaly/posterior/amax.hpp:68-133,
 *                                                include/anomaly/posterior/large_z.hpp:293-349
 * How it differs from the reference (B200-first, not a port): the reference's
 * adaptive tanh-sinh + exp-sinh rules in `a` (several hundred integrand
 * evaluations, data-dependent) are replaced by a FIXED-node rule: Newton for
 * the mode, a window where the log-integrand has dropped RHFQ_DROP nats found by
 * monotone Newton on the concave log-integrand, and two Gauss-Legendre panels
 * split at the mode.  One thread evaluates one (sample set, z) node; all
 * control flow is bounded and nearly warp-uniform.
 */
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cfloat>
#include <cstring>

#if defined(__CUDACC__)
#define RHFQ_HD __host__ __device__ __forceinline__
#define RHFQ_HDN __host__ __device__
/* Out-of-line on the device: the special-function bodies are called from many
 * places (mode search, window search, two quadrature panels, large-z branch);
 * inlining them everywhere made the node kernel ~300 KB of SASS and the warps
 * stalled on instruction fetch (ncu: stalled_no_instruction 11.7 per issue). */
#define RHFQ_HDC __host__ __device__ __noinline__
#else
#define RHFQ_HD inline
#define RHFQ_HDN inline
#define RHFQ_HDC inline
#endif

namespace rhfq {

/* Status codes (shared with include/reheatfunq_b200.h) */
enum : int {
	ST_OK = 0,
	ST_RUNTIME = 1,        /* generic failure (NaN, inf, ...)                      */
	ST_DOMAIN = 2,         /* q<=0, Qmax==inf, duplicate Qmax, bad weight          */
	ST_NOT_NORMALIZABLE = 3, /* n < v or degenerate n == v                           */
	ST_PRECISION = 4,      /* quadrature did not reach its tolerance               */
	ST_ROOT = 5            /* bracketed root search failed                         */
};

#define RHFQ_DROP 40.0

/*
 * Gauss-Legendre tables for the three node counts in use, concatenated:
 * [0,24) K=24, [24,56) K=32, [56,104) K=48, [104,124) K=20.  Host copy for the CPU unit tests
 * of this header, constant-memory copy for the kernels: every thread of a warp
 * reads the same node index at the same time, the broadcast case constant
 * memory is built for.
 */
static const double h_GLX[124] = {
#include "rhfq_gl_x.inc"
};
static const double h_GLW[124] = {
#include "rhfq_gl_w.inc"
};
#if defined(__CUDACC__)
static __constant__ double d_GLX[124] = {
#include "rhfq_gl_x.inc"
};
static __constant__ double d_GLW[124] = {
#include "rhfq_gl_w.inc"
};
#endif
#if defined(__CUDA_ARCH__)
#define GLX(j) d_GLX[j]
#define GLW(j) d_GLW[j]
#else
#define GLX(j) h_GLX[j]
#define GLW(j) h_GLW[j]
#endif

constexpr double LOG_2PI = 1.8378770664093454835606594728112;
constexpr double HALF_PI = 1.5707963267948966192313216916398;
constexpr double DBL_EPS = 2.220446049250313e-16;
constexpr double SQRT_EPS = 1.4901161193847656e-08;
#define RHFQ_INF (__builtin_huge_val())

RHFQ_HD bool is_nan(double x) { return x != x; }
RHFQ_HD bool is_inf(double x) { return fabs(x) == RHFQ_INF; }
RHFQ_HD bool is_fin(double x) { return fabs(x) < RHFQ_INF; }

/*
 * Per sample set state (SoA would not help: each task reads one record).
 */
/* SetLocals::flags */
enum { SET_Z_UNCONVERGED = 1 };   /* the z-quadrature of the norm ran out of levels (localsandnorm.hpp:221-226
                                      takes Boost's last estimate without a check; so does this path) */

struct SetLocals {
	/* Locals (locals.hpp:75-87) */
	double lp, ls, n, v, amin, Qmax;
	double h0, h1, h2, h3;
	double w, lh0, l1p_w, lv;
	/* Large-z polynomial coefficients in a (large_z.hpp:82-176): C_m(a) = sum_j cm[j] a^j */
	double c1[2], c2[3], c3[4];
	/* LocalsAndLogScale / LocalsAndNorm (localsandnorm.hpp:60-62,147-148) */
	double log_scale, ymax, ztrans;
	double S, taylor, norm;
	/* Posterior-level weight (posterior.hpp:582-625) and log(W/(Qmax_j * norm_global)) */
	double weight, log_fac;
	long long k_off;   /* offset of this set's k_i in the packed array */
	int N;
	int imax;
	int status;
	int flags;         /* SET_Z_UNCONVERGED, ... (diagnostics that are not errors in the reference) */
	/* table of H(a) = lgamma(v a) - n lgamma(a) - a v ln v for this set's (n, v), or null */
	const double* gtab;
	/* the set's Stirling coefficient block (gcoef_fill), or null */
	const double* gco;
	/* ln and sqrt of the machine epsilon of the reference's arithmetic type (large_z.hpp:314,
	 * localsandnorm.hpp:210): of double, or of long double for the "double" keyword */
	double log_eps, sqrt_eps;
};

/* ------------------------------------------------------------------ *
 *  Polygamma for x > 0 (Newton only; needs ~1e-12)                    *
 * ------------------------------------------------------------------ */
RHFQ_HD double digamma_pos(double x)
{
	double res = 0.0;
	while (x < 12.0) {
		res -= 1.0 / x;
		x += 1.0;
	}
	const double xi = 1.0 / x, x2 = xi * xi;
	const double ser = x2 * (1.0 / 12 - x2 * (1.0 / 120 - x2 * (1.0 / 252 - x2 * (1.0 / 240
	                   - x2 * (1.0 / 132 - x2 * (691.0 / 32760 - x2 * (1.0 / 12)))))));
	return res + log(x) - 0.5 * xi - ser;
}

RHFQ_HD double trigamma_pos(double x)
{
	double res = 0.0;
	while (x < 12.0) {
		res += 1.0 / (x * x);
		x += 1.0;
	}
	const double xi = 1.0 / x, x2 = xi * xi;
	const double ser = xi * x2 * (1.0 / 6 - x2 * (1.0 / 30 - x2 * (1.0 / 42 - x2 * (1.0 / 30
	                   - x2 * (5.0 / 66 - x2 * (691.0 / 2730 - x2 * (7.0 / 6)))))));
	return res + xi + 0.5 * x2 + ser;
}

/* ------------------------------------------------------------------ *
 *  g(a) = lgamma(v a) - n lgamma(a) + C a                             *
 *  a >= 12: same leading-term arrangement as integrand.hpp:137-147    *
 *  (the big a*ln(a) pieces of the two lgammas are combined            *
 *  analytically), with 7 Stirling terms instead of 4 + lgamma         *
 *  fallback: error < n * 0.03 / 12^15 ~ 2e-15 (n = 1000).             *
 *  a < 12: lgamma(a) = lgamma(a + 12) - ln (a)_12 (the rising-        *
 *  factorial shift of integrand.hpp:120-132, by 12 instead of 47),    *
 *  the two gammas evaluated separately (3 logs instead of 4).         *
 *  Everything that depends only on (n, v, C) is hoisted into GCoef:   *
 *  it is constant for all ~80 evaluations of one (set, z) node.       *
 * ------------------------------------------------------------------ */
constexpr double HALF_LOG_2PI = 0.91893853320467274178032973640562;

/*
 * lgamma-difference table.  H(a) = lgamma(v a) - n lgamma(a) - a v ln v depends on the
 * sample set only through (n, v) = prior + data count, and the quadrature in a evaluates
 * it ~50 times per (set, z) node.  It is tabulated once per distinct (n, v) of the batch,
 * together with H', H'' and H''', on [1, 1024): GTAB_CELLS uniform cells per BINADE, so that
 * the cell of a is read off the bits of the double (exponent and top seven mantissa bits)
 * and the relative cell width is constant.  All reads are cell-local two-point Hermite
 * interpolations from the records at the cell's ends:
 *   H    quintic from (H, H', H''): error h^6 H^(6) / 46080 with H^(6) ~ 24 (v - n) / a^5
 *        + 60 (1 - n) / a^6: 3e-16 n at a = 1 (as on the uniform h = 1/128 grid it replaces)
 *        and no larger in any later binade -- below the rounding of H everywhere;
 *   H'   cubic from (H', H''), H'' cubic from (H'', H'''): the Newton steps of the mode /
 *        window searches, which stop at 1e-9.
 * Round 1 tabulated [1, 12) only and used the Stirling series (a log, a division, 7 terms)
 * beyond: in the batched workloads the shape parameters of the regions spread over
 * a = 5 ... 500, the lanes of a warp sat on both sides of a = 12 and the Gauss-Legendre
 * kernel executed BOTH branches with ~15 of 32 lanes each (ncu: 18.1 active lanes per
 * instruction, `log(a)` alone 21 % of the kernel's instructions at 14.5 lanes).  With one
 * path for [1, 1024) the warp stays converged; the series remains for a >= 1024.
 */
constexpr int GTAB_CELL_BITS = 7;
constexpr int GTAB_CELLS = 1 << GTAB_CELL_BITS;         /* per binade */
constexpr int GTAB_BINADES = 10;
constexpr double GTAB_A0 = 1.0, GTAB_A1 = 1024.0;
constexpr int GTAB_SIZE = GTAB_BINADES * GTAB_CELLS + 1;
/* One 32-byte record per node: (H, H' = v psi(v a) - n psi(a) - v ln v, H'', H''').  The
 * quintic Hermite read of H needs the first three values of two neighbouring nodes: 64
 * contiguous bytes, four 16-byte loads from two 32-byte sectors -- as separate arrays it was
 * six 8-byte loads from six sectors (ncu: 5.6 long-scoreboard stalls per issued instruction
 * in the Gauss-Legendre kernel, L1 hit rate 71 %). */
constexpr int GTAB_REC = 4;
constexpr int GTAB_STRIDE = GTAB_REC * GTAB_SIZE;
struct alignas(16) GtabPair { double x, y; };
RHFQ_HD GtabPair gtab_load2(const double* p) { return *reinterpret_cast<const GtabPair*>(p); }
RHFQ_HD long long gtab_bits(double a)
{
#if defined(__CUDA_ARCH__)
	return __double_as_longlong(a);
#else
	long long b;
	std::memcpy(&b, &a, sizeof(b));
	return b;
#endif
}
RHFQ_HD double gtab_from_bits(long long b)
{
#if defined(__CUDA_ARCH__)
	return __longlong_as_double(b);
#else
	double a;
	std::memcpy(&a, &b, sizeof(a));
	return a;
#endif
}
constexpr int GTAB_SHIFT = 52 - GTAB_CELL_BITS;
RHFQ_HD double gtab_node(int j)
{
	return gtab_from_bits(((long long)j + ((long long)1023 << GTAB_CELL_BITS)) << GTAB_SHIFT);
}
/* cell j of a in [1, 1024), position s in [0, 1) inside it, cell width h */
struct GtabCell { int j; double s, h; };
RHFQ_HD GtabCell gtab_cell(double a)
{
	const long long b = gtab_bits(a);
	GtabCell c;
	c.j = (int)(b >> GTAB_SHIFT) - (1023 << GTAB_CELL_BITS);
	const double node = gtab_from_bits(b & ~(((long long)1 << GTAB_SHIFT) - 1));
	const long long E = b >> 52;                                    /* biased exponent (a > 0) */
	c.h = gtab_from_bits((E - GTAB_CELL_BITS) << 52);               /* 2^e / cells */
	c.s = (a - node) * gtab_from_bits((2046 + GTAB_CELL_BITS - E) << 52);   /* exact */
	return c;
}

/* 1 / prod_{c != a} (a - c) = (-1)^(P-1-a) / (a! (P-1-a)!) for P = 12 equispaced nodes */
#define RHFQ_LAGRANGE12_INV_DEN                                                             \
	{-1.0 / 39916800.0, 1.0 / 3628800.0, -1.0 / 725760.0, 1.0 / 241920.0, -1.0 / 120960.0,   \
	 1.0 / 86400.0, -1.0 / 86400.0, 1.0 / 120960.0, -1.0 / 241920.0, 1.0 / 725760.0,        \
	 -1.0 / 3628800.0, 1.0 / 39916800.0}
#if defined(__CUDACC__)
__constant__ double d_LAGRANGE12_INV_DEN[12] = RHFQ_LAGRANGE12_INV_DEN;
#endif
static const double h_LAGRANGE12_INV_DEN[12] = RHFQ_LAGRANGE12_INV_DEN;
#if defined(__CUDA_ARCH__)
#define LAGRANGE12_INV_DEN(a) d_LAGRANGE12_INV_DEN[a]
#else
#define LAGRANGE12_INV_DEN(a) h_LAGRANGE12_INV_DEN[a]
#endif

/* 12-point Lagrange interpolation on unit-spaced nodes f[0..11] at position s (node units
 * from f[0]; s in [5, 6] when centred) */
RHFQ_HD double lagrange12(const double* __restrict__ f, double s)
{
	double t[12], w[12];
	double acc = 1.0;
#pragma unroll
	for (int a = 0; a < 12; ++a) {
		t[a] = s - (double)a;
		w[a] = acc * LAGRANGE12_INV_DEN(a);
		acc *= t[a];
	}
	double suf = 1.0, sum = 0.0;
#pragma unroll
	for (int a = 11; a >= 0; --a) {
		sum = fma(w[a] * suf, f[a], sum);
		suf *= t[a];
	}
	return sum;
}

/* H(a), a in [GTAB_A0, GTAB_A1): quintic Hermite on the cell of a from H, H', H'' at its ends:
 *   p = u^3 [(1 + 3s + 6s^2) H0 + s (1 + 3s) h H0' + s^2 h^2 H0''/2]
 *     + s^3 [(1 + 3u + 6u^2) H1 - u (1 + 3u) h H1' + u^2 h^2 H1''/2],   u = 1 - s */
RHFQ_HD double gtab_eval_h(const double* __restrict__ tab, double a)
{
	const GtabCell c = gtab_cell(a);
	const double s = c.s, u = 1.0 - s, h = c.h, hh = 0.5 * h * h;
	const double* __restrict__ rec = tab + GTAB_REC * c.j;
	const GtabPair hd0 = gtab_load2(rec), e0 = gtab_load2(rec + 2);
	const GtabPair hd1 = gtab_load2(rec + GTAB_REC), e1 = gtab_load2(rec + GTAB_REC + 2);
	const double A = fma(fma(6.0, s, 3.0), s, 1.0) * hd0.x
	                 + s * (fma(3.0, s, 1.0) * (h * hd0.y) + s * (hh * e0.x));
	const double B = fma(fma(6.0, u, 3.0), u, 1.0) * hd1.x
	                 - u * (fma(3.0, u, 1.0) * (h * hd1.y) - u * (hh * e1.x));
	return (u * u * u) * A + (s * s * s) * B;
}

/* cubic Hermite of a tabulated derivative from (f, f') = (rec[k], rec[k + 1]) at the ends of
 * the cell: error h^4 f^(4) / 384 */
RHFQ_HD double gtab_eval_cubic(const double* __restrict__ tab, double a, int k)
{
	const GtabCell c = gtab_cell(a);
	const double s = c.s, u = 1.0 - s, h = c.h;
	const double* __restrict__ rec = tab + GTAB_REC * c.j + k;
	return (u * u) * (fma(2.0, s, 1.0) * rec[0] + s * (h * rec[1]))
	       + (s * s) * (fma(2.0, u, 1.0) * rec[GTAB_REC] - u * (h * rec[GTAB_REC + 1]));
}
/* H'(a) for the Newton steps of the mode / window searches (error < 1.2e-10 n / a; the
 * searches stop at 1e-9) */
RHFQ_HD double gtab_eval_d1(const double* __restrict__ tab, double a) { return gtab_eval_cubic(tab, a, 1); }
/* H''(a), the Newton derivative (its accuracy only sets the convergence rate) and the
 * curvature that sizes the first window guess */
RHFQ_HD double gtab_eval_d2(const double* __restrict__ tab, double a) { return gtab_eval_cubic(tab, a, 2); }

/* The 13 Stirling coefficients of the a >= 12 branch depend on the sample set only (n, v):
 *   d_k = c_k (v^-(2k-1) - n), c_k = B_2k / (2k (2k-1)), k = 1..7       at [0..6]
 *   e_k = (B_2k / 2k) (n - v^-(2k-1)), k = 1..6 (for the derivative)    at [7..12]
 * They are computed ONCE per set into a block in global memory (SetLocals::gco) that the node
 * kernels read with warp-uniform loads.  As members of the per-thread context they did not fit
 * the 64 registers of the Gauss-Legendre kernel: ncu showed their spill reloads (LDL 4 % of the
 * instructions) behind a third of its stall samples and an L1 hit rate of 47 %. */
constexpr int GCO_N = 16;
RHFQ_HD void gcoef_fill(double* dc, double n, double v)
{
	const double iv = 1.0 / v, iv2 = iv * iv;
	double ivp = iv;
	dc[0] = (1.0 / 12) * (ivp - n);        dc[7] = (1.0 / 12) * (n - ivp);       ivp *= iv2;
	dc[1] = (-1.0 / 360) * (ivp - n);      dc[8] = (-1.0 / 120) * (n - ivp);     ivp *= iv2;
	dc[2] = (1.0 / 1260) * (ivp - n);      dc[9] = (1.0 / 252) * (n - ivp);      ivp *= iv2;
	dc[3] = (-1.0 / 1680) * (ivp - n);     dc[10] = (-1.0 / 240) * (n - ivp);    ivp *= iv2;
	dc[4] = (1.0 / 1188) * (ivp - n);      dc[11] = (1.0 / 132) * (n - ivp);     ivp *= iv2;
	dc[5] = (-691.0 / 360360) * (ivp - n); dc[12] = (-691.0 / 32760) * (n - ivp); ivp *= iv2;
	dc[6] = (1.0 / 156) * (ivp - n);
}

struct GCoef {
	const double* tab;  /* H, H', H'' tables of (n, v), or null */
	const double* dc;   /* the set's coefficient block, or null: buf holds them */
	double n, v, lv, C;
	double nv1;        /* n - v                       */
	double vlvC;       /* v ln v + C                  */
	double buf[13];    /* only written / read when no per-set block is bound */
};
RHFQ_HD const double* gcoef_dc(const GCoef& g) { return g.dc ? g.dc : g.buf; }

RHFQ_HD void gcoef_init(GCoef& g, double n, double v, double lv, double C, const double* set_block = nullptr)
{
	g.tab = nullptr;
	g.n = n; g.v = v; g.lv = lv; g.C = C;
	g.nv1 = n - v;
	g.vlvC = v * lv + C;
	g.dc = set_block;
	if (!set_block) gcoef_fill(g.buf, n, v);
}

/* Stirling tail of lgamma(y) for y >= 12: lgamma(y) = (y - 1/2) ln y - y + ln(2 pi)/2 + this */
RHFQ_HD double stirling_tail(double iy)
{
	const double x2 = iy * iy;
	return iy * (1.0 / 12 - x2 * (1.0 / 360 - x2 * (1.0 / 1260 - x2 * (1.0 / 1680
	       - x2 * (1.0 / 1188 - x2 * (691.0 / 360360 - x2 * (1.0 / 156)))))));
}

RHFQ_HDC double lgdiff_c(double a, const GCoef& g)
{
	if (g.tab && a >= GTAB_A0 && a < GTAB_A1)
		return gtab_eval_h(g.tab, a) + a * g.vlvC;
	if (a >= 12.0) {
		const double la = log(a);
		const double ia = 1.0 / a, ia2 = ia * ia;
		const double* __restrict__ d = gcoef_dc(g);
		const double ser = ia * (d[0] + ia2 * (d[1] + ia2 * (d[2] + ia2 * (d[3] + ia2 * (d[4]
		                   + ia2 * (d[5] + ia2 * d[6]))))));
		return a * (g.nv1 * (1.0 - la) + g.vlvC)
		       + 0.5 * ((g.n - 1.0) * (la - LOG_2PI) - g.lv) + ser;
	}
	/* lgamma(a) = lgamma(a + 12) - ln (a)_12 */
	double ra = a;
#pragma unroll
	for (int i = 1; i < 12; ++i)
		ra *= (a + i);
	const double as = a + 12.0;
	const double lgam_a = (as - 0.5) * log(as) - as + HALF_LOG_2PI + stirling_tail(1.0 / as)
	                      - log(ra);
	const double va = g.v * a;
	double lgam_va;
	if (va >= 12.0) {
		lgam_va = (va - 0.5) * log(va) - va + HALF_LOG_2PI + stirling_tail(1.0 / va);
	} else {
		double rva = va;
#pragma unroll
		for (int i = 1; i < 12; ++i)
			rva *= (va + i);
		const double vas = va + 12.0;
		lgam_va = (vas - 0.5) * log(vas) - vas + HALF_LOG_2PI + stirling_tail(1.0 / vas)
		          - log(rva);
	}
	return lgam_va - g.n * lgam_a + g.C * a;
}

/* The same for the Gauss-Legendre loops, with the two common branches inline: called out of
 * line, lgdiff_c reads the coefficient struct through memory on every evaluation (ncu: LDL 4 %
 * of the quadrature kernel's instructions, a third of its stall samples behind them); the
 * quadrature kernel is small enough to hold the two branches twice. */
RHFQ_HD double lgdiff_q(double a, const GCoef& g)
{
	if (g.tab && a >= GTAB_A0 && a < GTAB_A1)
		return gtab_eval_h(g.tab, a) + a * g.vlvC;
	if (a >= 12.0) {
		const double la = log(a);
		const double ia = 1.0 / a, ia2 = ia * ia;
		const double* __restrict__ d = gcoef_dc(g);
		const double ser = ia * (d[0] + ia2 * (d[1] + ia2 * (d[2] + ia2 * (d[3] + ia2 * (d[4]
		                   + ia2 * (d[5] + ia2 * d[6]))))));
		return a * (g.nv1 * (1.0 - la) + g.vlvC)
		       + 0.5 * ((g.n - 1.0) * (la - LOG_2PI) - g.lv) + ser;
	}
	return lgdiff_c(a, g);
}

/* d/da of lgamma(v a) - n lgamma(a)  (integrand.hpp:166-169 extended to 7 terms) */
RHFQ_HDC double digdiff_c(double a, const GCoef& g)
{
	if (g.tab && a >= GTAB_A0 && a < GTAB_A1)
		return gtab_eval_d1(g.tab, a) + g.v * g.lv;
	if (a >= 12.0 && g.v * a >= 12.0) {
		const double ia = 1.0 / a, ia2 = ia * ia;
		const double* __restrict__ e = gcoef_dc(g) + 7;
		return -g.nv1 * log(a) + g.v * g.lv + 0.5 * (g.n - 1.0) * ia
		       + ia2 * (e[0] + ia2 * (e[1] + ia2 * (e[2] + ia2 * (e[3] + ia2 * (e[4] + ia2 * e[5])))));
	}
	return g.v * digamma_pos(g.v * a) - g.n * digamma_pos(a);
}

/* inline fast branches for the window / mode kernels (see lgdiff_q) */
RHFQ_HD double digdiff_q(double a, const GCoef& g)
{
	if (g.tab && a >= GTAB_A0 && a < GTAB_A1)
		return gtab_eval_d1(g.tab, a) + g.v * g.lv;
	if (a >= 12.0 && g.v * a >= 12.0) {
		const double ia = 1.0 / a, ia2 = ia * ia;
		const double* __restrict__ e = gcoef_dc(g) + 7;
		return -g.nv1 * log(a) + g.v * g.lv + 0.5 * (g.n - 1.0) * ia
		       + ia2 * (e[0] + ia2 * (e[1] + ia2 * (e[2] + ia2 * (e[3] + ia2 * (e[4] + ia2 * e[5])))));
	}
	return digdiff_c(a, g);
}

/* Convenience forms for the per-set (not per-node) code paths */
RHFQ_HD double lgdiff(double a, double n, double v, double C, double lv)
{
	GCoef g;
	gcoef_init(g, n, v, lv, C);
	return lgdiff_c(a, g);
}

RHFQ_HD double digdiff(double a, double n, double v)
{
	GCoef g;
	gcoef_init(g, n, v, log(v), 0.0);
	return digdiff_c(a, g);
}

RHFQ_HDC double trigdiff(double a, double n, double v, const double* tab = nullptr)
{
	if (tab && a >= GTAB_A0 && a < GTAB_A1)
		return gtab_eval_d2(tab, a);
	if (a >= 12.0 && v * a >= 12.0) {
		const double ia = 1.0 / a, ia2 = ia * ia;
		const double iv = 1.0 / v, iv2 = iv * iv;
		double ivp = iv;
		const double e1 = (1.0 / 6) * (ivp - n); ivp *= iv2;
		const double e2 = (-1.0 / 30) * (ivp - n); ivp *= iv2;
		const double e3 = (1.0 / 42) * (ivp - n); ivp *= iv2;
		const double e4 = (-1.0 / 30) * (ivp - n); ivp *= iv2;
		const double e5 = (5.0 / 66) * (ivp - n);
		return ia * ((v - n) + 0.5 * ia * (1.0 - n))
		       + ia * ia2 * (e1 + ia2 * (e2 + ia2 * (e3 + ia2 * (e4 + ia2 * e5))));
	}
	return v * v * trigamma_pos(v * a) - n * trigamma_pos(a);
}

/* psi''(x), x > 0 (third derivative of lgamma): asymptotic series from x >= 12, the
 * recurrence psi''(x) = psi''(x + 1) - 2 / x^3 below; relative error < 1e-12.  Only the
 * tables use it (H''' for the cubic Hermite read of H''). */
RHFQ_HD double tetragamma_pos(double x)
{
	double acc = 0.0;
	while (x < 12.0) {
		acc -= 2.0 / (x * x * x);
		x += 1.0;
	}
	const double ix = 1.0 / x, x2 = ix * ix;
	return acc - x2 * (1.0 + ix + x2 * (0.5 - x2 * (1.0 / 6 - x2 * (1.0 / 6 - x2 * (0.3 - x2 * (5.0 / 6))))));
}

/* record j of the table of (n, v): H, H', H'', H''' at the node (direct formulas) */
RHFQ_HD void gtab_fill_record(double n, double v, int j, double* rec)
{
	const double a = gtab_node(j);
	const double lv = log(v);
	GCoef g;
	gcoef_init(g, n, v, lv, 0.0);
	/* lgdiff_c with C = 0 is lgamma(v a) - n lgamma(a) */
	rec[0] = lgdiff_c(a, g) - a * (v * lv);
	rec[1] = digdiff_c(a, g) - v * lv;
	rec[2] = trigdiff(a, n, v);
	rec[3] = v * v * v * tetragamma_pos(v * a) - n * tetragamma_pos(a);
}

/* ------------------------------------------------------------------ *
 *  Brent's minimiser (Brent 1973) — bounded iterations, device safe.  *
 *  Used where the reference uses it (amax.hpp:114, large_z.hpp:282,   *
 *  integrand.hpp:274): start at the upper end, tolerance 2^-25.       *
 * ------------------------------------------------------------------ */
template<typename F>
RHFQ_HDC void brent_minimize(F f, double lo, double hi, int max_iter, double& xmin, double& fmin)
{
	const double tolerance = 2.98023223876953125e-08; /* 2^(1-26): bits capped at digits/2 */
	const double golden = 0.3819660;
	double x, w, v, u, fx, fw, fv, fu;
	double delta = 0.0, delta2 = 0.0;
	x = w = v = hi;
	fx = fw = fv = f(x);
	int count = max_iter;
	do {
		const double mid = 0.5 * (lo + hi);
		const double fract1 = tolerance * fabs(x) + 0.25 * tolerance;
		const double fract2 = 2.0 * fract1;
		if (fabs(x - mid) <= (fract2 - 0.5 * (hi - lo)))
			break;
		bool golden_step = true;
		if (fabs(delta2) > fract1) {
			double r = (x - w) * (fx - fv);
			double q = (x - v) * (fx - fw);
			double p = (x - v) * q - (x - w) * r;
			q = 2.0 * (q - r);
			if (q > 0.0)
				p = -p;
			q = fabs(q);
			const double td = delta2;
			delta2 = delta;
			if (is_fin(p) && is_fin(q) && !(fabs(p) >= fabs(0.5 * q * td))
			    && !(p <= q * (lo - x)) && !(p >= q * (hi - x)))
			{
				delta = p / q;
				u = x + delta;
				if ((u - lo) < fract2 || (hi - u) < fract2)
					delta = (mid - x) < 0.0 ? -fabs(fract1) : fabs(fract1);
				golden_step = false;
			}
		}
		if (golden_step) {
			delta2 = (x >= mid) ? lo - x : hi - x;
			delta = golden * delta2;
		}
		u = (fabs(delta) >= fract1) ? x + delta
		    : (delta > 0.0 ? x + fabs(fract1) : x - fabs(fract1));
		fu = f(u);
		if (fu <= fx) {
			if (u >= x) lo = x; else hi = x;
			v = w; w = x; x = u;
			fv = fw; fw = fx; fx = fu;
		} else {
			if (u < x) lo = u; else hi = u;
			if (fu <= fw || w == x) {
				v = w; w = u; fv = fw; fw = fu;
			} else if (fu <= fv || v == x || v == w) {
				v = u; fv = fu;
			}
		}
	} while (--count);
	xmin = x;
	fmin = fx;
}

/* ------------------------------------------------------------------ *
 *  Algorithm 748 (Alefeld, Potra, Shi 1995), bounded, device safe.    *
 * ------------------------------------------------------------------ */
namespace t748 {

RHFQ_HD double safe_div(double num, double denom, double r)
{
	if (fabs(denom) < 1.0) {
		if (fabs(denom * DBL_MAX) <= fabs(num))
			return r;
	}
	return num / denom;
}

RHFQ_HD double secant(double a, double b, double fa, double fb)
{
	const double tol = DBL_EPS * 5;
	double c = a - (fa / (fb - fa)) * (b - a);
	if (c <= a + fabs(a) * tol || c >= b - fabs(b) * tol)
		return 0.5 * (a + b);
	return c;
}

RHFQ_HD double quadratic(double a, double b, double d, double fa, double fb, double fd, int count)
{
	double B = safe_div(fb - fa, b - a, DBL_MAX);
	double A = safe_div(fd - fb, d - b, DBL_MAX);
	A = safe_div(A - B, d - a, 0.0);
	if (A == 0.0)
		return secant(a, b, fa, fb);
	double c = ((A < 0.0) == (fa < 0.0)) ? a : b;
	for (int i = 1; i <= count; ++i)
		c -= safe_div(fa + (B + A * (c - b)) * (c - a), B + A * (2 * c - a - b), 1 + c - a);
	if (c <= a || c >= b)
		c = secant(a, b, fa, fb);
	return c;
}

RHFQ_HD double cubic(double a, double b, double d, double e, double fa, double fb, double fd, double fe)
{
	double q11 = (d - e) * fd / (fe - fd);
	double q21 = (b - d) * fb / (fd - fb);
	double q31 = (a - b) * fa / (fb - fa);
	double d21 = (b - d) * fd / (fd - fb);
	double d31 = (a - b) * fb / (fb - fa);
	double q22 = (d21 - q11) * fb / (fe - fb);
	double q32 = (d31 - q21) * fa / (fd - fa);
	double d32 = (d31 - q21) * fd / (fd - fa);
	double q33 = (d32 - q22) * fa / (fe - fa);
	double c = q31 + q32 + q33 + a;
	if (!(c > a) || !(c < b))
		c = quadratic(a, b, d, fa, fb, fd, 3);
	return c;
}

template<typename F>
RHFQ_HD void bracket(F& f, double& a, double& b, double c, double& fa, double& fb, double& d, double& fd)
{
	const double tol = DBL_EPS * 2;
	if ((b - a) < 2 * tol * a)
		c = a + 0.5 * (b - a);
	else if (c <= a + fabs(a) * tol)
		c = a + fabs(a) * tol;
	else if (c >= b - fabs(b) * tol)
		c = b - fabs(b) * tol;
	double fc = f(c);
	if (fc == 0.0) {
		a = c; fa = 0.0; d = 0.0; fd = 0.0;
		return;
	}
	if ((fa < 0.0) != (fc < 0.0) && fa != 0.0) {
		d = b; fd = fb; b = c; fb = fc;
	} else {
		d = a; fd = fa; a = c; fa = fc;
	}
}

RHFQ_HD bool tol_reached(double a, double b, double eps)
{
	return fabs(a - b) <= eps * fmin(fabs(a), fabs(b));
}

} // namespace t748

/*
 * Solves f(x) = 0 on [a,b] given fa, fb of opposite sign; stops when
 * |a-b| <= eps * min(|a|,|b|) (the eps_tolerance rule of the reference's
 * dependency) or after max_iter function evaluations.  Returns the number of
 * evaluations used; (a,b) is the final bracket.
 */
template<typename F>
RHFQ_HDN int toms748(F f, double& a, double& b, double fa, double fb, double eps, int max_iter)
{
	using namespace t748;
	int count = max_iter;
	if (tol_reached(a, b, eps) || fa == 0.0 || fb == 0.0) {
		if (fa == 0.0) b = a;
		else if (fb == 0.0) a = b;
		return 0;
	}
	const double mu = 0.5;
	double d = 0, e = 1e5, fd = 1e5, fe = 1e5, c;
	c = secant(a, b, fa, fb);
	bracket(f, a, b, c, fa, fb, d, fd);
	--count;
	if (count && fa != 0.0 && !tol_reached(a, b, eps)) {
		c = quadratic(a, b, d, fa, fb, fd, 2);
		e = d; fe = fd;
		bracket(f, a, b, c, fa, fb, d, fd);
		--count;
	}
	while (count && fa != 0.0 && !tol_reached(a, b, eps)) {
		const double a0 = a, b0 = b;
		const double min_diff = DBL_MIN * 32;
		bool prof = (fabs(fa - fb) < min_diff) || (fabs(fa - fd) < min_diff)
		         || (fabs(fa - fe) < min_diff) || (fabs(fb - fd) < min_diff)
		         || (fabs(fb - fe) < min_diff) || (fabs(fd - fe) < min_diff);
		c = prof ? quadratic(a, b, d, fa, fb, fd, 2) : cubic(a, b, d, e, fa, fb, fd, fe);
		e = d; fe = fd;
		bracket(f, a, b, c, fa, fb, d, fd);
		if (--count == 0 || fa == 0.0 || tol_reached(a, b, eps))
			break;
		prof = (fabs(fa - fb) < min_diff) || (fabs(fa - fd) < min_diff)
		    || (fabs(fa - fe) < min_diff) || (fabs(fb - fd) < min_diff)
		    || (fabs(fb - fe) < min_diff) || (fabs(fd - fe) < min_diff);
		c = prof ? quadratic(a, b, d, fa, fb, fd, 3) : cubic(a, b, d, e, fa, fb, fd, fe);
		bracket(f, a, b, c, fa, fb, d, fd);
		if (--count == 0 || fa == 0.0 || tol_reached(a, b, eps))
			break;
		double u, fu;
		if (fabs(fa) < fabs(fb)) { u = a; fu = fa; }
		else { u = b; fu = fb; }
		c = u - 2 * (fu / (fb - fa)) * (b - a);
		if (fabs(c - u) > 0.5 * (b - a))
			c = a + 0.5 * (b - a);
		e = d; fe = fd;
		bracket(f, a, b, c, fa, fb, d, fd);
		if (--count == 0 || fa == 0.0 || tol_reached(a, b, eps))
			break;
		if ((b - a) < mu * (b0 - a0))
			continue;
		e = d; fe = fd;
		bracket(f, a, b, a + 0.5 * (b - a), fa, fb, d, fd);
		--count;
	}
	if (fa == 0.0) b = a;
	else if (fb == 0.0) a = b;
	return max_iter - count;
}

/* ------------------------------------------------------------------ *
 *  Mode of the a-integrand at fixed z (integrand.hpp:201-284).        *
 *  C = lp - v (ls + log1p(-w z)) + sum_i log1p(-k_i z)                *
 * ------------------------------------------------------------------ */
/*
 * One integrand context for both branches, so that the mode search, the window
 * search and the quadrature exist ONCE in the kernel (the node kernel is bound
 * by instruction fetch when its code outgrows the instruction cache):
 *   lz = 0  regular node:  phi(a) = g(a; C(z)) - (lp + sum)                    (integrand.hpp:337-341)
 *   lz = 1  large-z node:  phi(a) = g(a; C_y) - (lp + lh0) + ln P(a), P = sum_m |C_m| y^(m-1)
 *   lz = 2  large-z, y-integrated: P = sum_m |C_m| y^m / (a + m)              (large_z.hpp:190-261,442-473)
 */
struct PhiCtx {
	GCoef g;
	double off;
	int lz;
	double y, y2, y3, ly;
	const double* cz;     /* c1[2], c2[3], c3[4] of the set (SetLocals, contiguous): read where used */
};

RHFQ_HDC double phi_eval(double a, const PhiCtx& c)
{
	double r = lgdiff_c(a, c.g) - c.off;
	if (c.lz) {
		const double* __restrict__ z = c.cz;
		const double C1 = fabs(z[0] + a * z[1]);
		const double C2 = fabs(z[2] + a * (z[3] + a * z[4]));
		const double C3 = fabs(z[5] + a * (z[6] + a * (z[7] + a * z[8])));
		if (c.lz == 2)
			r += log(1.0 / a + C1 * c.y / (a + 1.0) + C2 * c.y2 / (a + 2.0) + C3 * c.y3 / (a + 3.0));
		else
			r += log(1.0 + C1 * c.y + C2 * c.y2 + C3 * c.y3) - c.ly;
	}
	return r;
}

/* The same split as exp(phi) = exp(r) * mult for the quadrature sums: the large-z factor P(a)
 * multiplies the exponential instead of entering the exponent through a logarithm (that log
 * was 17 % of the Gauss-Legendre kernel's instructions at the large-z nodes). */
RHFQ_HD double phi_eval_split(double a, const PhiCtx& c, double& mult)
{
	double r = lgdiff_q(a, c.g) - c.off;
	mult = 1.0;
	if (c.lz) {
		const double* __restrict__ z = c.cz;
		const double C1 = fabs(z[0] + a * z[1]);
		const double C2 = fabs(z[2] + a * (z[3] + a * z[4]));
		const double C3 = fabs(z[5] + a * (z[6] + a * (z[7] + a * z[8])));
		if (c.lz == 2)
			mult = 1.0 / a + C1 * c.y / (a + 1.0) + C2 * c.y2 / (a + 2.0) + C3 * c.y3 / (a + 3.0);
		else {
			mult = 1.0 + C1 * c.y + C2 * c.y2 + C3 * c.y3;
			r -= c.ly;
		}
	}
	return r;
}

/* derivative of the dominant part (the slowly varying ln P is ignored: it only places the window) */
RHFQ_HD double dphi_eval(double a, const PhiCtx& c)
{
	return digdiff_c(a, c.g) + c.g.C - (c.lz == 2 ? 1.0 / a : 0.0);
}
/* phi and its derivative with the common branches inline (window search of the dedicated
 * kernel: nested out-of-line calls re-read the context through local memory every time) */
RHFQ_HD double phi_eval_q(double a, const PhiCtx& c)
{
	double mult;
	const double r = phi_eval_split(a, c, mult);
	return c.lz ? r + log(mult) : r;
}
RHFQ_HD double dphi_eval_q(double a, const PhiCtx& c)
{
	return digdiff_q(a, c.g) + c.g.C - (c.lz == 2 ? 1.0 / a : 0.0);
}

/* ------------------------------------------------------------------ *
 *  Mode of the a-integrand (integrand.hpp:201-284; large_z.hpp:264-290*
 *  for the large-z branch, where the reference uses Brent only).      *
 * ------------------------------------------------------------------ */
RHFQ_HD int mode_newton_inl(const SetLocals& L, const PhiCtx& c, double& amax, double& curv, int* iters)
{
	const GCoef& g = c.g;
	if (L.n < L.v)
		return ST_NOT_NORMALIZABLE;
	const double K = g.vlvC;
	if (c.lz == 0 && L.n == L.v && (g.C >= 0.0 || !(K < 0.0)))
		return ST_NOT_NORMALIZABLE;
	/* All callers integrate over [amin, inf) and only distinguish "mode inside" from "mode at
	 * the lower limit": if the integrand already decreases at amin (it is concave), the
	 * unconstrained mode below amin is not needed (it used to cost ~15 Newton iterations per
	 * large-z node, with the digamma recurrences of small arguments). */
	if (dphi_eval_q(L.amin, c) <= 0.0) {
		amax = L.amin;
		curv = 0.0;
		if (iters)
			*iters = 1;
		return ST_OK;
	}
	const double shift = (c.lz == 2) ? 2.0 : 0.0;   /* the 1/a factor lowers the exponent by one */
	/* Start from the root of the first three terms of the large-a expansion of the
	 * derivative (n == v):  K + (n-1)/(2a) + (n - 1/v)/(12 a^2) = 0.  The reference
	 * starts from a = 1; the root is unique (g is concave), so the start only
	 * changes the iteration count (1-2 instead of ~8). */
	double a = 1.0;
	if (K < 0.0) {
		const double qa = (L.n - 1.0 / L.v) / 12.0, qb = 0.5 * (L.n - 1.0 - shift);
		const double disc = qb * qb - 4.0 * qa * K;
		const double x = (disc > 0.0 && qa > 0.0 && qb > 0.0) ? (-qb + sqrt(disc)) / (2.0 * qa)
		                                                     : -K / fmax(qb, 0.5);
		const double a0 = 1.0 / x;
		if (a0 > 1.0 && a0 < 1e12)
			a = a0;
	}
	if (g.tab) {
		/* The derivative is tabulated on [1, 1024]: if its root lies there, bracket it by
		 * bisection over the table and start Newton from the linear interpolation of the
		 * bracket (2-3 iterations instead of up to ~10 from the large-a estimate). */
		/* (the 1/a term of the y-integrated branch is kept out of the other branches' way:
		 * "0.0 / node" takes the slow path of the FP64 division, 130 instructions, and was
		 * 40 % of the mode kernel) */
		const double* hp = g.tab + 1;          /* H' of node j at hp[GTAB_REC * j] */
		const bool sh = c.lz == 2;
		auto fnode = [&](int j) -> double {
			const double f = hp[GTAB_REC * j] + K;
			return sh ? f - 1.0 / gtab_node(j) : f;
		};
		int lo = 0, hi = GTAB_SIZE - 1;
		double flo = fnode(lo), fhi = fnode(hi);
		if (flo > 0.0 && fhi < 0.0) {
			while (hi - lo > 1) {
				const int mid = (lo + hi) >> 1;
				const double fm = fnode(mid);
				if (fm > 0.0) { lo = mid; flo = fm; } else { hi = mid; fhi = fm; }
			}
			a = gtab_node(lo) + (flo / (flo - fhi)) * (gtab_node(hi) - gtab_node(lo));
		}
	}
	bool success = false;
	int it = 0;
	double f1 = -1.0;
	for (; it < 30; ++it) {
		const double f0 = dphi_eval_q(a, c);
		f1 = trigdiff(a, L.n, L.v, g.tab) + (c.lz == 2 ? 1.0 / (a * a) : 0.0);
		double da = -f0 / f1;
		da = fmax(fmin(da, 1e3 * a), -0.999 * a);
		const double anext = fmax(a + da, 1e-8);
		success = fabs(a - anext) <= 1e-9 * a;
		a = anext;
		if (success)
			break;
	}
	if (iters)
		*iters = it + 1;
	if (!success || !is_fin(a)) {
		/* Brent fallback on x = amin / a, as integrand.hpp:256-282 */
		const double amin = L.amin;
		const PhiCtx* cp = &c;
		auto cost = [=](double x) -> double {
			const double aa = amin / x;
			if (is_inf(aa))
				return RHFQ_INF;
			return -phi_eval(aa, *cp);
		};
		double xm, fm;
		brent_minimize(cost, 0.0, 1.0, 200, xm, fm);
		a = amin / xm;
		f1 = trigdiff(a, L.n, L.v, g.tab) + (c.lz == 2 ? 1.0 / (a * a) : 0.0);
	}
	amax = a;
	curv = -f1;
	return ST_OK;
}
RHFQ_HDC int mode_newton(const SetLocals& L, const PhiCtx& c, double& amax, double& curv, int* iters)
{
	return mode_newton_inl(L, c, amax, curv, iters);
}

/* Used by the per-set search for the global maximum (amax.hpp:68-133) */
RHFQ_HD int amax_newton(const SetLocals& L, const GCoef& g, double& amax, int* iters = nullptr)
{
	PhiCtx c;
	c.g = g;
	c.off = 0.0;
	c.lz = 0;
	double curv;
	return mode_newton(L, c, amax, curv, iters);
}

/* ------------------------------------------------------------------ *
 *  Fixed-node integral  S = int_{amin}^inf exp(phi(a) - phi_ref) da   *
 *  for a concave phi given by a functor returning phi(a); dphi(a) is  *
 *  an (approximate) derivative used only to place the window.         *
 *  Returns log(S) + phi_ref, phi_ref = phi(mode).                     *
 * ------------------------------------------------------------------ */
struct QuadInfo {
	double a_lo = 0.0, a_mode = 0.0, a_hi = 0.0;   /* window */
	int evals = 0;               /* number of phi evaluations */
	int newton = 0;              /* Newton iterations spent on the mode */
	int bad = 0;                 /* window search failed (see window_end_acceptable) */
};

/* Nodes per panel by the shape of the integrand (~ a^((n-1)/2) e^(-|K| a)):
 * the smaller n, the more skewed.  Measured against the long-double oracle
 * (tests/test_hostmath.py): n >= 50 -> 20 nodes (same error as 24: 4e-12 at N = 50,
 * 8e-11 at N = 1000, where the lgamma rounding dominates), 10 <= n < 50 -> 24
 * (<= 1e-11; 20 nodes would give 2e-10 at N = 10), 4 <= n < 10 -> 32, n < 4 -> 48. */
RHFQ_HD void gl_select(double n, int& off, int& K)
{
	if (n >= 50.0) { off = 104; K = 20; }
	else if (n >= 10.0) { off = 0; K = 24; }
	else if (n >= 4.0) { off = 24; K = 32; }
	else { off = 56; K = 48; }
}

/* Window ends from the closed form of the two leading terms of phi around its mode a_m
 * (u = a / a_m, s = (n-1)/2, delta = n - v):
 *     phi(a) - phi(a_m) ~ s (ln u - u + 1) - delta a_m (u ln u - u + 1)
 * solved for a drop of RHFQ_DROP + 2 nats by scalar Newton (no special functions of the
 * data).  The caller verifies the guess with one evaluation of phi and falls back to the
 * monotone Newton search on phi itself when it is off. */
RHFQ_HDC void window_guess(double s, double da_m, double& u_lo, double& u_hi)
{
	const double D = RHFQ_DROP + 2.0;
	const double curv = s + da_m;
	/* right end: Newton in u from the Gaussian estimate (h is concave, start inside) */
	double u = 1.0 + sqrt(2.0 * D / curv);
	for (int i = 0; i < 6; ++i) {
		const double lu = log(u);
		const double h = s * (lu - u + 1.0) - da_m * (u * lu - u + 1.0) + D;
		const double hp = s * (1.0 / u - 1.0) - da_m * lu;
		const double un = u - h / hp;
		const bool done = fabs(un - u) <= 1e-3 * u;
		u = (un > 1.0) ? un : 0.5 * (1.0 + u);
		if (done) break;
	}
	u_hi = u;
	/* left end: Newton in w = ln u */
	double w = -sqrt(2.0 * D / curv);
	for (int i = 0; i < 8; ++i) {
		const double ew = exp(w);
		const double h = s * (w - ew + 1.0) - da_m * (ew * w - ew + 1.0) + D;
		const double hp = s * (1.0 - ew) - da_m * ew * w;
		const double wn = w - h / hp;
		const bool done = fabs(wn - w) <= 1e-3 * fabs(w);
		w = (wn < 0.0) ? wn : 0.5 * w;
		if (done) break;
	}
	u_lo = exp(w);
}

/* Outcome of the window search: everything the Gauss-Legendre stage needs.  The two
 * stages can run in different kernels (the engine does that for the node evaluations:
 * the mode / window search and the quadrature loop are different code, and warps that
 * sit in different phases evict each other's instructions). */
struct QuadPlan {
	double a_ref, a_lo, a_hi, phi_ref;
	int interior;
	int evals;
	int bad;      /* a window search ran out of iterations far from its target drop */
};

/* A window end whose search hit the iteration cap is accepted if the drop of phi there is
 * still within (DROP - 12, DROP + 60) nats: the truncated tail is < e^-28 of the integral, the
 * widened window costs the Gauss-Legendre rule less than its margin.  Anything else is the
 * analogue of the reference's PrecisionError (integrand.hpp:459-466): status PRECISION. */
RHFQ_HD bool window_end_acceptable(double r) { return r <= 12.0 && r >= -60.0; }

RHFQ_HD void quad_plan_inl(const PhiCtx& ctx, double amin, double amode,
                           double curv /* -phi''(mode) > 0 or <= 0 if unknown */,
                           double nshape, double delta /* n - v, < 0: no closed-form guess */,
                           QuadPlan& plan)
{
	auto phi = [&ctx](double a) -> double { return phi_eval_q(a, ctx); };
	auto dphi = [&ctx](double a) -> double { return dphi_eval_q(a, ctx); };
	int evals = 0;
	const bool interior = amode > amin;
	const double a_ref = interior ? amode : amin;
	const double phi_ref = phi(a_ref);
	++evals;

	/* Window [a_lo, a_hi] with a drop of phi of DROP-2 ... DROP+6 nats at each end
	 * (clipped at amin).  First try: closed-form guess, one verification each. */
	double u_lo = 0.0, u_hi = 0.0;
	const bool guess = interior && delta >= 0.0 && nshape > 1.0;
	if (guess)
		window_guess(0.5 * (nshape - 1.0), delta * amode, u_lo, u_hi);

	/* Right end.  Fallback: Newton on the concave function converges monotonically
	 * from outside the window. */
	double step;
	if (interior && curv > 0.0)
		step = sqrt(2.0 * RHFQ_DROP / curv);
	else {
		const double d = dphi(a_ref);
		step = (d < 0.0) ? RHFQ_DROP / (-d) : a_ref;
	}
	double a_hi = guess ? amode * u_hi : a_ref + step;
	int bad = 0;
	bool found = false;
	for (int it = 0; it < 12; ++it) {
		const double r = phi(a_hi) - phi_ref + RHFQ_DROP;   /* want r in [-6, 2] (guess) / [-3, 0] */
		++evals;
		if (it == 0 && guess && r <= 2.0 && r >= -6.0) { found = true; break; }
		if (r <= 0.0 && r >= -3.0) { found = true; break; }
		const double d = dphi(a_hi);
		if (!(d < 0.0)) {                /* still left of where phi decreases */
			a_hi = a_ref + 2.0 * (a_hi - a_ref);
			continue;
		}
		double anew = a_hi - (r + 1.0) / d;   /* aim at r = -1 */
		if (!(anew > a_ref))
			anew = 0.5 * (a_ref + a_hi);
		a_hi = anew;
	}
	if (!found) {
		++evals;
		if (!window_end_acceptable(phi(a_hi) - phi_ref + RHFQ_DROP)) bad = 1;
	}

	double a_lo = amin;
	if (interior) {
		/* Left end, clipped at amin. */
		bool need_search = true;
		if (guess && amode * u_lo > amin) {
			const double a = amode * u_lo;
			const double r = phi(a) - phi_ref + RHFQ_DROP;
			++evals;
			if (r <= 2.0 && r >= -6.0) { a_lo = a; need_search = false; }
		}
		if (need_search) {
			double r = phi(amin) - phi_ref + RHFQ_DROP;
			++evals;
			if (r < -3.0) {
				double a = amode - step;
				if (!(a > amin))
					a = 0.5 * (amin + amode);
				bool found_lo = false;
				for (int it = 0; it < 12; ++it) {
					r = phi(a) - phi_ref + RHFQ_DROP;
					++evals;
					if (r <= 0.0 && r >= -3.0) { found_lo = true; break; }
					const double d = dphi(a);
					if (!(d > 0.0)) {
						a = 0.5 * (a + amode);
						continue;
					}
					double anew = a - (r + 1.0) / d;
					if (!(anew < amode))
						anew = 0.5 * (a + amode);
					if (!(anew > amin)) {
						anew = amin;
						a = anew;
						found_lo = true;          /* clipped at the lower limit */
						break;
					}
					a = anew;
				}
				if (!found_lo) {
					++evals;
					if (!window_end_acceptable(phi(a) - phi_ref + RHFQ_DROP)) bad = 1;
				}
				a_lo = a;
			}
		}
	}

	plan.a_ref = a_ref; plan.a_lo = a_lo; plan.a_hi = a_hi; plan.phi_ref = phi_ref;
	plan.interior = interior ? 1 : 0;
	plan.evals = evals;
	plan.bad = bad;
}

RHFQ_HDC void quad_plan(const PhiCtx& ctx, double amin, double amode, double curv, double nshape,
                        double delta, QuadPlan& plan)
{
	quad_plan_inl(ctx, amin, amode, curv, nshape, delta, plan);
}

/* Partial Gauss-Legendre sum of the planned integral over the nodes j = lane, lane + nlanes, ...
 * of every panel (sum over the lanes = S, the integral / exp(phi_ref)); `split`: every panel cut
 * in two -- twice the nodes, none shared with the plain rule: the embedded error estimate of the
 * sampled precision check, evaluated by the lanes of a warp together. */
/* Ends of the Gauss-Legendre panels, ascending: [a_lo, a_ref] (interior modes only) and
 * [a_ref, a_hi], cut once more at the root of C1(a) where it lies inside.  At large-z nodes the
 * integrand carries |C1(a)| y (the reference drops the signs of the expansion's terms,
 * large_z.hpp:215-262): a KINK at C1 = 0, which sits inside the window when the shape
 * parameter is small.  A rule that straddles it converges like h^2 only -- measured 8e-9 in
 * the integral (N = 44, alpha = 10, y = 1e-7; 1.4e-8 for the y-integrated form), enough to
 * flip a refinement decision of the interpolant that sat at 0.998 of its tolerance; the
 * reference's adaptive rule resolves the kink.  (|C2| y^2 and |C3| y^3 are 1e-7 of that.)
 * Returns the number of panels. */
/* The kink costs a rule that straddles it ~1.4e-4 of the |C1| y term (measured): it is cut out
 * only where that term can exceed 3e-6 of the integrand -- error < 5e-10 otherwise (most large-z
 * nodes of an interpolant sit at far smaller y; the extra panel is 50 % more evaluations for the
 * node and costs the Gauss-Legendre kernels 3.03 -> 3.4 ms per step: 0.17 ms for the third copy
 * of the node loop being there, the rest for warps in which some lane takes it.  A second pass
 * over flagged nodes with a thread per node was slower still: 4.4 ms). */
#ifndef RHFQ_KINK_MIN
#define RHFQ_KINK_MIN 3e-6
#endif
RHFQ_HD bool quad_kink_matters(const PhiCtx& ctx, double a_hi)
{
#ifdef RHFQ_NO_KINK
	(void)ctx; (void)a_hi;
	return false;
#else
	return ctx.lz && ctx.cz[1] != 0.0 && ctx.y * fabs(ctx.cz[1]) * a_hi > RHFQ_KINK_MIN;
#endif
}
RHFQ_HD int quad_panel_ends(const PhiCtx& ctx, const QuadPlan& plan, double* bp)
{
	int n = 0;
	bp[n++] = plan.interior ? plan.a_lo : plan.a_ref;
	if (plan.interior) bp[n++] = plan.a_ref;
	bp[n++] = plan.a_hi;
	if (quad_kink_matters(ctx, plan.a_hi)) {
		const double a0 = -ctx.cz[0] / ctx.cz[1];
		for (int i = 0; i + 1 < n; ++i) {
			/* strictly inside, and not within rounding of an end */
			if (a0 > bp[i] && a0 < bp[i + 1] && (a0 - bp[i]) > 1e-12 * a0 && (bp[i + 1] - a0) > 1e-12 * a0) {
				for (int j = n; j > i + 1; --j) bp[j] = bp[j - 1];
				bp[i + 1] = a0;
				++n;
				break;
			}
		}
	}
	return n - 1;
}

RHFQ_HDC double quad_panels_partial(const PhiCtx& ctx, const QuadPlan& plan, double nshape, bool split,
                                    int lane, int nlanes, int* evals_out)
{
	int gl_off, gl_K;
	gl_select(nshape, gl_off, gl_K);
	const double phi_ref = plan.phi_ref;
	double ends[4];
	const int np = quad_panel_ends(ctx, plan, ends);
	int evals = 0;
	double S = 0.0;
	const int pieces = split ? 2 : 1;
	for (int pnl = 0; pnl < np; ++pnl) {
		const double lo = ends[pnl], w = (ends[pnl + 1] - lo) / pieces;
		for (int piece = 0; piece < pieces; ++piece) {
			const double x0 = lo + piece * w, x1 = (piece + 1 == pieces) ? ends[pnl + 1] : x0 + w;
			const double hw = 0.5 * (x1 - x0), mid = 0.5 * (x1 + x0);
			double s = 0.0;
			for (int j = lane; j < gl_K; j += nlanes) {
				double mult;
				const double r = phi_eval_split(mid + hw * GLX(gl_off + j), ctx, mult);
				s += (GLW(gl_off + j) * mult) * exp(r - phi_ref);
				++evals;
			}
			S += hw * s;
		}
	}
	if (evals_out)
		*evals_out = evals;
	return S;
}

/* Gauss-Legendre panels split at the mode (and at the kink of a large-z integrand); returns
 * log(S) + phi_ref.  (_inl: inlined into the dedicated quadrature kernel, where the context then
 * lives in registers instead of being read through local memory by an out-of-line callee.) */
#ifndef RHFQ_GL_UNROLL
#define RHFQ_GL_UNROLL 1
#endif
#if defined(__CUDACC__)
#define RHFQ_PRAGMA(x) _Pragma(#x)
#define RHFQ_UNROLL_N(n) RHFQ_PRAGMA(unroll n)
#else
#define RHFQ_UNROLL_N(n)
#endif
RHFQ_HD double quad_panels_inl(const PhiCtx& ctx, const QuadPlan& plan, double nshape, int* evals_out)
{
	int gl_off, gl_K;
	gl_select(nshape, gl_off, gl_K);
	const double a_ref = plan.a_ref, phi_ref = plan.phi_ref, a_hi = plan.a_hi;
	const double a_lo = plan.interior ? plan.a_lo : a_ref;
	/* the kink of a large-z integrand (quad_panel_ends), kept in scalars: this is the hot loop
	 * of the Gauss-Legendre kernels, an array of panel ends cost them 16 % */
	bool ku = false, kl = false;
	double kk = a_hi;
	if (quad_kink_matters(ctx, a_hi)) {
		const double a0 = -ctx.cz[0] / ctx.cz[1], tiny = 1e-12 * a0;
		ku = a0 - a_ref > tiny && a_hi - a0 > tiny;
		kl = a0 - a_lo > tiny && a_ref - a0 > tiny;
		kk = a0;
	}
	int evals = 0;
	double S = 0.0;
	/* (three explicit copies of the node loop: as ONE loop over the panels the kernel was 10 %
	 * slower even without a kink anywhere, 3.33 against 3.03 ms per step) */
	auto panel = [&](double x0, double x1) {
		const double hw = 0.5 * (x1 - x0), mid = 0.5 * (x1 + x0);
		double s = 0.0;
		RHFQ_UNROLL_N(RHFQ_GL_UNROLL)
		for (int j = 0; j < gl_K; ++j) {
			const double a = mid + hw * GLX(gl_off + j);
			double mult;
			const double r = phi_eval_split(a, ctx, mult);
			s += (GLW(gl_off + j) * mult) * exp(r - phi_ref);
		}
		evals += gl_K;
		S += hw * s;
	};
	if (ku || kl) panel(kk, ku ? a_hi : a_ref);          /* the piece above the kink */
	panel(a_ref, ku ? kk : a_hi);                        /* above the mode */
	if (plan.interior) panel(a_lo, kl ? kk : a_ref);     /* below the mode */
	if (evals_out)
		*evals_out = evals;
	return log(S) + phi_ref;
}

RHFQ_HDC double quad_panels(const PhiCtx& ctx, const QuadPlan& plan, double nshape, int* evals_out)
{
	return quad_panels_inl(ctx, plan, nshape, evals_out);
}
RHFQ_HD double log_integral_concave(const PhiCtx& ctx, double amin, double amode, double curv,
                                    double nshape, double delta, QuadInfo* info)
{
	QuadPlan plan;
	quad_plan(ctx, amin, amode, curv, nshape, delta, plan);
	int evals = 0;
	const double r = quad_panels(ctx, plan, nshape, &evals);
	if (info)
		info->bad = plan.bad;
	if (info) {
		info->a_lo = plan.a_lo;
		info->a_mode = plan.a_ref;
		info->a_hi = plan.a_hi;
		info->evals = plan.evals + evals;
	}
	return r;
}

/*
 * log of the un-scaled outer integrand:
 *   log int_{amin}^inf exp( g(a; C) - (lp + sum) ) da
 * (= log(S) + logImax of integrand.hpp:381-480; subtract log_scale to get
 * log_outer_integrand).  lsum = sum_i log1p(-k_i z), l1p_wz = log1p(-w z).
 */
RHFQ_HD void phi_ctx_regular(const SetLocals& L, double C, double off, PhiCtx& c)
{
	gcoef_init(c.g, L.n, L.v, L.lv, C, L.gco);
	c.g.tab = L.gtab;
	c.off = off;
	c.lz = 0;
}

RHFQ_HD int log_outer_unscaled(const SetLocals& L, double lsum, double l1p_wz,
                               double& result, QuadInfo* info)
{
	const double C = L.lp - L.v * (L.ls + l1p_wz) + lsum;
	PhiCtx c;
	phi_ctx_regular(L, C, L.lp + lsum, c);
	double amax, curv;
	int newton = 0;
	const int st = mode_newton(L, c, amax, curv, &newton);
	if (info)
		info->newton = newton;
	if (st != ST_OK) {
		result = nan("");
		return st;
	}
	QuadInfo local;
	QuadInfo* qi = info ? info : &local;
	result = log_integral_concave(c, L.amin, amax, (amax > L.amin) ? curv : 0.0, L.n, L.n - L.v, qi);
	if (is_nan(result))
		return ST_RUNTIME;
	if (qi->bad)
		return ST_PRECISION;
	return ST_OK;
}

/* ------------------------------------------------------------------ *
 *  Large-z branch (large_z.hpp).                                      *
 * ------------------------------------------------------------------ */
RHFQ_HD void large_z_coefficients(SetLocals& L)
{
	/* Polynomial coefficients of C1, C2, C3 in a (large_z.hpp:88-175) */
	const double h1h0 = L.h1 / L.h0, h2h0 = L.h2 / L.h0, h3h0 = L.h3 / L.h0;
	const double v = L.v, w = L.w;
	L.c1[1] = h1h0 + v * w / (w - 1.0);
	L.c1[0] = -h1h0;
	{
		const double nrm = 1.0 / (w * w - 2 * w + 1);
		L.c2[2] = (v * v * w * w + h1h0 * (2 * v * w * (w - 1) + h1h0 * (1 + w * (w - 2)))) * nrm;
		L.c2[1] = (v * w * w + 2 * h2h0 * (w * (w - 2) + 1)
		           - h1h0 * (2 * w * v * (w - 1) + 3 * h1h0 * (w * (w - 2) + 1))) * nrm;
		L.c2[0] = 2 * (h1h0 * h1h0 - h2h0);
	}
	{
		const double v2 = v * v, v3 = v2 * v, w2 = w * w, w3 = w2 * w;
		const double F = w * (w * (w - 3) + 3) - 1;
		const double nrm = 1.0 / F;
		L.c3[3] = (v3 * w3 + h1h0 * (3 * v2 * w2 * (w - 1)
		           + h1h0 * (3 * v * w * (w * (w - 2) + 1) + h1h0 * F))) * nrm;
		L.c3[2] = 3 * (v2 * w3 + 2 * h2h0 * v * w * (w - 1) * (w - 1)
		           + h1h0 * (v * w2 * (v - 1) * (1 - w) + 2 * h2h0 * F
		                     + h1h0 * (3 * v * w * (w * (2 - w) - 1) - 2 * h1h0 * F))) * nrm;
		L.c3[1] = (2 * v * w3 + 6 * h3h0 * F + 6 * h2h0 * v * w * (w * (2 - w) - 1)
		           + h1h0 * (3 * v * w2 * (1 - w) - 18 * h2h0 * F
		                     + h1h0 * (6 * v * w * (w2 - 2 * w + 1) + 11 * h1h0 * F))) * nrm;
		L.c3[0] = 6 * (h1h0 * (2 * h2h0 - h1h0 * h1h0) - h3h0);
	}
}

RHFQ_HD double large_z_Cm(const SetLocals& L, int m, double a)
{
	switch (m) {
	case 1: return L.c1[0] + a * L.c1[1];
	case 2: return L.c2[0] + a * (L.c2[1] + a * L.c2[2]);
	case 3: return L.c3[0] + a * (L.c3[1] + a * (L.c3[2] + a * L.c3[3]));
	default: return 1.0;
	}
}

/* large_z.hpp:190-261 for a single order m */
/* g: gcoef_init(n, v, lv, lp + lh0 - v (ls + l1p_w) + ly), hoisted out of the Brent loops */
RHFQ_HD double large_z_log_term(const SetLocals& L, const GCoef& g, int m, bool y_integrated,
                                double a, double ly)
{
	double lC = (m == 0) ? 0.0 : log(fabs(large_z_Cm(L, m, a)));
	if (y_integrated)
		lC += m * ly - log(a + m);
	else
		lC += (m - 1) * ly;
	lC += lgdiff_c(a, g);
	lC -= L.lp + L.lh0;
	if (is_inf(lC))
		return -RHFQ_INF;
	return lC;
}

/* large_z.hpp:264-290 */
RHFQ_HD double large_z_amax(const SetLocals& L, const GCoef& g, int m, bool y_integrated, double ly)
{
	const double amin = L.amin;
	auto cost = [&L, &g, m, y_integrated, ly, amin](double x) -> double {
		const double a = amin / x;
		if (is_inf(a))
			return RHFQ_INF;
		return -large_z_log_term(L, g, m, y_integrated, a, ly);
	};
	double xm, fm;
	brent_minimize(cost, 0.0, 1.0, 200, xm, fm);
	return amin / xm;
}

/* One order m of the transition test (large_z.hpp:293-315): the maximum over a of the
 * y-integrated term, as (log term at the maximum, log of the maximiser). */
RHFQ_HD void y_transition_term(const SetLocals& L, double y, int m, double& s, double& la)
{
	const double ly = log(y);
	GCoef g;
	gcoef_init(g, L.n, L.v, L.lv, L.lp + L.lh0 - L.v * (L.ls + L.l1p_w) + ly, L.gco);
	g.tab = L.gtab;
	const double a = large_z_amax(L, g, m, true, ly);
	s = large_z_log_term(L, g, m, true, a, ly);
	la = log(a);
}
RHFQ_HD double y_transition_combine(double s0, double la0, double s3, double la3, double log_eps)
{
	return s3 + la3 - s0 - la0 - log_eps;
}

/* large_z.hpp:293-315 */
RHFQ_HD double y_transition_root_fn(const SetLocals& L, double y)
{
	double s0, la0, s3, la3;
	y_transition_term(L, y, 0, s0, la0);
	y_transition_term(L, y, 3, s3, la3);
	return y_transition_combine(s0, la0, s3, la3, L.log_eps);
}

/* large_z.hpp:320-349 */
RHFQ_HD int y_taylor_transition(const SetLocals& L, double& ymax)
{
	/* The reference doubles y from 1e-20 until the root function is non-negative
	 * (up to 67 evaluations, each with two Brent searches).  Below its first
	 * zero crossing the root function increases with y (the y^3 term gains on the
	 * y^0 term), so the same y_r = 1e-20 * 2^k is found by a stride-8 scan over
	 * the exponent followed by a walk through the last stride: <= 16 evaluations.
	 * (It is NOT monotone up to y = 1, hence no global bisection.) */
	double yr = 1e-20;
	double val = y_transition_root_fn(L, yr);
	if (val < 0.0 || is_nan(val)) {
		int klo = 0, khi = -1;
		double vhi = val;
		for (int k = 8; k <= 64; k += 8) {
			const double v = y_transition_root_fn(L, ldexp(1e-20, k));
			if (!(v < 0.0 || is_nan(v))) { khi = k; vhi = v; break; }
			klo = k;
		}
		if (khi < 0) {
			/* not found on the coarse grid: finish like the reference, one doubling at a time */
			yr = ldexp(1e-20, klo);
			while (val < 0.0 || is_nan(val)) {
				yr = fmin(2.0 * yr, 1.0);
				if (yr == 1.0)
					break;
				val = y_transition_root_fn(L, yr);
			}
		} else {
			/* the first non-negative exponent inside the stride or the one before it, as the
			 * reference's doubling finds it (the function can dip below zero again after its
			 * first zero, and the dip can hide that zero from the stride-8 scan) */
			for (int k = (klo >= 8 ? klo - 7 : 1); k < khi; ++k) {
				if (k == klo) continue;          /* negative on the coarse grid */
				const double vm = y_transition_root_fn(L, ldexp(1e-20, k));
				if (!(vm < 0.0 || is_nan(vm))) { khi = k; vhi = vm; break; }
			}
			yr = ldexp(1e-20, khi);
			val = vhi;
		}
	}
	double a = 1e-32, b = yr;
	const double fa = y_transition_root_fn(L, a);
	const double fb = (yr == 1.0) ? y_transition_root_fn(L, b) : val;
	if (is_nan(fa) || is_nan(fb) || ((fa < 0.0) == (fb < 0.0) && fa != 0.0 && fb != 0.0))
		return ST_ROOT;
	auto f = [&L](double y) -> double { return y_transition_root_fn(L, y); };
	const int used = toms748(f, a, b, fa, fb, 0.5 /* eps_tolerance(2) */, 98);
	if (used >= 98)
		return ST_ROOT;
	ymax = 0.5 * (a + b);
	return ST_OK;
}

RHFQ_HD void phi_ctx_large_z(const SetLocals& L, double y, bool y_integrated, PhiCtx& c)
{
	const double ly = log(y);
	gcoef_init(c.g, L.n, L.v, L.lv, L.lp + L.lh0 - L.v * (L.ls + L.l1p_w) + ly, L.gco);
	c.g.tab = L.gtab;
	c.off = L.lp + L.lh0;
	c.lz = y_integrated ? 2 : 1;
	c.y = y; c.y2 = y * y; c.y3 = c.y2 * y; c.ly = ly;
	static_assert(offsetof(SetLocals, c2) == offsetof(SetLocals, c1) + 2 * sizeof(double)
	              && offsetof(SetLocals, c3) == offsetof(SetLocals, c2) + 3 * sizeof(double),
	              "c1, c2, c3 are read as one array");
	c.cz = L.c1;
}

/*
 * log of the un-scaled large-z a-integral (large_z.hpp:368-506):
 *   log int_{amin}^inf exp(G(a)) * P(a) da,
 *   G(a) = g(a; lp + lh0 - v (ls + l1p_w) + ly) - (lp + lh0)
 *   P(a) = sum_m |C_m(a)| y^m / (a + m)      (y integrated)
 *        = sum_m |C_m(a)| y^(m-1)            (not integrated)
 * The signs of C_1..C_3 are dropped exactly as the reference does
 * (large_z.hpp:442-473 sums exp(log_abs)).
 */
RHFQ_HD int log_large_z_unscaled(const SetLocals& L, double y, bool y_integrated,
                                 double& result, QuadInfo* info)
{
	if (y == 0.0) {
		result = -RHFQ_INF;
		return ST_OK;
	}
	if (L.n < L.v)
		return ST_NOT_NORMALIZABLE;
	PhiCtx c;
	phi_ctx_large_z(L, y, y_integrated, c);
	/* Mode of G(a) [- ln a]: Newton on the dominant m = 0 term. */
	double a, curv;
	int newton = 0;
	const int st = mode_newton(L, c, a, curv, &newton);
	if (info)
		info->newton = newton;
	if (st != ST_OK) {
		result = nan("");
		return st;
	}
	QuadInfo local;
	QuadInfo* qi = info ? info : &local;
	result = log_integral_concave(c, L.amin, a, (a > L.amin) ? curv : 0.0,
	                              y_integrated ? L.n - 2.0 : L.n, L.n - L.v, qi);
	if (is_nan(result))
		return ST_RUNTIME;
	if (qi->bad)
		return ST_PRECISION;
	return ST_OK;
}

/* ------------------------------------------------------------------ *
 *  Per-set sufficient statistics (locals.hpp:94-113,193-329).         *
 *  q, c: this set's data; k_out: where to write k_i.                  *
 * ------------------------------------------------------------------ */
RHFQ_HD void kahan_add(double& S, double& corr, double x)
{
	const double dS = x - corr;
	const double next = S + dS;
	corr = (next - S) - dS;
	S = next;
}

RHFQ_HDN int compute_locals(const double* q, const double* c, int N, double p, double s,
                            double n, double v, double amin, double* k_out, SetLocals& L)
{
	double A = 0, Ac = 0, B = 0, Bc = 0, lq = 0, lqc = 0;
	double Qmax = RHFQ_INF;
	int imax = 0;
	for (int i = 0; i < N; ++i) {
		if (!(q[i] > 0.0))
			return ST_DOMAIN;
		kahan_add(A, Ac, q[i]);
		kahan_add(B, Bc, c[i]);
		kahan_add(lq, lqc, log(q[i]));
		const double Q = q[i] / c[i];
		if (Q > 0.0 && Q < Qmax) {
			Qmax = Q;
			imax = i;
		}
	}
	if (is_inf(Qmax))
		return ST_DOMAIN;
	for (int i = 0; i < N; ++i) {
		const double Q = q[i] / c[i];
		if (Q == Qmax && i != imax)
			return ST_DOMAIN;
	}
	const double Asum = s + A;
	L.lp = log(p) + lq;
	L.ls = log(Asum);
	L.n = n + N;
	L.v = v + N;
	L.amin = amin;
	L.Qmax = Qmax;
	double h0 = 1, h1 = 0, h2 = 0, h3 = 0;
	for (int i = 0; i < N; ++i) {
		const double ki = (c[i] * Qmax) / q[i];
		if (!is_fin(ki))
			return ST_RUNTIME;
		k_out[i] = ki;
		if (i == imax)
			continue;
		const double d0 = 1.0 - ki;
		h3 = d0 * h3 + ki * h2;
		h2 = d0 * h2 + ki * h1;
		h1 = d0 * h1 + ki * h0;
		h0 *= d0;
	}
	if (!is_fin(h0) || !is_fin(h1) || !is_fin(h2) || !is_fin(h3))
		return ST_RUNTIME;
	L.h0 = h0; L.h1 = h1; L.h2 = h2; L.h3 = h3;
	L.w = B * Qmax / Asum;
	L.lh0 = log(h0);
	L.l1p_w = log1p(-L.w);
	L.lv = log(L.v);
	L.N = N;
	L.imax = imax;
	large_z_coefficients(L);
	return ST_OK;
}

#if defined(__CUDACC__)
/*
 * compute_locals by the 32 lanes of a warp with the SAME arithmetic in the SAME order: the
 * element-wise work -- log q_i, q_i / c_i, k_i = c_i Qmax / q_i, the divisions and logarithms
 * that made one thread per set a 0.4 ms launch for 10^4 sets -- is dealt out to the lanes
 * (lgq, k_out as scratch), the order-dependent parts (Kahan sums, first minimum, the h
 * recurrences) are then run by every lane redundantly over the stored values (uniform
 * addresses: broadcast loads), so all lanes hold identical results and no value is reduced
 * across lanes.  Bit-identical with compute_locals.
 */
__device__ inline int compute_locals_warp(const double* q, const double* c, int N, double p, double s,
                                          double n, double v, double amin, double* k_out,
                                          double* lgq, SetLocals& L, int lane)
{
	for (int i = lane; i < N; i += 32) {
		lgq[i] = log(q[i]);
		k_out[i] = q[i] / c[i];
	}
	__syncwarp();
	double A = 0, Ac = 0, B = 0, Bc = 0, lq = 0, lqc = 0;
	double Qmax = RHFQ_INF;
	int imax = 0;
	for (int i = 0; i < N; ++i) {
		if (!(q[i] > 0.0))
			return ST_DOMAIN;
		kahan_add(A, Ac, q[i]);
		kahan_add(B, Bc, c[i]);
		kahan_add(lq, lqc, lgq[i]);
		const double Q = k_out[i];
		if (Q > 0.0 && Q < Qmax) {
			Qmax = Q;
			imax = i;
		}
	}
	if (is_inf(Qmax))
		return ST_DOMAIN;
	for (int i = 0; i < N; ++i) {
		const double Q = k_out[i];
		if (Q == Qmax && i != imax)
			return ST_DOMAIN;
	}
	__syncwarp();      /* every lane has read the quotients before k_out is overwritten */
	bool fin = true;
	for (int i = lane; i < N; i += 32) {
		const double ki = (c[i] * Qmax) / q[i];
		fin = fin && is_fin(ki);
		k_out[i] = ki;
	}
	if (!__all_sync(0xffffffffu, fin))
		return ST_RUNTIME;
	__syncwarp();
	const double Asum = s + A;
	L.lp = log(p) + lq;
	L.ls = log(Asum);
	L.n = n + N;
	L.v = v + N;
	L.amin = amin;
	L.Qmax = Qmax;
	double h0 = 1, h1 = 0, h2 = 0, h3 = 0;
	for (int i = 0; i < N; ++i) {
		if (i == imax)
			continue;
		const double ki = k_out[i];
		const double d0 = 1.0 - ki;
		h3 = d0 * h3 + ki * h2;
		h2 = d0 * h2 + ki * h1;
		h1 = d0 * h1 + ki * h0;
		h0 *= d0;
	}
	if (!is_fin(h0) || !is_fin(h1) || !is_fin(h2) || !is_fin(h3))
		return ST_RUNTIME;
	L.h0 = h0; L.h1 = h1; L.h2 = h2; L.h3 = h3;
	L.w = B * Qmax / Asum;
	L.lh0 = log(h0);
	L.l1p_w = log1p(-L.w);
	L.lv = log(L.v);
	L.N = N;
	L.imax = imax;
	large_z_coefficients(L);
	return ST_OK;
}
#endif

/* 1 - k z with k z rounded first, exactly the argument log1p(-k * z) of
 * integrand.hpp:393-396 sees (no FMA contraction: for k z -> 1 the rounding of
 * the product is what conditions the result). */
RHFQ_HD double one_minus_kz(double k, double z)
{
#if defined(__CUDA_ARCH__)
	return 1.0 - __dmul_rn(k, z);
#else
	volatile double t = k * z;
	return 1.0 - t;
#endif
}

/*
 * sum_i log1p(-k_i z) (integrand.hpp:393-396) as sum over groups of sixteen of
 * log(prod (1 - k_i z)): every factor lies in [2^-53, 1] for 0 <= k_i <= 1 (or slightly
 * above 1 for negative c_i) unless k_i z rounds to 1 (then it is 0 and the sum is -inf either
 * way), so a product of sixteen is >= 2^-848 and can neither overflow nor underflow, and one
 * log per sixteen data replaces sixteen log1p -- the data sum was a quarter of the node
 * kernel's instructions.
 */
RHFQ_HD double sum_log1p(const double* k, int N, double z)
{
	double s = 0.0;
	int i = 0;
	for (; i + 16 <= N; i += 16) {
		double p[4];
#pragma unroll
		for (int g = 0; g < 4; ++g) {
			const int o = i + 4 * g;
			p[g] = (one_minus_kz(k[o], z) * one_minus_kz(k[o + 1], z))
			       * (one_minus_kz(k[o + 2], z) * one_minus_kz(k[o + 3], z));
		}
		s += log((p[0] * p[1]) * (p[2] * p[3]));
	}
	if (i < N) {
		double p = 1.0;
		for (; i < N; ++i)
			p *= one_minus_kz(k[i], z);
		s += log(p);
	}
	return s;
}

/* amax.hpp:68-133: value of the global maximum of the log-integrand over (a,z) */
RHFQ_HDN int log_integrand_max(const SetLocals& L, const double* k, double& log_scale)
{
	int status = ST_OK;
	auto cost = [&](double z) -> double {
		if (z == 1.0)
			return RHFQ_INF;
		const double lsum = sum_log1p(k, L.N, z);
		const double l1p_wz = log1p(-L.w * z);
		const double C = L.lp - L.v * (L.ls + l1p_wz) + lsum;
		GCoef g;
		gcoef_init(g, L.n, L.v, L.lv, C, L.gco);
		g.tab = L.gtab;
		double amax;
		const int st = amax_newton(L, g, amax);
		if (st != ST_OK) {
			status = st;
			return RHFQ_INF;
		}
		amax = fmax(amax, L.amin);
		return -(lgdiff_c(amax, g) - (L.lp + lsum));
	};
	double zm, fm;
	brent_minimize(cost, 0.0, 1.0, 200, zm, fm);
	if (status != ST_OK)
		return status;
	if (is_nan(zm) || is_nan(fm) || is_inf(fm))
		return ST_RUNTIME;
	log_scale = -fm;
	return ST_OK;
}

} // namespace rhfq
