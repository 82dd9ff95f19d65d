/*
 * reheatfunq_b200 — gamma-conjugate-prior integrals on the device (replaces
 * external/pdtoolbox/src/gamma_conjugate_prior.cpp:100-712 for the entry points
 * bound by reheatfunq/regional/backend.pyx:240-301,497-528).
 *
 *   ln Phi(lp, ls, n, v; amin) = ln int_{amin}^inf exp(ln F(a)) da,
 *   ln F(a) = (a-1) lp + lgamma(v a) - n lgamma(a) - v a ls            (:113-118)
 *   KL(i || ref) = ln Phi_ref - ln Phi_i
 *                  + E_i[(C1 + C3 psi(a v_i)) a + C0 + C2 lgamma(a)]      (:588-642)
 *
 * B200-first re-design: the reference integrates the combined KL integrand for
 * every (candidate reference, sample tuple) pair — N adaptive exp-sinh integrals
 * per cost evaluation.  The integrand is LINEAR in the candidate-dependent
 * constants C0..C3, so three moments per sample tuple,
 *      m1 = E_i[a],  m2 = E_i[a psi(a v_i)],  m3 = E_i[lgamma(a)],
 * are computed once (one thread per tuple, fixed-node rule around the mode) and
 * every candidate then costs one ln Phi_ref integral plus O(N) arithmetic:
 *      KL_ki = ln Phi_ref(k) - ln Phi_i + C1 m1 + C3 m2 + C0 + C2 m3.
 */
namespace rhfq {

/* Polygamma for x > 0 including tiny x: upward recurrence + asymptotic series */
RHFQ_HDC double digamma_full(double x)
{
	double res = 0.0;
	while (x < 12.0) { res -= 1.0 / x; x += 1.0; }
	const double xi = 1.0 / x, x2 = xi * xi;
	const double ser = x2 * (1.0 / 12 - x2 * (1.0 / 120 - x2 * (1.0 / 252 - x2 * (1.0 / 240
	                   - x2 * (1.0 / 132 - x2 * (691.0 / 32760 - x2 * (1.0 / 12)))))));
	return res + log(x) - 0.5 * xi - ser;
}
RHFQ_HDC double trigamma_full(double x)
{
	double res = 0.0;
	while (x < 12.0) { res += 1.0 / (x * x); x += 1.0; }
	const double xi = 1.0 / x, x2 = xi * xi;
	const double ser = xi * x2 * (1.0 / 6 - x2 * (1.0 / 30 - x2 * (1.0 / 42 - x2 * (1.0 / 30
	                   - x2 * (5.0 / 66 - x2 * (691.0 / 2730 - x2 * (7.0 / 6)))))));
	return res + xi + 0.5 * x2 + ser;
}
RHFQ_HDC double tetragamma_full(double x)
{
	double res = 0.0;
	while (x < 12.0) { res -= 2.0 / (x * x * x); x += 1.0; }
	const double xi = 1.0 / x, x2 = xi * xi;
	const double ser = x2 * x2 * (0.5 - x2 * (1.0 / 6 - x2 * (1.0 / 6 - x2 * (0.3
	                   - x2 * (5.0 / 6 - x2 * (691.0 / 210 - x2 * 17.5))))));
	return res - x2 - xi * x2 - ser;
}

struct GcpTuple { double lp, s, ls, n, v; };

/* One ln Phi integral is ~400 evaluations of ln F behind a serial mode / window search; a
 * cost evaluation of the minimum-surprise fit has only a few thousand of them, far too few
 * threads for one thread each.  A warp shares one integral: every lane runs the (cheap,
 * identical) searches, the Gauss-Legendre nodes of each panel are dealt out to the lanes and
 * the partial sums are combined once at the end. */
struct GcpLanes {
	int lane, n;
};
#if defined(__CUDACC__) && !defined(RHFQ_EMU)
constexpr int GCP_LANES = 32;
#else
constexpr int GCP_LANES = 1;
#endif
RHFQ_HD double gcp_lane_sum(double x)
{
#if defined(__CUDA_ARCH__)
	for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
#endif
	return x;
}

RHFQ_HD double gcp_ln_F(double a, const GcpTuple& t)
{
	return (a - 1.0) * t.lp + lgamma(t.v * a) - t.n * lgamma(a) - t.v * a * t.ls;
}

/* gamma_conjugate_prior.cpp:123-148 */
RHFQ_HD double gcp_f0(double a, double v, double n)
{
	if (a > 0.1)
		return v * digamma_full(v * a) - n * digamma_full(a);
	return v * digamma_full(v * a + 1.0) - n * digamma_full(a + 1.0) + (n - 1.0) / a;
}
RHFQ_HD double gcp_f1(double a, double v, double n)
{
	if (a > 1e8) {
		const double ia = 1.0 / a;
		return (v - n) * ia + 0.5 * (1.0 - n) * ia * ia;
	}
	return v * v * trigamma_full(v * a) - n * trigamma_full(a);
}
RHFQ_HD double gcp_f2(double a, double v, double n)
{
	return v * v * v * tetragamma_full(v * a) - n * tetragamma_full(a);
}

/* :151-181 */
RHFQ_HDN int gcp_compute_amax(const GcpTuple& t, double amin, double& amax_out)
{
	const double C = t.lp - t.v * t.ls;
	if (!(gcp_f0(amin, t.v, t.n) + C > 0.0))
		return ST_RUNTIME;   /* "The derivative of the log-integrand is not positive at amin." */
	double a = (amin < 1.0) ? 1.0 : amin;
	double amax = a;
	int i = 0;
	while (gcp_f0(amax, t.v, t.n) + C > 0.0 && i < 160) { amax *= 100.0; ++i; }
	a = 0.5 * (amin + amax);
	for (i = 0; i < 200; ++i) {
		const double f0 = gcp_f0(a, t.v, t.n) + C;
		const double f1 = gcp_f1(a, t.v, t.n);
		const double da = fmin(fmax(-f0 / f1, -0.99 * (a - amin)), 0.99 * (amax - a));
		a += da;
		a = fmax(a, amin);
		if (fabs(da) <= 1e-12 * a)
			break;
	}
	amax_out = a;
	return ST_OK;
}

/* :207-260 */
RHFQ_HDN double gcp_derivative_amax(double v, double n)
{
	int iter = 0;
	double ar = 1.0, f1_r = 0.0;
	while (((f1_r = gcp_f1(ar, v, n)) > 0.0) && iter < 1100) {
		ar *= 2.0;
		++iter;
		if (is_inf(ar)) break;
	}
	if (f1_r > 0.0)
		return ar;
	iter = 0;
	while (gcp_f2(ar, v, n) > 0.0 && iter < 10000) {
		double ar_new = 0.5 * ar;
		while (gcp_f1(ar_new, v, n) > 0.0 && iter < 10000) {
			ar_new = 0.5 * (ar + ar_new);
			++iter;
		}
		ar = ar_new;
		++iter;
	}
	double a0 = 0.5 * ar;
	for (int i = 0; i < 100; ++i) {
		const double f0 = gcp_f1(a0, v, n);
		const double f1 = gcp_f2(a0, v, n);
		const double a1 = fmin(fmax(a0 - f0 / f1, 0.01 * a0), ar - 0.01 * (ar - a0));
		if (fabs(a1 - a0) < 1e-14 * a0) { a0 = a1; break; }
		a0 = a1;
	}
	return a0;
}

struct GcpIntegrals {
	double lnPhi;       /* ln int F */
	double m1, m2, m3;  /* E[a], E[a psi(a v)], E[lgamma(a)] under F / Phi */
	int status;
};

/* Accumulates int F {1, a, a psi(av), lgamma(a)} over [lo, hi] with GL-48 in a or in ln a */
RHFQ_HDN void gcp_panel(const GcpTuple& t, double lo, double hi, double lFref, bool logspace,
                        bool moments, double S[4], GcpLanes ln)
{
	const int off = 56, K = 48;
	const double A = logspace ? log(lo) : lo, B = logspace ? log(hi) : hi;
	const double hw = 0.5 * (B - A), mid = 0.5 * (B + A);
	for (int j = ln.lane; j < K; j += ln.n) {
		const double u = mid + hw * GLX(off + j);
		const double a = logspace ? exp(u) : u;
		const double lga = lgamma(a);
		const double lF = (a - 1.0) * t.lp + lgamma(t.v * a) - t.n * lga - t.v * a * t.ls;
		double w = hw * GLW(off + j) * exp(lF - lFref);
		if (logspace) w *= a;
		S[0] += w;
		if (moments) {
			S[1] += w * a;
			S[2] += w * a * digamma_full(a * t.v);
			S[3] += w * lga;
		}
	}
}

/* Integrates over [lo, hi]: one GL-48 panel if the range is narrow (hi/lo <= 3), else
 * geometric panels of ratio <= 2.5 in ln a — the integrand behaves like a power of a
 * times an exponential, and a wide linear panel puts its a -> 0 singularity too close to
 * the panel in relative terms (measured: 5e-7 for the predictive tuples with n ~ 1.2). */
RHFQ_HDN void gcp_range(const GcpTuple& t, double lo, double hi, double lFref, bool moments,
                        double S[4], GcpLanes ln)
{
	if (!(hi > lo)) return;
	const double ratio = hi / lo;
	if (!(lo > 0.0) || ratio <= 3.0) {
		gcp_panel(t, lo, hi, lFref, false, moments, S, ln);
		return;
	}
	int P = (int)ceil(log(ratio) / log(2.5));
	if (P < 2) P = 2;
	if (P > 64) P = 64;
	const double r = pow(ratio, 1.0 / P);
	double x0 = lo;
	for (int p = 0; p < P; ++p) {
		const double x1 = (p == P - 1) ? hi : x0 * r;
		gcp_panel(t, x0, x1, lFref, true, moments, S, ln);
		x0 = x1;
	}
}

/* Finds a with ln F(a) - lFref in [-(DROP+3), -DROP] on the monotone branch right of `from`
 * (dir = +1) or left of it (dir = -1, clipped at amin).  Returns false if clipped. */
RHFQ_HDN bool gcp_window_end(const GcpTuple& t, double from, double step, int dir, double amin,
                             double lFref, double& a_end)
{
	/* bracket by doubling, then bisection on the log-drop: robust for the
	 * non-concave (n < 1) shapes as well */
	double inner = from, outer = from;
	double st = step;
	for (int it = 0; it < 200; ++it) {
		double cand = from + dir * st;
		if (dir < 0 && cand <= amin) {
			const double r = gcp_ln_F(amin, t) - lFref + RHFQ_DROP;
			if (r >= -3.0) { a_end = amin; return false; }   /* clipped: no need to go further */
			cand = amin;
			outer = cand;
			break;
		}
		const double r = gcp_ln_F(cand, t) - lFref + RHFQ_DROP;
		if (r <= 0.0) { outer = cand; if (r >= -3.0) { a_end = cand; return true; } break; }
		inner = cand;
		st *= 2.0;
		outer = cand;
	}
	for (int it = 0; it < 100; ++it) {
		const double m = 0.5 * (inner + outer);
		const double r = gcp_ln_F(m, t) - lFref + RHFQ_DROP;
		if (r <= 0.0 && r >= -3.0) { a_end = m; return true; }
		if (r > 0.0) inner = m; else outer = m;
	}
	a_end = outer;
	return true;
}

/* ln_Phi_backend (:354-572) + the moments needed by the KL */
RHFQ_HDN GcpIntegrals gcp_integrals(const GcpTuple& t, double amin, bool moments,
                                    GcpLanes ln = GcpLanes{0, 1})
{
	GcpIntegrals R;
	R.lnPhi = R.m1 = R.m2 = R.m3 = nan("");
	R.status = ST_OK;
	if (is_inf(t.lp) && t.lp < 0) { R.lnPhi = -RHFQ_INF; return R; }
	if (amin < 0.0) { R.status = ST_DOMAIN; return R; }
	int n_extrema;
	double amax = 0.0, a_convex = 0.0;   /* left of a_convex the log-integrand is not concave */
	if (t.n > 1.0) {
		n_extrema = 1;
		R.status = gcp_compute_amax(t, 1e-12, amax);
	} else if (t.n == 1.0) {
		const double euler = 0.577215664901532860606512090082;
		if (t.lp - t.v * t.ls + (1.0 - t.v) * euler > 0.0) {
			n_extrema = 1;
			R.status = gcp_compute_amax(t, 1e-12, amax);
		} else n_extrema = 0;
	} else {
		a_convex = gcp_derivative_amax(t.v, t.n);
		if (gcp_f0(a_convex, t.v, t.n) + t.lp - t.v * t.ls > 0.0) {
			n_extrema = 2;
			R.status = gcp_compute_amax(t, a_convex, amax);
		} else n_extrema = 0;
	}
	if (R.status != ST_OK) return R;
	if (is_inf(amax) || is_nan(amax)) { R.status = ST_RUNTIME; return R; }

	double S[4] = {0, 0, 0, 0};
	double lFref;
	if (n_extrema >= 1 && amax > amin) {
		lFref = gcp_ln_F(amax, t);
		const double curv = fabs(gcp_f1(amax, t.v, t.n));
		if (amax > 1e10) {
			/* Laplace approximation (:404-435) */
			R.lnPhi = 0.5 * log(2.0 * 3.14159265358979323846) - 0.5 * log(curv) + lFref;
			if (moments) {
				R.m1 = amax;
				R.m2 = amax * digamma_full(amax * t.v);
				R.m3 = lgamma(amax);
			}
			return R;
		}
		const double step = sqrt(2.0 * RHFQ_DROP / curv);
		double a_hi, a_lo;
		gcp_window_end(t, amax, step, +1, amin, lFref, a_hi);
		const double lo_limit = fmax(amin, a_convex);
		const bool reached = gcp_window_end(t, amax, step, -1, lo_limit, lFref, a_lo);
		gcp_range(t, amax, a_hi, lFref, moments, S, ln);
		gcp_range(t, a_lo, amax, lFref, moments, S, ln);
		if (!reached && a_lo > amin) {
			/* n < 1 and the window ran into the non-concave region: add [amin, a_lo]
			 * (the integrand may rise again towards amin) */
			gcp_range(t, amin, a_lo, lFref, moments, S, ln);
		}
	} else {
		/* no interior maximum: monotone decay from amin (:515-560) */
		lFref = gcp_ln_F(amin > 0.0 ? amin : 1e-300, t);
		if (!is_fin(lFref)) lFref = gcp_ln_F(fmax(amin, 1.0), t);
		const double d = gcp_f0(fmax(amin, 1e-12), t.v, t.n) + t.lp - t.v * t.ls;
		const double step = (d < 0.0) ? RHFQ_DROP / (-d) : fmax(amin, 1.0);
		double a_hi;
		gcp_window_end(t, amin, step, +1, amin, lFref, a_hi);
		gcp_range(t, amin > 0.0 ? amin : 1e-300, a_hi, lFref, moments, S, ln);
	}
	if (ln.n > 1) {
		for (int i = 0; i < 4; ++i) S[i] = gcp_lane_sum(S[i]);
	}
	if (!(S[0] > 0.0) || !is_fin(S[0])) { R.status = ST_RUNTIME; return R; }
	R.lnPhi = lFref + log(S[0]);
	if (moments) { R.m1 = S[1] / S[0]; R.m2 = S[2] / S[0]; R.m3 = S[3] / S[0]; }
	return R;
}

struct GcpIntegralsBody {
	const double* params; double amin; int moments; GcpIntegrals* out;
	/* task = (tuple i, lane): GCP_LANES consecutive threads (one warp) share tuple i */
	RHFQ_HD void operator()(int64_t task) const {
		const int64_t i = task / GCP_LANES;
		const GcpLanes ln{(int)(task % GCP_LANES), GCP_LANES};
		GcpTuple t;
		t.lp = params[4 * i]; t.s = params[4 * i + 1]; t.ls = log(t.s);
		t.n = params[4 * i + 2]; t.v = params[4 * i + 3];
		const GcpIntegrals r = gcp_integrals(t, amin, moments != 0, ln);
		if (ln.lane == 0) out[i] = r;
	}
};

/* kullback_leibler_batch_max (:667-712) for K candidate references at once */
struct GcpKlMaxBody {
	const double* params; const GcpIntegrals* I; int64_t N;
	const double* refs; const GcpIntegrals* Iref; double* out; double* kl_all /* may be null */;
	RHFQ_HD void operator()(int64_t k) const {
		const double lp_ref = refs[4 * k], s_ref = refs[4 * k + 1], n_ref = refs[4 * k + 2],
		             v_ref = refs[4 * k + 3];
		if (Iref[k].status != ST_OK || is_nan(Iref[k].lnPhi)) { out[k] = RHFQ_INF; return; }
		double mx = -RHFQ_INF;
		for (int64_t i = 0; i < N; ++i) {
			double kl;
			if (I[i].status != ST_OK) kl = RHFQ_INF;
			else {
				const double lp = params[4 * i], s = params[4 * i + 1], n = params[4 * i + 2],
				             v = params[4 * i + 3];
				const double C0 = lp_ref - lp;
				const double C3 = v - v_ref;
				const double C1 = -C0 - v * (s - s_ref) / s - log(s) * C3;
				const double C2 = -(n - n_ref);
				kl = Iref[k].lnPhi - I[i].lnPhi + (C1 * I[i].m1 + C3 * I[i].m2 + C0 + C2 * I[i].m3);
				if (is_nan(kl)) kl = RHFQ_INF;
			}
			if (kl_all) kl_all[k * N + i] = kl;
			if (kl > mx) mx = kl;
		}
		out[k] = mx;
	}
};

#ifndef RHFQ_MATH_ONLY
inline void gcp_ln_phi(const double* h_params, int64_t K, double amin, double* h_out,
                       int32_t* h_status, Stream& st, unsigned long long* launches)
{
	DBuf<double> dp;
	DBuf<GcpIntegrals> dI;
	dp.alloc(4 * K); dI.alloc(K);
	h2d(dp.p, h_params, 4 * K * sizeof(double), st);
	launch(K * GCP_LANES, GcpIntegralsBody{dp.p, amin, 0, dI.p}, st, launches);
	std::vector<GcpIntegrals> hI(K);
	d2h(hI.data(), dI.p, K * sizeof(GcpIntegrals), st);
	for (int64_t k = 0; k < K; ++k) {
		h_out[k] = hI[k].lnPhi;
		if (h_status) h_status[k] = hI[k].status;
	}
}

inline void gcp_kl_batch_max(const double* h_params, int64_t N, const double* h_refs, int64_t K,
                             double amin, double* h_out, double* h_kl_all, Stream& st,
                             unsigned long long* launches)
{
	DBuf<double> dp, dr, dout, dall;
	DBuf<GcpIntegrals> dI, dIr;
	dp.alloc(4 * N); dr.alloc(4 * K); dout.alloc(K); dI.alloc(N); dIr.alloc(K);
	if (h_kl_all) dall.alloc(K * N);
	h2d(dp.p, h_params, 4 * N * sizeof(double), st);
	h2d(dr.p, h_refs, 4 * K * sizeof(double), st);
	launch(N * GCP_LANES, GcpIntegralsBody{dp.p, amin, 1, dI.p}, st, launches);
	launch(K * GCP_LANES, GcpIntegralsBody{dr.p, amin, 0, dIr.p}, st, launches);
	launch(K, GcpKlMaxBody{dp.p, dI.p, N, dr.p, dIr.p, dout.p, h_kl_all ? dall.p : nullptr}, st,
	       launches);
	d2h(h_out, dout.p, K * sizeof(double), st);
	if (h_kl_all) d2h(h_kl_all, dall.p, K * N * sizeof(double), st);
}
#endif /* RHFQ_MATH_ONLY */

} // namespace rhfq
