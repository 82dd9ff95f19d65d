/*
 * reheatfunq_b200 — the barycentric-Lagrange machinery on the reference's own known-answer
 * functions (src/testing/barylagrint.cpp:32-145, exposed there as
 * reheatfunq._testing.barylagrint, testing/test_numerics.py:28-95).
 *
 * The refinement rule (barylagrint.hpp:544-668) decides every posterior and every resilience
 * result, but the posterior engine only ever feeds it log-pdf samples.  This entry runs THE SAME
 * stage bodies -- BliCheckBody (FFT refinement check + keep / refine decision), BliCoeffBody
 * (Chebyshev coefficients) and the evaluation primitives cheb_eval / bli_interpolate -- on one
 * of four functions that the reference's test holds answers for:
 *   0  exp(-x)                       1  sin^2(x)
 *   2  Heaviside(x - 2)              3  exp(-(x - 2)^2 / (2 * 0.5^2))
 */
namespace rhfq {

RHFQ_HD double bli_test_function(int kind, double x)
{
	switch (kind) {
	case 0: return exp(-x);
	case 1: { const double s = sin(x); return s * s; }
	case 2: return x > 2.0 ? 1.0 : 0.0;
	default: { const double d = x - 2.0; return exp(-0.5 * d * d / (0.5 * 0.5)); }
	}
}

/* new nodes of refinement level `level` (BliPointsBody without the posterior) */
struct BliTestSampleBody {
	int kind; double xmin, xmax; int level; int Rmax; ChebTable T; double* f;
	RHFQ_HD void operator()(int64_t jn) const {
		const int stride = 1 << (Rmax - level);
		const int fi = (level == 0) ? (int)jn * stride : (2 * (int)jn + 1) * stride;
		const int n_fine = (8 << Rmax) + 1;
		double xv;
		/* chebyshev_to_support (barylagrint.hpp:262-276): nodes descend from xmax (i = 0) */
		if (fi == 0) xv = xmax;
		else if (fi == n_fine - 1) xv = xmin;
		else xv = fmin(fmax(xmin + 0.5 * (1.0 + T.val[fi]) * (xmax - xmin), xmin), xmax);
		f[fi] = bli_test_function(kind, xv);
	}
};

/* operator() of the interpolator (barylagrint.hpp:132-160, 674-720) through the engine's
 * evaluation primitives: Clenshaw on the kept coefficients, the original barycentric sum
 * within sqrt(eps) of the range ends (bli_locate's rule), clamped to [fmin, fmax] */
struct BliTestEvalBody {
	const PostState* P; const double* f; const double* c; int Rmax; ChebTable T;
	double xmin, xmax, fmin_, fmax_; const double* x; double* out; int barycentric;
	RHFQ_HD void operator()(int64_t i) const {
		const int lev = P[0].level[0];
		const double Dx = xmax - xmin;
		const double zff = fmin((x[i] - xmin) / Dx, 1.0), zfb = fmin((xmax - x[i]) / Dx, 1.0);
		const int n_fine = (8 << Rmax) + 1;
		double v;
		if (zfb == 0.0) v = f[0];
		else if (zff == 0.0) v = f[n_fine - 1];
		else if (barycentric || pii_in_front(zff, zfb, T.se) || pii_in_back(zff, zfb, T.se)) {
			const double zv = fmin(fmax(2.0 * zff - 1.0, -1.0), 1.0);
			v = bli_interpolate(zv, zff, zfb, f, 1 << (Rmax - lev), (8 << lev) + 1, T);
		} else
			v = cheb_eval(c, 8 << lev, zff, zfb);
		out[i] = fmax(fmin(v, fmax_), fmin_);
	}
};

/* info[2]: samples kept, refinement launches.  barycentric != 0: evaluate every point with
 * the barycentric formula instead of Clenshaw (the two must agree). */
inline void bli_known_answer(int kind, double xmin, double xmax, double tol_rel, double tol_abs,
                             double fmin_, double fmax_, int max_refinements, const double* x,
                             int64_t n, double* out, int barycentric, int64_t* info, Stream& st)
{
	if (kind < 0 || kind > 3 || !(xmax > xmin) || max_refinements < 0 || max_refinements > 12)
		throw std::runtime_error("bli_known_answer: invalid argument.");
	const int Rmax = max_refinements;
	const int n_fine = (8 << Rmax) + 1;
	unsigned long long launches = 0;
	std::vector<double> tab, tw;
	make_cheb_table(Rmax, tab);
	make_fine_twiddle(Rmax, tw);
	DBuf<double> cheb, dtw, f, c, cscratch, dx, dout;
	DBuf<PostState> P;
	DBuf<int> active, flag;
	cheb.alloc(tab.size()); dtw.alloc(tw.size()); f.alloc(2 * (size_t)n_fine); c.alloc(2 * (size_t)n_fine);
	P.alloc(1); active.alloc(1); flag.alloc(1);
	h2d(cheb.p, tab.data(), tab.size() * sizeof(double), st);
	h2d(dtw.p, tw.data(), tw.size() * sizeof(double), st);
	dzero(f.p, 2 * (size_t)n_fine * sizeof(double), st);
	PostState ps;
	std::memset(&ps, 0, sizeof(ps));
	ps.status = ST_OK;
	ps.level[0] = ps.level[1] = -1;
	/* BliCheckBody tests (err_abs < rtol / Qmax) || (err_rel < rtol) */
	ps.Qmax = tol_abs > 0.0 ? tol_rel / tol_abs : RHFQ_INF;
	h2d(P.p, &ps, sizeof(ps), st);
	const int zero = 0;
	h2d(active.p, &zero, sizeof(int), st);
	const ChebTable T{cheb.p, cheb.p + n_fine, cheb.p + 2 * n_fine};
	const FineTwiddle W{dtw.p, FINE_SIGMA * (8 << Rmax)};
	int kept = -1;
	for (int level = 0; level <= Rmax; ++level) {
		const int n_new = (level == 0) ? 9 : (8 << (level - 1));
		launch(n_new, BliTestSampleBody{kind, xmin, xmax, level, Rmax, T, f.p}, st, &launches);
		if (level == 0) continue;
		cscratch.ensure((size_t)(8 << (level - 1)) + 1);
		launch_bli_check(1, BliCheckBody{active.p, level - 1, Rmax, W, f.p, cscratch.p, tol_rel, P.p,
		                                 flag.p}, st, &launches);
		int still = 0;
		d2h(&still, flag.p, sizeof(int), st);
		if (!still) { kept = level - 1; break; }
	}
	if (kept < 0) {
		kept = Rmax;          /* max_refinements reached: the finest evaluated level stays unused */
		launch(1, BliFinishBody{active.p, Rmax, P.p}, st, &launches);
		/* (the reference keeps the level BEFORE the last evaluated one; the loop above evaluates
		 * levels 0..Rmax, i.e. one fewer than the reference, so the kept level is the same) */
	}
	launch_bli_coeff(1, BliCoeffBody{P.p, f.p, c.p, Rmax, W}, st, &launches);
	dx.alloc(n); dout.alloc(n);
	h2d(dx.p, x, n * sizeof(double), st);
	launch(n, BliTestEvalBody{P.p, f.p, c.p, Rmax, T, xmin, xmax, fmin_, fmax_, dx.p, dout.p,
	                          barycentric}, st, &launches);
	d2h(out, dout.p, n * sizeof(double), st);
	if (info) { info[0] = (8 << kept) + 1; info[1] = (int64_t)launches; }
}

} // namespace rhfq
