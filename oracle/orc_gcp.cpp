/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_numerics.hpp).
 *
 * CPU restatement of the gamma-conjugate-prior integrals
 * (external/pdtoolbox/src/gamma_conjugate_prior.cpp:100-712):
 *   ln_F, amax_f0/f1/f2, compute_amax, derivative_amax        :113-260
 *   ln_Phi_backend (3 extremum cases, Laplace for amax > 1e10) :354-572
 *   kullback_leibler_base / kullback_leibler / _batch_max      :588-712
 * in long double like the reference, with the oracle's own adaptive
 * tanh-sinh / exp-sinh at tolerance sqrt(eps_double).
 */
#include "orc_numerics.hpp"
#include <vector>
#include <cstdint>
#include <cstring>

using namespace orc;

namespace {

typedef long double R;

R ln_F(R a, R lp, R ls, R n, R v)
{
	return (a - 1) * lp + std::lgamma(v * a) - n * std::lgamma(a) - v * a * ls;
}

/* :123-148 */
R amax_f0(R a, R v, R n)
{
	if (a > 0.1L)
		return v * digamma<R>(v * a) - n * digamma<R>(a);
	return v * digamma<R>(v * a + 1) - n * digamma<R>(a + 1) + (n - 1) / a;
}
R amax_f1(R a, R v, R n)
{
	if (a > 1e8L) {
		const R ia = 1 / a;
		return (v - n) * ia + 0.5L * (1 - n) * ia * ia;
	}
	return v * v * trigamma<R>(v * a) - n * trigamma<R>(a);
}
R amax_f2(R a, R v, R n) { return v * v * v * tetragamma<R>(v * a) - n * tetragamma<R>(a); }

/* :151-181 */
R compute_amax(R lp, R ls, R n, R v, R amin = 1e-12L)
{
	const R C = lp - v * ls;
	R a = (amin < 1.0L) ? 1.0L : amin;
	if (!(amax_f0(amin, v, n) + C > 0))
		throw std::runtime_error("The derivative of the log-integrand is not positive at amin.");
	size_t i = 0;
	R amax = a;
	while (amax_f0(amax, v, n) + C > 0 && i < 400) {
		amax *= 100;
		++i;   /* (the reference never increments i; the loop ends through the sign change) */
	}
	a = 0.5L * (amin + amax);
	for (i = 0; i < 200; ++i) {
		const R f0 = amax_f0(a, v, n) + C;
		const R f1 = amax_f1(a, v, n);
		R da = std::min<R>(std::max<R>(-f0 / f1, -0.99L * (a - amin)), 0.99L * (amax - a));
		a += da;
		a = std::max<R>(a, amin);
		if (std::fabs(da) <= 1e-12L * a)
			break;
	}
	return a;
}

/* :207-260 */
R derivative_amax(R v, R n)
{
	size_t iter = 0;
	R ar = 1.0L, f1_r = 0;
	const size_t imax = 16400;
	while (((f1_r = amax_f1(ar, v, n)) > 0) && (iter < imax)) {
		ar *= 2;
		++iter;
		if (std::isinf(ar)) break;
	}
	if (f1_r > 0)
		return ar;
	iter = 0;
	while (amax_f2(ar, v, n) > 0 && iter < 10000) {
		R ar_new = ar / 2;
		while (amax_f1(ar_new, v, n) > 0 && iter < 10000) {
			ar_new = (ar + ar_new) / 2;
			++iter;
		}
		ar = ar_new;
		++iter;
	}
	R a0 = ar / 2;
	for (int i = 0; i < 100; ++i) {
		const R f0 = amax_f1(a0, v, n);
		const R f1 = amax_f2(a0, v, n);
		const R a1 = std::min(std::max(a0 - f0 / f1, 0.01L * a0), ar - 0.01L * (ar - a0));
		if (std::fabs(a1 - a0) < 1e-14L * a0) { a0 = a1; break; }
		a0 = a1;
	}
	return a0;
}

/* :354-572 */
R ln_Phi(R lp, R ls, R n, R v, R amin)
{
	if (std::isinf(lp) && lp < 0)
		return -lim<R>::inf();
	if (amin < 0)
		throw std::runtime_error("Parameter amin needs to be positive or zero.");
	int n_extrema;
	R amax = 0;
	if (n > 1) {
		n_extrema = 1;
		amax = compute_amax(lp, ls, n, v);
	} else if (n == 1) {
		const R euler = 0.577215664901532860606512090082L;
		if (lp - v * ls + (1 - v) * euler > 0) { n_extrema = 1; amax = compute_amax(lp, ls, n, v); }
		else n_extrema = 0;
	} else {
		const R ad = derivative_amax(v, n);
		if (amax_f0(ad, v, n) + lp - v * ls > 0) { n_extrema = 2; amax = compute_amax(lp, ls, n, v, ad); }
		else n_extrema = 0;
	}
	if (std::isinf((double)amax))
		throw std::runtime_error("Cannot normalize the integration constant since its maximum "
		                         "`amax` cannot be expressed in double precision.");
	const R term = std::sqrt((R)std::numeric_limits<double>::epsilon());
	R err, L1;
	size_t lvl;
	if (n_extrema >= 1 && amax > amin) {
		const R lFmax = ln_F(amax, lp, ls, n, v);
		if (amax > 1e10L) {
			/* Laplace approximation (:404-435); the reference's consistency check
			 * (peak_integration_bounds) only guards against round-off and throws. */
			const R f2 = std::fabs(amax_f1(amax, v, n));
			return 0.5L * std::log(2 * pi_v<R>()) - 0.5L * std::log(f2) + lFmax;
		}
		auto f = [&](R a) -> R { return std::exp(ln_F(a, lp, ls, n, v) - lFmax); };
		R S = tanh_sinh_finite<R>(f, amin, amax, term, &err, &L1, &lvl);
		S += exp_sinh<R>(f, amax, term, &err, &L1, &lvl);
		return lFmax + std::log(S);
	}
	const R lFmax = ln_F(std::max<R>(amin, 1.0L), lp, ls, n, v);
	auto f = [&](R a) -> R { return std::exp(ln_F(a, lp, ls, n, v) - lFmax); };
	const R I = exp_sinh<R>(f, amin, term, &err, &L1, &lvl);
	return lFmax + std::log(I);
}

/* :588-642; returns +inf when err * 1e5 > L1 (CONDITION_INF / the batch's catch) */
double kl_base(double lp, double s, double ls, double n, double v, double lp_ref, double s_ref,
               double ls_ref, double n_ref, double v_ref, double amin, R lnPhi, R lnPhi_ref,
               bool* bad)
{
	const R C0 = (R)lp_ref - lp;
	const R C3 = (R)v - v_ref;
	const R C1 = -C0 - (R)v * ((R)s - s_ref) / s - (R)ls * C3;
	const R C2 = -((R)n - n_ref);
	auto f = [&](R a) -> R {
		const R lga = std::lgamma(a);
		return std::exp((a - 1) * lp - n * lga + std::lgamma(a * v) - v * a * ls - lnPhi)
		       * ((C1 + C3 * digamma<R>(a * v)) * a + C0 + C2 * lga);
	};
	const R term = std::sqrt((R)std::numeric_limits<double>::epsilon());
	R err, L1;
	size_t lvl;
	const R I = exp_sinh<R>(f, (R)amin, term, &err, &L1, &lvl, 3, 14);
	if (bad) *bad = err * 1e5L > L1;
	return (double)(lnPhi_ref - lnPhi + I);
}

} // namespace

extern "C" {

int orc_gcp_ln_phi(double lp, double s, double n, double v, double amin, double* out)
{
	try {
		*out = (double)ln_Phi(lp, std::log(s), n, v, amin);
		return 0;
	} catch (...) {
		return 1;
	}
}

/* GammaConjugatePriorBase::kullback_leibler (:646-664) with CONDITION_ERROR */
int orc_gcp_kl(double lp, double s, double n, double v, double lp_ref, double s_ref,
               double n_ref, double v_ref, double amin, double* out)
{
	try {
		const double ls = std::log(s), ls_ref = std::log(s_ref);
		const R a = ln_Phi(lp, ls, n, v, amin);
		const R b = ln_Phi(lp_ref, ls_ref, n_ref, v_ref, amin);
		bool bad = false;
		*out = kl_base(lp, s, ls, n, v, lp_ref, s_ref, ls_ref, n_ref, v_ref, amin, a, b, &bad);
		return bad ? 2 : 0;
	} catch (...) {
		return 1;
	}
}

/* posterior_predictive_pdf (:794-849): p(q) = Phi(lp + ln q, s + q, n + 1, v + 1) / Phi(lp, s, n, v) */
int orc_gcp_predictive_pdf(const double* q, int64_t Nq, double lp, double s, double n, double v,
                           double amin, double* out)
{
	try {
		const R lnPhi = ln_Phi(lp, std::log(s), n, v, amin);
		for (int64_t i = 0; i < Nq; ++i) {
			if (q[i] <= 0) { out[i] = 0.0; continue; }
			out[i] = (double)std::exp(ln_Phi(lp + std::log(q[i]), std::log(s + q[i]), n + 1, v + 1,
			                                 amin) - lnPhi);
		}
		return 0;
	} catch (...) {
		return 1;
	}
}

/* posterior_predictive_pdf_common_norm (:926-990) */
int orc_gcp_predictive_pdf_common_norm(const double* q, int64_t Nq, const double* params,
                                       int64_t M, double amin, double* out)
{
	try {
		std::vector<R> lnPhi(M);
		for (int64_t i = 0; i < M; ++i)
			lnPhi[i] = ln_Phi(params[4 * i], std::log(params[4 * i + 1]), params[4 * i + 2],
			                  params[4 * i + 3], amin);
		R mx = lnPhi[0];
		for (R x : lnPhi) if (x > mx) mx = x;
		R tot = 0;
		for (R x : lnPhi) tot += std::exp(x - mx);
		for (int64_t j = 0; j < Nq; ++j) {
			R o = 0;
			for (int64_t i = 0; i < M; ++i)
				o += std::exp(ln_Phi(params[4 * i] + std::log(q[j]), std::log(params[4 * i + 1] + q[j]),
				                     params[4 * i + 2] + 1, params[4 * i + 3] + 1, amin) - mx);
			out[j] = (double)(o / tot);
		}
		return 0;
	} catch (...) {
		return 1;
	}
}

} // extern "C"

/* Adaptive Gauss-Kronrod (7,15) as the reference's dependency documents it: Kronrod
 * estimate, error = |K15 - G7|, bisect while levels remain and the error exceeds both the
 * relative and the (halved) absolute tolerance. */
template<typename F>
double gk15(F& f, double a, double b, int max_levels, double rel_tol, double abs_tol)
{
	static const double xk[8] = {0.0, 0.207784955007898467600689403773245, 0.405845151377397166906606412076961,
		0.586087235467691130294144838258730, 0.741531185599394439863864773280788,
		0.864864423359769072789712788640926, 0.949107912342758524526189684047851,
		0.991455371120812639206854697526329};
	static const double wk[8] = {0.209482141084727828012999174891714, 0.204432940075298892414161999234649,
		0.190350578064785409913256402421014, 0.169004726639267902826583426598550,
		0.140653259715525918745189590510238, 0.104790010322250183839876322541518,
		0.063092092629978553290700663189204, 0.022935322010529224963732008058970};
	static const double wg[4] = {0.417959183673469387755102040816327, 0.381830050505118944950369775488975,
		0.279705391489276667901467771423780, 0.129484966168869693270611432679082};
	const double mid = 0.5 * (a + b), hw = 0.5 * (b - a);
	const double f0 = f(mid);
	double K = wk[0] * f0, G = wg[0] * f0;
	for (int i = 1; i < 8; ++i) {
		const double fl = f(mid - hw * xk[i]), fr = f(mid + hw * xk[i]);
		K += wk[i] * (fl + fr);
		if (i % 2 == 0) G += wg[i / 2] * (fl + fr);
	}
	const double est = K * hw, err = std::fabs((K - G) * hw);
	const double abs_tol1 = std::fabs(est * rel_tol);
	if (abs_tol == 0) abs_tol = abs_tol1;
	if (max_levels && abs_tol1 < err && abs_tol < err)
		return gk15(f, a, mid, max_levels - 1, rel_tol, abs_tol / 2)
		     + gk15(f, mid, b, max_levels - 1, rel_tol, abs_tol / 2);
	return est;
}

extern "C" {

/* posterior_predictive_cdf (:996-1050) / _common_norm (:1054-1157): params[M][4]; M == 1 is
 * the plain cdf */
int orc_gcp_predictive_cdf(const double* q, int64_t Nq, const double* params, int64_t M,
                           double amin, double* out)
{
	try {
		std::vector<R> lnPhi(M);
		for (int64_t i = 0; i < M; ++i)
			lnPhi[i] = ln_Phi(params[4 * i], std::log(params[4 * i + 1]), params[4 * i + 2],
			                  params[4 * i + 3], amin);
		R mx = lnPhi[0];
		for (R x : lnPhi) if (x > mx) mx = x;
		R tot = 0;
		for (R x : lnPhi) tot += std::exp(x - mx);
		auto pdf = [&](double qi) -> double {
			R o = 0;
			for (int64_t i = 0; i < M; ++i)
				o += std::exp(ln_Phi(params[4 * i] + std::log(qi), std::log(params[4 * i + 1] + qi),
				                     params[4 * i + 2] + 1, params[4 * i + 3] + 1, amin) - mx);
			return (double)(o / tot);
		};
		std::vector<int64_t> ord(Nq);
		for (int64_t i = 0; i < Nq; ++i) ord[i] = i;
		std::sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) { return q[a] < q[b]; });
		for (int64_t i = 0; i < Nq; ++i) {
			const double ql = (i == 0) ? 0.0 : q[ord[i - 1]];
			const double qr = q[ord[i]];
			out[ord[i]] = (qr == ql) ? 0.0 : gk15(pdf, ql, qr, 5, 1e-9, 0.0);
		}
		long double sum = 0;
		for (int64_t i = 0; i < Nq; ++i) { sum += out[ord[i]]; out[ord[i]] = (double)sum; }
		if (sum > 1.0)
			for (int64_t i = 0; i < Nq; ++i) out[i] /= (double)sum;
		return 0;
	} catch (...) {
		return 1;
	}
}

/* kullback_leibler_batch_max (:667-712): refs[K][4], out[K] */
int orc_gcp_kl_batch_max(const double* params, int64_t N, const double* refs, int64_t K,
                         double amin, double* out)
{
	std::vector<R> lnPhi(N);
	std::vector<int> okp(N, 1);
	#pragma omp parallel for
	for (int64_t i = 0; i < N; ++i) {
		try {
			lnPhi[i] = ln_Phi(params[4 * i], std::log(params[4 * i + 1]), params[4 * i + 2],
			                  params[4 * i + 3], amin);
		} catch (...) {
			okp[i] = 0;
		}
	}
	for (int64_t k = 0; k < K; ++k) {
		const double lp_ref = refs[4 * k], s_ref = refs[4 * k + 1], n_ref = refs[4 * k + 2],
		             v_ref = refs[4 * k + 3];
		R lnPhi_ref;
		try {
			lnPhi_ref = ln_Phi(lp_ref, std::log(s_ref), n_ref, v_ref, amin);
		} catch (...) {
			out[k] = std::numeric_limits<double>::infinity();
			continue;
		}
		double mx = -std::numeric_limits<double>::infinity();
		#pragma omp parallel for reduction(max : mx)
		for (int64_t i = 0; i < N; ++i) {
			double kl;
			if (!okp[i]) kl = std::numeric_limits<double>::infinity();
			else {
				try {
					bool bad = false;
					kl = kl_base(params[4 * i], params[4 * i + 1], std::log(params[4 * i + 1]),
					             params[4 * i + 2], params[4 * i + 3], lp_ref, s_ref,
					             std::log(s_ref), n_ref, v_ref, amin, lnPhi[i], lnPhi_ref, &bad);
					if (bad) kl = std::numeric_limits<double>::infinity();
				} catch (...) {
					kl = std::numeric_limits<double>::infinity();
				}
			}
			if (kl > mx) mx = kl;
		}
		out[k] = mx;
	}
	return 0;
}

} // extern "C"
