/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_numerics.hpp).
 * Plain C entry points over the templated CPU restatement so that tests/ and
 * bench.py's cpu_baseline leg can drive it through ctypes.
 */
#include "orc_posterior.hpp"
#include <cstring>
#include <cstdint>
#include <omp.h>

using namespace orc;

namespace {

struct Handle {
	int precision; /* 0 = double, 1 = long double */
	std::unique_ptr<Posterior<double>> pd;
	std::unique_ptr<Posterior<long double>> pl;
};

void set_err(char* err, size_t errlen, const char* msg)
{
	if (err && errlen) {
		std::strncpy(err, msg, errlen - 1);
		err[errlen - 1] = 0;
	}
}

template<typename Fun>
int guarded(char* err, size_t errlen, Fun fun)
{
	try {
		fun();
		return 0;
	} catch (const PrecisionError& e) {
		set_err(err, errlen, e.what());
		return 4;
	} catch (const std::domain_error& e) {
		set_err(err, errlen, e.what());
		return 2;
	} catch (const std::exception& e) {
		set_err(err, errlen, e.what());
		return 1;
	} catch (...) {
		set_err(err, errlen, "unknown error");
		return 1;
	}
}

} // namespace

extern "C" {

void orc_set_num_threads(int n) { omp_set_num_threads(n); }
int orc_get_max_threads() { return omp_get_max_threads(); }

/* ymax_override: nullptr, or one value per sample set (<= 0: compute as the reference does) */
void* orc_posterior_create_ymax(int precision, const double* q, const double* c,
                           const int64_t* offsets, const double* w, int64_t M,
                           double p, double s, double n, double v, double amin,
                           double rtol, int algorithm, int bli_max_refinements,
                           const double* ymax_override, char* err, size_t errlen);

void* orc_posterior_create(int precision, const double* q, const double* c,
                           const int64_t* offsets, const double* w, int64_t M,
                           double p, double s, double n, double v, double amin,
                           double rtol, int algorithm, int bli_max_refinements,
                           char* err, size_t errlen)
{
	return orc_posterior_create_ymax(precision, q, c, offsets, w, M, p, s, n, v, amin, rtol,
	                                 algorithm, bli_max_refinements, nullptr, err, errlen);
}

void* orc_posterior_create_ymax(int precision, const double* q, const double* c,
                           const int64_t* offsets, const double* w, int64_t M,
                           double p, double s, double n, double v, double amin,
                           double rtol, int algorithm, int bli_max_refinements,
                           const double* ymax_override, char* err, size_t errlen)
{
	Handle* h = new Handle();
	h->precision = precision;
	std::vector<SampleSet> sets((size_t)M);
	for (int64_t i = 0; i < M; ++i)
		sets[i] = {q + offsets[i], c + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), w[i]};
	int rc = guarded(err, errlen, [&]() {
		if (precision == 0)
			h->pd.reset(new Posterior<double>(sets, p, s, n, v, amin, rtol,
			            (unsigned)bli_max_refinements, (pdf_algorithm_t)algorithm, ymax_override));
		else
			h->pl.reset(new Posterior<long double>(sets, p, s, n, v, amin, rtol,
			            (unsigned)bli_max_refinements, (pdf_algorithm_t)algorithm, ymax_override));
	});
	if (rc != 0) {
		delete h;
		return nullptr;
	}
	return h;
}

void orc_posterior_destroy(void* hv) { delete static_cast<Handle*>(hv); }

double orc_posterior_qmax(void* hv)
{
	Handle* h = static_cast<Handle*>(hv);
	return h->precision == 0 ? (double)h->pd->Qmax : (double)h->pl->Qmax;
}

#define ORC_FORWARD(name, call)                                                         \
	int name(void* hv, double* x, int64_t nx, char* err, size_t errlen)                 \
	{                                                                                   \
		Handle* h = static_cast<Handle*>(hv);                                           \
		return guarded(err, errlen, [&]() {                                             \
			if (h->precision == 0) h->pd->call(x, (size_t)nx);                          \
			else h->pl->call(x, (size_t)nx);                                            \
		});                                                                             \
	}

ORC_FORWARD(orc_posterior_pdf, pdf)
ORC_FORWARD(orc_posterior_cdf, cdf)
ORC_FORWARD(orc_posterior_tail, tail)
ORC_FORWARD(orc_posterior_tail_quantiles, tail_quantiles)

/* out[27]: lp ls n v amin Qmax h0 h1 h2 h3 w lh0 l1p_w lv log_scale ymax ztrans S taylor
 *          norm weight global_norm global_Qmax imax z_levels z_err z_l1 */
int orc_posterior_locals(void* hv, int64_t l, double* out)
{
	Handle* h = static_cast<Handle*>(hv);
	auto fill = [&](auto& P) -> int {
		if (l < 0 || (size_t)l >= P.locals.size()) return 1;
		auto& L = P.locals[l];
		double vals[27] = {(double)L.lp, (double)L.ls, (double)L.n, (double)L.v, (double)L.amin,
			(double)L.Qmax, (double)L.h[0], (double)L.h[1], (double)L.h[2], (double)L.h[3],
			(double)L.w, (double)L.lh0, (double)L.l1p_w, (double)L.lv,
			(double)L.log_scale.log_integrand, (double)L.ymax, (double)L.ztrans, (double)L.S,
			(double)L.full_taylor_integral, (double)L.norm, (double)P.weights[l],
			(double)P.norm, (double)P.Qmax, (double)L.imax, (double)L.z_levels,
			(double)L.z_err, (double)L.z_l1};
		std::memcpy(out, vals, sizeof(vals));
		return 0;
	};
	return h->precision == 0 ? fill(*h->pd) : fill(*h->pl);
}

/* log of outer_integrand (z <= ztrans) or of the large-z branch (y given) for set l. */
int orc_posterior_log_outer(void* hv, int64_t l, const double* z, int64_t nz, double* out,
                            char* err, size_t errlen)
{
	Handle* h = static_cast<Handle*>(hv);
	return guarded(err, errlen, [&]() {
		auto run = [&](auto& P) {
			typedef std::decay_t<decltype(P.Qmax)> R;
			auto& L = P.locals.at(l);
			#pragma omp parallel for schedule(dynamic)
			for (int64_t i = 0; i < nz; ++i)
				out[i] = (double)log_outer_integrand<R>((R)z[i], L, L.log_scale.log_integrand);
		};
		if (h->precision == 0) run(*h->pd); else run(*h->pl);
	});
}

int orc_posterior_log_large_z(void* hv, int64_t l, const double* y, int64_t ny, double* out,
                              int y_integrated, char* err, size_t errlen)
{
	Handle* h = static_cast<Handle*>(hv);
	return guarded(err, errlen, [&]() {
		auto run = [&](auto& P) {
			typedef std::decay_t<decltype(P.Qmax)> R;
			auto& L = P.locals.at(l);
			for (int64_t i = 0; i < ny; ++i)
				out[i] = (double)log_a_integral_large_z<R>(y_integrated != 0, (R)y[i],
				                                           L.log_scale.log_integrand, L);
		};
		if (h->precision == 0) run(*h->pd); else run(*h->pl);
	});
}

/* BLI diagnostics: which = 0 low, 1 high.  Returns number of samples kept;
 * writes up to cap (x, f) pairs; info[4] = refinements, evaluations, err_abs, err_rel */
int64_t orc_posterior_bli(void* hv, int which, double* x, double* f, int64_t cap, double* info)
{
	Handle* h = static_cast<Handle*>(hv);
	auto run = [&](auto& P) -> int64_t {
		auto& B = which == 0 ? P.low : P.high;
		if (!B) return -1;
		typedef std::decay_t<decltype(P.Qmax)> R;
		const PII<R> pmin(B->xmin, 0, B->xmax - B->xmin);
		int64_t ns = (int64_t)B->samples.size();
		for (int64_t i = 0; i < ns && i < cap; ++i) {
			x[i] = (double)BLI<R>::chebyshev_to_support(B->samples[i].z, pmin, B->range_xmax).val;
			f[i] = (double)B->samples[i].f;
		}
		if (info) {
			info[0] = B->refinements;
			info[1] = (double)B->evaluations;
			info[2] = (double)B->last_err_abs;
			info[3] = (double)B->last_err_rel;
		}
		return ns;
	};
	return h->precision == 0 ? run(*h->pd) : run(*h->pl);
}

/* Simpson tree diagnostics: returns number of leaves (builds the tree if needed) */
int64_t orc_posterior_saq_leaves(void* hv, double* xmax, int64_t cap)
{
	Handle* h = static_cast<Handle*>(hv);
	auto run = [&](auto& P) -> int64_t {
		if (!P.saq) P.init_saq();
		int64_t nl = (int64_t)P.saq->iv.size();
		for (int64_t i = 0; i < nl && i < cap; ++i)
			xmax[i] = (double)P.saq->iv[i].xmax;
		return nl;
	};
	return h->precision == 0 ? run(*h->pd) : run(*h->pl);
}

/*
 * The reference's known-answer functions for its interpolator (src/testing/barylagrint.cpp:32-145):
 * kind 0 exp(-x), 1 sin^2 x, 2 Heaviside(x - 2), 3 exp(-(x-2)^2 / (2 * 0.5^2)), interpolated by the
 * oracle's BLI<R> with the given constructor arguments and evaluated at x[n] given as
 * PointInInterval(x, x - xmin, xmax_minus_x).  info[2]: samples kept, function evaluations.
 */
int orc_bli_known_answer(int kind, int precision, double xmin, double xmax, double tol_rel,
                         double tol_abs, double fmin, double fmax, int max_refinements,
                         const double* x, const double* xmax_minus_x, int64_t n, double* out,
                         int64_t* info, char* err, size_t errlen)
{
	return guarded(err, errlen, [&]() {
		auto run = [&](auto tag) {
			typedef decltype(tag) R;
			auto fun = [kind](const PII<R>& p) -> R {
				const R xv = p.val;
				switch (kind) {
				case 0: return std::exp(-xv);
				case 1: { const double sx = std::sin((double)xv); return (R)(sx * sx); }   /* as the reference: double sine */
				case 2: return xv > 2 ? R(1) : R(0);
				default: { const double dx = (double)xv - 2.0; return (R)std::exp(-0.5 * dx * dx / (0.5 * 0.5)); }
				}
			};
			BLI<R> bli(fun, (R)xmin, (R)xmax, (R)tol_rel, (R)tol_abs, (R)fmin, (R)fmax,
			           (unsigned)max_refinements);
			#pragma omp parallel for
			for (int64_t i = 0; i < n; ++i)
				out[i] = (double)bli(PII<R>((R)x[i], (R)x[i] - (R)xmin, (R)xmax_minus_x[i]));
			if (info) { info[0] = (int64_t)bli.samples.size(); info[1] = (int64_t)bli.evaluations; }
		};
		if (precision == 0) run(double(0)); else run((long double)0);
	});
}

/* Thread-local evaluation counters of the calling thread (single-threaded use). */
void orc_counters(int64_t* outer, int64_t* inner, int reset)
{
	*outer = (int64_t)counters().outer;
	*inner = (int64_t)counters().inner;
	if (reset) { counters().outer = 0; counters().inner = 0; }
}

} // extern "C"
