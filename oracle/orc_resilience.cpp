/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_numerics.hpp).
 *
 * CPU restatement of the resilience Monte Carlo
 * (src/resilience/resilience.cpp:38-575): synthetic data generation with the
 * reference's own use of libstdc++ <random>, and per realisation two
 * Posterior<long double> (informed and flat prior, BLI <= 5 refinements,
 * rtol = tol) + tail_quantiles.
 *
 * Work -> RNG stream mapping: the reference uses `schedule(dynamic)` over
 * batches of 10 with one mt19937_64 per OpenMP thread (resilience.cpp:392-421),
 * which is only reproducible for nthread == 1.  Here stream t owns batches
 * j = t, t + T, t + 2T, ... (identical to the reference for T == 1).
 */
#include "orc_posterior.hpp"
#include <random>
#include <cstring>
#include <cstdint>
#include <omp.h>

using namespace orc;

namespace {

/* resilience.cpp:38-67, 70-91 */
void ls_anomaly_W_m2(const std::vector<double>& x, std::vector<double>& q, double P_MW,
                     double L_km, double d_km)
{
	const double Qbar_x2_MW_km = P_MW / L_km;
	const double invd = 1.0 / d_km;
	const double Q = 2.0 * Qbar_x2_MW_km * invd;
	const double QoP = Q / M_PI;
	for (size_t i = 0; i < x.size(); ++i) {
		if (x[i] == 0)
			q[i] = QoP;
		else {
			const double z = x[i] * invd;
			q[i] = QoP * (1.0 - z * std::atan(1.0 / z));
		}
	}
}

/* resilience.cpp:94-127 */
void lateral_positions(std::vector<double>& x, double radius, std::mt19937_64& gen,
                       std::vector<double>* uniforms)
{
	for (size_t i = 0; i < x.size(); ++i) {
		const double C = std::uniform_real_distribution()(gen);
		if (uniforms) uniforms->push_back(C);
		double a, b;
		if (C > 0.5) { a = 0.0; b = 1.0; }
		else { a = -1.0; b = 0.0; }
		double y = 0.5 * (a + b);
		while (std::fabs(b - a) > 1e-13) {
			const double f0 = (y * std::sqrt(1.0 - y * y) + std::asin(y) + 0.5 * M_PI) / M_PI - C;
			if (f0 > 0) b = y;
			else if (f0 < 0) a = y;
			else if (f0 == 0) break;
			y = 0.5 * (a + b);
		}
		x[i] = radius * y;
	}
}

struct Dist {
	int kind;        /* 0 gamma(K,T), 1 mixture(x0,s0,a0,x1,s1,a1) */
	double p[6];
	/* resilience.cpp:135-194 */
	void generate(std::vector<double>& q, std::mt19937_64& gen) const {
		if (kind == 0) {
			for (size_t i = 0; i < q.size(); ++i)
				q[i] = std::gamma_distribution(p[0], p[1])(gen);
			return;
		}
		std::uniform_real_distribution<> uni;
		std::normal_distribution<> n0{p[0], p[1]};
		std::normal_distribution<> n1{p[3], p[4]};
		const double w0 = p[2] / (p[2] + p[5]);
		for (size_t i = 0; i < q.size(); ++i) {
			double v = 0.0;
			for (size_t j = 0; j < 1000; ++j) {
				if (uni(gen) <= w0) { v = n0(gen); if (v > 0) break; }
				else { v = n1(gen); if (v > 0) break; }
			}
			if (v <= 0)
				throw std::runtime_error("Could not determine a positive q.");
			q[i] = v;
		}
	}
};

/* resilience.cpp:198-231 */
void synthetic_dataset(std::vector<double>& x, std::vector<double>& c, std::vector<double>& q,
                       double P_MW, std::mt19937_64& gen, const Dist& dist,
                       std::vector<double>* uniforms, std::vector<double>* q0)
{
	lateral_positions(x, 80.0, gen, uniforms);
	dist.generate(q, gen);
	if (q0) q0->insert(q0->end(), q.begin(), q.end());
	ls_anomaly_W_m2(x, c, P_MW, 160.0, 14.0);
	for (size_t i = 0; i < x.size(); ++i) {
		c[i] *= 1e3;
		q[i] += c[i];
		c[i] /= 1e6 * P_MW;
	}
}

template<typename R>
int run(const Dist& dist, size_t N, size_t M, double P_MW, const double* quantiles, int Nq,
        const double* prior, double amin, double tol, size_t seed, int nthread,
        double* out, int64_t* failures, double* data_q, double* data_c,
        double* data_u, double* data_q0, int data_only, int32_t* bli_samples = nullptr)
{
	if (nthread <= 0) nthread = omp_get_num_procs();
	/* resilience.cpp:392-395 */
	std::seed_seq seq{seed};
	std::vector<size_t> seeds(nthread);
	seq.generate(seeds.begin(), seeds.end());
	constexpr size_t batch = 10;
	const size_t J = M / batch + (M % batch != 0);
	int64_t fail_proper = 0, fail_improper = 0;
	int fatal = 0;
	#pragma omp parallel num_threads(nthread) reduction(+ : fail_proper, fail_improper)
	{
		const int t = omp_get_thread_num();
		std::mt19937_64 gen(seeds[t]);
		std::vector<double> x(N), c(N), q(N), work(Nq);
		for (size_t j = t; j < J; j += nthread) {
			for (size_t b = 0; b < batch && j * batch + b < M; ++b) {
				const size_t m = j * batch + b;
				std::vector<double> uni, q0;
				synthetic_dataset(x, c, q, P_MW, gen, dist, data_u ? &uni : nullptr,
				                  data_q0 ? &q0 : nullptr);
				if (data_q) std::memcpy(data_q + m * N, q.data(), N * sizeof(double));
				if (data_c) std::memcpy(data_c + m * N, c.data(), N * sizeof(double));
				if (data_u) std::memcpy(data_u + m * N, uni.data(), N * sizeof(double));
				if (data_q0) std::memcpy(data_q0 + m * N, q0.data(), N * sizeof(double));
				if (data_only)
					continue;
				SampleSet ss{q.data(), c.data(), N, 1.0};
				std::vector<SampleSet> sets(1, ss);
				/* resilience.cpp:320-348: informed prior; failure is fatal there */
				try {
					Posterior<R> post(sets, prior[0], prior[1], prior[2], prior[3], amin, tol, 5,
					                  BARYCENTRIC_LAGRANGE);
					for (int l = 0; l < Nq; ++l) work[l] = quantiles[l];
					post.tail_quantiles(work.data(), Nq);
					for (int l = 0; l < Nq; ++l) out[(0 * Nq + l) * M + m] = work[l];
					if (bli_samples) {
						bli_samples[4 * m + 0] = post.low ? (int32_t)post.low->samples.size() : 0;
						bli_samples[4 * m + 1] = post.high ? (int32_t)post.high->samples.size() : 0;
					}
				} catch (...) {
					for (int l = 0; l < Nq; ++l) out[(0 * Nq + l) * M + m] = std::nan("");
					++fail_proper;
					#pragma omp atomic write
					fatal = 1;
				}
				/* resilience.cpp:350-368: flat prior; failure -> NaN row */
				try {
					Posterior<R> post(sets, 1.0, 0.0, 0.0, 0.0, amin, tol, 5, BARYCENTRIC_LAGRANGE);
					for (int l = 0; l < Nq; ++l) work[l] = quantiles[l];
					post.tail_quantiles(work.data(), Nq);
					for (int l = 0; l < Nq; ++l) out[(1 * Nq + l) * M + m] = work[l];
					if (bli_samples) {
						bli_samples[4 * m + 2] = post.low ? (int32_t)post.low->samples.size() : 0;
						bli_samples[4 * m + 3] = post.high ? (int32_t)post.high->samples.size() : 0;
					}
				} catch (...) {
					for (int l = 0; l < Nq; ++l) out[(1 * Nq + l) * M + m] = std::nan("");
					++fail_improper;
				}
			}
		}
	}
	if (failures) { failures[0] = fail_proper; failures[1] = fail_improper; }
	return fatal;
}

} // namespace

extern "C" {

/*
 * out: [2][Nq][M] (proper, improper) as res[:, i, :, :] of
 * zeal2022hfresil.pyx:318-324.  precision: 0 double, 1 long double (the
 * reference hard-codes long double, resilience.cpp:321).
 * data_*: optional [M][N] dumps of the generated data sets (q, c as passed to
 * the posterior; u = position uniforms; q0 = heat flow before the anomaly).
 * Returns 1 if an informed-prior posterior failed (the reference terminates).
 */
int orc_resilience(int kind, const double* params, int64_t N, int64_t M, double P_MW,
                   const double* quantiles, int Nq, const double* prior, double amin,
                   double tol, uint64_t seed, int nthread, int precision, double* out,
                   int64_t* failures, double* data_q, double* data_c, double* data_u,
                   double* data_q0, int data_only)
{
	Dist d;
	d.kind = kind;
	for (int i = 0; i < (kind == 0 ? 2 : 6); ++i) d.p[i] = params[i];
	try {
		if (precision == 0)
			return run<double>(d, N, M, P_MW, quantiles, Nq, prior, amin, tol, seed, nthread, out,
			                   failures, data_q, data_c, data_u, data_q0, data_only);
		return run<long double>(d, N, M, P_MW, quantiles, Nq, prior, amin, tol, seed, nthread, out,
		                        failures, data_q, data_c, data_u, data_q0, data_only);
	} catch (...) {
		return 2;
	}
}

/* The same run, additionally reporting the number of samples each barycentric interpolant kept:
 * bli_samples[M][4] = (informed low, informed high, flat low, flat high); 0 for a failed one.
 * The refinement decisions are where double and long double arithmetic can part ways. */
int orc_resilience_levels(int kind, const double* params, int64_t N, int64_t M, double P_MW,
                          const double* quantiles, int Nq, const double* prior, double amin,
                          double tol, uint64_t seed, int nthread, int precision, double* out,
                          int64_t* failures, int32_t* bli_samples)
{
	Dist d;
	d.kind = kind;
	for (int i = 0; i < (kind == 0 ? 2 : 6); ++i) d.p[i] = params[i];
	try {
		if (precision == 0)
			return run<double>(d, N, M, P_MW, quantiles, Nq, prior, amin, tol, seed, nthread, out,
			                   failures, nullptr, nullptr, nullptr, nullptr, 0, bli_samples);
		return run<long double>(d, N, M, P_MW, quantiles, Nq, prior, amin, tol, seed, nthread, out,
		                        failures, nullptr, nullptr, nullptr, nullptr, 0, bli_samples);
	} catch (...) {
		return 2;
	}
}

} // extern "C"
