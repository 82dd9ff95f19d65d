/*
 * reheatfunq_b200 — resilience Monte Carlo (replaces
 * src/resilience/resilience.cpp:260-575 and the loops of
 * reheatfunq/resilience/zeal2022hfresil.pyx:160-348).
 *
 * Split of the work (B200-first):
 *   host, per RNG stream (sequential by nature): the raw draws — one uniform per
 *     position and the heat-flow rejection samplers.  libstdc++'s algorithms
 *     (generate_canonical, Marsaglia polar normal with its cached variate,
 *     Marsaglia-Tsang gamma built per draw) are restated here on top of
 *     std::mt19937_64 so that sample indexing is bit-exact with the reference
 *     for equal seeds (resilience.cpp:94-194; bits/random.tcc);
 *   device, one thread per datum: inverse-CDF bisection of the lateral
 *     position, LS1980 anomaly, unit conversions (resilience.cpp:38-127,198-231)
 *     written straight into the packed (q, c) arrays of the posterior batch;
 *   device: the batched posterior engine with two posteriors per realisation
 *     (informed prior, flat prior), BLI <= 5 refinements, rtol = tol, and the
 *     batched TOMS748 tail-quantile solve (resilience.cpp:316-368).
 *
 * Stream mapping: stream t owns batches j = t, t + T, ... of 10 realisations
 * (the reference's schedule(dynamic) is reproducible only for T = 1, where both
 * agree).  Multi-GPU: rank r of G owns the STREAMS t = r, r + G, ... -- whole streams, so no
 * rank ever replays another rank's draws (host draw work M / G per rank) and the experiment
 * is a function of (seed, T) alone, whatever G.  A rank's realisations are the batches of its
 * streams in increasing order (ResilShard); its result block is [2][Nq][M_local] in that
 * order and the gather interleaves the blocks by column (reheatfunq_b200/sharding.py).
 */
#include <random>
#include <future>
#include <thread>
#include <new>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace rhfq {

struct StreamRng {
	std::mt19937_64 g;
	explicit StreamRng(uint64_t seed) : g(seed) {}
	/* generate_canonical<double,53> for a 64-bit engine (random.tcc:3346-3381) */
	double canonical() {
		const double r = (double)g() * 5.421010862427522170037e-20;   /* 2^-64 */
		return (r >= 1.0) ? 0.99999999999999988898 : r;
	}
};

/* std::normal_distribution (random.tcc:1806-1845): polar method with a cached variate */
struct PolarNormal {
	double mean, stddev, saved;
	bool have;
	PolarNormal(double m, double s) : mean(m), stddev(s), saved(0), have(false) {}
	double operator()(StreamRng& r) {
		double ret;
		if (have) {
			have = false;
			ret = saved;
		} else {
			double x, y, r2;
			do {
				x = 2.0 * r.canonical() - 1.0;
				y = 2.0 * r.canonical() - 1.0;
				r2 = x * x + y * y;
			} while (r2 > 1.0 || r2 == 0.0);
			const double mult = std::sqrt(-2 * std::log(r2) / r2);
			saved = x * mult;
			have = true;
			ret = y * mult;
		}
		return ret * stddev + mean;
	}
};

/* std::gamma_distribution(alpha, beta) constructed per draw (resilience.cpp:147-148;
 * random.tcc:2336-2392): Marsaglia-Tsang */
inline double gamma_draw(double alpha, double beta, StreamRng& r)
{
	const double malpha = alpha < 1.0 ? alpha + 1.0 : alpha;
	const double a1 = malpha - 1.0 / 3.0;
	const double a2 = 1.0 / std::sqrt(9.0 * a1);
	PolarNormal nd(0.0, 1.0);
	double u, v, n;
	do {
		do {
			n = nd(r);
			v = 1.0 + a2 * n;
		} while (v <= 0.0);
		v = v * v * v;
		u = r.canonical();
	} while (u > 1.0 - 0.0331 * n * n * n * n
	         && std::log(u) > 0.5 * n * n + a1 * (1.0 - v + std::log(v)));
	if (alpha == malpha)
		return a1 * v * beta;
	do
		u = r.canonical();
	while (u == 0.0);
	return std::pow(u, 1.0 / alpha) * a1 * v * beta;
}

struct ResilDist {
	int kind;      /* 0: gamma(K, T); 1: two-normal mixture (x0, s0, a0, x1, s1, a1) */
	double p[6];
};

/*
 * Which realisations of an M-realisation experiment on T streams belong to rank `rank` of
 * `world`: the batches (10 realisations) j = cyc * T + t of the owned streams t = rank + pos *
 * world, pos = 0 .. Tr - 1.  Local order = increasing global index:
 *   local = cyc * 10 Tr + pos * 10 + b   <->   global m = (cyc * T + rank + pos * world) * 10 + b
 * (only the last cycle can be incomplete, and its missing realisations are the highest ones,
 * so the local indices stay dense).
 */
struct ResilShard {
	static constexpr int64_t BATCH = 10;
	int64_t M; int T, rank, world, Tr;
	int64_t n_batches, n_cycles;
	ResilShard(int64_t M_, int T_, int rank_, int world_) : M(M_), T(T_), rank(rank_), world(world_)
	{
		Tr = (rank < T) ? (T - rank + world - 1) / world : 0;
		n_batches = (M + BATCH - 1) / BATCH;
		n_cycles = (n_batches + T - 1) / T;
	}
	int stream(int pos) const { return rank + pos * world; }
	int64_t batch(int64_t cyc, int pos) const { return cyc * T + stream(pos); }
	int64_t cycle_count(int64_t cyc) const {
		int64_t n = 0;
		for (int pos = 0; pos < Tr; ++pos) {
			const int64_t left = M - batch(cyc, pos) * BATCH;
			n += left <= 0 ? 0 : (left < BATCH ? left : BATCH);
		}
		return n;
	}
	/* first local index of cycle cyc (cyc == n_cycles: the local count) */
	int64_t local_begin(int64_t cyc) const {
		if (cyc <= 0 || Tr == 0) return 0;
		if (cyc < n_cycles) return cyc * BATCH * Tr;
		return (n_cycles - 1) * BATCH * Tr + cycle_count(n_cycles - 1);
	}
	int64_t local_count() const { return local_begin(n_cycles); }
	int64_t global_index(int64_t li) const {
		const int64_t per = BATCH * Tr;
		const int64_t cyc = li / per, rem = li % per;
		return batch(cyc, (int)(rem / BATCH)) * BATCH + rem % BATCH;
	}
};

/*
 * Raw draws of the owned streams: U[l*N + i] position uniforms, Q0[l*N + i] heat flow before
 * the anomaly (mW/m^2), l the local realisation index.  The streams are sequential, but a
 * stream only ever continues with its next batch, so the realisations are produced cycle by
 * cycle -- which lets the host draw block k + 1 while the device works on block k.  `ok`
 * turns false if the mixture sampler failed to produce a positive value 1000 times
 * (resilience.cpp:189-190).
 */
struct ResilDrawer {
	ResilDist dist;
	int64_t N;
	ResilShard sh;
	std::vector<StreamRng> rng;      /* one per OWNED stream */
	bool ok = true;
	static int resolve_streams(int nstreams) {
		if (nstreams > 0) return nstreams;
#ifdef _OPENMP
		return omp_get_num_procs();
#else
		return 1;
#endif
	}
	ResilDrawer(const ResilDist& d, int64_t N_, const ResilShard& sh_, uint64_t seed)
	    : dist(d), N(N_), sh(sh_)
	{
		/* std::seed_seq{seed}.generate into vector<size_t>(nstreams) (resilience.cpp:392-395) */
		std::seed_seq seq{seed};
		std::vector<size_t> seeds(sh.T);
		seq.generate(seeds.begin(), seeds.end());
		rng.reserve(sh.Tr);
		for (int pos = 0; pos < sh.Tr; ++pos) rng.emplace_back(seeds[sh.stream(pos)]);
	}
	/* one realisation: positions, heat flow (resilience.cpp:216-230 order) */
	bool draw_one(StreamRng& r, double* u, double* q) const
	{
		bool good = true;
		for (int64_t i = 0; i < N; ++i)
			u[i] = r.canonical();
		if (dist.kind == 0) {
			for (int64_t i = 0; i < N; ++i)
				q[i] = gamma_draw(dist.p[0], dist.p[1], r);
		} else {
			PolarNormal n0(dist.p[0], dist.p[1]), n1(dist.p[3], dist.p[4]);
			const double w0 = dist.p[2] / (dist.p[2] + dist.p[5]);
			for (int64_t i = 0; i < N; ++i) {
				double v = 0.0;
				for (int k = 0; k < 1000; ++k) {
					if (r.canonical() <= w0) { v = n0(r); if (v > 0) break; }
					else { v = n1(r); if (v > 0) break; }
				}
				if (v <= 0) good = false;
				q[i] = v;
			}
		}
		return good;
	}
	/* Cycles [c_lo, c_hi); consecutive calls must cover consecutive ranges.  U, Q0 are indexed
	 * from the first local realisation of cycle c_lo. */
	void draw(int64_t c_lo, int64_t c_hi, double* U, double* Q0)
	{
		/* One host thread per owned stream (at most one per core), plain std::thread: launchers
		 * such as torchrun export OMP_NUM_THREADS=1 for every rank, which serialised an OpenMP
		 * loop here, and OpenMP workers spin after their share -- with 8 ranks on a 32-thread
		 * host the spinning took the cores the other ranks' streams needed. */
		const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
		const int n_threads = std::max(1, std::min(sh.Tr, hw));
		std::vector<char> good_t(n_threads, 1);
		const int64_t per = ResilShard::BATCH * sh.Tr;
		auto work = [&](int tid) {
			bool good = true;
			for (int pos = tid; pos < sh.Tr; pos += n_threads) {
				StreamRng& r = rng[pos];
				for (int64_t cyc = c_lo; cyc < c_hi; ++cyc) {
					const int64_t m_first = sh.batch(cyc, pos) * ResilShard::BATCH;
					for (int64_t b = 0; b < ResilShard::BATCH && m_first + b < sh.M; ++b) {
						const int64_t l = (cyc - c_lo) * per + pos * ResilShard::BATCH + b;
						good = draw_one(r, U + l * N, Q0 + l * N) && good;
					}
				}
			}
			good_t[tid] = good ? 1 : 0;
		};
		if (n_threads == 1)
			work(0);
		else {
			std::vector<std::thread> pool;
			pool.reserve(n_threads - 1);
			for (int tid = 1; tid < n_threads; ++tid) pool.emplace_back(work, tid);
			work(0);
			for (auto& th : pool) th.join();
		}
		bool good = true;
		for (char g : good_t) good = good && g;
		if (!good) ok = false;
	}
};

/* resilience.cpp:94-127 (bisection), :38-67 (LS1980), :216-230 (units), one datum per thread.
 * Writes the datum for the informed-prior set and its copy for the flat-prior set. */
struct ResilDataBody {
	const double* U; const double* Q0; int64_t n_data; double P_MW;
	double* q; double* c;   /* [2 * n_data] */
	RHFQ_HD void operator()(int64_t i) const {
		const double C = U[i];
		double a, b;
		if (C > 0.5) { a = 0.0; b = 1.0; }
		else { a = -1.0; b = 0.0; }
		double y = 0.5 * (a + b);
		const double PI = 3.14159265358979323846;
		while (fabs(b - a) > 1e-13) {
			const double f0 = (y * sqrt(1.0 - y * y) + asin(y) + 0.5 * PI) / PI - C;
			if (f0 > 0) b = y;
			else if (f0 < 0) a = y;
			else break;
			y = 0.5 * (a + b);
		}
		const double x = 80.0 * y;
		const double invd = 1.0 / 14.0;
		const double Qb = 2.0 * (P_MW / 160.0) * invd;
		const double QoP = Qb / PI;
		double ci;
		if (x == 0) ci = QoP;
		else {
			const double z = x * invd;
			ci = QoP * (1.0 - z * atan(1.0 / z));
		}
		ci *= 1e3;
		const double qi = Q0[i] + ci;
		ci /= 1e6 * P_MW;
		q[i] = qi; c[i] = ci;
		q[n_data + i] = qi; c[n_data + i] = ci;
	}
};

struct ResilOffsetsBody {
	int64_t N; int64_t n_post; int64_t* set_off; int64_t* post_off; double* w; double* t_all;
	int64_t* t_off; const double* quantiles; int Nq; double* prior; const double* prior_in;
	int64_t half;
	RHFQ_HD void operator()(int64_t r) const {
		set_off[r] = r * N;
		post_off[r] = r;
		t_off[r] = r * Nq;
		if (r == n_post) return;
		w[r] = 1.0;
		for (int l = 0; l < Nq; ++l) t_all[r * Nq + l] = quantiles[l];
		if (r < half) {
			prior[4 * r] = prior_in[0]; prior[4 * r + 1] = prior_in[1];
			prior[4 * r + 2] = prior_in[2]; prior[4 * r + 3] = prior_in[3];
		} else {
			prior[4 * r] = 1.0; prior[4 * r + 1] = 0.0; prior[4 * r + 2] = 0.0; prior[4 * r + 3] = 0.0;
		}
	}
};

/* out_t[(p * Nq + l) * ld + col0 + r] = quantile l of realisation r of the chunk under prior p
 * (NaN if the posterior failed): the [2][Nq][M] layout of zeal2022hfresil.pyx:318-324, either
 * a chunk-sized staging block (ld = nreal, col0 = 0) or the caller's device array */
struct ResilTransposeBody {
	const double* qv; const PostState* P; int64_t nreal; int Nq; double* out_t; int* fail;
	int64_t ld; int64_t col0;
	RHFQ_HD void operator()(int64_t t) const {
		const int64_t r = t % nreal;
		const int64_t pl = t / nreal;          /* p * Nq + l */
		const int64_t p = pl / Nq, l = pl % Nq;
		const int64_t post = p * nreal + r;
		const bool ok = P[post].status == ST_OK;
		out_t[pl * ld + col0 + r] = ok ? qv[post * Nq + l] : nan("");
		if (!ok && l == 0) atomic_add_i32(&fail[p], 1);
	}
};

/* samples[r][4] = samples kept by the (informed low, informed high, flat low, flat high)
 * interpolants of realisation r of the chunk, 0 for a failed posterior (diagnostics) */
struct ResilLevelsBody {
	const PostState* P; int64_t nreal; int* samples;
	RHFQ_HD void operator()(int64_t t) const {
		const int64_t r = t >> 2;
		const int j = (int)(t & 3);
		const PostState& ps = P[(j >> 1) * nreal + r];
		samples[t] = (ps.status == ST_OK && ps.level[j & 1] >= 0) ? (8 << ps.level[j & 1]) + 1 : 0;
	}
};

struct ResilResult {
	int64_t fail_proper = 0, fail_improper = 0;
	Counters cnt{};
	unsigned long long launches = 0;
	double node_ms = 0.0;
	double datagen_ms = 0.0;
};

/*
 * Runs the realisations that rank `shard_rank` of `shard_world` owns (ResilShard) of an
 * M-realisation experiment on `nstreams` RNG streams.  out: [2][Nq][M_local] in local order,
 * on the host or (out_on_device) in device memory.
 */
inline void resilience_run(const ResilDist& dist, int64_t N, int64_t M, int shard_rank, int shard_world,
                           double P_MW, const double* quantiles, int Nq, const double* prior4,
                           double amin, double tol, uint64_t seed, int nstreams, int device,
                           int64_t chunk, double* out, bool out_on_device, ResilResult& res, Stream& st,
                           double* dump_q, double* dump_c, double* dump_u, double* dump_q0,
                           int* dump_samples)
{
	if (N <= 0 || M <= 0 || Nq <= 0 || shard_world <= 0 || shard_rank < 0 || shard_rank >= shard_world)
		throw std::runtime_error("resilience: invalid sizes.");
	const ResilShard sh(M, ResilDrawer::resolve_streams(nstreams), shard_rank, shard_world);
	const int64_t M_local = sh.local_count();
	if (M_local == 0) return;
	const int64_t per = ResilShard::BATCH * sh.Tr;          /* realisations per cycle */
	if (chunk <= 0) chunk = 32760;
	int64_t cyc_per_block = std::max<int64_t>(1, chunk / per);
	cyc_per_block = std::min<int64_t>(cyc_per_block, sh.n_cycles);
	chunk = cyc_per_block * per;
	/* Blocks of whole cycles; the first one is a quarter block, so that the device waits for
	 * few draws before it starts. */
	struct Block { int64_t c_lo, c_hi; };
	std::vector<Block> blocks;
	for (int64_t c = 0; c < sh.n_cycles;) {
		int64_t len = cyc_per_block;
		if (c == 0 && sh.n_cycles > cyc_per_block) len = std::max<int64_t>(1, cyc_per_block / 4);
		const int64_t hi = std::min<int64_t>(sh.n_cycles, c + len);
		blocks.push_back(Block{c, hi});
		c = hi;
	}
	ResilDrawer drawer(dist, N, sh, seed);
	/* double-buffered host blocks: draws of block k + 1 while the device works on block k.
	 * Page-locked and kept per host thread (grow-only): from pageable memory the 52 MB of a
	 * chunk at N = 100 went to the device at 10 GB/s, 5 ms of the 70 ms a chunk takes. */
#ifndef RHFQ_EMU
	static thread_local PinnedStage pin[4];
	double* pU[2]; double* pQ[2];
	std::unique_ptr<double[]> fallback[4];
	for (int i = 0; i < 4; ++i) {
		double* p = (double*)pin[i].get((size_t)chunk * N * sizeof(double));
		if (!p) { fallback[i].reset(new double[(size_t)chunk * N]); p = fallback[i].get(); }
		(i < 2 ? pU[i] : pQ[i - 2]) = p;
	}
#else
	std::unique_ptr<double[]> hU[2], hQ[2];     /* not value-initialised */
	for (int i = 0; i < 2; ++i) {
		hU[i].reset(new double[(size_t)chunk * N]);
		hQ[i].reset(new double[(size_t)chunk * N]);
	}
	double* const pU[2] = {hU[0].get(), hU[1].get()};
	double* const pQ[2] = {hQ[0].get(), hQ[1].get()};
#endif
	double draw_ms = 0.0;
	auto draw_block = [&](size_t k, int slot) {
		const auto t0 = std::chrono::steady_clock::now();
		drawer.draw(blocks[k].c_lo, blocks[k].c_hi, pU[slot], pQ[slot]);
		draw_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	};

	DBuf<double> dU, dQ0, dq, dc, dw, dprior, dprior_in, dquant, dt, dout, dT[2];
	/* host output: the result block of chunk k is fetched by a helper thread on a copy stream
	 * (page-locked pieces scattered straight into `out`) while the device works on chunk k + 1 */
	SideCopy* const side_p = &side_copy(device);
	DBuf<int> dfail;
	dfail.alloc(2);
	DBuf<int64_t> dset_off, dpost_off, dt_off;
	dprior_in.alloc(4); dquant.alloc(Nq);
	h2d(dprior_in.p, prior4, 4 * sizeof(double), st);
	h2d(dquant.p, quantiles, Nq * sizeof(double), st);
	Trace tr;
	/* the large scratch of the chunks' batches, reused from chunk to chunk and -- through the
	 * per-device cache -- from call to call */
	Workspace ws;
	workspace_cache().adopt(ws, device);
	struct Stash { Workspace& w; int device; Stream& st;
		~Stash() { try { dsync(st); } catch (...) {} workspace_cache().stash(w, device); } } stash_ws{ws, device, st};
	std::future<void> fdraw = std::async(std::launch::async, draw_block, (size_t)0, 0);
	std::future<void> fout;          /* scatter of the previous block's results into `out` */
	for (size_t kb = 0; kb < blocks.size(); ++kb) {
		const int slot = (int)(kb & 1);
		fdraw.get();
		if (!drawer.ok)
			throw std::runtime_error("Could not determine a positive q.");
		if (kb + 1 < blocks.size())
			fdraw = std::async(std::launch::async, draw_block, kb + 1, slot ^ 1);
		const int64_t l0 = sh.local_begin(blocks[kb].c_lo), l1 = sh.local_begin(blocks[kb].c_hi);
		const int64_t nreal = l1 - l0;
		tr.stage("resilience host draws (exposed)", st, nreal * N);
		if (nreal <= 0) continue;
		if (dump_u) std::memcpy(dump_u + l0 * N, pU[slot], (size_t)nreal * N * sizeof(double));
		if (dump_q0) std::memcpy(dump_q0 + l0 * N, pQ[slot], (size_t)nreal * N * sizeof(double));
		const int64_t n_data = nreal * N;
		const int64_t n_post = 2 * nreal;
		dU.ensure(n_data); dQ0.ensure(n_data); dq.ensure(2 * n_data); dc.ensure(2 * n_data);
		dw.ensure(n_post); dprior.ensure(4 * n_post); dt.ensure(n_post * Nq); dout.ensure(n_post * Nq);
		dset_off.ensure(n_post + 1); dpost_off.ensure(n_post + 1); dt_off.ensure(n_post + 1);
		h2d(dU.p, pU[slot], n_data * sizeof(double), st);
		h2d(dQ0.p, pQ[slot], n_data * sizeof(double), st);
		launch(n_data, ResilDataBody{dU.p, dQ0.p, n_data, P_MW, dq.p, dc.p}, st, &res.launches);
		launch(n_post + 1, ResilOffsetsBody{N, n_post, dset_off.p, dpost_off.p, dw.p, dt.p, dt_off.p,
		                                    dquant.p, Nq, dprior.p, dprior_in.p, nreal}, st,
		       &res.launches);
		if (dump_q || dump_c) {
			std::vector<double> tmp(n_data);
			if (dump_q) { d2h(tmp.data(), dq.p, n_data * sizeof(double), st);
			              std::memcpy(dump_q + l0 * N, tmp.data(), n_data * sizeof(double)); }
			if (dump_c) { d2h(tmp.data(), dc.p, n_data * sizeof(double), st);
			              std::memcpy(dump_c + l0 * N, tmp.data(), n_data * sizeof(double)); }
		}
		tr.stage("resilience device datagen", st, n_data);
		Batch B;
		B.device = device;
		B.st = st;
		B.exchange(ws);
		struct GiveBack { Batch& b; Workspace& w; ~GiveBack() { b.exchange(w); } } give_back{B, ws};
		B.create(dq.p, dc.p, dset_off.p, dw.p, dpost_off.p, n_post, dprior.p, 4, amin, tol,
		         Algo::BLI, 5, true);
		tr.stage("resilience posteriors", st, n_post);
		B.query(2, n_post * Nq, dt_off.p, dt.p, dout.p, true);
		tr.stage("resilience quantiles", st, n_post * Nq);
		if (dump_samples) {
			DBuf<int> dl;
			dl.alloc(4 * nreal);
			launch(4 * nreal, ResilLevelsBody{B.P.p, nreal, dl.p}, st, &res.launches);
			d2h(dump_samples + 4 * l0, dl.p, 4 * nreal * sizeof(int), st);
		}
		dzero(dfail.p, 2 * sizeof(int), st);
		if (out_on_device) {
			launch(n_post * Nq, ResilTransposeBody{dout.p, B.P.p, nreal, Nq, out, dfail.p, M_local, l0},
			       st, &res.launches);
		} else {
			if (fout.valid()) fout.get();        /* one fetch at a time: frees the stage and dT[slot] */
			dT[slot].ensure(n_post * Nq);
			launch(n_post * Nq, ResilTransposeBody{dout.p, B.P.p, nreal, Nq, dT[slot].p, dfail.p, nreal, 0},
			       st, &res.launches);
			side_p->mark(st);
			const double* src = dT[slot].p;
			const size_t bytes = (size_t)n_post * Nq * sizeof(double);
			fout = std::async(std::launch::async, [=]() {
				side_p->fetch(device, src, bytes, [=](size_t off, const void* p, size_t m) {
					/* piece = doubles [i0, i1) of the [2 Nq][nreal] block */
					const double* v = (const double*)p;
					int64_t i = (int64_t)(off / sizeof(double));
					const int64_t i1 = i + (int64_t)(m / sizeof(double)), i0 = i;
					while (i < i1) {
						const int64_t pl = i / nreal, col = i % nreal;
						const int64_t len = std::min<int64_t>(nreal - col, i1 - i);
						std::memcpy(out + pl * M_local + l0 + col, v + (i - i0), len * sizeof(double));
						i += len;
					}
				});
			});
		}
		int hfail[2] = {0, 0};
		d2h(hfail, dfail.p, 2 * sizeof(int), st);
		res.fail_proper += hfail[0];
		res.fail_improper += hfail[1];
		Counters cc;
		d2h(&cc, B.cnt.p, sizeof(cc), st);
		res.cnt.nodes += cc.nodes; res.cnt.lz_nodes += cc.lz_nodes; res.cnt.g_evals += cc.g_evals;
		res.cnt.newton += cc.newton; res.cnt.nsum += cc.nsum; res.cnt.z_unconverged += cc.z_unconverged;
		res.launches += B.launches;
		res.node_ms += B.node_timer.total_ms();
		tr.stage("resilience output", st, n_post * Nq);
	}
	if (fout.valid()) fout.get();
	res.datagen_ms = draw_ms;
	dsync(st);
}

} // namespace rhfq
