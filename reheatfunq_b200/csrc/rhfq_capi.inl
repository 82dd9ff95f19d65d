/*
 * C-ABI implementation over rhfq::Batch (see include/reheatfunq_b200.h).
 * Included by rhfq_engine.cu (product, symbols rhfq_*) and by the test-only
 * host emulation (symbols emu_rhfq_*).
 */
#include <mutex>
#ifndef RHFQ_API
#error "RHFQ_API(name) must be defined"
#endif

struct rhfq_batch {
	rhfq::Batch impl;
	/* Calls on one handle are serialised: evaluation shares the handle's stream, its scratch
	 * buffers and the lazily built Simpson trees.  The reference releases the GIL around its
	 * evaluation calls (postbackend.pyx:284-288), so several Python threads may be inside
	 * pdf / cdf / tail of the same object at once. */
	std::mutex mtx;
};

namespace {

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
		return fun();
	} catch (const std::bad_alloc&) {
		set_err(err, errlen, "out of host memory");
		return 1;
	} catch (const std::exception& e) {
		set_err(err, errlen, e.what());
		const bool cuda = std::strncmp(e.what(), "CUDA error", 10) == 0;
		return cuda ? 7 : 6;
	} catch (...) {
		set_err(err, errlen, "unknown error");
		return 1;
	}
}

/* Every entry point leaves the caller's current CUDA device as it found it (torch code in the
 * same process relies on it). */
struct DeviceScope {
#ifndef RHFQ_EMU
	int prev = -1;
	DeviceScope() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
	~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
#endif
};

/* offsets[0] == 0, non-decreasing, offsets[n] == total */
bool offsets_valid(const int64_t* off, int64_t n, int64_t total)
{
	if (off[0] != 0 || off[n] != total) return false;
	for (int64_t i = 0; i < n; ++i)
		if (off[i + 1] < off[i]) return false;
	return true;
}

#ifndef RHFQ_EMU
int select_device(int device, char* err, size_t errlen)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
		set_err(err, errlen, "reheatfunq_b200: no CUDA device available (there is no CPU fallback).");
		return 7;
	}
	if (device < 0 || device >= n) {
		set_err(err, errlen, "reheatfunq_b200: invalid device ordinal.");
		return 6;
	}
	if (cudaSetDevice(device) != cudaSuccess) {
		set_err(err, errlen, "reheatfunq_b200: cudaSetDevice failed.");
		return 7;
	}
	static thread_local int pool_ready = -1;
	if (pool_ready != device) {
		rhfq::pool_setup(device);
		pool_ready = device;
	}
	return 0;
}
#else
int select_device(int, char*, size_t) { return 0; }
#endif

using rhfq::workspace_cache;

void bind_stream(rhfq_batch* b)
{
#ifndef RHFQ_EMU
	rhfq::alloc_stream() = b->impl.st.s;
#else
	(void)b;
#endif
}

int do_query(rhfq_batch* b, int mode, const double* x, const int64_t* x_off, int64_t n_points,
             double* out, int on_device, int32_t* qstatus, char* err, size_t errlen)
{
	if (!b || !x_off || (n_points > 0 && (!out || !x)) || n_points < 0) {
		set_err(err, errlen, "null argument");
		return 6;
	}
	DeviceScope scope;
	std::lock_guard<std::mutex> lock(b->mtx);
	int rc = select_device(b->impl.device, err, errlen);
	if (rc) return rc;
	bind_stream(b);
	return guarded(err, errlen, [&]() -> int {
		/* the offsets decide how far x is read and out is written: check them against the
		 * declared number of points before anything is touched */
		const int64_t R = b->impl.R;
		std::vector<int64_t> h_off;
		const int64_t* off = x_off;
		if (on_device) {
			h_off.resize(R + 1);
			rhfq::d2h(h_off.data(), x_off, (R + 1) * sizeof(int64_t), b->impl.st);
			off = h_off.data();
		}
		if (!offsets_valid(off, R, n_points)) {
			set_err(err, errlen, "x_off must start at 0, be non-decreasing and end at the number of points.");
			return 6;
		}
		b->impl.query(mode, n_points, x_off, x, out, on_device != 0);
		if (qstatus) {
			for (int64_t r = 0; r < R; ++r) {
				int s = b->impl.h_P[r].status;
				if (s == 0 && n_points > 0) s = b->impl.h_query_err[r];
				qstatus[r] = s;
			}
		}
		return 0;
	});
}

} // namespace

extern "C" {

const char* RHFQ_API(version)(void) { return "reheatfunq_b200 0.1.0 (sm_100a)"; }

int RHFQ_API(device_count)(void)
{
#ifndef RHFQ_EMU
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
#else
	return 1;
#endif
}

const char* RHFQ_API(status_message)(int status)
{
	switch (status) {
	case 0: return "OK";
	case 1: return "Numerical failure (NaN or infinite result).";
	case 2: return "Input outside the model definition space (q <= 0, no positive c_i, two data "
	               "points with q_i/c_i equal to Qmax, NaN/zero/negative weight, or quantile "
	               "out of bounds [0,1]).";
	case 3: return "Posterior not normalizeable (n >= v is required; n == v degenerate).";
	case 4: return "Quadrature did not reach its tolerance (precision error).";
	case 5: return "Could not determine root.";
	case 6: return "Invalid argument.";
	case 7: return "No CUDA device available (there is no CPU fallback).";
	default: return "Unknown status.";
	}
}

#ifndef RHFQ_EMU
/* Streams of destroyed batches are kept (per device, a few) for the next batch: creating one
 * costs 20-130 us (470 us under a profiler) of the ~1.5 ms a single-posterior batch takes. */
struct StreamPool {
	std::mutex mtx;
	std::vector<cudaStream_t> free_[16];
	cudaStream_t take(int device) {
		if (device >= 0 && device < 16) {
			std::lock_guard<std::mutex> lk(mtx);
			auto& v = free_[device];
			if (!v.empty()) { cudaStream_t s = v.back(); v.pop_back(); return s; }
		}
		cudaStream_t s = nullptr;
		RHFQ_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
		return s;
	}
	/* s has been synchronised */
	void give(int device, cudaStream_t s) {
		if (!s) return;
		if (device >= 0 && device < 16) {
			std::lock_guard<std::mutex> lk(mtx);
			auto& v = free_[device];
			if (v.size() < 8) { v.push_back(s); return; }
		}
		cudaStreamDestroy(s);
	}
	void release(int device) {
		std::lock_guard<std::mutex> lk(mtx);
		for (int d = 0; d < 16; ++d) {
			if (device >= 0 && d != device) continue;
			for (auto s : free_[d]) cudaStreamDestroy(s);
			free_[d].clear();
		}
	}
};
inline StreamPool& stream_pool() { static StreamPool* p = new StreamPool(); return *p; }
#endif

int RHFQ_API(batch_create)(rhfq_batch** out, const double* q, const double* c,
                           const int64_t* set_off, const double* w, const int64_t* post_off,
                           int64_t R, int64_t n_sets, int64_t n_data, const double* prior,
                           int prior_stride, double amin,
                           double rtol, int algorithm, int bli_max_refinements, int device,
                           int flags, char* err, size_t errlen)
{
	if (!out || !q || !c || !set_off || !w || !post_off || !prior) {
		set_err(err, errlen, "null argument");
		return 6;
	}
	if (prior_stride != 0 && prior_stride != 4) {
		set_err(err, errlen, "prior_stride must be 0 or 4");
		return 6;
	}
	if (algorithm < 0 || algorithm > 2) {
		set_err(err, errlen, "`pdf_algorithm` must be one of 'explicit', 'barycentric_lagrange', "
		                     "or 'adaptive_simpson'.");
		return 6;
	}
	*out = nullptr;
	if (R <= 0 || n_sets < 0 || n_data < 0) {
		set_err(err, errlen, R <= 0 ? "No posteriors given." : "negative size");
		return 6;
	}
	DeviceScope scope;
	int rc = select_device(device, err, errlen);
	if (rc) return rc;
	return guarded(err, errlen, [&]() -> int {
		std::unique_ptr<rhfq_batch> b(new rhfq_batch());
		b->impl.device = device;
#ifndef RHFQ_EMU
		b->impl.st.s = stream_pool().take(device);
#endif
		bind_stream(b.get());
		workspace_cache().adopt(b->impl, device);
		if (flags & 2) b->impl.long_double_thresholds();
		try {
			b->impl.create(q, c, set_off, w, post_off, R, prior, prior_stride, amin, rtol, algorithm,
			               bli_max_refinements, (flags & 1) != 0, n_sets, n_data);
		} catch (...) {
#ifndef RHFQ_EMU
			/* buffers go back to the pool in stream order, then the stream itself */
			cudaStream_t s = b->impl.st.s;
			if (s) cudaStreamSynchronize(s);
			b.reset();
			if (s) { cudaStreamSynchronize(s); stream_pool().give(device, s); }
			rhfq::alloc_stream() = nullptr;
#endif
			throw;
		}
		*out = b.release();
		return 0;
	});
}

void RHFQ_API(batch_destroy)(rhfq_batch* b)
{
	if (!b) return;
	DeviceScope scope;
#ifndef RHFQ_EMU
	cudaSetDevice(b->impl.device);
	cudaStream_t s = b->impl.st.s;
	bind_stream(b);
	if (s) cudaStreamSynchronize(s);
	workspace_cache().stash(b->impl, b->impl.device);
	const int dev = b->impl.device;
	delete b;            /* buffers go back to the pool, ordered on s */
	if (s) { cudaStreamSynchronize(s); stream_pool().give(dev, s); }
	rhfq::alloc_stream() = nullptr;
#else
	delete b;
#endif
}

/* Frees the scratch kept from destroyed batches (device < 0: on every device). */
void RHFQ_API(cache_release)(int device)
{
	DeviceScope scope;
#ifndef RHFQ_EMU
	for (int d = 0; d < 16; ++d) {
		if (device >= 0 && d != device) continue;
		if (cudaSetDevice(d) != cudaSuccess) { cudaGetLastError(); continue; }
		rhfq::alloc_stream() = nullptr;
		workspace_cache().release(d);
		cudaDeviceSynchronize();
	}
	stream_pool().release(device);
#else
	workspace_cache().release(device);
#endif
}

int64_t RHFQ_API(batch_size)(const rhfq_batch* b) { return b ? b->impl.R : 0; }

int RHFQ_API(batch_qmax)(const rhfq_batch* b, double* qmax)
{
	if (!b || !qmax) return 6;
	for (int64_t r = 0; r < b->impl.R; ++r) qmax[r] = b->impl.h_P[r].Qmax;
	return 0;
}

int RHFQ_API(batch_status)(const rhfq_batch* b, int32_t* status)
{
	if (!b || !status) return 6;
	for (int64_t r = 0; r < b->impl.R; ++r) status[r] = b->impl.h_P[r].status;
	return 0;
}

int RHFQ_API(batch_pdf)(rhfq_batch* b, const double* x, const int64_t* x_off, int64_t n_points,
                        double* out, int on_device, int32_t* qstatus, char* err, size_t errlen)
{
	return do_query(b, 4, x, x_off, n_points, out, on_device, qstatus, err, errlen);
}

int RHFQ_API(batch_cdf)(rhfq_batch* b, const double* x, const int64_t* x_off, int64_t n_points,
                        double* out, int on_device, int32_t* qstatus, char* err, size_t errlen)
{
	return do_query(b, 0, x, x_off, n_points, out, on_device, qstatus, err, errlen);
}

int RHFQ_API(batch_tail)(rhfq_batch* b, const double* x, const int64_t* x_off, int64_t n_points,
                        double* out, int on_device, int32_t* qstatus, char* err, size_t errlen)
{
	return do_query(b, 1, x, x_off, n_points, out, on_device, qstatus, err, errlen);
}

int RHFQ_API(batch_tail_quantiles)(rhfq_batch* b, const double* t, const int64_t* t_off,
                                   int64_t n_points, double* out, int on_device, int32_t* qstatus,
                                   char* err, size_t errlen)
{
	return do_query(b, 2, t, t_off, n_points, out, on_device, qstatus, err, errlen);
}

int RHFQ_API(batch_get_locals)(const rhfq_batch* bc, int64_t s, double* out)
{
	rhfq_batch* b = const_cast<rhfq_batch*>(bc);
	if (!b || !out || s < 0 || s >= b->impl.S) return 6;
	DeviceScope scope;
	std::lock_guard<std::mutex> lock(b->mtx);
	if (select_device(b->impl.device, nullptr, 0)) return 7;
	rhfq::SetLocals l;
	try {
		rhfq::d2h(&l, b->impl.L.p + s, sizeof(l), b->impl.st);
	} catch (...) {
		return 7;
	}
	const double v[36] = {l.lp, l.ls, l.n, l.v, l.amin, l.Qmax, l.h0, l.h1, l.h2, l.h3, l.w, l.lh0,
		l.l1p_w, l.lv, l.log_scale, l.ymax, l.ztrans, l.S, l.taylor, l.norm, l.weight, l.log_fac,
		(double)l.N, (double)l.imax, (double)l.status, (double)l.k_off, (double)l.flags,
		l.c1[0], l.c1[1], l.c2[0], l.c2[1], l.c2[2], l.c3[0], l.c3[1], l.c3[2], l.c3[3]};
	std::memcpy(out, v, sizeof(v));
	return 0;
}

/* k_i = c_i Qmax / q_i of sample set s (Locals::ki, locals.hpp:75-87); returns N or < 0 */
int64_t RHFQ_API(batch_get_k)(const rhfq_batch* bc, int64_t s, double* k, int64_t cap)
{
	rhfq_batch* b = const_cast<rhfq_batch*>(bc);
	if (!b || !k || s < 0 || s >= b->impl.S) return -1;
	DeviceScope scope;
	std::lock_guard<std::mutex> lock(b->mtx);
	if (select_device(b->impl.device, nullptr, 0)) return -1;
	try {
		rhfq::SetLocals l;
		rhfq::d2h(&l, b->impl.L.p + s, sizeof(l), b->impl.st);
		const int64_t n = std::min<int64_t>(l.N, cap);
		if (n > 0) rhfq::d2h(k, b->impl.k.p + l.k_off, n * sizeof(double), b->impl.st);
		return l.N;
	} catch (...) {
		return -1;
	}
}

int64_t RHFQ_API(batch_get_bli)(const rhfq_batch* bc, int64_t r, int which, double* f, int64_t cap)
{
	rhfq_batch* b = const_cast<rhfq_batch*>(bc);
	if (!b || r < 0 || r >= b->impl.R || which < 0 || which > 1) return -1;
	if (b->impl.algorithm != rhfq::Algo::BLI || !b->impl.bli_f.p) return -1;
	DeviceScope scope;
	std::lock_guard<std::mutex> lock(b->mtx);
	if (select_device(b->impl.device, nullptr, 0)) return -1;
	const int lev = b->impl.h_P[r].level[which];
	if (lev < 0) return -1;
	const int Rmax = b->impl.Rmax;
	const int64_t n_fine = (8 << Rmax) + 1;
	const int64_t n = (8 << lev) + 1;
	const int64_t stride = 1 << (Rmax - lev);
	std::vector<double> fine(n_fine);
	try {
		rhfq::d2h(fine.data(), b->impl.bli_f.p + (r * 2 + which) * n_fine, n_fine * sizeof(double),
		          b->impl.st);
	} catch (...) {
		return -1;
	}
	for (int64_t i = 0; i < n && i < cap; ++i) f[i] = fine[i * stride];
	return n;
}

int64_t RHFQ_API(batch_saq_leaves)(const rhfq_batch* b) { return b ? b->impl.n_leaves : 0; }

int RHFQ_API(batch_counters)(const rhfq_batch* bc, uint64_t* counters)
{
	rhfq_batch* b = const_cast<rhfq_batch*>(bc);
	if (!b || !counters) return 6;
	DeviceScope scope;
	std::lock_guard<std::mutex> lock(b->mtx);
	if (select_device(b->impl.device, nullptr, 0)) return 7;
	rhfq::Counters c;
	try {
		rhfq::d2h(&c, b->impl.cnt.p, sizeof(c), b->impl.st);
	} catch (...) {
		return 7;
	}
	counters[0] = c.nodes; counters[1] = c.lz_nodes; counters[2] = c.g_evals;
	counters[3] = c.newton; counters[4] = c.nsum; counters[5] = b->impl.launches;
	try {
		rhfq::dsync(b->impl.st);
	} catch (...) {
		return 7;
	}
	counters[6] = (uint64_t)(b->impl.node_timer.total_ms() * 1e6);   /* ns in NodeEvalBody */
	counters[7] = b->impl.node_timer.count;
	counters[8] = (uint64_t)(b->impl.norm_timer.total_ms() * 1e6);   /* ns in NormEvalBody */
	counters[9] = b->impl.norm_timer.count;
	counters[10] = c.cheb_terms; counters[11] = b->impl.saq_steps;
	counters[12] = (uint64_t)(b->impl.saq_timer.total_ms() * 1e6);   /* ns in SaqStepBliBody */
	counters[13] = b->impl.saq_timer.count;
	/* fine-grid evaluations: two per interval except the counted exceptions */
	counters[14] = (b->impl.use_fine && b->impl.algorithm == rhfq::Algo::BLI && 2 * counters[11] >= c.saq_other)
	                   ? 2 * counters[11] - c.saq_other : 0;
	counters[15] = c.fine_bad;
	counters[16] = (uint64_t)(b->impl.plan_timer.total_ms() * 1e6);   /* ns in the mode/window kernels */
	counters[17] = b->impl.plan_timer.count;
	counters[18] = c.g_evals_quad;
	counters[19] = c.z_unconverged;
	counters[20] = c.quad_checks;
	return 0;
}

} // extern "C"

extern "C" int64_t RHFQ_API(resilience_shard_count)(int64_t M, int nstreams, int shard_rank, int shard_world)
{
	if (M <= 0 || shard_world <= 0 || shard_rank < 0 || shard_rank >= shard_world) return -1;
	return rhfq::ResilShard(M, rhfq::ResilDrawer::resolve_streams(nstreams), shard_rank, shard_world).local_count();
}

extern "C" int RHFQ_API(resilience_shard_index)(int64_t M, int nstreams, int shard_rank, int shard_world,
                                                int64_t* index)
{
	if (M <= 0 || shard_world <= 0 || shard_rank < 0 || shard_rank >= shard_world || !index) return 6;
	const rhfq::ResilShard sh(M, rhfq::ResilDrawer::resolve_streams(nstreams), shard_rank, shard_world);
	const int64_t n = sh.local_count();
	for (int64_t l = 0; l < n; ++l) index[l] = sh.global_index(l);
	return 0;
}

extern "C" int RHFQ_API(resilience_run)(int dist_kind, const double* dist_params, int64_t N,
                                        int64_t M, int shard_rank, int shard_world, double P_MW,
                                        const double* quantiles, int Nq, const double* prior,
                                        double amin, double tol, uint64_t seed, int nstreams,
                                        int device, int64_t chunk, double* out, int out_on_device,
                                        int64_t* failures, uint64_t* stats, double* dump_q,
                                        double* dump_c, double* dump_u, double* dump_q0,
                                        int32_t* dump_samples, char* err, size_t errlen)
{
	if (!dist_params || !quantiles || !prior) {
		set_err(err, errlen, "null argument");
		return 6;
	}
	if (!out && RHFQ_API(resilience_shard_count)(M, nstreams, shard_rank, shard_world) != 0) {
		set_err(err, errlen, "null output");
		return 6;
	}
	if (dist_kind != 0 && dist_kind != 1) {
		set_err(err, errlen, "dist_kind must be 0 (gamma) or 1 (mixture)");
		return 6;
	}
	DeviceScope scope;
	int rc = select_device(device, err, errlen);
	if (rc) return rc;
	return guarded(err, errlen, [&]() -> int {
		rhfq::Stream st;
#ifndef RHFQ_EMU
		RHFQ_CUDA_CHECK(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
		rhfq::alloc_stream() = st.s;
#endif
		rhfq::ResilDist d;
		d.kind = dist_kind;
		for (int i = 0; i < (dist_kind == 0 ? 2 : 6); ++i) d.p[i] = dist_params[i];
		rhfq::ResilResult res;
		int ret = 0;
		try {
			rhfq::resilience_run(d, N, M, shard_rank, shard_world, P_MW, quantiles, Nq, prior, amin,
			                     tol, seed, nstreams, device, chunk, out, out_on_device != 0, res, st,
			                     dump_q, dump_c, dump_u, dump_q0, dump_samples);
		} catch (...) {
#ifndef RHFQ_EMU
			cudaStreamSynchronize(st.s);
			cudaStreamDestroy(st.s);
			rhfq::alloc_stream() = nullptr;
#endif
			throw;
		}
#ifndef RHFQ_EMU
		cudaStreamSynchronize(st.s);
		cudaStreamDestroy(st.s);
		rhfq::alloc_stream() = nullptr;
#endif
		if (failures) { failures[0] = res.fail_proper; failures[1] = res.fail_improper; }
		if (stats) {
			stats[0] = res.cnt.nodes; stats[1] = res.cnt.lz_nodes; stats[2] = res.cnt.g_evals;
			stats[3] = res.cnt.newton; stats[4] = res.cnt.nsum; stats[5] = res.launches;
			stats[6] = (uint64_t)(res.node_ms * 1e6); stats[7] = (uint64_t)(res.datagen_ms * 1e6);
			stats[8] = res.cnt.z_unconverged;
		}
		return ret;
	});
}

namespace {
template<typename Fun>
int with_stream(int device, char* err, size_t errlen, Fun fun)
{
	DeviceScope scope;
	int rc = select_device(device, err, errlen);
	if (rc) return rc;
	return guarded(err, errlen, [&]() -> int {
		rhfq::Stream st;
#ifndef RHFQ_EMU
		RHFQ_CUDA_CHECK(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
		rhfq::alloc_stream() = st.s;
#endif
		try {
			fun(st);
		} catch (...) {
#ifndef RHFQ_EMU
			cudaStreamSynchronize(st.s);
			cudaStreamDestroy(st.s);
			rhfq::alloc_stream() = nullptr;
#endif
			throw;
		}
#ifndef RHFQ_EMU
		cudaStreamSynchronize(st.s);
		cudaStreamDestroy(st.s);
		rhfq::alloc_stream() = nullptr;
#endif
		return 0;
	});
}
} // namespace

extern "C" int RHFQ_API(bli_known_answer)(int kind, double xmin, double xmax, double tol_rel,
                                          double tol_abs, double fmin, double fmax,
                                          int max_refinements, const double* x, int64_t n,
                                          double* out, int barycentric, int64_t* info, int device,
                                          char* err, size_t errlen)
{
	if (!x || !out || n <= 0) { set_err(err, errlen, "invalid argument"); return 6; }
	return with_stream(device, err, errlen, [&](rhfq::Stream& st) {
		rhfq::bli_known_answer(kind, xmin, xmax, tol_rel, tol_abs, fmin, fmax, max_refinements, x, n,
		                       out, barycentric, info, st);
	});
}

extern "C" int RHFQ_API(gcp_ln_phi)(const double* params, int64_t K, double amin, double* out,
                                    int32_t* status, int device, char* err, size_t errlen)
{
	if (!params || !out || K <= 0) { set_err(err, errlen, "invalid argument"); return 6; }
	return with_stream(device, err, errlen, [&](rhfq::Stream& st) {
		unsigned long long launches = 0;
		rhfq::gcp_ln_phi(params, K, amin, out, status, st, &launches);
	});
}

extern "C" int RHFQ_API(gcp_kl_batch_max)(const double* params, int64_t N, const double* refs,
                                          int64_t K, double amin, double* out, double* kl_all,
                                          int device, char* err, size_t errlen)
{
	if (!params || !refs || !out || N <= 0 || K <= 0) {
		set_err(err, errlen, "invalid argument");
		return 6;
	}
	return with_stream(device, err, errlen, [&](rhfq::Stream& st) {
		unsigned long long launches = 0;
		rhfq::gcp_kl_batch_max(params, N, refs, K, amin, out, kl_all, st, &launches);
	});
}

/* ---- d_min-restricted sample enumeration (host only, rhfq_coverings.inl) ---- */

struct rhfq_samples {
	rhfq::coverings::RestrictedSamples impl;
};

extern "C" int RHFQ_API(restricted_samples)(rhfq_samples** out, const double* xy, int64_t N,
                                            double dmin, int64_t max_samples, int64_t max_iter,
                                            void* numpy_bitgen, char* err, size_t errlen)
{
	if (!out || !xy || N <= 0 || max_samples < 0 || max_iter < 0) {
		set_err(err, errlen, N <= 0 ? "Empty coordinates." : "invalid argument");
		return 6;
	}
	*out = nullptr;
	return guarded(err, errlen, [&]() {
		std::unique_ptr<rhfq_samples> h(new rhfq_samples());
		rhfq::coverings::determine_restricted_samples(
		    xy, (size_t)N, dmin, (size_t)max_samples, (size_t)max_iter,
		    (rhfq::coverings::bitgen_t*)numpy_bitgen, h->impl);
		*out = h.release();
		return 0;
	});
}

extern "C" int RHFQ_API(bootstrap_data_selection)(rhfq_samples** out, const double* xy, int64_t N,
                                                  double dmin, int64_t B, void* numpy_bitgen,
                                                  char* err, size_t errlen)
{
	if (!out || !xy || N <= 0 || B <= 0 || !numpy_bitgen) {
		set_err(err, errlen, "invalid argument");
		return 6;
	}
	*out = nullptr;
	return guarded(err, errlen, [&]() {
		std::unique_ptr<rhfq_samples> h(new rhfq_samples());
		rhfq::coverings::bootstrap_data_selection(xy, (size_t)N, dmin, (size_t)B,
		                                          (rhfq::coverings::bitgen_t*)numpy_bitgen, h->impl);
		*out = h.release();
		return 0;
	});
}

extern "C" int64_t RHFQ_API(samples_count)(const rhfq_samples* s) { return s ? (int64_t)s->impl.w.size() : 0; }
extern "C" int64_t RHFQ_API(samples_index_count)(const rhfq_samples* s) { return s ? (int64_t)s->impl.idx.size() : 0; }

extern "C" int RHFQ_API(samples_get)(const rhfq_samples* s, double* w, int64_t* off, int64_t* idx)
{
	if (!s || !w || !off || !idx)
		return 6;
	std::copy(s->impl.w.begin(), s->impl.w.end(), w);
	std::copy(s->impl.off.begin(), s->impl.off.end(), off);
	std::copy(s->impl.idx.begin(), s->impl.idx.end(), idx);
	return 0;
}

extern "C" void RHFQ_API(samples_destroy)(rhfq_samples* s) { delete s; }

extern "C" int RHFQ_API(global_phmax)(const double* xy, const double* phmax_i, int64_t N,
                                      double dmin, double* out, char* err, size_t errlen)
{
	if (!xy || !phmax_i || !out || N <= 0) {
		set_err(err, errlen, N <= 0 ? "Empty coordinates." : "invalid argument");
		return 6;
	}
	return guarded(err, errlen, [&]() {
		*out = rhfq::coverings::global_PHmax(xy, phmax_i, (size_t)N, dmin);
		return 0;
	});
}
