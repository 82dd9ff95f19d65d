/*
 * reheatfunq_b200 — batched anomaly-posterior engine.
 *
 * Level-synchronous re-design of the reference's per-posterior, recursive,
 * adaptive C++ (include/anomaly/posterior.hpp:88-1030): every stage is one
 * data-parallel launch over ALL posteriors / sample sets / nodes of the batch,
 * and adaptivity (quadrature levels in z, barycentric refinement levels,
 * Simpson bisection levels) is expressed as "which items are still active"
 * between launches.  Stage -> reference:
 *
 *   LocalsBody      Locals ctor                     locals.hpp:94-113,193-329
 *   Scale/YScan/YBis/YMaxBody  log_integrand_max, y_max  amax.hpp:68-133, large_z.hpp:320-349
 *   NormEval/Reduce compute_integrals (tanh-sinh z) localsandnorm.hpp:208-233
 *   TaylorBody      Taylor tail + norm              localsandnorm.hpp:238-250,159-164
 *   PostBody        init_weights/compute_norm/Qmax  posterior.hpp:582-648
 *   NodeEvalBody    pdf_single_explicit_j           posterior.hpp:838-881
 *   CombineBody     log_pdf_t::sum / pdf_t::sum     posterior.hpp:773-780,814-832
 *   BliErr/Decide   determine_ranges refinement     barylagrint.hpp:576-668
 *   PdfBliBody      pdf_single<BLI>                 posterior.hpp:965-1008, barylagrint.hpp:674-720
 *   Saq*Body        SimpsonAdaptiveQuadrature       simpson.hpp:107-360
 *   Cdf/Tail/Quant  cdf/tail/tail_quantiles         posterior.hpp:150-337
 *
 * This file is compiled twice: by nvcc into the product (kernels for sm_100a)
 * and, with RHFQ_EMU defined, by g++ into a TEST-ONLY helper that runs each
 * "kernel" as a host loop so the pipeline logic can be debugged without a GPU.
 * The product library never contains the emulation.
 */
#include "rhfq_math.cuh"
#include <vector>
#include <string>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <stdexcept>
#include <algorithm>
#include <memory>
#include <chrono>
#include <map>
#include <tuple>
#include <mutex>

namespace rhfq {

/* RHFQ_TRACE=1: per-stage wall times (stream synchronised) on stderr */
#ifndef RHFQ_EMU
/* reserved size of the device's default memory pool (a stage that makes it grow pays for the
 * driver mapping new physical memory) */
inline double trace_pool_mb()
{
	cudaMemPool_t pool;
	int dev = 0;
	unsigned long long v = 0;
	if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
		cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &v);
	return (double)v / 1048576.0;
}
#else
inline double trace_pool_mb() { return 0.0; }
#endif
struct Trace {
	bool on;
	std::chrono::steady_clock::time_point t0;
	Trace() : on(std::getenv("RHFQ_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
	template<typename S>
	void stage(const char* name, S& st, long long items = -1) {
		if (!on) return;
		dsync(st);
		auto t1 = std::chrono::steady_clock::now();
		std::fprintf(stderr, "[rhfq] %-28s %9.3f ms  items=%lld  pool=%.0f MB\n", name,
		             std::chrono::duration<double, std::milli>(t1 - t0).count(), items, trace_pool_mb());
		t0 = t1;
	}
};

/* ------------------------------------------------------------------ *
 *  Launch / memory abstraction                                        *
 * ------------------------------------------------------------------ */
#ifdef RHFQ_EMU

struct Stream {};
inline void* dmalloc(size_t n) { return std::calloc(1, n ? n : 1); }
inline void dfree(void* p) { std::free(p); }
inline void h2d(void* d, const void* h, size_t n, Stream&) { if (n) std::memcpy(d, h, n); }
inline void d2h(void* h, const void* d, size_t n, Stream&) { if (n) std::memcpy(h, d, n); }
inline void d2h_async(void* h, const void* d, size_t n, Stream&) { if (n) std::memcpy(h, d, n); }
struct SideCopy {
	void mark(Stream&) {}
	template<typename F> void fetch(int, const void* d, size_t n, F sink) { if (n) sink((size_t)0, d, n); }
};
inline SideCopy& side_copy(int) { static thread_local SideCopy c; return c; }
inline void d2d(void* d, const void* s, size_t n, Stream&) { if (n) std::memcpy(d, s, n); }
inline void dzero(void* d, size_t n, Stream&) { if (n) std::memset(d, 0, n); }
inline void dsync(Stream&) {}
/* A launch sized from an upper bound whose true extent lives in device memory: task i is
 * live if (i mod period) < *count * per (period 0: i itself).  Lets the host queue the next
 * refinement level without reading the number of still-active items back first. */
struct Limit {
	const int* count = nullptr;
	int64_t per = 1;
	int64_t period = 0;
	RHFQ_HD bool live(int64_t i) const {
		if (!count) return true;
		const int64_t p = period ? i % period : i;
		return p < (int64_t)*count * per;
	}
};
template<typename B>
inline void launch(int64_t n, const B& b, Stream&, unsigned long long* launches, Limit lim = Limit())
{
	if (n <= 0) return;
	if (launches) ++*launches;
	for (int64_t i = 0; i < n; ++i) if (lim.live(i)) b(i);
}
struct KernelTimer {
	double total_ms() { return 0.0; }
	unsigned long long count = 0;
	bool enabled = true;
};
template<typename B>
inline void launch_timed(int64_t n, const B& b, Stream& st, unsigned long long* launches, KernelTimer& t,
                         Limit lim = Limit())
{
	if (n > 0) ++t.count;
	launch(n, b, st, launches, lim);
}
/* device counts mirrored to the host with a lag (see the CUDA version) */
struct CountMirror {
	int h[32] = {};
	explicit CountMirror(int = 0) {}
	void record(int slot, const int* d, Stream&) { h[slot] = *d; }
	int get(int slot) { return h[slot]; }
};
RHFQ_HD void atomic_add_u64(unsigned long long* p, unsigned long long v) { *p += v; }
RHFQ_HD int atomic_add_i32(int* p, int v) { int o = *p; *p += v; return o; }
RHFQ_HD void atomic_max_i32(int* p, int v) { if (v > *p) *p = v; }
RHFQ_HD unsigned long long atomic_add_u64_ret(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
RHFQ_HD int atomic_cas_i32(int* p, int expect, int v) { const int o = *p; if (o == expect) *p = v; return o; }
RHFQ_HD void atomic_add_u64_plain(unsigned long long* p, unsigned long long v) { *p += v; }
RHFQ_HD int claim_pair_i32(int* counter) { const int o = *counter; *counter += 2; return o; }
/* out[i] = sum of in[j] over j < i (forward) or j > i (backward) within runs of equal keys */
inline void segmented_exclusive_sums(const int* keys, const double* in, double* fw, double* bw,
                                     int64_t n, Stream&, void*&, size_t&)
{
	double acc = 0.0;
	for (int64_t i = 0; i < n; ++i) {
		if (i == 0 || keys[i] != keys[i - 1]) acc = 0.0;
		fw[i] = acc; acc += in[i];
	}
	for (int64_t i = n - 1; i >= 0; --i) {
		if (i == n - 1 || keys[i] != keys[i + 1]) acc = 0.0;
		bw[i] = acc; acc += in[i];
	}
}
inline void exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, Stream&, void*&, size_t&)
{
	int64_t acc = 0;
	for (int64_t i = 0; i < n; ++i) { out[i] = acc; acc += in[i]; }
	out[n] = acc;
}

#else  /* CUDA */

struct Stream { cudaStream_t s = nullptr; };

#define RHFQ_CUDA_CHECK(expr)                                                        \
	do {                                                                             \
		cudaError_t e__ = (expr);                                                    \
		if (e__ != cudaSuccess)                                                      \
			throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e__) \
			                         + " at " #expr);                                \
	} while (0)

/* Stream-ordered allocation from the device's default memory pool: the many
 * per-level scratch buffers of the level-synchronous pipeline are recycled by
 * the pool instead of paying cudaMalloc/cudaFree (which synchronise) each time. */
inline cudaStream_t& alloc_stream()
{
	static thread_local cudaStream_t s = nullptr;
	return s;
}
/* Large requests are rounded up to size classes m * 2^k, m in {4..7} (<= 25 % slack).  Many
 * scratch sizes are data dependent (active sets, frontier lengths); without classes the
 * pool splits its big free blocks for slightly different requests and the next multi-GB
 * request makes the driver re-map physical memory (measured: 20-230 ms spikes per step). */
inline size_t alloc_size_class(size_t n)
{
	if (n < ((size_t)1 << 20)) return n ? n : 8;
	int k = 0;
	while ((n >> k) >= 8) ++k;          /* n >> k in [4, 8) */
	const size_t m = (n + (((size_t)1 << k) - 1)) >> k;
	return m << k;
}
inline void* dmalloc(size_t n)
{
	void* p = nullptr;
	RHFQ_CUDA_CHECK(cudaMallocAsync(&p, alloc_size_class(n), alloc_stream()));
	return p;
}
inline void dfree(void* p) { if (p) cudaFreeAsync(p, alloc_stream()); }
inline void pool_setup(int device)
{
	cudaMemPool_t pool;
	if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
		unsigned long long thr = ~0ull;
		cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
	}
}
inline void h2d(void* d, const void* h, size_t n, Stream& st)
{
	if (n) RHFQ_CUDA_CHECK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, st.s));
}
/* Read-backs of 16 KB and more (posterior records, offsets, table keys, result blocks) go
 * through a 4 MB page-locked staging buffer kept per host thread: a pageable destination makes the driver
 * copy at 3-6 GB/s in small pieces (0.2 ms for the 720 KB of posterior records, three times
 * per batch). */
struct PinnedStage {
	void* p = nullptr;
	size_t n = 0;
	~PinnedStage() { if (p) cudaFreeHost(p); }
	void* get(size_t need) {
		if (need > n) {
			if (p) cudaFreeHost(p);
			p = nullptr; n = 0;
			const size_t cap = std::max<size_t>(need, (size_t)4 << 20);
			if (cudaMallocHost(&p, cap) != cudaSuccess) { cudaGetLastError(); p = nullptr; return nullptr; }
			n = cap;
		}
		return p;
	}
};
inline void d2h(void* h, const void* d, size_t n, Stream& st)
{
	if (!n) return;
	if (n >= ((size_t)16 << 10)) {
		/* larger blocks (the resilience result block, 21 MB) in pieces of the stage size:
		 * page-locking a buffer of their own size costs more than it saves unless the call is
		 * repeated many times (measured: ~10 ms per MB on the bench hosts) */
		static thread_local PinnedStage stage;
		const size_t piece = (size_t)4 << 20;
		if (void* s = stage.get(std::min(n, piece))) {
			for (size_t o = 0; o < n; o += piece) {
				const size_t m = std::min(piece, n - o);
				RHFQ_CUDA_CHECK(cudaMemcpyAsync(s, (const char*)d + o, m, cudaMemcpyDeviceToHost, st.s));
				RHFQ_CUDA_CHECK(cudaStreamSynchronize(st.s));
				std::memcpy((char*)h + o, s, m);
			}
			return;
		}
	}
	RHFQ_CUDA_CHECK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, st.s));
	RHFQ_CUDA_CHECK(cudaStreamSynchronize(st.s));
}
/* Device block -> host on a copy stream of its own, in page-locked pieces handed to `sink`
 * (offset, pointer, bytes): a helper thread fetches a result block while the producing stream
 * already works on the next one.  mark() is called by the producer after the kernel that fills
 * the block, fetch() by the helper. */
struct SideCopy {
	cudaStream_t s = nullptr;
	cudaEvent_t ready = nullptr;
	cudaEvent_t done[2] = {nullptr, nullptr};
	PinnedStage stage[2];
	~SideCopy() {
		if (ready) cudaEventDestroy(ready);
		for (auto e : done) if (e) cudaEventDestroy(e);
		if (s) cudaStreamDestroy(s);
	}
	void mark(Stream& producer) {
		if (!s) RHFQ_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
		if (!ready) RHFQ_CUDA_CHECK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
		for (auto& e : done)
			if (!e) RHFQ_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		for (auto& st : stage) st.get(PIECE);
		RHFQ_CUDA_CHECK(cudaEventRecord(ready, producer.s));
	}
	static constexpr size_t PIECE = (size_t)4 << 20;
	/* two staging buffers: piece k + 1 is on its way while the sink works on piece k */
	template<typename F>
	void fetch(int device, const void* d, size_t n, F sink) {
		RHFQ_CUDA_CHECK(cudaSetDevice(device));
		RHFQ_CUDA_CHECK(cudaStreamWaitEvent(s, ready, 0));
		if (!stage[0].p || !stage[1].p) throw std::runtime_error("no page-locked staging buffer");
		auto issue = [&](size_t o, int b) {
			const size_t m = std::min(PIECE, n - o);
			RHFQ_CUDA_CHECK(cudaMemcpyAsync(stage[b].p, (const char*)d + o, m, cudaMemcpyDeviceToHost, s));
			RHFQ_CUDA_CHECK(cudaEventRecord(done[b], s));
		};
		if (n) issue(0, 0);
		int b = 0;
		for (size_t o = 0; o < n; o += PIECE, b ^= 1) {
			if (o + PIECE < n) issue(o + PIECE, b ^ 1);
			RHFQ_CUDA_CHECK(cudaEventSynchronize(done[b]));
			sink(o, (const void*)stage[b].p, std::min(PIECE, n - o));
		}
	}
};
/* The copy stream, events and staging buffers of the calling host thread for `device`, kept
 * across calls: creating them took 4.2 ms per resilience call (cudaMallocHost of the stage). */
inline SideCopy& side_copy(int device)
{
	static thread_local std::map<int, std::unique_ptr<SideCopy>> cache;
	auto& p = cache[device];
	if (!p) p.reset(new SideCopy());
	return *p;
}
/* no synchronisation: h must stay valid (page-locked for a truly asynchronous copy) until the
 * stream has been synchronised */
inline void d2h_async(void* h, const void* d, size_t n, Stream& st)
{
	if (n) RHFQ_CUDA_CHECK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, st.s));
}
inline void d2d(void* d, const void* s, size_t n, Stream& st)
{
	if (n) RHFQ_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st.s));
}
inline void dzero(void* d, size_t n, Stream& st)
{
	if (n) RHFQ_CUDA_CHECK(cudaMemsetAsync(d, 0, n, st.s));
}
inline void dsync(Stream& st) { RHFQ_CUDA_CHECK(cudaStreamSynchronize(st.s)); }

/* A launch sized from an upper bound whose true extent lives in device memory: task i is
 * live if (i mod period) < *count * per (period 0: i itself).  Lets the host queue the next
 * refinement level without reading the number of still-active items back first. */
struct Limit {
	const int* count = nullptr;
	int64_t per = 1;
	int64_t period = 0;
	RHFQ_HD bool live(int64_t i) const {
		if (!count) return true;
		const int64_t p = period ? i % period : i;
		return p < (int64_t)*count * per;
	}
};

/* One generic kernel template; the functor type names the stage in ncu. */
template<typename B>
__global__ void __launch_bounds__(128) rhfq_kernel(int64_t n, B b, Limit lim)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && lim.live(i)) b(i);
}

template<typename B>
inline void launch(int64_t n, const B& b, Stream& st, unsigned long long* launches, Limit lim = Limit())
{
	if (n <= 0) return;
	if (launches) ++*launches;
	const int block = 128;
	const int64_t grid = (n + block - 1) / block;
	rhfq_kernel<B><<<(unsigned)grid, block, 0, st.s>>>(n, b, lim);
	RHFQ_CUDA_CHECK(cudaGetLastError());
}

/* Device-side counts mirrored to the host WITH A LAG: record() queues a 4-byte copy into
 * page-locked memory and an event behind the kernel that produced the count; get() waits for
 * that event only.  The refinement loops use the count of two levels ago as the bound of the
 * level they are about to queue -- by then the copy has long completed, so the host never
 * drains the stream (it used to, ~25 times per batch). */
struct CountMirror {
	static constexpr int SLOTS = 32;
	int* h = nullptr;
	cudaEvent_t ev[SLOTS] = {};
	~CountMirror() {
		for (auto e : ev) if (e) cudaEventDestroy(e);
	}
	/* page-locked slots: one small allocation per host thread, kept (two mirrors can be alive
	 * at a time: the norm loop's and the interpolant loop's never overlap, `which` separates
	 * a nested use) */
	static int* slots(int which) {
		static thread_local int* base = nullptr;
		if (!base) RHFQ_CUDA_CHECK(cudaHostAlloc((void**)&base, 4 * SLOTS * sizeof(int), cudaHostAllocPortable));
		return base + which * SLOTS;
	}
	explicit CountMirror(int which = 0) { h = slots(which); }
	void record(int slot, const int* d, Stream& st) {
		if (!ev[slot]) RHFQ_CUDA_CHECK(cudaEventCreateWithFlags(&ev[slot], cudaEventDisableTiming));
		RHFQ_CUDA_CHECK(cudaMemcpyAsync(h + slot, d, sizeof(int), cudaMemcpyDeviceToHost, st.s));
		RHFQ_CUDA_CHECK(cudaEventRecord(ev[slot], st.s));
	}
	int get(int slot) {
		RHFQ_CUDA_CHECK(cudaEventSynchronize(ev[slot]));
		return h[slot];
	}
};

/* CUDA-event timing of one stage's launches on the launching stream (the
 * roofline's "average launch duration", measured live, not under a profiler). */
struct KernelTimer {
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
	std::vector<cudaEvent_t> pool;
	double ms = 0.0;
	unsigned long long count = 0;
	/* Event pairs around a launch cost 4-6 us of device idle time each (CUPTI timeline,
	 * profiles/r02_timeline_v7.txt): the ~60 node / norm launches of a batch are timed only on
	 * request (RHFQ_KERNEL_TIMERS=1, read when the batch is created); the one Simpson launch
	 * per block always is. */
	bool enabled = true;
	cudaEvent_t get() {
		if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
		cudaEvent_t e;
		RHFQ_CUDA_CHECK(cudaEventCreate(&e));
		return e;
	}
	/* call after the stream has been synchronised */
	double total_ms() {
		for (auto& pr : pending) {
			float t = 0.f;
			if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) ms += t;
			pool.push_back(pr.first);
			pool.push_back(pr.second);
		}
		pending.clear();
		return ms;
	}
	~KernelTimer() {
		for (auto& pr : pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
		for (auto e : pool) cudaEventDestroy(e);
	}
};

template<typename B>
inline void launch_timed(int64_t n, const B& b, Stream& st, unsigned long long* launches, KernelTimer& t,
                         Limit lim = Limit())
{
	if (n <= 0) return;
	if (!t.enabled) { launch(n, b, st, launches, lim); ++t.count; return; }
	cudaEvent_t e0 = t.get(), e1 = t.get();
	RHFQ_CUDA_CHECK(cudaEventRecord(e0, st.s));
	launch(n, b, st, launches, lim);
	RHFQ_CUDA_CHECK(cudaEventRecord(e1, st.s));
	t.pending.emplace_back(e0, e1);
	++t.count;
}

/* The stage bodies are __host__ __device__ (so the test-only emulation can
 * compile them); in the product they only ever run on the device. */
RHFQ_HD void atomic_add_u64(unsigned long long* p, unsigned long long v)
{
#if defined(__CUDA_ARCH__)
	/* warp-aggregated: one atomic per converged group (counters are < 2^32 per thread) */
	const unsigned mask = __activemask();
	const unsigned long long total = __reduce_add_sync(mask, (unsigned)(v & 0xffffffffu));
	const int leader = __ffs(mask) - 1;
	if ((int)(threadIdx.x & 31) == leader)
		atomicAdd(p, total);
#else
	*p += v;
#endif
}
RHFQ_HD int atomic_add_i32(int* p, int v)
{
#if defined(__CUDA_ARCH__)
	return atomicAdd(p, v);
#else
	int o = *p; *p += v; return o;
#endif
}
RHFQ_HD void atomic_max_i32(int* p, int v)
{
#if defined(__CUDA_ARCH__)
	atomicMax(p, v);
#else
	if (v > *p) *p = v;
#endif
}
RHFQ_HD int atomic_cas_i32(int* p, int expect, int v)
{
#if defined(__CUDA_ARCH__)
	return atomicCAS(p, expect, v);
#else
	const int o = *p; if (o == expect) *p = v; return o;
#endif
}
RHFQ_HD unsigned long long atomic_add_u64_ret(unsigned long long* p, unsigned long long v)
{
#if defined(__CUDA_ARCH__)
	return atomicAdd(p, v);
#else
	unsigned long long o = *p; *p += v; return o;
#endif
}
RHFQ_HD void atomic_add_u64_plain(unsigned long long* p, unsigned long long v)
{
#if defined(__CUDA_ARCH__)
	atomicAdd(p, v);
#else
	*p += v;
#endif
}
/* two consecutive slots from a (shared-memory) counter, one atomic per converged group, slots
 * in lane order */
RHFQ_HD int claim_pair_i32(int* counter)
{
#if defined(__CUDA_ARCH__)
	const unsigned mask = __activemask();
	const int lane = (int)(threadIdx.x & 31);
	const int leader = __ffs(mask) - 1;
	const int rank = __popc(mask & ((1u << lane) - 1u));
	int base = 0;
	if (lane == leader) base = atomicAdd(counter, 2 * __popc(mask));
	base = __shfl_sync(mask, base, leader);
	return base + 2 * rank;
#else
	const int o = *counter; *counter += 2; return o;
#endif
}

inline void exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, Stream& st,
                               void*& tmp, size_t& tmp_bytes)
{
	/* out has n+1 entries: out[n] = total.  Scan n+1 items (the last input is padding). */
	size_t need = 0;
	RHFQ_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, n + 1, st.s));
	if (need > tmp_bytes) {
		dfree(tmp);
		tmp = dmalloc(need);
		tmp_bytes = need;
	}
	RHFQ_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, need, in, out, n + 1, st.s));
}

inline void segmented_exclusive_sums(const int* keys, const double* in, double* fw, double* bw,
                                     int64_t n, Stream& st, void*& tmp, size_t& tmp_bytes)
{
	typedef thrust::reverse_iterator<const int*> rk_t;
	typedef thrust::reverse_iterator<const double*> rv_t;
	typedef thrust::reverse_iterator<double*> ro_t;
	size_t need1 = 0, need2 = 0;
	RHFQ_CUDA_CHECK(cub::DeviceScan::ExclusiveSumByKey(nullptr, need1, keys, in, fw, n,
	                                                   cub::Equality(), st.s));
	RHFQ_CUDA_CHECK(cub::DeviceScan::ExclusiveSumByKey(nullptr, need2, rk_t(keys + n), rv_t(in + n),
	                                                   ro_t(bw + n), n, cub::Equality(), st.s));
	size_t need = need1 > need2 ? need1 : need2;
	if (need > tmp_bytes) {
		dfree(tmp);
		tmp = dmalloc(need);
		tmp_bytes = need;
	}
	RHFQ_CUDA_CHECK(cub::DeviceScan::ExclusiveSumByKey(tmp, need, keys, in, fw, n,
	                                                   cub::Equality(), st.s));
	RHFQ_CUDA_CHECK(cub::DeviceScan::ExclusiveSumByKey(tmp, need, rk_t(keys + n), rv_t(in + n),
	                                                   ro_t(bw + n), n, cub::Equality(), st.s));
}

#endif

template<typename T>
struct DBuf {
	T* p = nullptr;
	size_t n = 0;
	DBuf() {}
	DBuf(const DBuf&) = delete;
	DBuf& operator=(const DBuf&) = delete;
	~DBuf() { dfree(p); }
	/* n is the CAPACITY: a buffer that is large enough is kept (contents undefined, as after a
	 * fresh allocation) */
	void alloc(size_t count) {
		if (p && count <= n) return;
		dfree(p); p = nullptr; n = count; p = (T*)dmalloc(count * sizeof(T));
	}
	void ensure(size_t count) { if (count > n) alloc(count + count / 4); }
	void swap(DBuf& o) { std::swap(p, o.p); std::swap(n, o.n); }
};

/* The large scratch of a batch -- Simpson node pool, fine grids, frontiers, node plans,
 * interpolant samples -- as an object that can outlive it: the resilience loop builds one batch
 * per chunk of realisations and hands the same buffers from chunk to chunk.  (Returned to the
 * memory pool and requested again, the 8-13 GB node pool of a chunk came back in a slightly
 * different size class every few chunks and the driver re-mapped pool memory: stalls of
 * 0.2-0.9 s, at random.) */
struct Workspace {
	DBuf<double> tree_fmid, tree_S;
	DBuf<int64_t> tree_child;
	DBuf<double> fine, ff[3];
	DBuf<unsigned> fj;
	DBuf<double> plan_d[7];
	DBuf<int> plan_flags, pt_post;
	DBuf<double> vbuf, pt_val, pt_fb, lpdf, bli_f, bli_c;
};

/* ------------------------------------------------------------------ *
 *  Device-side records                                                *
 * ------------------------------------------------------------------ */
struct PostState {
	double Qmax, norm, lsmax, tmin, tmax;
	double inv_lo, inv_hi;   /* 1 / (0.9 Qmax), 1 / (tmax - tmin): ranges of the two interpolants */
	int status;
	int level[2];      /* kept BLI level (samples = 8 * 2^level + 1) of low / high */
	int pad;
};

struct Counters {
	unsigned long long nodes;      /* regular (set, z) node evaluations                */
	unsigned long long lz_nodes;   /* large-z node evaluations                          */
	unsigned long long g_evals;    /* evaluations of the a log-integrand (K in F_node)  */
	unsigned long long newton;     /* Newton iterations (I in F_node)                   */
	unsigned long long nsum;       /* sum of N over regular node evaluations            */
	unsigned long long cheb_terms; /* Chebyshev terms evaluated by the Simpson-tree kernel */
	unsigned long long saq_steps;  /* frontier intervals processed by the Simpson-tree kernel */
	unsigned long long saq_other;  /* Simpson-tree evaluations that did not use the fine theta grid */
	unsigned long long g_evals_quad; /* part of g_evals spent in the Gauss-Legendre (phase B) kernels */
	unsigned long long fine_bad;   /* interpolants whose fine grid failed the probe check        */
	unsigned long long z_unconverged; /* sets whose z-quadrature ran out of levels (value kept)   */
	unsigned long long quad_checks;   /* nodes re-evaluated by the sampled precision check         */
};

struct Algo { enum { EXPLICIT = 0, BLI = 1, SIMPSON = 2 }; };

/* Chebyshev-Lobatto node table on the finest grid (barylagrint.hpp:225-238) */
struct ChebTable {
	const double* val;
	const double* ff;
	const double* fb;
	/* PointInInterval's threshold sqrt(eps) of the reference's `real` (intervall.hpp:58):
	 * 1.49e-8 for its double build, 3.29e-10 for its long double build */
	double se = SQRT_EPS;
};

RHFQ_HD bool pii_in_front(double ff, double fb, double se) { return ff < se * fb; }
RHFQ_HD bool pii_in_back(double ff, double fb, double se) { return fb < se * ff; }
/* intervall.hpp:106-117 */
RHFQ_HD double pii_minus(double v, double ff, double fb, double ov, double off, double ofb, double se)
{
	if (pii_in_front(ff, fb, se) && pii_in_front(off, ofb, se))
		return ff - off;
	if (pii_in_back(ff, fb, se) && pii_in_back(off, ofb, se))
		return ofb - fb;
	return v - ov;
}

/*
 * Second barycentric formula on Chebyshev-Lobatto samples (barylagrint.hpp:674-720).
 * Samples i = 0..n-1 live at fine-grid indices i*stride of f and of the table.
 */
RHFQ_HD double bli_interpolate(double zv, double zff, double zfb, const double* f, int stride,
                               int n, const ChebTable& T)
{
	if (zv == T.val[0])
		return f[0];
	double nom = 0.0, den = 0.0;
	double wi = 0.5 / pii_minus(zv, zff, zfb, T.val[0], T.ff[0], T.fb[0], T.se);
	nom += wi * f[0];
	den += wi;
	double sign = -1.0;
	for (int i = 1; i < n - 1; ++i) {
		const int fi = i * stride;
		const double zi = T.val[fi];
		if (zv == zi)
			return f[fi];
		wi = sign / pii_minus(zv, zff, zfb, zi, T.ff[fi], T.fb[fi], T.se);
		nom += wi * f[fi];
		den += wi;
		sign = -sign;
	}
	const int fl = (n - 1) * stride;
	{
		/* PointInInterval::operator== (intervall.hpp:84-97) */
		const bool infr = pii_in_front(zff, zfb, T.se), inba = pii_in_back(zff, zfb, T.se);
		const bool oinfr = pii_in_front(T.ff[fl], T.fb[fl], T.se), oinba = pii_in_back(T.ff[fl], T.fb[fl], T.se);
		if (infr == oinfr && inba == oinba) {
			const bool eq = infr ? (zff == T.ff[fl]) : (inba ? (zfb == T.fb[fl]) : (zv == T.val[fl]));
			if (eq)
				return f[fl];
		}
	}
	wi = sign * 0.5 / pii_minus(zv, zff, zfb, T.val[fl], T.ff[fl], T.fb[fl], T.se);
	nom += wi * f[fl];
	den += wi;
	return nom / den;
}

/*
 * The same formula for a query that is not within sqrt(eps) of either end of
 * [-1, 1] (true for every Chebyshev node that is not an end node): then
 * PointInInterval::operator- always takes its `val - other.val` branch
 * (intervall.hpp:116) and the end-coordinate bookkeeping drops out.
 */
RHFQ_HD double bli_interpolate_interior(double zv, const double* f, int stride, int n,
                                        const double* tval)
{
	double nom = 0.0, den = 0.0;
	double wi = 0.5 / (zv - tval[0]);
	nom += wi * f[0];
	den += wi;
	double sign = -1.0;
	for (int i = 1; i < n - 1; ++i) {
		const int fi = i * stride;
		const double zi = tval[fi];
		if (zv == zi)
			return f[fi];
		wi = sign / (zv - zi);
		nom = fma(wi, f[fi], nom);
		den += wi;
		sign = -sign;
	}
	const int fl = (n - 1) * stride;
	wi = sign * 0.5 / (zv - tval[fl]);
	nom = fma(wi, f[fl], nom);
	den += wi;
	return nom / den;
}

/* ------------------------------------------------------------------ *
 *  Stage bodies                                                       *
 * ------------------------------------------------------------------ */
/* Lanes that share one reduction (a warp on the device, one "lane" in the host emulation) */
#if defined(__CUDACC__) && !defined(RHFQ_EMU)
constexpr int REDUCE_LANES = 32;
#else
constexpr int REDUCE_LANES = 1;
#endif
RHFQ_HD double lane_sum(double x)
{
#if defined(__CUDA_ARCH__)
	for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
#endif
	return x;
}

/* task = (set, lane): REDUCE_LANES consecutive threads (one warp on the device) share a set */
struct LocalsBody {
	const double* q; const double* c; const int64_t* set_off; const int* set_post;
	const double* prior; int prior_stride; double amin;
	double* k; SetLocals* L; const double* w;
	double* lgq;      /* scratch, packed like q */
	int sequential;   /* RHFQ_LOCALS_SEQ=1: lane 0 alone runs compute_locals (for the comparison) */
	double* gco;      /* [S][GCO_N] Stirling coefficient blocks of the sets (gcoef_fill), or null */
	double log_eps, sqrt_eps;   /* of the reference's arithmetic type (Batch::eps_real) */
	RHFQ_HD void operator()(int64_t task) const {
		const int64_t s = task / REDUCE_LANES;
		const int lane = (int)(task % REDUCE_LANES);
		SetLocals l;
		const int64_t o = set_off[s];
		const int N = (int)(set_off[s + 1] - o);
		const double* pr = prior + (int64_t)set_post[s] * prior_stride;
		int st;
		if (!(w[s] > 0.0))
			st = ST_DOMAIN;     /* NaN, zero, or negative weight (posterior.hpp:539-541) */
		else if (N <= 0)
			st = ST_DOMAIN;
		else {
#if defined(__CUDA_ARCH__)
			if (sequential) {
				if (lane != 0) return;
				st = compute_locals(q + o, c + o, N, pr[0], pr[1], pr[2], pr[3], amin, k + o, l);
			} else
				st = compute_locals_warp(q + o, c + o, N, pr[0], pr[1], pr[2], pr[3], amin, k + o,
				                         lgq + o, l, lane);
#else
			st = compute_locals(q + o, c + o, N, pr[0], pr[1], pr[2], pr[3], amin, k + o, l);
#endif
		}
		if (lane != 0) return;
		l.k_off = o;
		l.N = N;
		l.status = st;
		l.log_scale = 0; l.ymax = 0; l.ztrans = 0; l.S = 0; l.taylor = 0; l.norm = 0;
		l.weight = 0; l.log_fac = 0; l.flags = 0; l.gtab = nullptr;
		l.log_eps = log_eps; l.sqrt_eps = sqrt_eps;
		l.gco = nullptr;
		if (st == ST_OK && gco) { gcoef_fill(gco + s * GCO_N, l.n, l.v); l.gco = gco + s * GCO_N; }
		L[s] = l;
	}
};

/* --- lgamma-difference tables (rhfq_math.cuh: GTAB) --- */
/* The distinct (n, v) of the batch are found ON THE DEVICE: an open-addressing table of
 * GTAB_SLOTS slots keyed by the bit patterns of (n, v); the first set to claim a slot owns it
 * and its table is built at the slot's index.  (The keys used to travel to the host, 2 S
 * doubles, be hashed there and travel back: one of the ~25 stream drains per batch.)  Sets that
 * find no slot within GTAB_PROBES steps -- more distinct (n, v) than a batch with shared or
 * few priors has -- keep a null table and use the direct formula. */
constexpr int GTAB_SLOTS = 1024, GTAB_PROBES = 64;
RHFQ_HD unsigned gtab_hash(double n, double v)
{
	unsigned long long a, b;
#if defined(__CUDA_ARCH__)
	a = (unsigned long long)__double_as_longlong(n); b = (unsigned long long)__double_as_longlong(v);
#else
	std::memcpy(&a, &n, 8); std::memcpy(&b, &v, 8);
#endif
	unsigned long long h = a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15ull + (a << 6) + (a >> 2));
	h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
	return (unsigned)h;
}
struct GtabClaimBody {
	SetLocals* L; int* owner; double* tab; int* n_classes;
	RHFQ_HD void operator()(int64_t s) const {
		SetLocals& l = L[s];
		l.gtab = nullptr;
		if (l.status != ST_OK) return;
		const double n = l.n, v = l.v;
		unsigned h = gtab_hash(n, v) & (GTAB_SLOTS - 1);
		for (int probe = 0; probe < GTAB_PROBES; ++probe, h = (h + 1) & (GTAB_SLOTS - 1)) {
			int o = atomic_cas_i32(&owner[h], -1, (int)s);
			if (o == -1) { o = (int)s; atomic_add_i32(n_classes, 1); }
			/* n and v of the owner were written by LocalsBody, a launch ago */
			if (L[o].n == n && L[o].v == v) {
				l.gtab = tab + (int64_t)h * GTAB_STRIDE;
				return;
			}
		}
	}
};

struct GtabBuildBody {
	const SetLocals* L; const int* owner; double* tab;
	RHFQ_HD void operator()(int64_t t) const {
		const int64_t c = t / GTAB_SIZE;
		const int j = (int)(t % GTAB_SIZE);
		const int o = owner[c];
		if (o < 0) return;
		gtab_fill_record(L[o].n, L[o].v, j, tab + c * GTAB_STRIDE + GTAB_REC * j);
	}
};

/*
 * Per-set setup (log_scale, y_max) in four launches.  One thread per set ran ~32 Brent
 * searches back to back (10^4 threads, 2.5 ms of pure latency); the searches of the
 * transition y_max are independent of each other wherever the algorithm only asks "where
 * does the sign change", so they are dealt out to threads:
 *   scale   one thread per set: log_integrand_max                       (amax.hpp:68-133)
 *   yscan   20 threads per set: both orders (m = 0, 3) of the root function at
 *           y = 1e-20 * 2^(8c), c = 0..8, and at the lower bracket end 1e-32
 *   ybis    14 threads per set: the 7 exponents inside the stride that holds the sign change
 *   ymax    2 threads per set (one per order, values exchanged by shuffle): the reference's
 *           y_r from the stored values, then TOMS748                    (large_z.hpp:320-349)
 * Same values, same decisions as the sequential y_taylor_transition (rhfq_math.cuh).
 */
struct ScaleBody {
	SetLocals* L; const double* k;
	RHFQ_HD void operator()(int64_t s) const {
		SetLocals l = L[s];
		if (l.status != ST_OK)
			return;
		double ls = 0;
		const int st = log_integrand_max(l, k + l.k_off, ls);
		L[s].log_scale = (st == ST_OK) ? ls : 0.0;
		L[s].status = st;
	}
};

constexpr int YSCAN_C = 10;       /* candidates c = 0..8: y = 1e-20 * 2^(8c); c = 9: y = 1e-32 */
/* inner exponents of the stride [klo, khi] that holds the first non-negative coarse value, AND
 * of the stride before it: the root function can rise above zero, dip below again and rise for
 * good (1 of 20 000 random posteriors with a double epsilon, 1 of 3000 with the long double
 * one) -- the reference's doubling stops at the FIRST non-negative exponent, which then lies
 * just before a negative coarse value.  slots 0..6: klo - 7 .. klo - 1, slots 7..13: klo + 1 ..
 * klo + 7. */
constexpr int YBIS_C = 14;
RHFQ_HD int ybis_exponent(int klo, int i) { return i < 7 ? klo - 7 + i : klo + 1 + (i - 7); }
RHFQ_HD double yscan_y(int c) { return c < 9 ? ldexp(1e-20, 8 * c) : 1e-32; }
RHFQ_HD bool yroot_negative(double v) { return v < 0.0 || is_nan(v); }
/* root function from the stored (log term, log maximiser) pairs of m = 0 and m = 3 */
RHFQ_HD double yroot_stored(const double* T, double log_eps) { return y_transition_combine(T[0], T[1], T[2], T[3], log_eps); }

struct YScanBody {
	const SetLocals* L; double* T;    /* T[s][c][m][2] */
	RHFQ_HD void operator()(int64_t t) const {
		const int64_t s = t / (2 * YSCAN_C);
		const int j = (int)(t % (2 * YSCAN_C));
		if (L[s].status != ST_OK) return;
		double v, la;
		y_transition_term(L[s], yscan_y(j >> 1), (j & 1) ? 3 : 0, v, la);
		T[2 * t] = v; T[2 * t + 1] = la;
	}
};

/* the stride [klo, khi] of the exponent scan that holds the first sign change; false if the
 * root function is already non-negative at 1e-20 (khi = 0) or nowhere on the grid (khi = -1) */
RHFQ_HD bool yscan_stride(const double* Ts, double log_eps, int& klo, int& khi, double& vhi)
{
	vhi = yroot_stored(Ts, log_eps);
	klo = 0; khi = 0;
	if (!yroot_negative(vhi)) return false;
	khi = -1;
	for (int c = 1; c <= 8; ++c) {
		const double v = yroot_stored(Ts + 4 * c, log_eps);
		if (!yroot_negative(v)) { khi = 8 * c; vhi = v; return true; }
		klo = 8 * c;
	}
	return false;
}

struct YBisBody {
	const SetLocals* L; const double* T; double* Tb;    /* Tb[s][i][m][2], exponent klo + 1 + i */
	RHFQ_HD void operator()(int64_t t) const {
		const int64_t s = t / (2 * YBIS_C);
		const int j = (int)(t % (2 * YBIS_C));
		if (L[s].status != ST_OK) return;
		int klo, khi; double vhi;
		if (!yscan_stride(T + s * 4 * YSCAN_C, L[s].log_eps, klo, khi, vhi)) return;
		const int k = ybis_exponent(klo, j >> 1);
		if (k <= 0) return;               /* no stride before the first one */
		double v, la;
		y_transition_term(L[s], ldexp(1e-20, k), (j & 1) ? 3 : 0, v, la);
		Tb[2 * t] = v; Tb[2 * t + 1] = la;
	}
};

/* The root function for TOMS748, evaluated by the two threads of a set: each computes its
 * order's term, the pair exchanges the values and both continue with the same number. */
struct YRootPair {
	const SetLocals* l; int m3;
	RHFQ_HD double operator()(double y) const {
#if defined(__CUDA_ARCH__)
		double v, la;
		y_transition_term(*l, y, m3 ? 3 : 0, v, la);
		const unsigned mask = 3u << ((threadIdx.x & 31u) & ~1u);
		const double vo = __shfl_xor_sync(mask, v, 1), lao = __shfl_xor_sync(mask, la, 1);
		return m3 ? y_transition_combine(vo, lao, v, la, l->log_eps) : y_transition_combine(v, la, vo, lao, l->log_eps);
#else
		return y_transition_root_fn(*l, y);
#endif
	}
};

struct YMaxBody {
	SetLocals* L; const double* T; const double* Tb;
	RHFQ_HD void operator()(int64_t t) const {
		const int64_t s = t >> 1;
		const int m3 = (int)(t & 1);
#if !defined(__CUDA_ARCH__)
		if (m3) return;           /* host emulation: one call per set does both orders */
#endif
		if (L[s].status != ST_OK) return;
		const SetLocals l = L[s];
		const YRootPair f{&l, m3};
		const double* Ts = T + s * 4 * YSCAN_C;
		int klo, khi; double val;
		double yr = 1e-20;
		if (yscan_stride(Ts, l.log_eps, klo, khi, val)) {
			/* the reference doubles y until the root function is non-negative: the FIRST such
			 * exponent of the stride (all seven inner values are stored; a bisection over them
			 * would pick another sign change where the function is not monotone -- seen with
			 * the long-double threshold, where it dips below zero again after its first zero) */
			const double* Tbs = Tb + s * 4 * YBIS_C;
			for (int i = 0; i < YBIS_C; ++i) {
				const int k = ybis_exponent(klo, i);
				if (k <= 0) continue;
				const double vm = yroot_stored(Tbs + 4 * i, l.log_eps);
				if (!yroot_negative(vm)) { khi = k; val = vm; break; }
			}
			yr = ldexp(1e-20, khi);
		} else if (khi < 0) {
			/* no sign change on the coarse grid: one doubling at a time like the reference */
			yr = ldexp(1e-20, klo);
			while (yroot_negative(val)) {
				yr = fmin(2.0 * yr, 1.0);
				if (yr == 1.0)
					break;
				val = f(yr);
			}
		}
		double a = 1e-32, b = yr;
		const double fa = yroot_stored(Ts + 4 * 9, l.log_eps);
		const double fb = (yr == 1.0) ? f(b) : val;
		int st = ST_OK;
		double ym = 0;
		if (is_nan(fa) || is_nan(fb) || ((fa < 0.0) == (fb < 0.0) && fa != 0.0 && fb != 0.0))
			st = ST_ROOT;
		else {
			const int used = toms748(f, a, b, fa, fb, 0.5 /* eps_tolerance(2) */, 98);
			if (used >= 98) st = ST_ROOT;
			else ym = 0.5 * (a + b);
		}
		if (m3 == 0) {
			L[s].ymax = ym;
			L[s].ztrans = 1.0 - ym;
			L[s].status = st;
		}
	}
};

RHFQ_HD void count_node(Counters* cnt, const QuadInfo& qi, int N, bool large_z)
{
	if (!cnt) return;
	if (large_z) atomic_add_u64(&cnt->lz_nodes, 1ull);
	else atomic_add_u64(&cnt->nodes, 1ull);
	atomic_add_u64(&cnt->g_evals, (unsigned long long)qi.evals);
	atomic_add_u64(&cnt->newton, (unsigned long long)qi.newton);
	if (!large_z) atomic_add_u64(&cnt->nsum, (unsigned long long)N);
}

/*
 * Node evaluation in two launches.  Phase A finds the mode and the window of the
 * a-integral (Newton, closed-form guess, verification: branchy, different for every
 * node); phase B is the Gauss-Legendre loop (the same ~600 instructions ~45 times per
 * node).  In one kernel the warps of an SM sit in different phases and evict each other's
 * instructions (ncu: stalled_no_instruction 3.95 per issue, 97 KB of SASS); split, phase B
 * runs out of the instruction cache.  The plan (8 values per node) goes through global
 * memory.
 */
struct NodePlanView {
	double* C; double* off; double* y;
	double* a_ref; double* a_lo; double* a_hi; double* phi_ref;
	int* flags;      /* PLAN_DONE: the output already holds the final value */
};
constexpr int PLAN_LZ = 1, PLAN_INTERIOR = 2, PLAN_DONE = 4, PLAN_WINDOW_BAD = 8;
/* one node task in QUAD_CHECK_STRIDE is re-evaluated with every Gauss-Legendre panel cut in two
 * and compared (the fixed-node rule has no error estimate of its own; the reference's adaptive
 * rule throws PrecisionError when its estimate exceeds 1e-5 of the integral,
 * integrand.hpp:459-466). */
constexpr int QUAD_CHECK_STRIDE = 127;
constexpr double QUAD_CHECK_TOL = 1e-5;

/* phase B of one node: log of the a-integral from the stored plan */
RHFQ_HD double plan_quadrature(const NodePlanView& pv, int64_t t, const SetLocals& l, int* evals)
{
	PhiCtx c;
	const int fl = pv.flags[t];
	if (fl & PLAN_LZ) phi_ctx_large_z(l, pv.y[t], false, c);
	else phi_ctx_regular(l, pv.C[t], pv.off[t], c);
	QuadPlan q;
	q.a_ref = pv.a_ref[t]; q.a_lo = pv.a_lo[t]; q.a_hi = pv.a_hi[t]; q.phi_ref = pv.phi_ref[t];
	q.interior = (fl & PLAN_INTERIOR) ? 1 : 0;
	q.evals = 0;
	q.bad = 0;
	return quad_panels_inl(c, q, l.n, evals);
}

/* ctx of a planned node from its stored constants */
RHFQ_HD void plan_ctx(const NodePlanView& pv, int64_t t, const SetLocals& l, PhiCtx& c)
{
	if (pv.flags[t] & PLAN_LZ) phi_ctx_large_z(l, pv.y[t], false, c);
	else phi_ctx_regular(l, pv.C[t], pv.off[t], c);
}

/* How a task index maps to its sample set, error slot and output: node evaluations of the
 * pdf (task = (set m of the point's posterior, point p)) or nodes of the norm's z-quadrature
 * (task = (active set a, node j)). */
struct NodeTaskMap {
	const int64_t* post_off; const int* pt_post; int64_t n_pts; double* vbuf; int* post_err;
	RHFQ_HD int64_t set(int64_t t) const { return post_off[pt_post[t % n_pts]] + t / n_pts; }
	RHFQ_HD void fail(int64_t t, int st) const {
		atomic_max_i32(&post_err[pt_post[t % n_pts]], st);
		vbuf[t] = nan("");
	}
	RHFQ_HD void finish(int64_t t, const SetLocals& l, double res) const {
		vbuf[t] = res - l.log_scale + l.log_fac;
	}
};
struct NormTaskMap {
	const int* active; int n_nodes; double* fbuf; int* err_status;   /* fbuf[a][n_nodes]: this launch's nodes */
	RHFQ_HD int64_t set(int64_t t) const { return active[t / n_nodes]; }
	RHFQ_HD double& out(int64_t t) const { return fbuf[t]; }
	RHFQ_HD void fail(int64_t t, int st) const {
		atomic_max_i32(&err_status[active[t / n_nodes]], st);
		out(t) = 0.0;
	}
	RHFQ_HD void finish(int64_t t, const SetLocals& l, double res) const {
		out(t) = exp(res - l.log_scale);
	}
};

/* phase A2: mode of the a-integrand (Newton); amode -> a_ref, curvature -> a_lo (scratch) */
template<typename Map>
struct PlanModeBody {
	const SetLocals* L; Map map; Counters* cnt; NodePlanView pv;
	RHFQ_HD void operator()(int64_t t) const {
		if (pv.flags[t] & PLAN_DONE) return;
		const SetLocals& l = L[map.set(t)];
		int st = ST_OK;
		if ((pv.flags[t] & PLAN_LZ) && l.n < l.v) st = ST_NOT_NORMALIZABLE;
		double amode = 0.0, curv = 0.0;
		int newton = 0;
		if (st == ST_OK) {
			PhiCtx c;
			plan_ctx(pv, t, l, c);
			st = mode_newton_inl(l, c, amode, curv, &newton);
		}
		if (cnt) atomic_add_u64(&cnt->newton, (unsigned long long)newton);
		if (st != ST_OK) {
			pv.flags[t] |= PLAN_DONE;
			map.fail(t, st);
			return;
		}
		pv.a_ref[t] = amode;
		pv.a_lo[t] = curv;
	}
};

/* phase A3: window of the a-integral (closed-form guess, verification, Newton fallback) */
template<typename Map>
struct PlanWindowBody {
	const SetLocals* L; Map map; Counters* cnt; NodePlanView pv;
	RHFQ_HD void operator()(int64_t t) const {
		if (pv.flags[t] & PLAN_DONE) return;
		const SetLocals& l = L[map.set(t)];
		PhiCtx c;
		plan_ctx(pv, t, l, c);
		const double amode = pv.a_ref[t], curv = pv.a_lo[t];
		QuadPlan q;
		quad_plan_inl(c, l.amin, amode, (amode > l.amin) ? curv : 0.0, l.n, l.n - l.v, q);
		if (cnt) atomic_add_u64(&cnt->g_evals, (unsigned long long)q.evals);
		pv.a_ref[t] = q.a_ref; pv.a_lo[t] = q.a_lo; pv.a_hi[t] = q.a_hi; pv.phi_ref[t] = q.phi_ref;
		if (q.interior) pv.flags[t] |= PLAN_INTERIOR;
		if (q.bad) pv.flags[t] |= PLAN_WINDOW_BAD;
	}
};

/* phase B: the Gauss-Legendre panels (cut at the |C1| kink of a large-z node where it matters) */
template<typename Map>
struct PlanQuadBody {
	const SetLocals* L; Map map; Counters* cnt; NodePlanView pv;
	RHFQ_HD void operator()(int64_t t) const {
		const int fl = pv.flags[t];
		if (fl & PLAN_DONE) return;
		const SetLocals& l = L[map.set(t)];
		if (fl & PLAN_WINDOW_BAD) { map.fail(t, ST_PRECISION); return; }
		int evals = 0;
		const double res = plan_quadrature(pv, t, l, &evals);
		if (cnt) {
			atomic_add_u64(&cnt->g_evals, (unsigned long long)evals);
			atomic_add_u64(&cnt->g_evals_quad, (unsigned long long)evals);
		}
		if (is_nan(res)) { map.fail(t, ST_RUNTIME); return; }
		map.finish(t, l, res);
	}
};

/* phase B with a WARP per node: the lanes share the Gauss-Legendre nodes and combine their sums
 * by shuffles.  For small batches (a single posterior is 9 ... 512 nodes per launch) a thread per
 * node leaves the device empty while every thread walks through ~90 evaluations of the
 * integrand one after the other -- the launch takes as long as for 10^4 posteriors; 32 lanes per
 * node cut that latency by ~20x.  Large batches keep one thread per node (no idle lanes). */
template<typename Map>
struct PlanQuadWarpBody {
	const SetLocals* L; Map map; Counters* cnt; NodePlanView pv;
	RHFQ_HD void operator()(int64_t task) const {
		const int64_t t = task / REDUCE_LANES;
		const int lane = (int)(task % REDUCE_LANES);
		const int fl = pv.flags[t];
		if (fl & PLAN_DONE) return;                                  /* uniform over the warp */
		const SetLocals& l = L[map.set(t)];
		if (fl & PLAN_WINDOW_BAD) { if (lane == 0) map.fail(t, ST_PRECISION); return; }
		PhiCtx c;
		plan_ctx(pv, t, l, c);
		QuadPlan q;
		q.a_ref = pv.a_ref[t]; q.a_lo = pv.a_lo[t]; q.a_hi = pv.a_hi[t]; q.phi_ref = pv.phi_ref[t];
		q.interior = (fl & PLAN_INTERIOR) ? 1 : 0;
		q.evals = 0; q.bad = 0;
		int evals = 0;
		const double S = lane_sum(quad_panels_partial(c, q, l.n, false, lane, REDUCE_LANES, &evals));
		if (cnt) {
			atomic_add_u64(&cnt->g_evals, (unsigned long long)evals);
			atomic_add_u64(&cnt->g_evals_quad, (unsigned long long)evals);
		}
		if (lane != 0) return;
		const double res = log(S) + q.phi_ref;
		if (is_nan(res)) { map.fail(t, ST_RUNTIME); return; }
		map.finish(t, l, res);
	}
};
/* node tasks per launch below which the Gauss-Legendre stage runs with a warp per node */
constexpr int64_t QUAD_WARP_BELOW = 16384;

/* Sampled precision check: a launch of its own over one node task in QUAD_CHECK_STRIDE, a warp
 * per sampled task.  (Inside the quadrature kernel the second rule cost that kernel 50 % of its
 * speed -- registers, instruction cache; one thread per sampled task made every launch wait
 * for ~270 serial evaluations of the integrand: 0.12 ms x 13 launches per batch.)  The lanes
 * share the Gauss-Legendre nodes of the plain rule and of the rule with every panel cut in two
 * and compare the sums. */
template<typename Map>
struct PlanQuadCheckBody {
	const SetLocals* L; Map map; Counters* cnt; NodePlanView pv; int64_t n_tasks; Limit live;
	RHFQ_HD void operator()(int64_t task) const {
		const int64_t t = (task / REDUCE_LANES) * QUAD_CHECK_STRIDE;
		const int lane = (int)(task % REDUCE_LANES);
		if (t >= n_tasks || !live.live(t)) return;                   /* uniform over the warp */
		const int fl = pv.flags[t];
		if (fl & (PLAN_DONE | PLAN_WINDOW_BAD)) return;
		const SetLocals& l = L[map.set(t)];
		PhiCtx c;
		plan_ctx(pv, t, l, c);
		QuadPlan q;
		q.a_ref = pv.a_ref[t]; q.a_lo = pv.a_lo[t]; q.a_hi = pv.a_hi[t]; q.phi_ref = pv.phi_ref[t];
		q.interior = (fl & PLAN_INTERIOR) ? 1 : 0;
		q.evals = 0; q.bad = 0;
		int e1 = 0, e2 = 0;
		const double S1 = lane_sum(quad_panels_partial(c, q, l.n, false, lane, REDUCE_LANES, &e1));
		const double S2 = lane_sum(quad_panels_partial(c, q, l.n, true, lane, REDUCE_LANES, &e2));
		if (cnt) atomic_add_u64(&cnt->g_evals, (unsigned long long)(e1 + e2));
		if (lane != 0) return;
		if (cnt) atomic_add_u64(&cnt->quad_checks, 1ull);
		/* relative difference of the two integrals */
		if (!is_nan(S1) && !(fabs(S2 - S1) <= QUAD_CHECK_TOL * fabs(S1))) map.fail(t, ST_PRECISION);
	}
};
inline int64_t quad_check_tasks(int64_t n) { return (n + QUAD_CHECK_STRIDE - 1) / QUAD_CHECK_STRIDE * REDUCE_LANES; }

/* tanh-sinh nodes in z (localsandnorm.hpp:221-226): node j has complement xc_j = 1 - |x_j|,
 * weight w_j and side (+1 right, -1 left, 0 centre).  z = ztrans * (1 + x) / 2. */
struct NormEvalBody {
	const SetLocals* L; const double* k; const int* active; int n_nodes; int node0;
	const double* xc; const double* wt; const signed char* side;
	double* fbuf; Counters* cnt; int* err_status;
	RHFQ_HD void operator()(int64_t t) const {
		const int a = (int)(t / n_nodes);
		const int j = node0 + (int)(t % n_nodes);
		const int s = active[a];
		const SetLocals& l = L[s];
		const double hw = 0.5 * l.ztrans;
		double z;
		if (side[j] == 0) z = hw;
		else if (side[j] > 0) z = l.ztrans - hw * xc[j];
		else z = hw * xc[j];
		double val = 0.0;
		const bool ev = (side[j] == 0) || (side[j] > 0 ? z < l.ztrans : z > 0.0);
		if (ev) {
			QuadInfo qi;
			const double* ks = k + l.k_off;
			const double lsum = sum_log1p(ks, l.N, z);
			const double l1p_wz = log1p(-l.w * z);
			double lo;
			const int st = log_outer_unscaled(l, lsum, l1p_wz, lo, &qi);
			count_node(cnt, qi, l.N, false);
			if (st != ST_OK) {
				atomic_max_i32(&err_status[s], st);
				val = 0.0;
			} else {
				val = exp(lo - l.log_scale);
			}
		}
		fbuf[t] = val;
	}
};

/* phase A1 for the nodes of the z-quadrature of the norm */
struct NormPrepBody {
	const SetLocals* L; const double* k; const int* active; int n_nodes; int node0;
	const double* xc; const signed char* side;
	double* fbuf; Counters* cnt; NodePlanView pv;
	RHFQ_HD void operator()(int64_t t) const {
		const int a = (int)(t / n_nodes);
		const int j = node0 + (int)(t % n_nodes);
		const SetLocals& l = L[active[a]];
		const double hw = 0.5 * l.ztrans;
		double z;
		if (side[j] == 0) z = hw;
		else if (side[j] > 0) z = l.ztrans - hw * xc[j];
		else z = hw * xc[j];
		const bool ev = (side[j] == 0) || (side[j] > 0 ? z < l.ztrans : z > 0.0);
		if (!ev) {
			pv.flags[t] = PLAN_DONE;
			fbuf[t] = 0.0;
			return;
		}
		const double* ks = k + l.k_off;
		const double lsum = sum_log1p(ks, l.N, z);
		const double l1p_wz = log1p(-l.w * z);
		pv.C[t] = l.lp - l.v * (l.ls + l1p_wz) + lsum;
		pv.off[t] = l.lp + lsum;
		pv.y[t] = 0.0;
		pv.flags[t] = 0;
		if (cnt) {
			atomic_add_u64(&cnt->nodes, 1ull);
			atomic_add_u64(&cnt->nsum, (unsigned long long)l.N);
		}
	}
};

/* One warp per active set: the lanes stride over the level's nodes (coalesced reads of the
 * set's row of fbuf), the partial sums are combined by shuffles and every lane takes the same
 * convergence decision; one thread per set walked over up to 456 strided values. */
struct NormReduceBody {
	SetLocals* L; const int* active; const double* fbuf; int fstride;   /* row = nodes of levels level_a..level_b */
	const double* wt; const int* level_off; int level_a; int level_b;
	double* acc;      /* per set: sum, l1, I_prev */
	int* converged;   /* per active slot */
	const int* err_status;
	RHFQ_HD void operator()(int64_t task) const {
		const int64_t a = task / REDUCE_LANES;
		const int lane = (int)(task % REDUCE_LANES);
		const int s = active[a];
		SetLocals& l = L[s];
		double sum = acc[3 * s], l1 = acc[3 * s + 1], Iprev = acc[3 * s + 2];
		const double hw = 0.5 * l.ztrans;
		int conv = 0;
		if (err_status[s] != ST_OK) {
			if (lane == 0) {
				l.status = err_status[s];
				converged[a] = 1;
			}
			return;
		}
		double S = 0.0;
		for (int lev = level_a; lev <= level_b; ++lev) {
			double ps = 0.0, pl = 0.0;
			for (int j = level_off[lev] + lane; j < level_off[lev + 1]; j += REDUCE_LANES) {
				const double f = fbuf[(int64_t)a * fstride + (j - level_off[level_a])];
				ps += wt[j] * f;
				pl += wt[j] * fabs(f);
			}
			sum += lane_sum(ps);
			l1 += lane_sum(pl);
			const double h = ldexp(1.0, -lev);
			const double I = sum * h;
			const double err = fabs(I - Iprev);
			Iprev = I;
			S = I * hw;
			if (lev >= 3 && err <= l.sqrt_eps * (l1 * h)) {
				conv = 1;
				break;
			}
		}
		if (lane == 0) {
			l.S = S;
			acc[3 * s] = sum; acc[3 * s + 1] = l1; acc[3 * s + 2] = Iprev;
			converged[a] = conv;
		}
	}
};

struct TaylorBody {
	SetLocals* L; Counters* cnt;
	RHFQ_HD void operator()(int64_t s) const {
		SetLocals& l = L[s];
		if (l.status != ST_OK)
			return;
		QuadInfo qi;
		double lt;
		int st = log_large_z_unscaled(l, l.ymax, true, lt, &qi);
		count_node(cnt, qi, l.N, true);
		if (st != ST_OK) { l.status = st; return; }
		l.taylor = exp(lt - l.log_scale);
		l.norm = l.S + l.taylor;
		if (is_nan(l.S) || is_nan(l.taylor) || !is_fin(l.norm) || l.norm == 0.0)
			l.status = ST_RUNTIME;
	}
};

struct PostBody {
	SetLocals* L; const int64_t* post_off; const double* w; PostState* P;
	RHFQ_HD void operator()(int64_t r) const {
		PostState ps;
		ps.status = ST_OK;
		ps.level[0] = ps.level[1] = -1;
		ps.pad = 0;
		const int64_t s0 = post_off[r], s1 = post_off[r + 1];
		double lsmax = -RHFQ_INF, Qmax = 0.0;
		if (s1 <= s0) ps.status = ST_DOMAIN;   /* "No samples given." */
		for (int64_t s = s0; s < s1; ++s) {
			if (L[s].status != ST_OK) { if (ps.status == ST_OK) ps.status = L[s].status; continue; }
			if (L[s].log_scale > lsmax) lsmax = L[s].log_scale;
			if (L[s].Qmax > Qmax) Qmax = L[s].Qmax;
		}
		double norm = 0.0, corr = 0.0;
		if (ps.status == ST_OK) {
			for (int64_t s = s0; s < s1; ++s) {
				const double wi = w[s] * exp(L[s].log_scale - lsmax);
				L[s].weight = wi;
				kahan_add(norm, corr, L[s].norm * wi);
			}
			for (int64_t s = s0; s < s1; ++s)
				L[s].log_fac = log(L[s].weight / (L[s].Qmax * norm));
		}
		ps.Qmax = Qmax;
		ps.norm = norm;
		ps.lsmax = lsmax;
		ps.tmin = log(DBL_EPS * Qmax);
		ps.tmax = log((1.0 - 0.9) * Qmax);
		ps.inv_lo = 1.0 / (0.9 * Qmax - 0.0);
		ps.inv_hi = 1.0 / (ps.tmax - ps.tmin);
		P[r] = ps;
	}
};

/* sum_i log1p(-k_i z) of a pdf node.  With the long-double thresholds the Taylor branch only
 * starts at y = 1 - z ~ 1e-7 (instead of ~1e-6) and regular nodes come that close to z = 1, where
 * the factor of the datum that defines Qmax, 1 - k_imax z = 1 - z (k_imax = 1), loses
 * (a - 1) * 1e-16 / y of relative accuracy in double: it is taken from the node's distance to
 * Qmax instead, which the interpolant's nodes carry exactly (x_fb = exp(t)).  The reference's
 * long double build has 1e-19 / y there. */
RHFQ_HD double node_sum_log1p(const SetLocals& l, const double* ks, double zi, double x_fb, double Qp)
{
	const double lsum = sum_log1p(ks, l.N, zi);
	if (l.log_eps < -40.0 && l.Qmax == Qp) {
		const double yi = x_fb / Qp;
		if (yi > 0.0 && yi < 1e-3)
			return lsum - log(one_minus_kz(ks[l.imax], zi)) + log(yi);
	}
	return lsum;
}

/*
 * log of one set's contribution to the pdf at the point (x_val, x_fb)
 * (posterior.hpp:838-881 with log_pdf_t, :785-812).
 */
RHFQ_HD double log_pdf_term(const SetLocals& l, const double* k, double x_val, double x_fb,
                            double Qmax_post, int& status, Counters* cnt)
{
	const double zi = x_val / l.Qmax;
	if (is_nan(zi)) { status = ST_RUNTIME; return nan(""); }
	if (zi > 1.0 || zi < 0.0)
		return -RHFQ_INF;
	double res;
	QuadInfo qi;
	if (zi <= l.ztrans) {
		const double* ks = k + l.k_off;
		const double lsum = node_sum_log1p(l, ks, zi, x_fb, Qmax_post);
		const double l1p_wz = log1p(-l.w * zi);
		const int st = log_outer_unscaled(l, lsum, l1p_wz, res, &qi);
		count_node(cnt, qi, l.N, false);
		if (st != ST_OK) { status = st; return nan(""); }
	} else {
		const double yi = (l.Qmax == Qmax_post) ? x_fb / Qmax_post : 1.0 - zi;
		const int st = log_large_z_unscaled(l, yi, false, res, &qi);
		count_node(cnt, qi, l.N, true);
		if (st != ST_OK) { status = st; return nan(""); }
	}
	return res - l.log_scale + l.log_fac;
}

/* phase A1 of a pdf node: range checks, the data sum and the constants of the integrand
 * (the front half of log_pdf_term / log_outer_unscaled / log_large_z_unscaled) */
struct NodePrepBody {
	const SetLocals* L; const double* k; const int64_t* post_off; const PostState* P;
	const int* pt_post; const double* pt_val; const double* pt_fb; int64_t n_pts; int Mmax;
	double* vbuf; int* post_err; Counters* cnt; NodePlanView pv;
	RHFQ_HD void operator()(int64_t t) const {
		const int64_t m = t / n_pts;
		const int64_t p = t % n_pts;
		const int r = pt_post[p];
		const int64_t s0 = post_off[r];
		const int64_t M = post_off[r + 1] - s0;
		pv.flags[t] = PLAN_DONE;
		if (m >= M) return;
		if (P[r].status != ST_OK) { vbuf[t] = nan(""); return; }
		const SetLocals& l = L[s0 + m];
		const double zi = pt_val[p] / l.Qmax;
		if (is_nan(zi)) { atomic_max_i32(&post_err[r], ST_RUNTIME); vbuf[t] = nan(""); return; }
		if (zi > 1.0 || zi < 0.0) { vbuf[t] = -RHFQ_INF; return; }
		if (zi <= l.ztrans) {
			const double* ks = k + l.k_off;
			const double lsum = node_sum_log1p(l, ks, zi, pt_fb[p], P[r].Qmax);
			const double l1p_wz = log1p(-l.w * zi);
			pv.C[t] = l.lp - l.v * (l.ls + l1p_wz) + lsum;
			pv.off[t] = l.lp + lsum;
			pv.y[t] = 0.0;
			pv.flags[t] = 0;
			if (cnt) {
				atomic_add_u64(&cnt->nodes, 1ull);
				atomic_add_u64(&cnt->nsum, (unsigned long long)l.N);
			}
		} else {
			const double Qp = P[r].Qmax;
			const double yi = (l.Qmax == Qp) ? pt_fb[p] / Qp : 1.0 - zi;
			if (cnt) atomic_add_u64(&cnt->lz_nodes, 1ull);
			if (yi == 0.0) { vbuf[t] = -RHFQ_INF; return; }
			pv.y[t] = yi;
			pv.flags[t] = PLAN_LZ;
		}
	}
};

/*
 * Evaluates log-pdf contributions for a list of points.  Point p belongs to
 * posterior pt_post[p]; task (p, m) evaluates set m of that posterior.  Layout
 * of vbuf: [m][p] so that consecutive threads share a set only when Mmax == 1,
 * and share k_i through L1 otherwise.
 */
struct NodeEvalBody {
	const SetLocals* L; const double* k; const int64_t* post_off; const PostState* P;
	const int* pt_post; const double* pt_val; const double* pt_fb; int64_t n_pts; int Mmax;
	double* vbuf; int* post_err; Counters* cnt;
	RHFQ_HD void operator()(int64_t t) const {
		const int64_t m = t / n_pts;
		const int64_t p = t % n_pts;
		const int r = pt_post[p];
		const int64_t s0 = post_off[r];
		const int64_t M = post_off[r + 1] - s0;
		if (m >= M) return;
		if (P[r].status != ST_OK) { vbuf[t] = nan(""); return; }
		int st = ST_OK;
		const double v = log_pdf_term(L[s0 + m], k, pt_val[p], pt_fb[p], P[r].Qmax, st, cnt);
		if (st != ST_OK) atomic_max_i32(&post_err[r], st);
		vbuf[t] = v;
	}
};

/* log-sum-exp over the sets of a posterior (posterior.hpp:814-832) */
struct CombineBody {
	const int64_t* post_off; const int* pt_post; int64_t n_pts; const double* vbuf;
	double* out;   /* log pdf per point */
	RHFQ_HD void operator()(int64_t p) const {
		const int r = pt_post[p];
		const int64_t M = post_off[r + 1] - post_off[r];
		double mx = -RHFQ_INF;
		bool anynan = false;
		for (int64_t m = 0; m < M; ++m) {
			const double v = vbuf[m * n_pts + p];
			if (is_nan(v)) anynan = true;
			if (v > mx) mx = v;
		}
		if (anynan) { out[p] = nan(""); return; }
		if (is_inf(mx)) { out[p] = mx; return; }
		double S = 0.0, corr = 0.0;
		for (int64_t m = 0; m < M; ++m)
			kahan_add(S, corr, exp(vbuf[m * n_pts + p] - mx));
		out[p] = mx + log(S);
	}
};

/* --- BLI construction (barylagrint.hpp:544-668, posterior.hpp:674-719) --- */
struct BliPointsBody {
	/* Generates the new nodes of refinement level `level` for each active unit. */
	const PostState* P; const int* active_units; int level; int Rmax; int n_new; ChebTable T;
	int* pt_post; double* pt_val; double* pt_fb;
	RHFQ_HD void operator()(int64_t t) const {
		const int a = (int)(t / n_new);
		const int jn = (int)(t % n_new);
		const int u = active_units[a];
		const int r = u >> 1, which = u & 1;
		/* fine-grid index of the jn-th new node of this level */
		const int stride = 1 << (Rmax - level);
		const int fi = (level == 0) ? jn * stride : (2 * jn + 1) * stride;
		const int n_fine = (8 << Rmax) + 1;
		const PostState& ps = P[r];
		double xv, xfb;
		if (which == 0) {
			/* low interpolant on [0, 0.9 Qmax]; nodes descend from xmax (i=0) to xmin */
			const double xmax = 0.9 * ps.Qmax;
			if (fi == 0) xv = xmax;
			else if (fi == n_fine - 1) xv = 0.0;
			else xv = fmin(fmax(0.0 + 0.5 * (1.0 + T.val[fi]) * xmax, 0.0), xmax);
			xfb = ps.Qmax - xv;
		} else {
			const double Dt = ps.tmax - ps.tmin;
			double tv;
			if (fi == 0) tv = ps.tmax;
			else if (fi == n_fine - 1) tv = ps.tmin;
			else tv = fmin(fmax(ps.tmin + 0.5 * (1.0 + T.val[fi]) * Dt, ps.tmin), ps.tmax);
			const double xb = exp(tv);
			xv = ps.Qmax - xb;
			xfb = xb;
		}
		pt_post[t] = r;
		pt_val[t] = xv;
		pt_fb[t] = xfb;
	}
};

struct BliStoreBody {
	const int* active_units; int level; int Rmax; int n_new; const double* lpdf;
	double* bli_f; int* post_err; const int* pt_post;
	RHFQ_HD void operator()(int64_t t) const {
		const int a = (int)(t / n_new);
		const int jn = (int)(t % n_new);
		const int u = active_units[a];
		const int stride = 1 << (Rmax - level);
		const int fi = (level == 0) ? jn * stride : (2 * jn + 1) * stride;
		const int n_fine = (8 << Rmax) + 1;
		const double v = lpdf[t];
		/* NaN / INF function evaluation is an error (barylagrint.hpp:346-368) */
		if (!is_fin(v)) atomic_max_i32(&post_err[pt_post[t]], ST_RUNTIME);
		bli_f[(int64_t)u * n_fine + fi] = v;
	}
};

struct BliErrBody {
	/* Error of the level-`level` interpolant at the new nodes of level+1 */
	const int* active_units; int level; int Rmax; int n_new; ChebTable T; const double* bli_f;
	double* err_abs; double* err_rel;
	RHFQ_HD void operator()(int64_t t) const {
		const int a = (int)(t / n_new);
		const int jn = (int)(t % n_new);
		const int u = active_units[a];
		const int n_fine = (8 << Rmax) + 1;
		const int stride_old = 1 << (Rmax - level);
		const int stride_new = stride_old >> 1;
		const int fi = (2 * jn + 1) * stride_new;
		const double* f = bli_f + (int64_t)u * n_fine;
		const int n_old = (8 << level) + 1;
		/* new nodes are interior Chebyshev nodes: plain differences are exact here */
		const double fint = bli_interpolate_interior(T.val[fi], f, stride_old, n_old, T.val);
		const double fref = f[fi];
		double ea = fabs(fint - fref), er;
		if (fref == 0.0) er = (fint != 0.0) ? 1.0 : 0.0;
		else er = ea / fabs(fref);
		if (is_nan(ea)) { ea = RHFQ_INF; er = RHFQ_INF; }
		err_abs[t] = ea;
		err_rel[t] = er;
	}
};

struct BliDecideBody {
	const int* active_units; int level; int n_new; const double* err_abs; const double* err_rel;
	double rtol; PostState* P; int* still_active;
	RHFQ_HD void operator()(int64_t a) const {
		const int u = active_units[a];
		const int r = u >> 1, which = u & 1;
		double ea = 0.0, er = 0.0;
		for (int j = 0; j < n_new; ++j) {
			const double x = err_abs[(int64_t)a * n_new + j], y = err_rel[(int64_t)a * n_new + j];
			if (x > ea) ea = x;
			if (y > er) er = y;
		}
		const double tol_rel = rtol, tol_abs = rtol / P[r].Qmax;
		const bool ok = (ea < tol_abs) || (er < tol_rel);
		if (ok) {
			P[r].level[which] = level;
			still_active[a] = 0;
		} else {
			still_active[a] = 1;
		}
	}
};

struct BliFinishBody {
	const int* active_units; int level; PostState* P;
	RHFQ_HD void operator()(int64_t a) const {
		const int u = active_units[a];
		P[u >> 1].level[u & 1] = level;
	}
};

struct CompactBody {
	const int* in; const int* flag; int* out; int* counter;
	RHFQ_HD void operator()(int64_t i) const {
		if (flag[i]) out[atomic_add_i32(counter, 1)] = in[i];
	}
};

struct MarkBody {
	SetLocals* L; const int* e;
	RHFQ_HD void operator()(int64_t s) const {
		if (L[s].status == ST_OK && e[s] != ST_OK) L[s].status = e[s];
	}
};

struct FlagListBody {
	const int* list; SetLocals* L; int flag; Counters* cnt;
	RHFQ_HD void operator()(int64_t a) const {
		L[list[a]].flags |= flag;
		if (cnt) atomic_add_u64(&cnt->z_unconverged, 1ull);
	}
};

struct MarkListBody {
	const int* list; int* e; int code;
	RHFQ_HD void operator()(int64_t a) const { atomic_max_i32(&e[list[a]], code); }
};

struct StatusFlagBody {
	const SetLocals* L; int* flag; int* idx;
	RHFQ_HD void operator()(int64_t s) const { flag[s] = (L[s].status == ST_OK) ? 1 : 0; idx[s] = (int)s; }
};

struct NotBody {
	const int* in; int* out;
	RHFQ_HD void operator()(int64_t i) const { out[i] = in[i] ? 0 : 1; }
};

struct FillBody {
	const PostState* P; const int* post; const double* x; double* v; double* fb;
	RHFQ_HD void operator()(int64_t i) const { v[i] = x[i]; fb[i] = P[post[i]].Qmax - x[i]; }
};

struct OneBody {
	int64_t* oc;
	RHFQ_HD void operator()(int64_t i) const { oc[i] = 1; }
};

struct PostErrBody {
	PostState* P; const int* post_err;
	RHFQ_HD void operator()(int64_t r) const {
		if (P[r].status == ST_OK && post_err[r] != ST_OK)
			P[r].status = post_err[r];
	}
};

/*
 * Chebyshev coefficients of the kept interpolant (type-I DCT of the samples at
 * the Chebyshev-Lobatto nodes): p(z) = sum'' c_k T_k(z) is the SAME polynomial
 * the barycentric formula of barylagrint.hpp:674-720 evaluates, but Clenshaw's
 * recurrence needs no division and no node table per term.  Used for the many
 * interpolant evaluations of the Simpson tree and of pdf().
 */
/* Clenshaw-Reinsch evaluation of sum_k c_k T_k(z); zff = (1+z)/2, zfb = (1-z)/2
 * are given separately so that the recurrence is stable at both ends. */
RHFQ_HD double cheb_eval(const double* c, int N, double zff, double zfb)
{
	double b1 = 0.0, d = 0.0;
	if (zfb <= zff) {
		/* z >= 0: differences d_k = b_k - b_{k+1}, delta = 2 (z - 1) = -4 zfb */
		const double delta = -4.0 * zfb;
		for (int k = N; k >= 1; --k) {
			d = fma(delta, b1, d) + c[k];
			b1 = d + b1;
		}
		d = fma(delta, b1, d) + c[0];
		return d - 0.5 * delta * b1;
	}
	/* z < 0: sums d_k = b_k + b_{k+1}, delta = 2 (z + 1) = 4 zff */
	const double delta = 4.0 * zff;
	for (int k = N; k >= 1; --k) {
		d = fma(delta, b1, -d) + c[k];
		b1 = d - b1;
	}
	d = fma(delta, b1, -d) + c[0];
	return d - 0.5 * delta * b1;
}

/* Two evaluation points of the same series at once: two independent recurrences
 * in flight hide the FP64 dependency latency (the Simpson tree always needs the
 * two quarter points of an interval). */
template<bool UPA, bool UPB>
RHFQ_HD void cheb_eval2_t(const double* __restrict__ c, int N, double dA, double dB,
                          double& outA, double& outB)
{
	double bA = 0.0, eA = 0.0, bB = 0.0, eB = 0.0;
#pragma unroll 4
	for (int k = N; k >= 1; --k) {
		const double ck = c[k];
		eA = fma(dA, bA, UPA ? eA : -eA) + ck;
		eB = fma(dB, bB, UPB ? eB : -eB) + ck;
		bA = UPA ? eA + bA : eA - bA;
		bB = UPB ? eB + bB : eB - bB;
	}
	eA = fma(dA, bA, UPA ? eA : -eA) + c[0];
	eB = fma(dB, bB, UPB ? eB : -eB) + c[0];
	outA = eA - 0.5 * dA * bA;
	outB = eB - 0.5 * dB * bB;
}

RHFQ_HD void cheb_eval2(const double* c, int N, double zffA, double zfbA, double zffB,
                        double zfbB, double& outA, double& outB)
{
	const bool upA = zfbA <= zffA, upB = zfbB <= zffB;
	const double dA = upA ? -4.0 * zfbA : 4.0 * zffA;
	const double dB = upB ? -4.0 * zfbB : 4.0 * zffB;
	if (upA && upB) cheb_eval2_t<true, true>(c, N, dA, dB, outA, outB);
	else if (!upA && !upB) cheb_eval2_t<false, false>(c, N, dA, dB, outA, outB);
	else if (upA) cheb_eval2_t<true, false>(c, N, dA, dB, outA, outB);
	else cheb_eval2_t<false, true>(c, N, dA, dB, outA, outB);
}

/* ------------------------------------------------------------------ *
 *  Fine theta grid: O(1) evaluation of a kept interpolant             *
 * ------------------------------------------------------------------ *
 * The Simpson tree evaluates each interpolant ~10^4 times; a Clenshaw sum costs
 * n (up to 1024) terms per point.  p(cos theta) = sum_k c_k cos(k theta) is a
 * band-limited even function of theta, so it is tabulated ONCE on a uniform theta
 * grid with FINE_SIGMA-fold oversampling (G = FINE_SIGMA * n cells; one DCT-I by
 * an in-shared-memory FFT, O(G log G)) and then evaluated by FINE_PTS-point
 * Lagrange interpolation in theta: the interpolation error is
 * <= (pi / FINE_SIGMA)^12 / C(12,6) = 1.4e-8 of the HIGHEST-order coefficients
 * (~1e-8 themselves after refinement), i.e. below the rounding of the sum.  The
 * build kernel verifies the table against Clenshaw at probe points and flags the
 * interpolant for the direct evaluation if they disagree (unresolved functions).
 */
constexpr int FINE_SIGMA = 8;
constexpr int FINE_PTS = 12;
constexpr int FINE_PROBES = 2;
constexpr double FINE_TOL = 2e-12;

/* twiddles e^{-2 pi i k / (2 Gmax)}, k < Gmax: re at tw[k], im at tw[Gmax + k] */
struct FineTwiddle {
	const double* tw;
	int Gmax;
};

inline void make_fine_twiddle(int Rmax, std::vector<double>& tw)
{
	const int Gmax = FINE_SIGMA * (8 << Rmax);
	tw.resize(2 * (size_t)Gmax);
	const long double pi = 3.14159265358979323846264338327950288L;
	for (int k = 0; k < Gmax; ++k) {
		const long double a = pi * k / Gmax;
		tw[k] = (double)cosl(a);
		tw[Gmax + k] = (double)(-sinl(a));
	}
}

RHFQ_HD int bit_reverse(int m, int bits)
{
#if defined(__CUDA_ARCH__)
	return (int)(__brev((unsigned)m) >> (32 - bits));
#else
	int r = 0;
	for (int b = 0; b < bits; ++b) { r = (r << 1) | (m & 1); m >>= 1; }
	return r;
#endif
}

struct alignas(16) FftCplx { double x, y; };
RHFQ_HD int fft_rev5(int v)
{
#if defined(__CUDA_ARCH__)
	return (int)(__brev((unsigned)v) >> 27);
#else
	return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
#endif
}

/*
 * Cooperative DCT-I of x_0..x_G, G = 2^bits:
 *   Y_m = x_0 + (-1)^m x_G + 2 sum_{j=1}^{G-1} x_j cos(pi j m / G),  m = 0..G.
 * The even extension a (length 2G, a_j = a_{2G-j} = x_j) is packed as z_j = a_{2j} + i a_{2j+1};
 * one radix-2 decimation-in-frequency FFT of length G in the scratch arrays re/im (shared
 * memory on the device), then the split of the real transform.  get(j) supplies x_j, put(m, Y_m)
 * takes the result; `tid`/`nt`: this thread and the number of cooperating threads; `sync` is
 * the barrier between passes.
 */
template<typename Get, typename Put, typename Sync>
RHFQ_HD void dct1_fft(Get get, Put put, int bits, double* re, double* im,
                      const FineTwiddle& W, int tid, int nt, Sync sync)
{
	const int G = 1 << bits;
	/* Storage swizzle: element i lives at i ^ rev(top five bits of i) ^ (bits 4..7 of i).  The
	 * transform ends in bit-reversed order, so consecutive outputs sit G/2, G/4, ... apart --
	 * all in one shared-memory bank without the first term; the late passes walk through the
	 * array in strides of 16 ... 128 elements -- 16 lanes on one bank without the second
	 * (ncu: 41 % excess shared-memory wavefronts with the first term alone).  Unit-stride
	 * accesses stay conflict free under both. */
	const int sw = bits >= 10 ? bits - 5 : 31;
	/* (the top five bits enter REVERSED: consecutive outputs m differ in the top bits of
	 * brev(m) in reversed order, and reversed again they land on consecutive banks) */
#define RHFQ_PH(i) ((i) ^ fft_rev5(((i) >> sw) & 31) ^ (bits >= 10 ? (((i) >> 4) & 15) : 0))
	/* The scratch holds the G complex values INTERLEAVED (16-byte elements: one 128-bit
	 * shared-memory access per value instead of two 64-bit ones -- the instruction queue of
	 * the shared-memory pipe was the kernel's largest stall, ncu mio_throttle 3.9 per issue).
	 * All callers pass im = re + cap with cap >= G, so re[0 .. 2G) is theirs. */
	(void)im;
	FftCplx* const z = reinterpret_cast<FftCplx*>(re);
	for (int j = tid; j < G; j += nt) {
		const int e = 2 * j, o = 2 * j + 1;
		FftCplx v;
		v.x = get(e <= G ? e : 2 * G - e);
		v.y = get(o <= G ? o : 2 * G - o);
		z[RHFQ_PH(j)] = v;
	}
	sync();
	/* decimation in frequency, THREE radix-2 stages per pass (radix 8 in registers: a third
	 * of the shared-memory traffic and barriers of single stages), then one radix-4 or
	 * radix-2 pass for the stages left over */
	int half = G >> 1;
	for (; half >= 4; half >>= 3) {
		const int e = half >> 2;                     /* eighth of the current block (size 2 half) */
		const int tstep = W.Gmax / half;             /* W_{2 half}^k = tw[k * tstep] */
		constexpr double R2 = 0.70710678118654752440;
		for (int b = tid; b < (G >> 3); b += nt) {
			const int pos = b & (e - 1);
			const int l0 = ((b - pos) << 3) + pos;
			int ix[8];
			double xr[8], xi[8];
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				ix[k] = RHFQ_PH(l0 + k * e);
				const FftCplx v = z[ix[k]];
				xr[k] = v.x; xi[k] = v.y;
			}
			/* W^pos from the table, W^(2 pos) and W^(4 pos) by squaring (two of the three pairs of
			 * table loads were a third of the kernel's long-scoreboard stalls; the squares are
			 * good to 2-3 ulp, the transform's own rounding is larger) */
			const double w1r = W.tw[pos * tstep], w1i = W.tw[W.Gmax + pos * tstep];
			const double w2r = (w1r - w1i) * (w1r + w1i), w2i = 2.0 * w1r * w1i;
			const double w4r = (w2r - w2i) * (w2r + w2i), w4i = 2.0 * w2r * w2i;
			/* stage 1: (k, k + 4), twiddle W^pos W_8^k */
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const double sr = xr[k] + xr[k + 4], si = xi[k] + xi[k + 4];
				double dr = xr[k] - xr[k + 4], di = xi[k] - xi[k + 4];
				if (k == 1) { const double t = R2 * (dr + di); di = R2 * (di - dr); dr = t; }        /* (1 - i) / sqrt 2 */
				else if (k == 2) { const double t = di; di = -dr; dr = t; }                          /* -i */
				else if (k == 3) { const double t = R2 * (di - dr); di = -R2 * (dr + di); dr = t; }  /* (-1 - i) / sqrt 2 */
				xr[k] = sr; xi[k] = si;
				xr[k + 4] = dr * w1r - di * w1i; xi[k + 4] = dr * w1i + di * w1r;
			}
			/* stage 2: (k, k + 2) inside each half, twiddle W^(2 pos) W_4^k */
#pragma unroll
			for (int g = 0; g < 8; g += 4) {
#pragma unroll
				for (int k = 0; k < 2; ++k) {
					const double sr = xr[g + k] + xr[g + k + 2], si = xi[g + k] + xi[g + k + 2];
					double dr = xr[g + k] - xr[g + k + 2], di = xi[g + k] - xi[g + k + 2];
					if (k == 1) { const double t = di; di = -dr; dr = t; }
					xr[g + k] = sr; xi[g + k] = si;
					xr[g + k + 2] = dr * w2r - di * w2i; xi[g + k + 2] = dr * w2i + di * w2r;
				}
			}
			/* stage 3: (k, k + 1) inside each quarter, twiddle W^(4 pos) */
#pragma unroll
			for (int k = 0; k < 8; k += 2) {
				const double dr = xr[k] - xr[k + 1], di = xi[k] - xi[k + 1];
				FftCplx a, d;
				a.x = xr[k] + xr[k + 1]; a.y = xi[k] + xi[k + 1];
				d.x = dr * w4r - di * w4i; d.y = dr * w4i + di * w4r;
				z[ix[k]] = a;
				z[ix[k + 1]] = d;
			}
		}
		sync();
	}
	if (half == 2) {
		/* two stages left: one radix-4 pass on blocks of four */
		for (int b = tid; b < (G >> 2); b += nt) {
			const int l0 = b << 2;
			const int i0 = RHFQ_PH(l0), i1 = RHFQ_PH(l0 + 1), i2 = RHFQ_PH(l0 + 2), i3 = RHFQ_PH(l0 + 3);
			const FftCplx x0 = z[i0], x1 = z[i1], x2 = z[i2], x3 = z[i3];
			const double a0r = x0.x + x2.x, a0i = x0.y + x2.y, a2r = x0.x - x2.x, a2i = x0.y - x2.y;
			const double a1r = x1.x + x3.x, a1i = x1.y + x3.y;
			const double a3r = x1.y - x3.y, a3i = -(x1.x - x3.x);          /* (x1 - x3) times -i */
			FftCplx o0, o1, o2, o3;
			o0.x = a0r + a1r; o0.y = a0i + a1i;
			o1.x = a0r - a1r; o1.y = a0i - a1i;
			o2.x = a2r + a3r; o2.y = a2i + a3i;
			o3.x = a2r - a3r; o3.y = a2i - a3i;
			z[i0] = o0; z[i1] = o1; z[i2] = o2; z[i3] = o3;
		}
		sync();
	}
	if (half == 1) {
		for (int b = tid; b < (G >> 1); b += nt) {
			const int i0 = RHFQ_PH(b << 1), i1 = RHFQ_PH((b << 1) + 1);
			const FftCplx a = z[i0], c = z[i1];
			FftCplx s2, d2;
			s2.x = a.x + c.x; s2.y = a.y + c.y;
			d2.x = a.x - c.x; d2.y = a.y - c.y;
			z[i0] = s2; z[i1] = d2;
		}
		sync();
	}
	/* bit-reversed output: Z_m sits at brev(m) */
	const int wstep = W.Gmax / G;
	for (int m = tid; m <= G; m += nt) {
		if (m == G) { const FftCplx z0 = z[0]; put(G, z0.x - z0.y); continue; }
		const FftCplx a = z[RHFQ_PH(bit_reverse(m, bits))],
		              c = z[RHFQ_PH(bit_reverse((G - m) & (G - 1), bits))];
		const double Er = 0.5 * (a.x + c.x);
		const double Or = 0.5 * (a.y + c.y), Oi = -0.5 * (a.x - c.x);
		const double wr = W.tw[m * wstep], wi = W.tw[W.Gmax + m * wstep];
		put(m, Er + (wr * Or - wi * Oi));
	}
	sync();
#undef RHFQ_PH
}

/* asin(sqrt(u)) for u in [0, 1/2] as s + s u Q(u), s = sqrt(u): Q(u) = (asin(s)/s - 1)/u is
 * analytic up to u = 1, so a degree-18 near-minimax polynomial in t = 4u - 1 (Chebyshev
 * interpolant of a 60-digit evaluation, truncation 2e-17) gives it to rounding; the leading
 * term is added last, so the result is within 1 ulp (checked against mpmath over [0, 1/2]
 * and down to 1e-300).  asin(sqrt()) of the CUDA math library -- two square roots, a
 * division-free but branchy 90 instructions -- was 19 % of the Simpson kernel. */
/* (coefficients in constant memory: an FP64 literal costs two UMOV per use, a constant-bank
 * pair one LDCU.128) */
#define RHFQ_ASIN_Q                                                                          \
	{1.88790204786390997e-01, 2.62157695789168553e-02, 4.97983937697808483e-03,              \
	 1.09883478140597869e-03, 2.64620112259498163e-04, 6.74725995719136634e-05,              \
	 1.79096108523544007e-05, 4.89752092154258992e-06, 1.37032947181502114e-06,              \
	 3.90468504432157323e-07, 1.12916240313059778e-07, 3.30300466119960349e-08,              \
	 9.76916043707669294e-09, 2.95420115919907182e-09, 8.88593343912337773e-10,              \
	 2.33131875074799646e-10, 7.07723870404404399e-11, 3.95578439507660028e-11,              \
	 1.22056722459746048e-11, 0.0}
#if defined(__CUDACC__)
__constant__ double d_ASIN_Q[20] = RHFQ_ASIN_Q;
#endif
static const double h_ASIN_Q[20] = RHFQ_ASIN_Q;
#if defined(__CUDA_ARCH__)
#define ASIN_Q(k) d_ASIN_Q[k]
#else
#define ASIN_Q(k) h_ASIN_Q[k]
#endif
RHFQ_HD double asin_sqrt_half(double u)
{
	const double t = 4.0 * u - 1.0, t2 = t * t;
	double E = ASIN_Q(18), O = ASIN_Q(17);
#pragma unroll
	for (int k = 16; k >= 2; k -= 2) {
		E = fma(E, t2, ASIN_Q(k));
		O = fma(O, t2, ASIN_Q(k - 1));
	}
	E = fma(E, t2, ASIN_Q(0));
	const double sq = sqrt(u);
	return fma(sq * u, fma(t, O, E), sq);
}

/* FINE_PTS-point Lagrange interpolation in theta on the fine grid (G cells);
 * zff = (1+z)/2 = cos^2(theta/2), zfb = (1-z)/2 = sin^2(theta/2).
 * NP independent points are evaluated in lockstep, branch-free, so that their FP64
 * dependency chains interleave.  The node window is centred on the point's cell and
 * shifted inwards at the two ends of the grid (one-sided there: the error constant
 * grows ~100x, still ~1e-14 for resolved interpolants; the probe check covers it). */
template<int NP>
RHFQ_HD void fine_eval_n(const double* const* fine, const int* G, const double* zff,
                         const double* zfb, double* out)
{
	static_assert(FINE_PTS == 12, "the denominators are tabulated for 12 points");
	constexpr int H = FINE_PTS / 2;
	double s[NP];
	const double* f[NP];
#pragma unroll
	for (int h = 0; h < NP; ++h) {
		/* theta / pi without cancellation at either end: theta/2 = asin(sqrt(zfb)) */
		const bool up = zfb[h] <= zff[h];
		const double ah = asin_sqrt_half(fmin(fmax(up ? zfb[h] : zff[h], 0.0), 0.5));   /* [0, pi/4] */
		const double scale = (double)G[h] * 0.63661977236758134308;    /* G / (pi/2) */
		const double p = up ? ah * scale : (double)G[h] - ah * scale;
		int j0 = (int)p - (H - 1);
		j0 = j0 < 0 ? 0 : j0;
		j0 = j0 > G[h] - (FINE_PTS - 1) ? G[h] - (FINE_PTS - 1) : j0;
		s[h] = p - (double)j0;                  /* position relative to the first node */
		f[h] = fine[h] + j0;
	}
	double t[NP][FINE_PTS], w[NP][FINE_PTS];
	double acc[NP];
#pragma unroll
	for (int h = 0; h < NP; ++h) acc[h] = 1.0;
#pragma unroll
	for (int a = 0; a < FINE_PTS; ++a) {
#pragma unroll
		for (int h = 0; h < NP; ++h) {
			t[h][a] = s[h] - (double)a;
			w[h][a] = acc[h] * LAGRANGE12_INV_DEN(a);
			acc[h] *= t[h][a];
		}
	}
	double suf[NP], sum[NP];
#pragma unroll
	for (int h = 0; h < NP; ++h) { suf[h] = 1.0; sum[h] = 0.0; }
#pragma unroll
	for (int a = FINE_PTS - 1; a >= 0; --a) {
#pragma unroll
		for (int h = 0; h < NP; ++h) {
			sum[h] = fma(w[h][a] * suf[h], f[h][a], sum[h]);
			suf[h] *= t[h][a];
		}
	}
#pragma unroll
	for (int h = 0; h < NP; ++h) out[h] = sum[h];
}

RHFQ_HD double fine_eval(const double* fine, int G, double zff, double zfb)
{
	double out;
	fine_eval_n<1>(&fine, &G, &zff, &zfb, &out);
	return out;
}

struct CtaSync {
	RHFQ_HD void operator()() const {
#if defined(__CUDA_ARCH__)
		__syncthreads();
#endif
	}
};

/* Chebyshev coefficients of the kept interpolants (the DCT-I of the samples,
 * barylagrint.hpp's nodes z_j = cos(pi j / N)): c_k = (2/N) sum'' f_j cos(pi j k / N), with
 * the halving of the first and last term of the series folded into c_0 and c_N.
 * One cooperating group per interpolant, the same FFT as the fine grid. */
/* Reduction of two values per thread over the cooperating group (sum or max): warp shuffles,
 * then the warp results through shared memory; the totals are valid in thread 0 .. 31.
 * (A single thread walking over the nt partial values kept the CTA -- and with 128 KB of
 * shared memory the whole SM -- waiting for ~1000 dependent additions.) */
template<bool MAX>
RHFQ_HD void cta_reduce2(double& a, double& b, double* re, double* im, int tid, int nt)
{
#if defined(__CUDA_ARCH__)
	for (int o = 16; o > 0; o >>= 1) {
		const double ao = __shfl_xor_sync(0xffffffffu, a, o), bo = __shfl_xor_sync(0xffffffffu, b, o);
		if (MAX) { a = ao > a ? ao : a; b = bo > b ? bo : b; }
		else { a += ao; b += bo; }
	}
	if (nt <= 32) return;
	if ((tid & 31) == 0) { re[tid >> 5] = a; im[tid >> 5] = b; }
	__syncthreads();
	if (tid < 32) {
		const int nw = nt >> 5;
		a = tid < nw ? re[tid] : (MAX ? -RHFQ_INF : 0.0);
		b = tid < nw ? im[tid] : (MAX ? -RHFQ_INF : 0.0);
		for (int o = 16; o > 0; o >>= 1) {
			const double ao = __shfl_xor_sync(0xffffffffu, a, o), bo = __shfl_xor_sync(0xffffffffu, b, o);
			if (MAX) { a = ao > a ? ao : a; b = bo > b ? bo : b; }
			else { a += ao; b += bo; }
		}
	}
#else
	(void)a; (void)b; (void)re; (void)im; (void)tid; (void)nt;   /* host emulation: nt == 1 */
#endif
}

struct BliCoeffBody {
	const PostState* P; const double* bli_f; double* bli_c; int Rmax; FineTwiddle W;
	RHFQ_HD void run(int64_t u, double* re, double* im, int tid, int nt) const {
		const PostState& ps = P[u >> 1];
		if (ps.status != ST_OK) return;
		const int n_fine = (8 << Rmax) + 1;
		const int lev = ps.level[u & 1];
		const int N = 8 << lev;
		const int stride = 1 << (Rmax - lev);
		const double* f = bli_f + u * n_fine;
		double* c = bli_c + u * n_fine;
		const double inv = 1.0 / N;
		dct1_fft([=](int j) { return f[j * stride]; },
		         [=](int k, double y) { c[k] = (k == 0 || k == N) ? 0.5 * inv * y : inv * y; },
		         3 + lev, re, im, W, tid, nt, CtaSync());
	}
};

/* Refinement check of one active interpolant (barylagrint.hpp:576-668): the level-`level`
 * interpolant evaluated at the new nodes of level + 1 -- the odd points of the twice finer
 * Chebyshev-Lobatto grid -- by two FFTs (samples -> coefficients -> zero-padded synthesis),
 * O(N log N) instead of one O(N) barycentric sum per new node, followed by the max-error
 * reduction and the keep / refine decision.  One cooperating group per active unit. */
struct BliCheckBody {
	const int* active_units; int level; int Rmax; FineTwiddle W; const double* bli_f;
	double* cscratch;      /* [n_active][(8 << level) + 1] coefficients */
	double rtol; PostState* P; int* still_active;
	RHFQ_HD void run(int64_t a, double* re, double* im, int tid, int nt) const {
		const int u = active_units[a];
		const int r = u >> 1, which = u & 1;
		const int n_fine = (8 << Rmax) + 1;
		const int N = 8 << level;
		const int stride_old = 1 << (Rmax - level), stride_new = stride_old >> 1;
		const double* f = bli_f + (int64_t)u * n_fine;
		double* c = cscratch + a * (int64_t)(N + 1);
		const double inv = 1.0 / N;
		dct1_fft([=](int j) { return f[j * stride_old]; },
		         [=](int k, double y) { c[k] = (k == 0 || k == N) ? 0.5 * inv * y : inv * y; },
		         3 + level, re, im, W, tid, nt, CtaSync());
		double ea = 0.0, er = 0.0;
		dct1_fft([=](int k) { return k == 0 ? c[0] : (k <= N ? 0.5 * c[k] : 0.0); },
		         [&](int m, double fint) {
			         if (!(m & 1)) return;
			         const double fref = f[m * stride_new];
			         double e = fabs(fint - fref), rel;
			         if (fref == 0.0) rel = (fint != 0.0) ? 1.0 : 0.0;
			         else rel = e / fabs(fref);
			         if (is_nan(e)) { e = RHFQ_INF; rel = RHFQ_INF; }
			         if (e > ea) ea = e;
			         if (rel > er) er = rel;
		         },
		         4 + level, re, im, W, tid, nt, CtaSync());
		cta_reduce2<true>(ea, er, re, im, tid, nt);
		if (tid == 0) {
			const double tol_rel = rtol, tol_abs = rtol / P[r].Qmax;
			const bool ok = (ea < tol_abs) || (er < tol_rel);
			if (ok) P[r].level[which] = level;
			still_active[a] = ok ? 0 : 1;
		}
	}
};

/* One cooperating group (a CTA on the device) per kept interpolant of the posterior
 * block starting at r0: fine grid by FFT, then the probe check against Clenshaw. */
struct FineBuildBody {
	const PostState* P; const double* bli_c; int Rmax; FineTwiddle W;
	double* fine; int* fine_bad; int64_t r0; int bits_lo, bits_hi; Counters* cnt;
	RHFQ_HD static int64_t stride(int Rmax) { return (int64_t)FINE_SIGMA * (8 << Rmax) + 1; }
	RHFQ_HD void run(int64_t u, double* re, double* im, int tid, int nt) const {
		const int64_t r = r0 + (u >> 1);
		const int which = (int)(u & 1);
		const PostState& ps = P[r];
		if (ps.status != ST_OK) return;
		const int lev = ps.level[which];
		const int n = 8 << lev;
		int bits = 3 + lev;
		for (int sg = FINE_SIGMA; sg > 1; sg >>= 1) ++bits;      /* G = FINE_SIGMA * n */
		if (bits < bits_lo || bits > bits_hi) return;
		const int n_fine = (8 << Rmax) + 1;
		const double* c = bli_c + (r * 2 + which) * n_fine;
		double* out = fine + u * stride(Rmax);
		/* p(cos theta_m) = sum_k c_k cos(k theta_m): x_0 = c_0, x_k = c_k / 2, zero beyond n */
		dct1_fft([=](int k) { return k == 0 ? c[0] : (k <= n ? 0.5 * c[k] : 0.0); },
		         [=](int m, double y) { out[m] = y; }, bits, re, im, W, tid, nt, CtaSync());
		const int G = 1 << bits;
		/* probe check: the series summed term by term (all threads) at two points
		 * between grid nodes against the table */
		double zff[FINE_PROBES], zfb[FINE_PROBES], part[FINE_PROBES];
		for (int i = 0; i < FINE_PROBES; ++i) {
			const double theta = 2.0 * HALF_PI * (i == 0 ? 0.2113 : 0.7887) + (i + 0.37) * (2.0 * HALF_PI / G);
			double acc = 0.0;
			for (int k = tid; k <= n; k += nt) acc = fma(c[k], cos(k * theta), acc);
			part[i] = acc;
			if (tid == i) {
				const double ch = cos(0.5 * theta), sh = sin(0.5 * theta);
				zff[i] = ch * ch; zfb[i] = sh * sh;
			}
		}
		static_assert(FINE_PROBES == 2, "cta_reduce2 carries two values");
		cta_reduce2<false>(part[0], part[1], re, im, tid, nt);
		if (tid < FINE_PROBES) {
			const double direct = part[tid];
			const double tab = fine_eval(out, G, zff[tid], zfb[tid]);
			if (!(fabs(direct - tab) <= FINE_TOL * fmax(1.0, 0.01 * fabs(direct)))) {
				if (atomic_add_i32(&fine_bad[u], 1) == 0 && cnt) atomic_add_u64(&cnt->fine_bad, 1ull);
			}
		}
	}
};

/* pdf through the two interpolants (posterior.hpp:965-1008) */
RHFQ_HD bool bli_locate(const PostState& ps, const double* bli_f_r, int Rmax, const ChebTable& T,
                        double x, int& which, double& zff, double& zfb, double& val);

RHFQ_HD double pdf_bli(const PostState& ps, const double* bli_f_r /* 2 * n_fine */,
                       const double* bli_c_r /* 2 * n_fine */, int Rmax,
                       const ChebTable& T, double x)
{
	int which = 0;
	double zff, zfb, val;
	if (!bli_locate(ps, bli_f_r, Rmax, T, x, which, zff, zfb, val))
		return val;
	const int n_fine = (8 << Rmax) + 1;
	return exp(cheb_eval(bli_c_r + which * n_fine, 8 << ps.level[which], zff, zfb));
}

struct PdfBliBody {
	const PostState* P; const double* bli_f; const double* bli_c; int Rmax; ChebTable T;
	const int* pt_post; const double* x; double* out; int range_check;
	RHFQ_HD void operator()(int64_t i) const {
		const PostState& ps = P[pt_post[i]];
		if (ps.status != ST_OK) { out[i] = nan(""); return; }
		const double xi = x[i];
		if (range_check && (xi < 0.0 || xi >= ps.Qmax)) { out[i] = 0.0; return; }
		const int n_fine = (8 << Rmax) + 1;
		const int64_t o = (int64_t)pt_post[i] * 2 * n_fine;
		out[i] = pdf_bli(ps, bli_f + o, bli_c + o, Rmax, T, xi);
	}
};

struct ExpRangeBody {
	/* explicit pdf: exp of the combined log value with the out-of-range shortcuts
	 * (posterior.hpp:130-136, 1021-1030) */
	const PostState* P; const int* pt_post; const double* x; const double* lpdf; double* out;
	RHFQ_HD void operator()(int64_t i) const {
		const PostState& ps = P[pt_post[i]];
		if (ps.status != ST_OK) { out[i] = nan(""); return; }
		if (x[i] < 0.0 || x[i] >= ps.Qmax) { out[i] = 0.0; return; }
		out[i] = exp(lpdf[i]);
	}
};

/* --- Simpson tree (simpson.hpp:218-359), breadth-first over the whole batch --- */
struct SimpsonRule {
	double C0, C1, C2, I;
	RHFQ_HD SimpsonRule(double dx, double f1, double f2, double f3) {
		C0 = f1;
		C1 = (-3 * f1 + 4 * f2 - f3) / dx;
		C2 = 2 / (dx * dx) * (f1 - 2 * f2 + f3);
		I = dx / 6 * (f1 + 4 * f2 + f3);
		const double I2 = integral(dx);
		if (I2 > 0) {
			const double scale = I / I2;
			C0 *= scale; C1 *= scale; C2 *= scale;
		}
	}
	RHFQ_HD double integral(double x) const { return x * (C0 + x * (C1 / 2 + x * C2 / 3)); }
};

/*
 * The Simpson tree of the whole batch lives in one node pool.  A node is an interval
 * of the bisection; it stores the pdf at its midpoint (its end values are the
 * parent's), the index of its first child (-1: leaf) and the Simpson mass of its
 * subtree.  The pool is filled level by level: one launch evaluates the FRONTIER
 * (the intervals still being bisected) and each splitting interval claims two
 * consecutive pool slots through a warp-aggregated counter -- no scan, no
 * compaction, no final reordering.  The frontier carries the interval ends and end
 * values so that the pool never has to be read while it is built.
 *
 * The reference keeps the leaves in x-order with prefix sums of their masses
 * (simpson.hpp:321-355) and answers a query by binary search; here a query descends
 * from the root (<= 24 levels), adding the subtree masses of the siblings passed on
 * the left (forward integral) or on the right (backward integral): the same sums,
 * associated as a tree instead of sequentially.
 */
struct SaqTree {
	double* fmid;      /* pdf at the midpoint of the node's interval */
	double* S;         /* leaf: I = dx/6 (f1 + 4 f2 + f3); inner node: S[child] + S[child + 1] */
	int64_t* child;    /* first child, -1 for a leaf */
	/* Trees built by one CTA per posterior (SaqPostBody) record a child RELATIVE to the pool
	 * index of its level's first node, lv[r][depth of the child] -- the level's pool range is
	 * claimed only when the level above is complete.  nullptr: child indices are absolute. */
	const long long* lv = nullptr;
};
constexpr int SAQ_POST_LEVELS = 26;
/* pool index of the first child of node `id` of posterior r, given the child's depth */
RHFQ_HD int64_t saq_child(const SaqTree& tree, int64_t r, int child_depth, int64_t c)
{
	return tree.lv ? (int64_t)tree.lv[r * SAQ_POST_LEVELS + child_depth] + c : c;
}

/* A frontier entry is 32 bytes: the three pdf values and a key (posterior << 32 | j), j the
 * position of the interval among the 2^L intervals of its level: [j, j + 1] Qmax / 2^L.  The
 * pool index of entry k is the level's first index + k (children are appended to the next
 * frontier and to the pool in the same order), so neither the ends nor the index are stored
 * -- the step kernel moves 130 B per interval from / to HBM at 2.7 TB/s, and the entry used
 * to be 52 of them. */
struct SaqFrontier {
	double* f1; double* f2; double* f3;
	unsigned long long* key;
	int64_t n;
};
RHFQ_HD unsigned long long saq_key(int post, int64_t j) { return ((unsigned long long)(unsigned)post << 32) | (unsigned long long)j; }
/* x = Qmax * num / 2^lev (num / 2^lev is exact) */
RHFQ_HD double saq_x(double Qmax, int64_t num, double inv_pow) { return Qmax * ((double)num * inv_pow); }
RHFQ_HD double saq_inv_pow(int lev) { return ldexp(1.0, -lev); }

/* Device-side control block of one posterior block's tree build: the level launches read the
 * frontier length and the pool position from here, so that the host queues all levels of a
 * block without reading anything back (one read-back per block instead of one per level). */
constexpr int SAQ_CTL_LEVELS = 32;
struct SaqCtl {
	unsigned long long count[SAQ_CTL_LEVELS];   /* frontier length of tree level L */
	unsigned long long first[SAQ_CTL_LEVELS];   /* pool index of level L's first node */
	int overflow;                               /* a frontier or the pool was too small: redo */
	int pad;
};

RHFQ_HD int64_t saq_claim_pair(unsigned long long* counter)
{
#if defined(__CUDA_ARCH__)
	/* one atomic per converged group; slots in lane order, so that the x-order of a
	 * warp's intervals (and the locality of the next level) is kept */
	const unsigned mask = __activemask();
	const int lane = (int)(threadIdx.x & 31);
	const int leader = __ffs(mask) - 1;
	const int rank = __popc(mask & ((1u << lane) - 1u));
	unsigned long long base = 0;
	if (lane == leader) base = atomicAdd(counter, 2ull * (unsigned)__popc(mask));
	base = __shfl_sync(mask, base, leader);
	return (int64_t)base + 2 * rank;
#else
	const unsigned long long o = *counter;
	*counter += 2;
	return (int64_t)o;
#endif
}

struct SaqRootPointsBody {
	const PostState* P; int* pt_post; double* pt_x; int64_t r0;
	RHFQ_HD void operator()(int64_t t) const {
		const int r = (int)(r0 + t / 3), j = (int)(t % 3);
		pt_post[t] = r;
		const double Q = P[r].Qmax;
		pt_x[t] = (j == 0) ? 0.0 : (j == 1 ? (Q + 0.0) / 2 : Q);
	}
};

/* root intervals (simpson.hpp:243-256): pool nodes base..base+Rb-1 */
struct SaqRootBody {
	const PostState* P; const double* fv; SaqFrontier fr; SaqTree tree; int64_t r0; int64_t base;
	double* root_f1; double* root_f3; int64_t* root_id;
	RHFQ_HD void operator()(int64_t i) const {
		fr.f1[i] = fv[3 * i]; fr.f2[i] = fv[3 * i + 1]; fr.f3[i] = fv[3 * i + 2];
		fr.key[i] = saq_key((int)(r0 + i), 0);
		root_f1[r0 + i] = fv[3 * i]; root_f3[r0 + i] = fv[3 * i + 2];
		root_id[r0 + i] = base + i;
	}
};

struct SaqPointsBody {
	SaqFrontier fr; int64_t first; const PostState* P; int depth; int* pt_post; double* pt_x;
	RHFQ_HD void operator()(int64_t t) const {
		const int64_t i = first + (t >> 1);
		const unsigned long long key = fr.key[i];
		const int post = (int)(key >> 32);
		const int64_t j = (int64_t)(key & 0xffffffffull);
		pt_post[t] = post;
		pt_x[t] = saq_x(P[post].Qmax, 4 * j + ((t & 1) ? 3 : 1), saq_inv_pow(depth + 2));
	}
};

/* simpson.hpp:258-311: accept / split decision of one frontier interval.  The reference
 * compares the Simpson masses of the two halves, Il and Ir, with the integrals of the parent's
 * parabola over the same halves, Ic and I - Ic.  All four are dx/24 times a combination of
 * the five pdf values:
 *   Il = dx/24 * 2 (f1 + 4 a12 + f2)     Ic     = dx/24 * (5 f1 + 8 f2 - f3)
 *   Ir = dx/24 * 2 (f2 + 4 a23 + f3)     I - Ic = dx/24 * (-f1 + 8 f2 + 5 f3)
 * (the reference's rescaling of the parabola by I/I2 is 1 up to rounding), so the test needs
 * neither dx nor a division -- the nine divisions of three SimpsonRule constructors were 15 %
 * of the Simpson kernel.  Decisions can differ from the reference's arithmetic only when a
 * difference sits within rounding of its threshold. */
RHFQ_HD bool saq_accept(double f1, double f2, double f3, double a12, double a23, int level,
                        int min_levels, int max_levels, double tol = SQRT_EPS)
{
	const double Il = 2.0 * (f1 + 4.0 * a12 + f2), Ir = 2.0 * (f2 + 4.0 * a23 + f3);
	const double Icl = 5.0 * f1 + 8.0 * f2 - f3, Icr = 5.0 * f3 + 8.0 * f2 - f1;
	if ((fabs(Il - Icl) <= tol * fabs(Il)) && (fabs(Ir - Icr) <= tol * fabs(Ir))
	    && level >= min_levels)
		return true;
	return level >= max_levels;
}

/* Outcome of one frontier interval given the pdf at its quarter points: a leaf keeps its
 * Simpson mass; a split claims two pool slots and appends the halves to the next frontier. */
struct SaqSplit {
	SaqFrontier fr; SaqFrontier nx; SaqTree tree; unsigned long long* next_count;
	int64_t next_base;   /* pool index of the next level's first node */
	int64_t first;       /* this launch covers the frontier entries first, first + 1, ... */
	int level; int max_levels; int min_levels;   /* level = depth of the intervals + 1 */
	int64_t level_base = 0;   /* pool index of this level's first node */
	double inv_pow = 1.0;     /* 2^-depth of this level's intervals (set with `level`) */
	/* device-controlled levels: capacities of the next frontier and of the pool; a claim beyond
	 * them raises *overflow (host-controlled levels size both beforehand: overflow == nullptr) */
	int64_t nx_cap = 0, pool_cap = 0;
	int* overflow = nullptr;
	double tol = SQRT_EPS;    /* simpson.hpp:222: sqrt(eps) of the reference's arithmetic type */
	/* all fields of a frontier entry, loaded up front so that the loads are in flight
	 * while the quarter points are evaluated */
	struct Item { double f1, f2, f3; int post; int64_t j; int64_t id; };
	RHFQ_HD Item load(int64_t i) const {
		Item it;
		const unsigned long long key = fr.key[i];
		it.f1 = fr.f1[i]; it.f2 = fr.f2[i]; it.f3 = fr.f3[i];
		it.post = (int)(key >> 32); it.j = (int64_t)(key & 0xffffffffull);
		it.id = level_base + i;
		return it;
	}
	/* every interval writes its OWN pool entry (its slot was claimed by the parent one
	 * level earlier), so the pool only ever needs room for levels whose size is known */
	RHFQ_HD void dead(const Item& it) const {
		tree.child[it.id] = -1;
		tree.fmid[it.id] = 0.0;
		tree.S[it.id] = 0.0;
	}
	/* dx = Qmax / 2^depth: width of the intervals of this level */
	RHFQ_HD void finish(const Item& it, double dx, double a12, double a23) const {
		tree.fmid[it.id] = it.f2;
		if (saq_accept(it.f1, it.f2, it.f3, a12, a23, level, min_levels, max_levels, tol)) {
			tree.child[it.id] = -1;
			tree.S[it.id] = dx / 6 * (it.f1 + 4 * it.f2 + it.f3);
			return;
		}
		const int64_t k = saq_claim_pair(next_count);
		const int64_t c = next_base + k;
		if (overflow && (k + 2 > nx_cap || c + 2 > pool_cap)) {
			*overflow = 1;
			tree.child[it.id] = -1;
			tree.S[it.id] = 0.0;
			return;
		}
		tree.child[it.id] = c;
		nx.f1[k] = it.f1; nx.f2[k] = a12; nx.f3[k] = it.f2;
		nx.key[k] = saq_key(it.post, 2 * it.j);
		nx.f1[k + 1] = it.f2; nx.f2[k + 1] = a23; nx.f3[k + 1] = it.f3;
		nx.key[k + 1] = saq_key(it.post, 2 * it.j + 1);
	}
};

/* explicit / adaptive_simpson algorithms: the quarter-point values come from the node kernels */
struct SaqDecideBody {
	SaqSplit sp; const PostState* P; const double* fv;
	RHFQ_HD void operator()(int64_t t) const {
		const SaqSplit::Item it = sp.load(sp.first + t);
		if (P[it.post].status != ST_OK) { sp.dead(it); return; }
		sp.finish(it, P[it.post].Qmax * sp.inv_pow, fv[2 * t], fv[2 * t + 1]);
	}
};

/* Chebyshev argument of x in the interpolant `which` (pdf_single<BLI>, posterior.hpp:965-1008);
 * returns false when the pdf is identically zero or an end sample applies (value in `val`). */
RHFQ_HD bool bli_locate(const PostState& ps, const double* bli_f_r, int Rmax, const ChebTable& T,
                        double x, int& which, double& zff, double& zfb, double& val)
{
	const double x_fb = ps.Qmax - x;
	val = 0.0;
	if (x < 0.0 || x_fb <= DBL_EPS) return false;
	const double t = log(x_fb);
	if (t <= ps.tmin) return false;
	const int n_fine = (8 << Rmax) + 1;
	which = (t >= ps.tmax) ? 0 : 1;
	const double xmin = which ? ps.tmin : 0.0;
	const double xmax = which ? ps.tmax : 0.9 * ps.Qmax;
	const double xv = which ? t : x;
	const double Dx = xmax - xmin;
	zff = fmin((xv - xmin) / Dx, 1.0);
	zfb = fmin((xmax - xv) / Dx, 1.0);
	if (zfb == 0.0) { val = exp(bli_f_r[which * n_fine]); return false; }
	if (zff == 0.0) { val = exp(bli_f_r[which * n_fine + n_fine - 1]); return false; }
	if (pii_in_front(zff, zfb, T.se) || pii_in_back(zff, zfb, T.se)) {
		/* Within sqrt(eps) of an end of the interpolation range the reference's
		 * barycentric sum weighs the end sample with the end-distance coordinate
		 * (intervall.hpp:106-117), which is NOT the interpolating polynomial
		 * (deviation up to ~1e-5 in log-pdf): reproduce it with the original formula. */
		const int lev = ps.level[which];
		const double zv = fmin(fmax(2.0 * zff - 1.0, -1.0), 1.0);
		val = exp(bli_interpolate(zv, zff, zfb, bli_f_r + which * n_fine, 1 << (Rmax - lev),
		                          (8 << lev) + 1, T));
		return false;
	}
	return true;
}

/* Fused Simpson step for the barycentric-Lagrange algorithm: quarter points,
 * interpolant evaluation (two Clenshaw recurrences interleaved) and the accept
 * test in one kernel, one thread per frontier interval. */
struct SaqStepBliBody {
	SaqSplit sp; const PostState* P; const double* bli_f; const double* bli_c; int Rmax;
	ChebTable T; Counters* cnt;
	/* fine theta grids of the posterior block starting at r0 (nullptr: Clenshaw only) */
	const double* fine; const int* fine_bad; int64_t r0;

	/* Rare cases of pdf_single<BLI> (posterior.hpp:965-1008) that the lockstep path
	 * below defers to: end samples, the intervall.hpp:106-117 quirk zone (see
	 * bli_locate) and interpolants without a usable fine grid. */
	RHFQ_HD double eval_special(const PostState& ps, int r, const double* fr, int which,
	                            double zff, double zfb, unsigned long long& terms) const {
		const int n_fine = (8 << Rmax) + 1;
		if (zfb == 0.0) return exp(fr[which * n_fine]);
		if (zff == 0.0) return exp(fr[which * n_fine + n_fine - 1]);
		const int lev = ps.level[which];
		if (pii_in_front(zff, zfb, T.se) || pii_in_back(zff, zfb, T.se)) {
			const double zv = fmin(fmax(2.0 * zff - 1.0, -1.0), 1.0);
			return exp(bli_interpolate(zv, zff, zfb, fr + which * n_fine, 1 << (Rmax - lev),
			                           (8 << lev) + 1, T));
		}
		terms += (unsigned long long)(8 << lev);
		return exp(cheb_eval(bli_c + ((int64_t)r * 2 + which) * n_fine, 8 << lev, zff, zfb));
	}

	/* pdf at NP of the two quarter points of frontier interval i, starting with point h0.
	 * NP = 2: both in lockstep by one thread (one point per thread, two threads per interval
	 * exchanging values by shuffle, was measured 15-30 % slower).  Returns false if the
	 * posterior is not healthy. */
	template<int NP>
	RHFQ_HD bool eval_quarter_points(const SaqSplit::Item& it, int h0, double* v) const {
		const int r = it.post;
		const PostState& ps = P[r];
		if (ps.status != ST_OK) return false;
		const double Qmax = ps.Qmax, tmin = ps.tmin, tmax = ps.tmax;
		const double inv_pow = 0.25 * sp.inv_pow;      /* quarter points: depth + 2 */
		double x[NP];
#pragma unroll
		for (int h = 0; h < NP; ++h) x[h] = saq_x(Qmax, 4 * it.j + ((h0 + h == 0) ? 1 : 3), inv_pow);
		const int n_fine = (8 << Rmax) + 1;
		const double* fr = bli_f + (int64_t)r * 2 * n_fine;
		/* Same decisions as bli_locate, arranged so that the low interpolant needs no
		 * logarithm: t = log(x_fb) is only compared with tmax there, i.e. x_fb with
		 * (1 - 0.9) Qmax; within 1e-12 of that threshold the logarithm decides. */
		const double thr = (1.0 - 0.9) * Qmax * (1.0 + 1e-12);
		double x_fb[NP], t[NP];
		bool dead[NP], need_log[NP];
		bool any_log = false;
#pragma unroll
		for (int h = 0; h < NP; ++h) {
			t[h] = 0.0;
			x_fb[h] = Qmax - x[h];
			dead[h] = x[h] < 0.0 || x_fb[h] <= DBL_EPS;
			need_log[h] = !(x_fb[h] > thr);
			any_log = any_log || need_log[h];
		}
		if (any_log) {
#pragma unroll
			for (int h = 0; h < NP; ++h) t[h] = log(x_fb[h]);
		}
		int which[NP];
		double zff[NP], zfb[NP];
		bool special[NP];
		const double* tab[NP];
		int G[NP];
#pragma unroll
		for (int h = 0; h < NP; ++h) {
			which[h] = (need_log[h] && !(t[h] >= tmax)) ? 1 : 0;
			dead[h] = dead[h] || (need_log[h] && t[h] <= tmin);
			const double xmin = which[h] ? tmin : 0.0;
			const double xmax = which[h] ? tmax : 0.9 * Qmax;
			const double xv = which[h] ? t[h] : x[h];
			const double inv_dx = which[h] ? ps.inv_hi : ps.inv_lo;     /* 1 / (xmax - xmin) */
			zff[h] = fmin((xv - xmin) * inv_dx, 1.0);
			zfb[h] = fmin((xmax - xv) * inv_dx, 1.0);
			const int64_t u = 2 * ((int64_t)r - r0) + which[h];
			special[h] = zfb[h] == 0.0 || zff[h] == 0.0 || pii_in_front(zff[h], zfb[h], T.se)
			             || pii_in_back(zff[h], zfb[h], T.se) || !fine || fine_bad[fine ? u : 0] != 0;
			tab[h] = fine + (fine ? u * FineBuildBody::stride(Rmax) : 0);
			G[h] = FINE_SIGMA * (8 << ps.level[which[h]]);
		}
		unsigned long long terms = 0, other = 0;
		if (fine) {
			double lp[NP];
			fine_eval_n<NP>(tab, G, zff, zfb, lp);
#pragma unroll
			for (int h = 0; h < NP; ++h) v[h] = exp(lp[h]);
		}
#pragma unroll
		for (int h = 0; h < NP; ++h) {
			if (dead[h]) { v[h] = 0.0; ++other; }
			else if (special[h]) { v[h] = eval_special(ps, r, fr, which[h], zff[h], zfb[h], terms); ++other; }
		}
		/* the common case (fine-grid evaluations) is counted on the host from the frontier
		 * sizes; only the exceptions cost an atomic */
		if (cnt && other) {
			atomic_add_u64(&cnt->cheb_terms, terms);
			atomic_add_u64(&cnt->saq_other, other);
		}
		return true;
	}

	RHFQ_HD void operator()(int64_t tix) const {
		const SaqSplit::Item it = sp.load(sp.first + tix);
		double v[2] = {0.0, 0.0};
		if (!eval_quarter_points<2>(it, 0, v)) { sp.dead(it); return; }
		sp.finish(it, P[it.post].Qmax * sp.inv_pow, v[0], v[1]);
	}
};

/*
 * Simpson tree, one posterior per CTA (persistent CTAs).
 *
 * The level-synchronous kernel above streams every frontier interval of every posterior
 * through HBM once per level and gathers the fine theta grids through L2 (ncu: 3.1 long-
 * scoreboard stalls per issued instruction).  Here a CTA owns a posterior from root to leaves:
 *   - small fine theta grids (<= 20 KB per CTA: up to five refinements) are staged in shared
 *     memory, the large ones are read through L1 / L2 (staging them made no difference);
 *   - the frontier ping-pongs between two CTA-private buffers that the CTA reuses for one
 *     posterior after the other: most of it stays in the 126 MB L2;
 *   - slots of the next level are claimed through a shared-memory counter, the pool range of a
 *     level through ONE global atomic per level, one barrier per level;
 *   - the subtree masses are summed bottom-up afterwards by SaqPostSumBody (inside this kernel
 *     the latency-bound walk over the levels held 13 % of its stall samples).
 * Same evaluations, same accept test, same leaf masses as the level-synchronous build; only the
 * pool placement differs.
 */
#ifndef RHFQ_SAQP_MINB
#define RHFQ_SAQP_MINB 6
#endif
#ifndef RHFQ_SAQP_THREADS
#define RHFQ_SAQP_THREADS 128
#endif
#define RHFQ_SAQP_MINB_HOST RHFQ_SAQP_MINB
#define RHFQ_SAQP_THREADS_HOST RHFQ_SAQP_THREADS
struct SaqPostShared {            /* head of the CTA's shared memory */
	long long base[SAQ_POST_LEVELS];   /* pool index of the first node of each level */
	int count[SAQ_POST_LEVELS];
	int next_count[2];                 /* children claimed so far, levels alternate */
	int arrived;                       /* warps that have finished the current level */
	int depth;                         /* levels that hold nodes */
	int r;
	int overflow;
};
RHFQ_HD void warp_sync()
{
#if defined(__CUDA_ARCH__)
	__syncwarp();
#endif
}
RHFQ_HD void fence_block()
{
#if defined(__CUDA_ARCH__)
	__threadfence_block();
#endif
}

struct SaqPostBody {
	SaqStepBliBody b;          /* interpolants, fine grids (global), counters; b.sp.tree = the pool */
	const double* fv;          /* [3 Rb] pdf at 0, Qmax / 2, Qmax of the block's posteriors */
	int64_t Rb;
	double* root_f1; double* root_f3; int64_t* root_id; double* norm;
	unsigned long long* pool_ptr; int64_t pool_cap;     /* pool_ptr[1]: intervals processed */
	/* CTA-private frontier ping-pong: buffer s of CTA c starts at (2 c + s) * cap */
	double* ff1; double* ff2; double* ff3; unsigned* fj; int64_t cap;
	int* next_post; int* overflow;
	int max_levels, min_levels;
	int smem_doubles;          /* room for fine grids behind the SaqPostShared header */
	/* per posterior of the block: pool index of the first node and node count of every tree
	 * level, [Rb][SAQ_POST_LEVELS] each, for the bottom-up mass kernel (SaqPostSumBody) */
	long long* lv_base; int* lv_count;

	template<typename Sync>
	RHFQ_HD void run(int cta, int tid, int nt, SaqPostShared* sh, double* sgrid, Sync sync) const
	{
		const SaqTree& tree = b.sp.tree;
		unsigned long long steps = 0;
		for (;;) {
			if (tid == 0) sh->r = atomic_add_i32(next_post, 1);
			sync();
			const int64_t rl = sh->r;
			if (rl >= Rb) break;
			const int64_t r = b.r0 + rl;
			const PostState& ps = b.P[r];
			if (ps.status != ST_OK) {
				if (tid == 0) {
					const int64_t id = (int64_t)atomic_add_u64_ret(pool_ptr, 1ull);
					if (id + 1 > pool_cap) { *overflow = 1; }
					else { tree.child[id] = -1; tree.fmid[id] = 0.0; tree.S[id] = 0.0; }
					root_id[r] = id < pool_cap ? id : 0;
					root_f1[r] = fv[3 * rl]; root_f3[r] = fv[3 * rl + 2];
					for (int L = 0; L < SAQ_POST_LEVELS; ++L) {
						lv_base[rl * SAQ_POST_LEVELS + L] = (L == 0 && id < pool_cap) ? id : 0;
						lv_count[rl * SAQ_POST_LEVELS + L] = (L == 0 && id < pool_cap) ? 1 : 0;
					}
				}
				sync();
				continue;
			}
			/* per-posterior constants */
			const double Qmax = ps.Qmax, tmin = ps.tmin, tmax = ps.tmax;
			const double inv_lo = ps.inv_lo, inv_hi = ps.inv_hi;
			const int lev0 = ps.level[0], lev1 = ps.level[1];
			const int G0 = FINE_SIGMA * (8 << lev0), G1 = FINE_SIGMA * (8 << lev1);
			const int64_t u0 = 2 * rl;
			const bool bad0 = !b.fine || b.fine_bad[u0] != 0, bad1 = !b.fine || b.fine_bad[u0 + 1] != 0;
			const double* g0 = b.fine ? b.fine + u0 * FineBuildBody::stride(b.Rmax) : nullptr;
			const double* g1 = b.fine ? b.fine + (u0 + 1) * FineBuildBody::stride(b.Rmax) : nullptr;
			/* stage the grids: the high interpolant first (it is the one at the refinement cap),
			 * whatever does not fit is read from global memory */
			const double* t1 = g1; const double* t0 = g0;
			int used = 0;
			if (!bad1 && G1 + 1 <= smem_doubles) {
				for (int k = tid; k <= G1; k += nt) sgrid[k] = g1[k];
				t1 = sgrid; used = G1 + 1;
			}
			if (!bad0 && used + G0 + 1 <= smem_doubles) {
				for (int k = tid; k <= G0; k += nt) sgrid[used + k] = g0[k];
				t0 = sgrid + used;
			}
			if (tid == 0) {
				const int64_t id = (int64_t)atomic_add_u64_ret(pool_ptr, 1ull);
				sh->overflow = (id + 1 > pool_cap) ? 1 : 0;
				sh->base[0] = id; sh->count[0] = 1;
				sh->next_count[0] = 0; sh->arrived = 0; sh->depth = 1;
				root_id[r] = id < pool_cap ? id : 0;
				root_f1[r] = fv[3 * rl]; root_f3[r] = fv[3 * rl + 2];
				/* root interval (simpson.hpp:243-256) */
				const int64_t o = (int64_t)(2 * cta) * cap;
				ff1[o] = fv[3 * rl]; ff2[o] = fv[3 * rl + 1]; ff3[o] = fv[3 * rl + 2]; fj[o] = 0u;
			}
			sync();
			int cur = 0;
			double inv_pow = 1.0;                     /* 2^-L: width of the level's intervals / Qmax */
			const int n_fine = (8 << b.Rmax) + 1;
			const double* fr = b.bli_f + r * 2 * n_fine;
			const double thr = (1.0 - 0.9) * Qmax * (1.0 + 1e-12);
			const int n_warps = (nt + 31) / 32;
			/* ONE barrier per level: the last warp to finish a level (arrival counter) knows the
			 * number of children, claims their pool range and publishes it before it joins the
			 * barrier.  Children are recorded relative to that range (SaqTree::lv). */
			for (int L = 0; L < max_levels; ++L) {
				const int n = sh->count[L];
				if (n == 0 || sh->overflow) break;
				int* const ctr = &sh->next_count[L & 1];
				const int64_t oc = (int64_t)(2 * cta + cur) * cap, on = (int64_t)(2 * cta + (cur ^ 1)) * cap;
				const int64_t base = sh->base[L];
				const double dx = Qmax * inv_pow, ip4 = 0.25 * inv_pow;
				for (int i = tid; i < n; i += nt) {
					const double f1 = ff1[oc + i], f2 = ff2[oc + i], f3 = ff3[oc + i];
					const unsigned j = fj[oc + i];
					const int64_t id = base + i;
					/* the two quarter points in lockstep (SaqStepBliBody::eval_quarter_points with
					 * the posterior's constants in registers and its grids in shared memory) */
					double x[2], x_fb[2], t[2], v[2];
					bool dead[2], need_log[2];
					bool any_log = false;
#pragma unroll
					for (int h = 0; h < 2; ++h) {
						x[h] = saq_x(Qmax, 4 * (int64_t)j + (h == 0 ? 1 : 3), ip4);
						t[h] = 0.0;
						x_fb[h] = Qmax - x[h];
						dead[h] = x[h] < 0.0 || x_fb[h] <= DBL_EPS;
						need_log[h] = !(x_fb[h] > thr);
						any_log = any_log || need_log[h];
					}
					if (any_log) {
#pragma unroll
						for (int h = 0; h < 2; ++h) t[h] = log(x_fb[h]);
					}
					int which[2], G[2];
					double zff[2], zfb[2];
					bool special[2];
					const double* tab[2];
#pragma unroll
					for (int h = 0; h < 2; ++h) {
						which[h] = (need_log[h] && !(t[h] >= tmax)) ? 1 : 0;
						dead[h] = dead[h] || (need_log[h] && t[h] <= tmin);
						const double xmin = which[h] ? tmin : 0.0;
						const double xmax = which[h] ? tmax : 0.9 * Qmax;
						const double xv = which[h] ? t[h] : x[h];
						const double inv_dx = which[h] ? inv_hi : inv_lo;
						zff[h] = fmin((xv - xmin) * inv_dx, 1.0);
						zfb[h] = fmin((xmax - xv) * inv_dx, 1.0);
						special[h] = zfb[h] == 0.0 || zff[h] == 0.0 || pii_in_front(zff[h], zfb[h], b.T.se)
						             || pii_in_back(zff[h], zfb[h], b.T.se) || (which[h] ? bad1 : bad0);
						tab[h] = which[h] ? t1 : t0;
						G[h] = which[h] ? G1 : G0;
					}
					unsigned long long terms = 0, other = 0;
					if (b.fine) {
						/* a rejected grid is never dereferenced beyond its first window; its value
						 * is replaced below */
						double lp[2];
						fine_eval_n<2>(tab, G, zff, zfb, lp);
						v[0] = exp(lp[0]); v[1] = exp(lp[1]);
					}
#pragma unroll
					for (int h = 0; h < 2; ++h) {
						if (dead[h]) { v[h] = 0.0; ++other; }
						else if (special[h]) { v[h] = b.eval_special(ps, (int)r, fr, which[h], zff[h], zfb[h], terms); ++other; }
					}
					if (b.cnt && other) {
						atomic_add_u64(&b.cnt->cheb_terms, terms);
						atomic_add_u64(&b.cnt->saq_other, other);
					}
					tree.fmid[id] = f2;
					if (saq_accept(f1, f2, f3, v[0], v[1], L + 1, min_levels, max_levels, b.sp.tol)) {
						tree.child[id] = -1;
						tree.S[id] = dx / 6 * (f1 + 4 * f2 + f3);
					} else {
						const int k = claim_pair_i32(ctr);
						if ((int64_t)k + 2 > cap) {
							sh->overflow = 1;
							tree.child[id] = -1;
							tree.S[id] = 0.0;
						} else {
							tree.child[id] = k;           /* relative to the next level's pool range */
							ff1[on + k] = f1; ff2[on + k] = v[0]; ff3[on + k] = f2; fj[on + k] = 2u * j;
							ff1[on + k + 1] = f2; ff2[on + k + 1] = v[1]; ff3[on + k + 1] = f3; fj[on + k + 1] = 2u * j + 1u;
						}
					}
				}
				warp_sync();
				if ((tid & 31) == 0) {
					fence_block();
					if (atomic_add_i32(&sh->arrived, 1) == n_warps - 1) {
						const int n_next = atomic_add_i32(ctr, 0);
						long long nb = 0;
						if (n_next > 0 && !sh->overflow) {
							nb = (long long)atomic_add_u64_ret(pool_ptr, (unsigned long long)n_next);
							if (nb + n_next > pool_cap) sh->overflow = 1;
						}
						sh->base[L + 1] = nb; sh->count[L + 1] = n_next;
						sh->next_count[(L + 1) & 1] = 0;
						sh->arrived = 0;
						if (n_next > 0) sh->depth = L + 2;
					}
				}
				sync();
				steps += (tid == 0) ? (unsigned long long)n : 0ull;
				cur ^= 1;
				inv_pow *= 0.5;
			}
			const int D = sh->depth;
			if (sh->overflow) {
				if (tid == 0) *overflow = 1;
				sync();
				continue;          /* the host redoes the block with larger capacities */
			}
			/* the subtree masses are summed bottom-up by SaqPostSumBody (a latency-bound walk
			 * over the levels: 13 % of this kernel's stall samples when it ran here, at two
			 * resident CTAs per SM; on its own it runs at full occupancy) */
			for (int L = tid; L < SAQ_POST_LEVELS; L += nt) {
				lv_base[rl * SAQ_POST_LEVELS + L] = L < D ? sh->base[L] : 0;
				lv_count[rl * SAQ_POST_LEVELS + L] = L < D ? sh->count[L] : 0;
			}
			sync();
		}
		if (tid == 0 && steps) atomic_add_u64_plain(pool_ptr + 1, steps);   /* intervals of this attempt */
	}
};

/* Subtree masses of the trees built by SaqPostBody: one cooperating group per posterior walks
 * its levels from the deepest inner one up to the root, then records the total mass. */
struct SaqPostSumBody {
	SaqTree tree; const long long* lv_base; const int* lv_count; const int64_t* root_id;
	double* norm; int64_t r0;
	template<typename Sync>
	RHFQ_HD void run(int64_t rl, int tid, int nt, Sync sync) const {
		const long long* base = lv_base + rl * SAQ_POST_LEVELS;
		const int* count = lv_count + rl * SAQ_POST_LEVELS;
		int D = 0;
		while (D < SAQ_POST_LEVELS && count[D] > 0) ++D;
		for (int L = D - 2; L >= 0; --L) {
			const int64_t b0 = base[L];
			const int n = count[L];
			/* four nodes per round: their loads are in flight together */
			const int64_t b1 = base[L + 1];         /* children are relative to the next level's range */
			for (int i0 = 4 * tid; i0 < n; i0 += 4 * nt) {
				int64_t c[4];
				double a[4], b[4];
#pragma unroll
				for (int u = 0; u < 4; ++u) c[u] = (i0 + u < n) ? tree.child[b0 + i0 + u] : -1;
#pragma unroll
				for (int u = 0; u < 4; ++u) { a[u] = c[u] >= 0 ? tree.S[b1 + c[u]] : 0.0; b[u] = c[u] >= 0 ? tree.S[b1 + c[u] + 1] : 0.0; }
#pragma unroll
				for (int u = 0; u < 4; ++u) if (c[u] >= 0) tree.S[b0 + i0 + u] = a[u] + b[u];
			}
			sync();
		}
		if (tid == 0) {
			const int64_t id = root_id[r0 + rl];
			norm[r0 + rl] = (D > 0 && tree.child[id] >= 0) ? tree.S[id] : 1.0;
		}
	}
};

/* Level L of a device-controlled build: frontier length and pool position come from the
 * control block; `writer` (one thread of the launch) records where the next level starts. */
RHFQ_HD unsigned long long saq_ctl_level(SaqStepBliBody& b, SaqCtl* ctl, int L, bool writer)
{
	const unsigned long long n = ctl->count[L];
	const unsigned long long nb = ctl->first[L] + n;
	if (writer) ctl->first[L + 1] = nb;
	if (ctl->overflow) return 0;
	b.sp.first = 0;
	b.sp.level_base = (int64_t)ctl->first[L];
	b.sp.next_base = (int64_t)nb;
	b.sp.next_count = &ctl->count[L + 1];
	b.sp.overflow = &ctl->overflow;
	return n;
}

/* subtree masses bottom-up: one launch per level over the level's pool range */
struct SaqSumBody {
	SaqTree tree; int64_t first;
	RHFQ_HD void operator()(int64_t i) const {
		const int64_t c = tree.child[first + i];
		if (c >= 0) tree.S[first + i] = tree.S[c] + tree.S[c + 1];
	}
};

/* Unit-mass normalisation (simpson.hpp:346-355) is applied when the tree is read (the
 * queries divide by the posterior's total mass) instead of rewriting all nodes; a single
 * accepted leaf keeps its un-normalised rule (simpson.hpp:317-318). */
struct SaqNormBody {
	SaqTree tree; const int64_t* root_id; double* norm; int64_t r0;
	RHFQ_HD void operator()(int64_t i) const {
		const int64_t id = root_id[r0 + i];
		norm[r0 + i] = (tree.child[id] >= 0) ? tree.S[id] : 1.0;
	}
};

/* The leaf containing x (first leaf with xmax >= x, simpson.hpp:123-158) and the masses to
 * its left / right. */
struct SaqPath {
	double xl, xr, f1, f2, f3, left, right;
	int64_t j; int depth;     /* the leaf is interval j of the 2^depth intervals of its level */
	double dx;                /* Qmax / 2^depth, the width the build used for the leaf's mass */
};
RHFQ_HD SaqPath saq_descend(const SaqTree& tree, int64_t r, int64_t id, double Qmax, double f1,
                            double f3, double x)
{
	SaqPath p;
	p.xl = 0.0; p.xr = Qmax; p.f1 = f1; p.f3 = f3; p.left = 0.0; p.right = 0.0;
	p.j = 0; p.depth = 0;
	p.f2 = tree.fmid[id];
	double inv_pow = 1.0;
	for (int64_t c = tree.child[id]; c >= 0; c = tree.child[id]) {
		inv_pow *= 0.5;
		++p.depth;
		c = saq_child(tree, r, p.depth, c);
		const double x2 = saq_x(Qmax, 2 * p.j + 1, inv_pow);    /* as the build places it */
		if (x <= x2) {
			p.right += tree.S[c + 1];
			p.xr = x2; p.f3 = p.f2;
			p.j = 2 * p.j;
			id = c;
		} else {
			p.left += tree.S[c];
			p.xl = x2; p.f1 = p.f2;
			p.j = 2 * p.j + 1;
			id = c + 1;
		}
		p.f2 = tree.fmid[id];
	}
	p.dx = Qmax * inv_pow;
	return p;
}

/* The leaf in which the un-normalised backward integral (the mass to the right of x) passes
 * `target`: one descent that compares the mass to the right of every midpoint on the way with
 * the target.  The masses are non-negative (sums of Simpson rules of a non-negative pdf), so
 * the backward integral decreases from node boundary to node boundary. */
RHFQ_HD SaqPath saq_descend_mass(const SaqTree& tree, int64_t r, int64_t id, double Qmax, double f1,
                                 double f3, double target)
{
	SaqPath p;
	p.xl = 0.0; p.xr = Qmax; p.f1 = f1; p.f3 = f3; p.left = 0.0; p.right = 0.0;
	p.j = 0; p.depth = 0;
	p.f2 = tree.fmid[id];
	double inv_pow = 1.0;
	for (int64_t c = tree.child[id]; c >= 0; c = tree.child[id]) {
		inv_pow *= 0.5;
		++p.depth;
		c = saq_child(tree, r, p.depth, c);
		const double x2 = saq_x(Qmax, 2 * p.j + 1, inv_pow);
		const double right_of_mid = p.right + tree.S[c + 1];
		if (right_of_mid <= target) {
			/* the crossing lies at or left of the midpoint */
			p.right = right_of_mid;
			p.xr = x2; p.f3 = p.f2;
			p.j = 2 * p.j;
			id = c;
		} else {
			p.left += tree.S[c];
			p.xl = x2; p.f1 = p.f2;
			p.j = 2 * p.j + 1;
			id = c + 1;
		}
		p.f2 = tree.fmid[id];
	}
	p.dx = Qmax * inv_pow;
	return p;
}

RHFQ_HD double saq_integral(const SaqTree& tree, int64_t r, int64_t root, double Qmax, double f1,
                            double f3, double x, bool backward, double inv_norm)
{
	if (x <= 0.0)
		return backward ? tree.S[root] * inv_norm : 0.0;
	if (x >= Qmax)
		return backward ? 0.0 : tree.S[root] * inv_norm;
	const SaqPath p = saq_descend(tree, r, root, Qmax, f1, f3, x);
	const SimpsonRule rule(p.dx, p.f1, p.f2, p.f3);
	const double part = rule.integral(x - p.xl);
	if (backward)
		return (p.right + rule.I - part) * inv_norm;
	return (p.left + part) * inv_norm;
}

/* simpson.hpp:161-187 */
RHFQ_HD double saq_density(const SaqTree& tree, int64_t r, int64_t root, double Qmax, double f1,
                           double f3, double x, double inv_norm)
{
	if (x <= 0.0 || x >= Qmax)
		return 0.0;
	const SaqPath p = saq_descend(tree, r, root, Qmax, f1, f3, x);
	const SimpsonRule rule(p.dx, p.f1, p.f2, p.f3);
	const double d = x - p.xl;
	return (rule.C0 + d * (rule.C1 + d * rule.C2)) * inv_norm;
}

/* posterior r of point i: x_off[r] <= i < x_off[r + 1] (binary search; empty posteriors skipped) */
struct PointPostBody {
	const int64_t* x_off; int64_t R; int* post;
	RHFQ_HD void operator()(int64_t i) const {
		int64_t lo = 0, hi = R;          /* invariant: x_off[lo] <= i < x_off[hi] */
		while (hi - lo > 1) {
			const int64_t mid = (lo + hi) >> 1;
			if (x_off[mid] <= i) lo = mid; else hi = mid;
		}
		post[i] = (int)lo;
	}
};

struct QueryBody {
	/* mode: 0 cdf, 1 tail, 2 tail quantile, 3 density (adaptive_simpson pdf) */
	int mode; const PostState* P; SaqTree tree; const int64_t* root_id; const double* root_f1;
	const double* root_f3; const double* norm;
	const int* pt_post; const double* x; double* out; int* post_err;
	RHFQ_HD void operator()(int64_t i) const {
		const int r = pt_post[i];
		const PostState& ps = P[r];
		if (ps.status != ST_OK) { out[i] = nan(""); return; }
		const int64_t root = root_id[r];
		const double f1 = root_f1[r], f3 = root_f3[r];
		const double Qmax = ps.Qmax;
		const double xi = x[i];
		const double inv_norm = 1.0 / norm[r];
		if (mode == 0) {
			/* posterior.hpp:168-181 */
			if (xi <= 0.0) out[i] = 0.0;
			else if (xi >= Qmax) out[i] = 1.0;
			else out[i] = saq_integral(tree, r, root, Qmax, f1, f3, xi, false, inv_norm);
		} else if (mode == 1) {
			out[i] = saq_integral(tree, r, root, Qmax, f1, f3, xi, true, inv_norm);
		} else if (mode == 3) {
			if (xi < 0.0 || xi >= Qmax) out[i] = 0.0;
			else out[i] = saq_density(tree, r, root, Qmax, f1, f3, xi, inv_norm);
		} else {
			/* posterior.hpp:279-317 */
			if (xi == 0.0) out[i] = Qmax;
			else if (xi == 1.0) out[i] = 0.0;
			else if (xi > 0.0 && xi < 1.0) {
				const SaqTree tr = tree;
				/* eps_tolerance(digits - 2) -> 2^(1-51), not below 4 eps */
				const double tol = 8.881784197001252e-16;
				/* The reference brackets the root of tail(Qmax - u) - t over u in [0, Qmax]
				 * (TOMS748, ~10 evaluations = ~10 walks down the tree).  One walk that follows
				 * the masses finds the leaf in which the tail passes t; inside that leaf the
				 * tail is the leaf's cubic, evaluated exactly as saq_integral does it, and the
				 * same solver finishes on the leaf's ends without touching the tree again. */
				const SaqPath p = saq_descend_mass(tr, r, root, Qmax, f1, f3, xi * norm[r]);
				const SimpsonRule rule(p.dx, p.f1, p.f2, p.f3);
				auto leaf = [=](double back) -> double {
					const double part = rule.integral((Qmax - back) - p.xl);
					return (p.right + rule.I - part) * inv_norm - xi;
				};
				double lo = Qmax - p.xr, hi = Qmax - p.xl;
				const double flo = (p.xr >= Qmax) ? -xi : leaf(lo);
				const double fhi = (p.xl <= 0.0) ? 1.0 - xi : leaf(hi);
				if (flo <= 0.0 && fhi >= 0.0 && !(flo == 0.0 && fhi == 0.0)) {
					toms748(leaf, lo, hi, flo, fhi, tol, 100);
				} else {
					/* rounding at a node boundary (or a plateau of zero density) left the leaf
					 * without a sign change: the reference's bracket over the whole range */
					auto qf = [=](double back) -> double {
						return saq_integral(tr, r, root, Qmax, f1, f3, Qmax - back, true, inv_norm) - xi;
					};
					lo = 0.0; hi = Qmax;
					toms748(qf, lo, hi, -xi, 1.0 - xi, tol, 100);
				}
				out[i] = Qmax - 0.5 * (lo + hi);
			} else {
				out[i] = nan("");
				atomic_max_i32(&post_err[r], ST_DOMAIN);
			}
		}
	}
};

#ifndef RHFQ_EMU
/* The heavy stages get their own kernel names (readable in ncu / launch lists). */
#define RHFQ_NAMED_KERNEL(Body, kname, BLOCK) RHFQ_NAMED_KERNEL_MB(Body, kname, BLOCK, 1)
#define RHFQ_NAMED_KERNEL_MB(Body, kname, BLOCK, MINB)                                   \
	__global__ void __launch_bounds__(BLOCK, MINB) kname(int64_t n, Body b, Limit lim)   \
	{                                                                                    \
		const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;                \
		if (i < n && lim.live(i)) b(i);                                                  \
	}                                                                                    \
	inline void launch(int64_t n, const Body& b, Stream& st, unsigned long long* launches, \
	                   Limit lim = Limit())                                              \
	{                                                                                    \
		if (n <= 0) return;                                                              \
		if (launches) ++*launches;                                                       \
		const int block = BLOCK;                                                         \
		kname<<<(unsigned)((n + block - 1) / block), block, 0, st.s>>>(n, b, lim);       \
		RHFQ_CUDA_CHECK(cudaGetLastError());                                             \
	}
#ifndef RHFQ_NODE_MINB
#define RHFQ_NODE_MINB 5
#endif
RHFQ_NAMED_KERNEL_MB(NodeEvalBody, rhfq_node_eval, 128, RHFQ_NODE_MINB)
RHFQ_NAMED_KERNEL_MB(NormEvalBody, rhfq_norm_eval, 128, RHFQ_NODE_MINB)
/* resident CTAs per SM (register caps 64 / 64 / 96): measured on the bench batch, see
 * tools/ab_variants.sh -- quad 5 -> 8: 4.30 -> 4.09 ms, plan 5 -> 8: 3.30 -> 3.22 ms,
 * Simpson step 8 -> 10: 7.55 -> 6.85 ms (12: 6.93, 14: 6.76, 16: 7.25); again on the final
 * kernel (32-byte frontier entries, one-wave grid): 8: 6.88, 10: 6.24, 12: 6.06, 13: 6.41,
 * 14: 6.03, 16: 6.02 ms -> 12 (80 registers) */
#ifndef RHFQ_PLAN_MINB
#define RHFQ_PLAN_MINB 8
#endif
#ifndef RHFQ_QUAD_MINB
#define RHFQ_QUAD_MINB 6
#endif
RHFQ_NAMED_KERNEL_MB(NodePrepBody, rhfq_node_prep, 128, RHFQ_PLAN_MINB)
RHFQ_NAMED_KERNEL_MB(NormPrepBody, rhfq_norm_prep, 128, RHFQ_PLAN_MINB)
RHFQ_NAMED_KERNEL_MB(PlanModeBody<NodeTaskMap>, rhfq_node_mode, 128, RHFQ_PLAN_MINB)
RHFQ_NAMED_KERNEL_MB(PlanModeBody<NormTaskMap>, rhfq_norm_mode, 128, RHFQ_PLAN_MINB)
RHFQ_NAMED_KERNEL_MB(PlanWindowBody<NodeTaskMap>, rhfq_node_window, 128, RHFQ_PLAN_MINB)
RHFQ_NAMED_KERNEL_MB(PlanWindowBody<NormTaskMap>, rhfq_norm_window, 128, RHFQ_PLAN_MINB)
RHFQ_NAMED_KERNEL_MB(PlanQuadBody<NodeTaskMap>, rhfq_node_quad, 128, RHFQ_QUAD_MINB)
RHFQ_NAMED_KERNEL_MB(PlanQuadBody<NormTaskMap>, rhfq_norm_quad, 128, RHFQ_QUAD_MINB)
RHFQ_NAMED_KERNEL_MB(PlanQuadCheckBody<NodeTaskMap>, rhfq_node_quad_check, 128, 4)
RHFQ_NAMED_KERNEL_MB(PlanQuadWarpBody<NodeTaskMap>, rhfq_node_quad_warp, 128, 4)
RHFQ_NAMED_KERNEL_MB(PlanQuadWarpBody<NormTaskMap>, rhfq_norm_quad_warp, 128, 4)
RHFQ_NAMED_KERNEL_MB(PlanQuadCheckBody<NormTaskMap>, rhfq_norm_quad_check, 128, 4)
/* one thread per sample set with long serial searches: small blocks so that a
 * batch of 10^4 sets still spreads over all 148 SMs */
RHFQ_NAMED_KERNEL(ScaleBody, rhfq_set_scale, 32)
RHFQ_NAMED_KERNEL(YScanBody, rhfq_set_yscan, 64)
RHFQ_NAMED_KERNEL(YBisBody, rhfq_set_ybis, 64)
RHFQ_NAMED_KERNEL(YMaxBody, rhfq_set_ymax, 32)
RHFQ_NAMED_KERNEL(TaylorBody, rhfq_set_taylor, 32)
RHFQ_NAMED_KERNEL(PdfBliBody, rhfq_pdf_bli, 128)
RHFQ_NAMED_KERNEL(BliErrBody, rhfq_bli_err, 128)
RHFQ_NAMED_KERNEL(SaqDecideBody, rhfq_saq_decide, 128)
#ifndef RHFQ_SAQ_THREADS
#define RHFQ_SAQ_THREADS 64
#endif
#ifndef RHFQ_SAQ_MINB
#define RHFQ_SAQ_MINB 12
#endif
RHFQ_NAMED_KERNEL_MB(SaqStepBliBody, rhfq_saq_step_bli, RHFQ_SAQ_THREADS, RHFQ_SAQ_MINB)
/* The same step with the frontier length read on the device: a grid-stride loop over a grid
 * that is sized from the host's upper bound (capped at a few CTAs per resident slot). */
__global__ void __launch_bounds__(RHFQ_SAQ_THREADS, RHFQ_SAQ_MINB)
rhfq_saq_step_bli_ctl(SaqStepBliBody b, SaqCtl* ctl, int L)
{
	const int64_t n = (int64_t)saq_ctl_level(b, ctl, L, blockIdx.x == 0 && threadIdx.x == 0);
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b(i);
}
inline void launch_saq_ctl(int64_t bound, const SaqStepBliBody& b, SaqCtl* ctl, int L, int max_ctas,
                           Stream& st, unsigned long long* launches, KernelTimer& t)
{
	if (bound <= 0) bound = 1;
	const int block = RHFQ_SAQ_THREADS;
	const int64_t grid = std::min<int64_t>((bound + block - 1) / block, max_ctas);
	cudaEvent_t e0 = t.get(), e1 = t.get();
	RHFQ_CUDA_CHECK(cudaEventRecord(e0, st.s));
	if (launches) ++*launches;
	rhfq_saq_step_bli_ctl<<<(unsigned)grid, block, 0, st.s>>>(b, ctl, L);
	RHFQ_CUDA_CHECK(cudaGetLastError());
	RHFQ_CUDA_CHECK(cudaEventRecord(e1, st.s));
	t.pending.emplace_back(e0, e1);
	++t.count;
}
/* CTA shape (A/B on the B200, profiles/r02_ab_saqp_shape.txt): 6 x 128 threads per SM beat
 * 3 x 256 by 12 % on C3 and 5 % on the resilience trees -- fewer idle lanes in the levels that
 * are narrower than the CTA and in the last round of a level. */
#ifndef RHFQ_SAQP_THREADS
#define RHFQ_SAQP_THREADS 128
#endif
#ifndef RHFQ_SAQP_MINB
#define RHFQ_SAQP_MINB 6
#endif
__global__ void __launch_bounds__(RHFQ_SAQP_THREADS, RHFQ_SAQP_MINB) rhfq_saq_post(SaqPostBody b)
{
	extern __shared__ double s_saqp[];
	SaqPostShared* sh = reinterpret_cast<SaqPostShared*>(s_saqp);
	double* grid = s_saqp + (sizeof(SaqPostShared) + sizeof(double) - 1) / sizeof(double);
	b.run((int)blockIdx.x, (int)threadIdx.x, (int)blockDim.x, sh, grid, CtaSync());
}
inline size_t saq_post_header_doubles() { return (sizeof(SaqPostShared) + sizeof(double) - 1) / sizeof(double); }
inline void launch_saq_post(int n_cta, int threads, const SaqPostBody& b, Stream& st,
                            unsigned long long* launches, KernelTimer& t)
{
	const size_t smem = (saq_post_header_doubles() + (size_t)b.smem_doubles) * sizeof(double);
	static thread_local size_t configured = 0;
	if (smem > 48 * 1024 && smem > configured) {
		RHFQ_CUDA_CHECK(cudaFuncSetAttribute(rhfq_saq_post, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                                     (int)smem));
		configured = smem;
	}
	cudaEvent_t e0 = t.get(), e1 = t.get();
	RHFQ_CUDA_CHECK(cudaEventRecord(e0, st.s));
	if (launches) ++*launches;
	rhfq_saq_post<<<(unsigned)n_cta, threads, smem, st.s>>>(b);
	RHFQ_CUDA_CHECK(cudaGetLastError());
	RHFQ_CUDA_CHECK(cudaEventRecord(e1, st.s));
	t.pending.emplace_back(e0, e1);
	++t.count;
}
__global__ void __launch_bounds__(128) rhfq_saq_sum_post(SaqPostSumBody b)
{
	b.run((int64_t)blockIdx.x, (int)threadIdx.x, (int)blockDim.x, CtaSync());
}
inline void launch_saq_sum_post(int64_t Rb, const SaqPostSumBody& b, Stream& st, unsigned long long* launches)
{
	if (Rb <= 0) return;
	if (launches) ++*launches;
	rhfq_saq_sum_post<<<(unsigned)Rb, 128, 0, st.s>>>(b);
	RHFQ_CUDA_CHECK(cudaGetLastError());
}
/* One CTA per interpolant; the FFT scratch (2 x G doubles) lives in shared memory.  Two
 * instances per body: up to 512 threads (128 registers: the radix-8 butterflies keep their
 * eight complex values, indices and twiddles in registers) and 1024 threads (64 registers). */
template<typename B, int MAXT>
__global__ void __launch_bounds__(MAXT) rhfq_fft_kernel(B b, int cap, const int* live)
{
	extern __shared__ __align__(16) double s_fft[];
	if (live && (int)blockIdx.x >= *live) return;
	b.run((int64_t)blockIdx.x, s_fft, s_fft + cap, (int)threadIdx.x, (int)blockDim.x);
}
template<typename B, int MAXT>
inline void launch_fft_t(int64_t n_interp, const B& b, int cap, int threads, size_t smem, Stream& st,
                         const int* live)
{
	static thread_local size_t configured = 0;
	if (smem > 48 * 1024 && smem > configured) {
		RHFQ_CUDA_CHECK(cudaFuncSetAttribute(rhfq_fft_kernel<B, MAXT>,
		                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		configured = smem;
	}
	rhfq_fft_kernel<B, MAXT><<<(unsigned)n_interp, threads, smem, st.s>>>(b, cap, live);
	RHFQ_CUDA_CHECK(cudaGetLastError());
}
template<typename B>
inline void launch_fft(int64_t n_interp, const B& b, int bits_cap, int threads, Stream& st,
                       unsigned long long* launches, const int* live = nullptr)
{
	const int cap = std::max(1 << bits_cap, threads);
	const size_t smem = 2 * (size_t)cap * sizeof(double);
	if (launches) ++*launches;
	if (threads <= 512) launch_fft_t<B, 512>(n_interp, b, cap, threads, smem, st, live);
	else launch_fft_t<B, 1024>(n_interp, b, cap, threads, smem, st, live);
}
inline void launch_fine_build(int64_t n_interp, FineBuildBody b, Stream& st, unsigned long long* launches)
{
	if (n_interp <= 0) return;
	/* two size classes so that the small transforms are not throttled by the shared
	 * memory of the largest one (2 x 8192 doubles = 128 KB at 7 refinements) */
	const int bits_max = 3 + b.Rmax + 3;
	const int split = bits_max < 11 ? bits_max : 11;
	for (int cls = 0; cls < 2; ++cls) {
		b.bits_lo = cls == 0 ? 0 : split + 1;
		b.bits_hi = cls == 0 ? split : bits_max;
		if (b.bits_lo > b.bits_hi) continue;
		static const int big = [] { const char* e = std::getenv("RHFQ_FFT_BIG_THREADS");
		                            const int v = e ? std::atoi(e) : 0;
		                            return (v >= 128 && v <= 1024 && v % 32 == 0) ? v : 1024; }();
		launch_fft(n_interp, b, b.bits_hi, cls == 0 ? 256 : big, st, launches);
	}
}
inline void launch_bli_coeff(int64_t n_interp, const BliCoeffBody& b, Stream& st, unsigned long long* launches)
{
	if (n_interp <= 0) return;
	launch_fft(n_interp, b, 3 + b.Rmax, 128, st, launches);
}
inline void launch_bli_check(int64_t n_active, const BliCheckBody& b, Stream& st, unsigned long long* launches,
                             const int* live = nullptr)
{
	if (n_active <= 0) return;
	const int bits = 4 + b.level;
	launch_fft(n_active, b, bits, bits >= 9 ? 256 : (bits >= 7 ? 64 : 32), st, launches, live);
}
RHFQ_NAMED_KERNEL(SaqSumBody, rhfq_saq_sum, 256)
RHFQ_NAMED_KERNEL(QueryBody, rhfq_query, 64)
#else
inline void launch_saq_ctl(int64_t, SaqStepBliBody b, SaqCtl* ctl, int L, int, Stream&,
                           unsigned long long* launches, KernelTimer&)
{
	if (launches) ++*launches;
	const int64_t n = (int64_t)saq_ctl_level(b, ctl, L, true);
	for (int64_t i = 0; i < n; ++i) b(i);
}
inline size_t saq_post_header_doubles() { return (sizeof(SaqPostShared) + sizeof(double) - 1) / sizeof(double); }
inline void launch_saq_post(int n_cta, int, const SaqPostBody& b, Stream&, unsigned long long* launches,
                            KernelTimer& t)
{
	if (launches) ++*launches;
	++t.count;
	/* the CTAs one after the other, one "thread" each: the first takes every posterior */
	std::vector<double> grid((size_t)b.smem_doubles + 1);
	for (int c = 0; c < n_cta; ++c) {
		SaqPostShared sh;
		std::memset(&sh, 0, sizeof(sh));
		b.run(c, 0, 1, &sh, grid.data(), CtaSync());
	}
}
inline void launch_saq_sum_post(int64_t Rb, const SaqPostSumBody& b, Stream&, unsigned long long* launches)
{
	if (Rb <= 0) return;
	if (launches) ++*launches;
	for (int64_t r = 0; r < Rb; ++r) b.run(r, 0, 1, CtaSync());
}
inline void launch_fine_build(int64_t n_interp, FineBuildBody b, Stream&, unsigned long long* launches)
{
	if (n_interp <= 0) return;
	if (launches) ++*launches;
	b.bits_lo = 0;
	b.bits_hi = 6 + b.Rmax;
	const size_t cap = (size_t)1 << b.bits_hi;
	std::vector<double> scratch(2 * cap);
	for (int64_t u = 0; u < n_interp; ++u)
		b.run(u, scratch.data(), scratch.data() + cap, 0, 1);
}
inline void launch_bli_check(int64_t n_active, const BliCheckBody& b, Stream&, unsigned long long* launches,
                             const int* live = nullptr)
{
	if (n_active <= 0) return;
	if (launches) ++*launches;
	const size_t cap = (size_t)16 << b.level;
	std::vector<double> scratch(2 * cap);
	if (live && *live < n_active) n_active = *live;
	for (int64_t a = 0; a < n_active; ++a)
		b.run(a, scratch.data(), scratch.data() + cap, 0, 1);
}
inline void launch_bli_coeff(int64_t n_interp, const BliCoeffBody& b, Stream&, unsigned long long* launches)
{
	if (n_interp <= 0) return;
	if (launches) ++*launches;
	const size_t cap = (size_t)8 << b.Rmax;
	std::vector<double> scratch(2 * cap);
	for (int64_t u = 0; u < n_interp; ++u)
		b.run(u, scratch.data(), scratch.data() + cap, 0, 1);
}
#endif

/* ------------------------------------------------------------------ *
 *  Host-side engine                                                   *
 * ------------------------------------------------------------------ */
/* grow-only host array of posterior records; page-locked in the product */
struct HostMirror {
	PostState* p = nullptr;
	size_t n = 0, cap = 0;
	bool pinned = false;
	PostState& operator[](int64_t r) { return p[r]; }
	const PostState& operator[](int64_t r) const { return p[r]; }
	PostState* data() { return p; }
	size_t size() const { return n; }
	void release() {
		if (!p) return;
#ifndef RHFQ_EMU
		if (pinned) cudaFreeHost(p); else
#endif
		std::free(p);
		p = nullptr; n = cap = 0; pinned = false;
	}
	void reserve(size_t want) {
		release();
#ifndef RHFQ_EMU
		void* q = nullptr;
		if (cudaHostAlloc(&q, want * sizeof(PostState), cudaHostAllocPortable) == cudaSuccess) {
			p = (PostState*)q; cap = want; pinned = true;
			return;
		}
		cudaGetLastError();
#endif
		p = (PostState*)std::malloc(want * sizeof(PostState));
		if (!p) throw std::bad_alloc();
		cap = want;
	}
};

class Batch {
public:
	int device = 0;
	Stream st;
	int64_t R = 0, S = 0, total = 0;
	int Mmax = 1;
	double amin = 1.0, rtol = 1e-8;
	/* The reference's thresholds that are powers of the machine epsilon of ITS arithmetic type
	 * (y_max criterion large_z.hpp:314, tanh-sinh tolerance localsandnorm.hpp:210, Simpson
	 * tolerance simpson.hpp:222, PointInInterval intervall.hpp:58).  Default: its double build;
	 * long_double_thresholds() selects those of its long double build -- what the Python keyword
	 * precision="double" runs (postbackend.pyx:219-222) -- with the arithmetic still IEEE double. */
	double eps_real = DBL_EPS, sqrt_eps = SQRT_EPS;
	void long_double_thresholds() { eps_real = 1.0842021724855044e-19; sqrt_eps = 3.2927225399135963e-10; }
	int algorithm = Algo::BLI;
	int Rmax = 7;              /* bli_max_refinements */
	int64_t alloc_pts = 0;      /* allocation size (points) of the current refinement level, >= the launch size */
	bool saq_built = false;
	int64_t saq_block = 16384;  /* posteriors per Simpson-tree block (bounds transient memory) */
	int64_t saq_frontier_per_post = 4096;   /* frontier capacity estimate (RHFQ_SAQ_FRONTIER) */
	bool saq_host_ctl = false;  /* RHFQ_SAQ_HOSTCTL=1: the host reads every level's frontier length */
	int64_t saq_post_frontier = 32768;      /* intervals per CTA-private frontier buffer (a level cannot be wider
	                                         * than the tree has leaves: ~10^4 at rtol 1e-8) */
	bool saq_per_post = true;   /* one posterior per persistent CTA (RHFQ_SAQ_LEVELSYNC=1: level-synchronous launches) */
	int n_sm = 148;
	int64_t saq_nodes_per_post = 20480;     /* pool estimate of a device-controlled block */
	int saq_max_ctas = 148 * 10;            /* grid cap of the device-controlled level launches */

	DBuf<double> q, c, k, w, prior;
	DBuf<int64_t> set_off, post_off;
	DBuf<int> set_post;
	DBuf<SetLocals> L;
	DBuf<double> gco;           /* [S][GCO_N] per-set Stirling coefficient blocks */
	DBuf<PostState> P;
	DBuf<Counters> cnt;
	DBuf<int> post_err, set_err;
	DBuf<double> gtab;          /* lgamma-difference tables, one per slot of the (n, v) hash table */
	DBuf<int> gtab_owner;       /* [GTAB_SLOTS] set that owns the slot (-1: free), [GTAB_SLOTS]: classes */
	bool use_gtab = true;       /* RHFQ_NO_GTAB=1: direct lgamma differences everywhere */
	bool use_fft_check = true;  /* RHFQ_NO_FFT_CHECK=1: barycentric sums in the refinement check */
	int64_t n_gclasses = 0;
	void build_gtab();
	bool use_fine = true;       /* RHFQ_NO_FINE=1: Clenshaw evaluation in the Simpson tree */
	DBuf<double> bli_f;         /* R * 2 * n_fine: log pdf at the nodes */
	DBuf<double> bli_c;         /* R * 2 * n_fine: Chebyshev coefficients of the kept level */
	/* point-evaluation scratch */
	DBuf<int> pt_post;
	DBuf<double> pt_val, pt_fb, vbuf, lpdf;
	/* Simpson tree: node pool of the whole batch + per-posterior roots */
	DBuf<double> tree_fmid, tree_S;
	DBuf<int64_t> tree_child;
	DBuf<int64_t> root_id;
	DBuf<double> root_f1, root_f3;
	DBuf<double> saq_norm;      /* total Simpson mass per posterior */
	DBuf<double> saq_fine, saq_ff[3];   /* fine theta grids of the block being built; CTA-private frontiers */
	DBuf<unsigned> saq_fj;
	DBuf<long long> saq_lv;     /* [R][SAQ_POST_LEVELS] pool index of each tree level's first node (per-posterior build) */
	DBuf<int> saq_lvc;          /* [R][SAQ_POST_LEVELS] nodes per level */
	bool saq_relative = false;  /* child indices are relative to saq_lv (SaqTree::lv) */
	int64_t n_nodes = 0;        /* pool nodes in use */
	int64_t n_leaves = 0;
	unsigned long long saq_steps = 0;   /* frontier intervals processed (host count) */
	void* scan_tmp = nullptr;
	size_t scan_tmp_bytes = 0;
	unsigned long long launches = 0;
	KernelTimer node_timer, norm_timer, saq_timer;   /* node / norm-node / Simpson-step launches */
	KernelTimer plan_timer;                          /* phase A (mode + window) of the node launches */
	bool split_nodes = true;    /* RHFQ_FUSED_NODES=1: one kernel per node evaluation */
	bool quad_check = true;     /* RHFQ_NO_QUAD_CHECK=1: no sampled precision check of the alpha-quadrature */
	bool quad_warp = true;      /* RHFQ_NO_QUAD_WARP=1: one thread per node also for small launches */
	DBuf<double> plan_d[7];
	DBuf<int> plan_flags;
	NodePlanView plan_view(size_t n) {
		for (auto& b : plan_d) b.ensure(n);
		plan_flags.ensure(n);
		return NodePlanView{plan_d[0].p, plan_d[1].p, plan_d[2].p, plan_d[3].p, plan_d[4].p,
		                    plan_d[5].p, plan_d[6].p, plan_flags.p};
	}
	std::vector<int64_t> h_post_off;
	HostMirror h_P;               /* host mirror after create (page-locked, recycled: see sync_status) */
	std::vector<int> h_query_err; /* per-posterior status of the last query */

	~Batch();
	/* swaps the large scratch buffers with a workspace (call before create() and again when
	 * the batch is done: the second call hands the grown buffers back) */
	void exchange(Workspace& w) {
		tree_fmid.swap(w.tree_fmid); tree_S.swap(w.tree_S); tree_child.swap(w.tree_child);
		saq_fine.swap(w.fine); saq_fj.swap(w.fj);
		for (int i = 0; i < 3; ++i) saq_ff[i].swap(w.ff[i]);
		for (int i = 0; i < 7; ++i) plan_d[i].swap(w.plan_d[i]);
		plan_flags.swap(w.plan_flags); pt_post.swap(w.pt_post);
		vbuf.swap(w.vbuf); pt_val.swap(w.pt_val); pt_fb.swap(w.pt_fb); lpdf.swap(w.lpdf);
		bli_f.swap(w.bli_f); bli_c.swap(w.bli_c);
	}

	/* Constant tables (FFT twiddles, Chebyshev nodes) depend on bli_max_refinements only: built
	 * once per (device, Rmax) in the process and kept -- building the 8192 long-double twiddles on
	 * the host and uploading them was 1 ms of the 3 ms a single posterior took to construct. */
	const double* fine_tw_p = nullptr;
	const double* cheb_p = nullptr;
	void ensure_fine_twiddle() {
		if (!fine_tw_p) fine_tw_p = const_table(0);
	}
	const double* const_table(int kind);
	FineTwiddle fine_twiddle() const {
		return FineTwiddle{fine_tw_p, FINE_SIGMA * (8 << Rmax)};
	}
	ChebTable cheb_table() const {
		const int n_fine = (8 << Rmax) + 1;
		return ChebTable{cheb_p, cheb_p + n_fine, cheb_p + 2 * n_fine, sqrt_eps};
	}
	SaqTree tree() const { return SaqTree{tree_fmid.p, tree_S.p, tree_child.p, saq_relative ? saq_lv.p : nullptr}; }
	void tree_reserve(int64_t count);

	void create(const double* hq, const double* hc, const int64_t* hset_off, const double* hw,
	            const int64_t* hpost_off, int64_t R_, const double* hprior, int prior_stride,
	            double amin_, double rtol_, int algorithm_, int bli_max_refinements,
	            bool inputs_on_device, int64_t n_sets = -1, int64_t n_data = -1);
	void eval_points(int64_t n_pts, double* out_lpdf, Limit lim = Limit());   /* pt_post/pt_val/pt_fb filled */
	void build_bli();
	void build_saq();
	void pdf_values(int64_t n, const int* d_post, const double* d_x, double* d_out, bool range_check);
	void query(int mode, int64_t n, const int64_t* x_off, const double* x, double* out,
	           bool on_device);
	void sync_status();
};

inline void make_cheb_table(int Rmax, std::vector<double>& tab)
{
	/* barylagrint.hpp:225-238 */
	const int n = (8 << Rmax) + 1;
	tab.resize(3 * (size_t)n);
	const double pi = 3.14159265358979323846;
	for (int i = 0; i < n; ++i) {
		const double z = std::cos(i * pi / (n - 1));
		const double cha = std::cos(i * pi / (2 * (n - 1)));
		const double sha = std::sin(i * pi / (2 * (n - 1)));
		tab[i] = z;
		tab[n + i] = cha * cha;
		tab[2 * n + i] = sha * sha;
	}
}

/* kind 0: FFT twiddles, 1: Chebyshev node table; cached per (device, Rmax, kind) for the life
 * of the process */
inline const double* Batch::const_table(int kind)
{
	static std::mutex mtx;
	static std::map<std::tuple<int, int, int>, const double*> cache;
	std::lock_guard<std::mutex> lock(mtx);
	const auto key = std::make_tuple(device, Rmax, kind);
	auto it = cache.find(key);
	if (it != cache.end()) return it->second;
	std::vector<double> tab;
	if (kind == 0) make_fine_twiddle(Rmax, tab);
	else make_cheb_table(Rmax, tab);
#ifdef RHFQ_EMU
	double* d = new double[tab.size()];
	std::memcpy(d, tab.data(), tab.size() * sizeof(double));
#else
	double* d = nullptr;
	RHFQ_CUDA_CHECK(cudaMalloc((void**)&d, tab.size() * sizeof(double)));
	RHFQ_CUDA_CHECK(cudaMemcpy(d, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
#endif
	cache[key] = d;
	return d;
}

/* tanh-sinh node table in z: levels 0..LMAX */
struct TanhSinhTable {
	std::vector<double> xc, wt;
	std::vector<signed char> side;
	std::vector<int> level_off;
	static constexpr int LMAX = 12;     /* deepest level (the oracle's restatement stops there too) */
	static constexpr int LCOMMON = 8;   /* levels uploaded up front; deeper ones only when a set needs them */
	TanhSinhTable() {
		const double hpi = HALF_PI;
		level_off.push_back(0);
		for (int lev = 0; lev <= LMAX; ++lev) {
			const double h = std::ldexp(1.0, -lev);
			if (lev == 0) {
				xc.push_back(1.0); wt.push_back(hpi); side.push_back(0);
			}
			const double t0 = h, dt = (lev == 0) ? h : 2 * h;
			for (double t = t0; t <= 6.2; t += dt) {
				const double u = hpi * std::sinh(t);
				const double e2u = std::exp(2 * u);
				const double xcv = 2 / (1 + e2u);
				const double ch = std::cosh(u);
				const double w = hpi * std::cosh(t) / (ch * ch);
				/* Integrands here are bounded by ~1 after rescaling: nodes whose
				 * weight is below 1e-22 cannot change a double-precision sum. */
				if (w < 1e-22 || xcv == 0.0)
					break;
				xc.push_back(xcv); wt.push_back(w); side.push_back(+1);
				xc.push_back(xcv); wt.push_back(w); side.push_back(-1);
			}
			level_off.push_back((int)xc.size());
		}
	}
};

inline void Batch::create(const double* hq, const double* hc, const int64_t* hset_off,
                          const double* hw, const int64_t* hpost_off, int64_t R_,
                          const double* hprior, int prior_stride, double amin_, double rtol_,
                          int algorithm_, int bli_max_refinements, bool inputs_on_device,
                          int64_t n_sets, int64_t n_data)
{
	R = R_;
	amin = amin_; rtol = rtol_; algorithm = algorithm_; Rmax = bli_max_refinements;
	if (R <= 0) throw std::runtime_error("No posteriors given.");
	if (std::getenv("RHFQ_NO_FINE")) use_fine = false;
	if (std::getenv("RHFQ_NO_GTAB")) use_gtab = false;
	if (std::getenv("RHFQ_FUSED_NODES")) split_nodes = false;
	if (std::getenv("RHFQ_NO_QUAD_CHECK")) quad_check = false;
	if (std::getenv("RHFQ_NO_QUAD_WARP")) quad_warp = false;
	{
		const char* e = std::getenv("RHFQ_KERNEL_TIMERS");
		const bool on = e && e[0] == '1';
		node_timer.enabled = norm_timer.enabled = plan_timer.enabled = on;
	}
	if (std::getenv("RHFQ_NO_FFT_CHECK")) use_fft_check = false;
	if (const char* e = std::getenv("RHFQ_SAQ_FRONTIER")) {
		const long long v = std::atoll(e);
		if (v > 0) { saq_frontier_per_post = v; saq_post_frontier = std::max<long long>(v, 8); }
	}
	if (const char* e = std::getenv("RHFQ_SAQ_BLOCK")) {
		const long long v = std::atoll(e);
		if (v > 0) saq_block = v;
	}
	if (std::getenv("RHFQ_SAQ_HOSTCTL")) saq_host_ctl = true;
	if (std::getenv("RHFQ_SAQ_LEVELSYNC")) saq_per_post = false;
	if (const char* e = std::getenv("RHFQ_SAQ_NODES")) {
		const long long v = std::atoll(e);
		if (v > 0) saq_nodes_per_post = v;
	}
#ifndef RHFQ_EMU
	{
		n_sm = 0;
		/* CTAs per resident slot of the grid-stride level kernel (RHFQ_SAQ_WAVES); measured on
		 * the bench batch: 1 -> 6.24 ms, 2 -> 6.32, 4 -> 6.34, 16 -> 6.32, 64 -> 6.77 */
		int waves = 1;
		if (const char* e = std::getenv("RHFQ_SAQ_WAVES")) {
			const int v = std::atoi(e);
			if (v > 0) waves = v;
		}
		if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && n_sm > 0)
			saq_max_ctas = n_sm * RHFQ_SAQ_MINB * waves;
		else
			n_sm = 148;
	}
#endif
	if (Rmax < 0 || Rmax > 12) throw std::runtime_error("bli_max_refinements out of range [0,12].");

	/* offsets are needed on the host for sizing */
	h_post_off.resize(R + 1);
	if (inputs_on_device) d2h(h_post_off.data(), hpost_off, (R + 1) * sizeof(int64_t), st);
	else std::memcpy(h_post_off.data(), hpost_off, (R + 1) * sizeof(int64_t));
	/* The offsets decide how far every input array is read: they must start at 0, be
	 * non-decreasing and (when the caller declared the array lengths, as the C ABI does) end
	 * at them.  Checked before they index anything. */
	auto check_offsets = [](const std::vector<int64_t>& off, int64_t declared, const char* what) {
		bool ok = off.front() == 0 && (declared < 0 || off.back() == declared);
		for (size_t i = 0; ok && i + 1 < off.size(); ++i) ok = off[i + 1] >= off[i];
		if (!ok)
			throw std::runtime_error(std::string(what) + " must start at 0, be non-decreasing and "
			                         "end at the declared length.");
	};
	check_offsets(h_post_off, n_sets, "post_off");
	S = h_post_off[R];
	if (S <= 0) throw std::runtime_error("No samples given.");
	std::vector<int64_t> h_set_off(S + 1);
	if (inputs_on_device) d2h(h_set_off.data(), hset_off, (S + 1) * sizeof(int64_t), st);
	else std::memcpy(h_set_off.data(), hset_off, (S + 1) * sizeof(int64_t));
	check_offsets(h_set_off, n_data, "set_off");
	total = h_set_off[S];
	std::vector<int> h_set_post(S);
	Mmax = 1;
	for (int64_t r = 0; r < R; ++r) {
		const int64_t M = h_post_off[r + 1] - h_post_off[r];
		if (M > Mmax) Mmax = (int)M;
		for (int64_t s = h_post_off[r]; s < h_post_off[r + 1]; ++s) h_set_post[s] = (int)r;
	}

	q.alloc(total); c.alloc(total); k.alloc(total); w.alloc(S);
	set_off.alloc(S + 1); post_off.alloc(R + 1); set_post.alloc(S);
	const int64_t n_prior = (prior_stride == 0) ? 4 : 4 * R;
	prior.alloc(n_prior);
	L.alloc(S); P.alloc(R); cnt.alloc(1); post_err.alloc(R); set_err.alloc(S);
	gco.alloc((size_t)S * GCO_N);
	if (inputs_on_device) {
		d2d(q.p, hq, total * sizeof(double), st); d2d(c.p, hc, total * sizeof(double), st);
		d2d(w.p, hw, S * sizeof(double), st);
		d2d(prior.p, hprior, n_prior * sizeof(double), st);
	} else {
		h2d(q.p, hq, total * sizeof(double), st); h2d(c.p, hc, total * sizeof(double), st);
		h2d(w.p, hw, S * sizeof(double), st);
		h2d(prior.p, hprior, n_prior * sizeof(double), st);
	}
	h2d(set_off.p, h_set_off.data(), (S + 1) * sizeof(int64_t), st);
	h2d(post_off.p, h_post_off.data(), (R + 1) * sizeof(int64_t), st);
	h2d(set_post.p, h_set_post.data(), S * sizeof(int), st);
	dzero(cnt.p, sizeof(Counters), st);
	dzero(post_err.p, R * sizeof(int), st);
	dzero(set_err.p, S * sizeof(int), st);

	Trace tr;
	tr.stage("h2d inputs", st, total);
	/* 1. locals, 2. log_scale + y_max */
	{
		DBuf<double> lgq;
		lgq.alloc(total);
		launch(S * REDUCE_LANES, LocalsBody{q.p, c.p, set_off.p, set_post.p, prior.p, prior_stride,
		                                    amin, k.p, L.p, w.p, lgq.p,
		                                    std::getenv("RHFQ_LOCALS_SEQ") ? 1 : 0,
		                                    std::getenv("RHFQ_NO_GCO") ? nullptr : gco.p,
		                                    std::log(eps_real), sqrt_eps}, st, &launches);
	}
	tr.stage("locals", st, S);
	if (use_gtab) {
		build_gtab();
		tr.stage("lgamma tables", st, n_gclasses);
	}
	{
		DBuf<double> yT, yTb;
		yT.alloc((size_t)S * 4 * YSCAN_C); yTb.alloc((size_t)S * 4 * YBIS_C);
		launch(S, ScaleBody{L.p, k.p}, st, &launches);
		launch(S * 2 * YSCAN_C, YScanBody{L.p, yT.p}, st, &launches);
		launch(S * 2 * YBIS_C, YBisBody{L.p, yT.p, yTb.p}, st, &launches);
		launch(S * 2, YMaxBody{L.p, yT.p, yTb.p}, st, &launches);
	}
	tr.stage("setup (log_scale, ymax)", st, S);

	/* 3. norm: tanh-sinh in z over [0, ztrans], level doubling */
	{
		static const TanhSinhTable tab;
		const int n_all = (int)tab.xc.size();
		DBuf<double> d_xc, d_wt, fbuf, acc;
		DBuf<signed char> d_side;
		DBuf<int> d_loff, active, active2, conv, counter;
		d_xc.alloc(n_all); d_wt.alloc(n_all); d_side.alloc(n_all); d_loff.alloc(tab.level_off.size());
		/* the levels nearly every set converges within; the deeper ones (5x the bytes) follow
		 * only if a set asks for them */
		int n_up = tab.level_off[TanhSinhTable::LCOMMON + 1];
		h2d(d_xc.p, tab.xc.data(), n_up * sizeof(double), st);
		h2d(d_wt.p, tab.wt.data(), n_up * sizeof(double), st);
		h2d(d_side.p, tab.side.data(), n_up, st);
		h2d(d_loff.p, tab.level_off.data(), tab.level_off.size() * sizeof(int), st);
		acc.alloc(3 * S); dzero(acc.p, 3 * S * sizeof(double), st);
		active.alloc(S); active2.alloc(S); conv.alloc(S); counter.alloc(2);
		/* active = sets with status OK.  As in the interpolant loop (build_bli) the number of
		 * still-active sets stays on the device and the host sizes the launches of a level
		 * group from the count of the group before the last (CountMirror). */
		DBuf<int> flag, idx, nflag;
		flag.alloc(S); idx.alloc(S); nflag.alloc(S);
		launch(S, StatusFlagBody{L.p, flag.p, idx.p}, st, &launches);
		dzero(counter.p, 2 * sizeof(int), st);
		launch(S, CompactBody{idx.p, flag.p, active.p, counter.p}, st, &launches);
		CountMirror mirror(1);
		mirror.record(0, counter.p, st);
		int64_t bound = S;
		int lev_a = 0, lev_b = 3, g = 0;
		for (; lev_a <= TanhSinhTable::LMAX; ++g) {
			if (g >= 1) {
				bound = std::min<int64_t>(bound, mirror.get(g - 1));
				if (bound <= 0) break;
			}
			const int* live = counter.p + (g & 1);
			int* next = counter.p + ((g + 1) & 1);
			const int node0 = tab.level_off[lev_a];
			const int n_nodes = tab.level_off[lev_b + 1] - node0;
			if (node0 + n_nodes > n_up) {
				h2d(d_xc.p + n_up, tab.xc.data() + n_up, (n_all - n_up) * sizeof(double), st);
				h2d(d_wt.p + n_up, tab.wt.data() + n_up, (n_all - n_up) * sizeof(double), st);
				h2d(d_side.p + n_up, tab.side.data() + n_up, n_all - n_up, st);
				n_up = n_all;
			}
			const int64_t nt = bound * n_nodes;
			const Limit nlim{live, n_nodes, 0};
			const int64_t nt_alloc = nt;
			fbuf.ensure((size_t)nt_alloc);
			if (split_nodes) {
				const NodePlanView pv = plan_view((size_t)nt_alloc);
				const NormTaskMap map{active.p, n_nodes, fbuf.p, set_err.p};
				launch_timed(nt, NormPrepBody{L.p, k.p, active.p, n_nodes, node0, d_xc.p, d_side.p,
				                              fbuf.p, cnt.p, pv}, st, &launches, plan_timer, nlim);
				launch_timed(nt, PlanModeBody<NormTaskMap>{L.p, map, cnt.p, pv}, st, &launches, plan_timer, nlim);
				launch_timed(nt, PlanWindowBody<NormTaskMap>{L.p, map, cnt.p, pv}, st, &launches, plan_timer, nlim);
				if (nt < QUAD_WARP_BELOW && quad_warp)
					launch_timed(nt * REDUCE_LANES, PlanQuadWarpBody<NormTaskMap>{L.p, map, cnt.p, pv}, st,
					             &launches, norm_timer, Limit{live, (int64_t)n_nodes * REDUCE_LANES, 0});
				else
				launch_timed(nt, PlanQuadBody<NormTaskMap>{L.p, map, cnt.p, pv}, st, &launches, norm_timer, nlim);
				if (quad_check)
					launch(quad_check_tasks(nt), PlanQuadCheckBody<NormTaskMap>{L.p, map, cnt.p, pv, nt, nlim}, st, &launches);
			} else
			launch_timed(nt,
			       NormEvalBody{L.p, k.p, active.p, n_nodes, node0, d_xc.p, d_wt.p, d_side.p,
			                    fbuf.p, cnt.p, set_err.p}, st, &launches, norm_timer, nlim);
			launch(bound * REDUCE_LANES,
			       NormReduceBody{L.p, active.p, fbuf.p, n_nodes, d_wt.p, d_loff.p, lev_a, lev_b,
			                      acc.p, conv.p, set_err.p}, st, &launches, Limit{live, REDUCE_LANES, 0});
			/* keep the unconverged */
			launch(bound, NotBody{conv.p, nflag.p}, st, &launches, Limit{live, 1, 0});
			dzero(next, sizeof(int), st);
			launch(bound, CompactBody{active.p, nflag.p, active2.p, next}, st, &launches, Limit{live, 1, 0});
			mirror.record(g + 1, next, st);
			std::swap(active.p, active2.p);
			std::swap(active.n, active2.n);
			if (tr.on)
				std::fprintf(stderr, "[rhfq]   norm levels %d-%d: %d nodes x <= %lld sets\n",
				             lev_a, lev_b, n_nodes, (long long)bound);
			lev_a = lev_b + 1;
			lev_b = lev_a;
		}
		if (lev_a > TanhSinhTable::LMAX && bound > 0) {
			/* Not converged at the deepest level.  The reference takes whatever its tanh-sinh
			 * rule returns (localsandnorm.hpp:221-226: no check of `error`; Boost stops at its
			 * refinement cap without raising), so the last estimate is kept here too; the set
			 * is flagged (SET_Z_UNCONVERGED, rhfq_batch_get_locals / counters) so that callers
			 * CAN see it.  RHFQ_Z_STRICT=1 turns it into status PRECISION. */
			const Limit fl{counter.p + (g & 1), 1, 0};
			if (std::getenv("RHFQ_Z_STRICT"))
				launch(bound, MarkListBody{active.p, set_err.p, (int)ST_PRECISION}, st, &launches, fl);
			else
				launch(bound, FlagListBody{active.p, L.p, (int)SET_Z_UNCONVERGED, cnt.p}, st, &launches, fl);
		}
		launch(S, MarkBody{L.p, set_err.p}, st, &launches);
		dsync(st);
	}
	tr.stage("norm quadrature", st, S);
	launch(S, TaylorBody{L.p, cnt.p}, st, &launches);
	launch(R, PostBody{L.p, post_off.p, w.p, P.p}, st, &launches);
	tr.stage("taylor + post", st, R);

	cheb_p = const_table(1);          /* Chebyshev-Lobatto node table */
	if (algorithm == Algo::BLI)
		build_bli();
	else if (algorithm == Algo::SIMPSON)
		build_saq();
	tr.stage("pdf algorithm setup", st, R);
	sync_status();
}

/* The host mirror of the posterior records is page-locked and recycled from batch to batch
 * (per host thread): a fresh 4.7 MB vector for the 65 520 posteriors of a resilience chunk cost
 * 2.1 ms of page faults while the device sat idle, and going through the 4 MB staging buffer
 * another 0.35 ms of host memcpy between the two pieces (CUPTI timeline). */
inline std::vector<HostMirror>& host_mirror_pool()
{
	static thread_local struct Pool {
		std::vector<HostMirror> v;
		~Pool() { for (auto& m : v) m.release(); }
	} pool;
	return pool.v;
}
inline Batch::~Batch()
{
	dfree(scan_tmp);
	auto& pool = host_mirror_pool();
	if (h_P.cap && pool.size() < 4) { pool.push_back(h_P); h_P = HostMirror(); }
	h_P.release();
}
inline void Batch::sync_status()
{
	if ((int64_t)h_P.cap < R) {
		auto& pool = host_mirror_pool();
		int best = -1;
		for (int i = 0; i < (int)pool.size(); ++i) {
			const bool fits = (int64_t)pool[i].cap >= R;
			if (best < 0) { best = i; continue; }
			const bool best_fits = (int64_t)pool[best].cap >= R;
			if (fits ? (!best_fits || pool[i].cap < pool[best].cap)
			         : (!best_fits && pool[i].cap > pool[best].cap))
				best = i;
		}
		if (best >= 0) {
			h_P.release();
			h_P = pool[best];
			pool.erase(pool.begin() + best);
		}
		if ((int64_t)h_P.cap < R) h_P.reserve((size_t)R + (size_t)R / 8);
	}
	h_P.n = (size_t)R;
	if (h_P.pinned) {
		d2h_async(h_P.p, P.p, R * sizeof(PostState), st);
		dsync(st);
	} else {
		d2h(h_P.p, P.p, R * sizeof(PostState), st);
	}
}

/* Evaluates the log pdf at the n_pts points stored in pt_post / pt_val / pt_fb. */
inline void Batch::eval_points(int64_t n_pts, double* out_lpdf, Limit lim)
{
	/* lim: only the points p < *lim.count * lim.per are live (the launch is sized from a bound) */
	if (n_pts <= 0) return;
	const int64_t n_alloc = std::max<int64_t>(n_pts, alloc_pts);
	vbuf.ensure((size_t)n_alloc * Mmax);
	lim.period = lim.count ? n_pts : 0;        /* tasks are (set m, point p) = m * n_pts + p */
	if (split_nodes) {
		const NodePlanView pv = plan_view((size_t)n_alloc * Mmax);
		const int64_t nt = n_pts * Mmax;
		const NodeTaskMap map{post_off.p, pt_post.p, n_pts, vbuf.p, post_err.p};
		launch_timed(nt, NodePrepBody{L.p, k.p, post_off.p, P.p, pt_post.p, pt_val.p, pt_fb.p, n_pts,
		                              Mmax, vbuf.p, post_err.p, cnt.p, pv}, st, &launches, plan_timer, lim);
		launch_timed(nt, PlanModeBody<NodeTaskMap>{L.p, map, cnt.p, pv}, st, &launches, plan_timer, lim);
		launch_timed(nt, PlanWindowBody<NodeTaskMap>{L.p, map, cnt.p, pv}, st, &launches, plan_timer, lim);
		if (nt < QUAD_WARP_BELOW && quad_warp) {
			Limit wl = lim;          /* tasks are (node, lane): scale the live range by the lanes */
			wl.per *= REDUCE_LANES; wl.period *= REDUCE_LANES;
			launch_timed(nt * REDUCE_LANES, PlanQuadWarpBody<NodeTaskMap>{L.p, map, cnt.p, pv}, st, &launches,
			             node_timer, wl);
		} else {
			launch_timed(nt, PlanQuadBody<NodeTaskMap>{L.p, map, cnt.p, pv}, st, &launches, node_timer, lim);
		}
		if (quad_check) {
			/* the sampled tasks t = s * STRIDE must be live themselves */
			Limit cl;
			launch(quad_check_tasks(nt), PlanQuadCheckBody<NodeTaskMap>{L.p, map, cnt.p, pv, nt, lim}, st,
			       &launches, cl);
		}
	} else
	launch_timed(n_pts * Mmax,
	       NodeEvalBody{L.p, k.p, post_off.p, P.p, pt_post.p, pt_val.p, pt_fb.p, n_pts, Mmax,
	                    vbuf.p, post_err.p, cnt.p}, st, &launches, node_timer, lim);
	Limit pl = lim;
	pl.period = 0;
	launch(n_pts, CombineBody{post_off.p, pt_post.p, n_pts, vbuf.p, out_lpdf}, st, &launches, pl);
}

/* units (interpolants) of healthy posteriors, in any order */
struct BliActiveInitBody {
	const PostState* P; int* active; int* counter;
	RHFQ_HD void operator()(int64_t u) const {
		if (P[u >> 1].status == ST_OK) active[atomic_add_i32(counter, 1)] = (int)u;
	}
};

inline void Batch::build_bli()
{
	const int n_fine = (8 << Rmax) + 1;
	bli_f.alloc((size_t)R * 2 * n_fine);
	DBuf<int> active, active2, flag, counts;
	DBuf<double> err_abs, err_rel, cscratch;
	ensure_fine_twiddle();
	active.alloc(2 * R); active2.alloc(2 * R); flag.alloc(2 * R);
	/* The number of still-active interpolants stays on the device: counts[level & 1] is what
	 * the launches of a level read through their Limit.  The host sizes every launch from a
	 * bound -- the count of two levels ago, mirrored with a lag (CountMirror) -- so that it
	 * never waits for the level it has just queued. */
	counts.alloc(2);
	dzero(counts.p, 2 * sizeof(int), st);
	launch(2 * R, BliActiveInitBody{P.p, active.p, counts.p}, st, &launches);
	CountMirror mirror;
	mirror.record(0, counts.p, st);
	int64_t bound = 2 * R;
	const ChebTable T = cheb_table();

	Trace tr;
	for (int level = 0; level <= Rmax; ++level) {
		if (level >= 2) {
			bound = std::min<int64_t>(bound, mirror.get(level - 1));   /* units active after level - 2 */
			if (bound <= 0) break;
		}
		const int* live = counts.p + (level & 1);
		int* next = counts.p + ((level + 1) & 1);
		const int n_new = (level == 0) ? 9 : (8 << (level - 1));
		const int64_t n_pts = bound * n_new;
		const Limit lim{live, n_new, 0};
		tr.stage("  bli level begin", st, n_pts);
		alloc_pts = n_pts;
		pt_post.ensure(alloc_pts); pt_val.ensure(alloc_pts); pt_fb.ensure(alloc_pts); lpdf.ensure(alloc_pts);
		launch(n_pts, BliPointsBody{P.p, active.p, level, Rmax, n_new, T, pt_post.p, pt_val.p, pt_fb.p},
		       st, &launches, lim);
		eval_points(n_pts, lpdf.p, lim);
		launch(n_pts, BliStoreBody{active.p, level, Rmax, n_new, lpdf.p, bli_f.p, post_err.p, pt_post.p},
		       st, &launches, lim);
		if (level == 0) {
			/* every unit goes on to level 1: the count carries over */
			d2d(next, live, sizeof(int), st);
			mirror.record(1, next, st);
			continue;
		}
		/* check level-1 interpolant against the new nodes of `level` */
		if (use_fft_check) {
			cscratch.ensure((size_t)bound * ((8 << (level - 1)) + 1));
			launch_bli_check(bound, BliCheckBody{active.p, level - 1, Rmax, fine_twiddle(), bli_f.p,
			                                     cscratch.p, rtol, P.p, flag.p}, st, &launches, live);
		} else {
			err_abs.ensure(alloc_pts); err_rel.ensure(alloc_pts);
			launch(n_pts, BliErrBody{active.p, level - 1, Rmax, n_new, T, bli_f.p, err_abs.p, err_rel.p},
			       st, &launches, lim);
			launch(bound, BliDecideBody{active.p, level - 1, n_new, err_abs.p, err_rel.p, rtol, P.p, flag.p},
			       st, &launches, Limit{live, 1, 0});
		}
		dzero(next, sizeof(int), st);
		launch(bound, CompactBody{active.p, flag.p, active2.p, next}, st, &launches, Limit{live, 1, 0});
		mirror.record(level + 1, next, st);
		std::swap(active.p, active2.p);
		std::swap(active.n, active2.n);
	}
	alloc_pts = 0;
	/* whoever is still active keeps the finest level (max_refinements reached) */
	if (bound > 0)
		launch(bound, BliFinishBody{active.p, Rmax, P.p}, st, &launches,
		       Limit{counts.p + ((Rmax + 1) & 1), 1, 0});
	launch(R, PostErrBody{P.p, post_err.p}, st, &launches);
	tr.stage("  bli refinement done", st, R);
	bli_c.alloc((size_t)R * 2 * n_fine);
	ensure_fine_twiddle();
	launch_bli_coeff(2 * R, BliCoeffBody{P.p, bli_f.p, bli_c.p, Rmax, fine_twiddle()}, st, &launches);
	tr.stage("  bli coefficients", st, R);
	dsync(st);      /* the mirror's events and the scratch of this scope die here */
}

inline void Batch::pdf_values(int64_t n, const int* d_post, const double* d_x, double* d_out,
                              bool range_check)
{
	if (n <= 0) return;
	if (algorithm == Algo::BLI) {
		launch(n, PdfBliBody{P.p, bli_f.p, bli_c.p, Rmax, cheb_table(), d_post, d_x, d_out, range_check ? 1 : 0},
		       st, &launches);
	} else {
		/* explicit evaluation: x as PointInInterval(x, x, Qmax - x) */
		pt_val.ensure(n); pt_fb.ensure(n); lpdf.ensure(n);
		if (pt_post.p != d_post) { pt_post.ensure(n); d2d(pt_post.p, d_post, n * sizeof(int), st); }
		launch(n, FillBody{P.p, pt_post.p, d_x, pt_val.p, pt_fb.p}, st, &launches);
		eval_points(n, lpdf.p);
		launch(n, ExpRangeBody{P.p, pt_post.p, d_x, lpdf.p, d_out}, st, &launches);
	}
}

inline void Batch::build_gtab()
{
	/* distinct (n, v) of the batch: they are prior + data count, so a batch of 10^4 regions
	 * with a shared prior has < 100 of them */
	gtab_owner.alloc(GTAB_SLOTS + 1);
	gtab.alloc((size_t)GTAB_SLOTS * GTAB_STRIDE);
#ifndef RHFQ_EMU
	RHFQ_CUDA_CHECK(cudaMemsetAsync(gtab_owner.p, 0xff, GTAB_SLOTS * sizeof(int), st.s));
#else
	std::memset(gtab_owner.p, 0xff, GTAB_SLOTS * sizeof(int));
#endif
	dzero(gtab_owner.p + GTAB_SLOTS, sizeof(int), st);          /* number of classes (diagnostics) */
	launch(S, GtabClaimBody{L.p, gtab_owner.p, gtab.p, gtab_owner.p + GTAB_SLOTS}, st, &launches);
	launch((int64_t)GTAB_SLOTS * GTAB_SIZE, GtabBuildBody{L.p, gtab_owner.p, gtab.p}, st, &launches);
	n_gclasses = -1;       /* on the device (gtab_owner[GTAB_SLOTS]) */
}

inline void Batch::tree_reserve(int64_t count)
{
	if ((size_t)count <= tree_fmid.n) return;
	const size_t cap = std::max<size_t>((size_t)count, tree_fmid.n + tree_fmid.n / 2);
	DBuf<double> nf, ns;
	DBuf<int64_t> nc;
	nf.alloc(cap); ns.alloc(cap); nc.alloc(cap);
	if (n_nodes > 0) {
		d2d(nf.p, tree_fmid.p, n_nodes * sizeof(double), st);
		d2d(ns.p, tree_S.p, n_nodes * sizeof(double), st);
		d2d(nc.p, tree_child.p, n_nodes * sizeof(int64_t), st);
	}
	std::swap(nf.p, tree_fmid.p); std::swap(nf.n, tree_fmid.n);
	std::swap(ns.p, tree_S.p); std::swap(ns.n, tree_S.n);
	std::swap(nc.p, tree_child.p); std::swap(nc.n, tree_child.n);
	/* the old buffers are released in stream order after the copies */
}

inline void Batch::build_saq()
{
	if (saq_built) return;
	struct Frontier {
		DBuf<double> d[3];
		DBuf<unsigned long long> key;
		int64_t n = 0;
		int64_t capacity() const { return (int64_t)key.n; }
		/* contents are dead when this is called: no copy */
		void reserve(int64_t cap) {
			if ((size_t)cap <= key.n) return;
			for (auto& b : d) b.alloc(cap);
			key.alloc(cap);
		}
		/* rare: the frontier outgrew the estimate while it is being filled */
		void grow(int64_t used, Stream& st) {
			const size_t cap = 2 * key.n + 1024;
			for (auto& b : d) {
				DBuf<double> nb; nb.alloc(cap);
				d2d(nb.p, b.p, used * sizeof(double), st);
				std::swap(nb.p, b.p); std::swap(nb.n, b.n);
			}
			{ DBuf<unsigned long long> nb; nb.alloc(cap);
			  d2d(nb.p, key.p, used * sizeof(unsigned long long), st);
			  std::swap(nb.p, key.p); std::swap(nb.n, key.n); }
		}
		SaqFrontier view() {
			return SaqFrontier{d[0].p, d[1].p, d[2].p, key.p, n};
		}
	};
	DBuf<int> sp_post;
	DBuf<double> sp_x, fv;
	DBuf<unsigned long long> next_count;
	DBuf<SaqCtl> ctl;
	next_count.alloc(1);
	ctl.alloc(1);
	Frontier front[2];          /* ping-pong, kept across levels and blocks */
	Trace tr;
	const int max_levels = 24, min_levels = 5;
	root_id.alloc(R); root_f1.alloc(R); root_f3.alloc(R);
	saq_norm.alloc(R);
	n_nodes = 0;
	/* ~1.4e4 nodes per posterior at rtol 1e-8; grown on demand */
	tree_reserve(std::min<int64_t>(R, saq_block)
	             * ((algorithm == Algo::BLI && !saq_host_ctl) ? saq_nodes_per_post : 16384));
	tr.stage("  saq pool reserve", st, (long long)tree_fmid.n);

	/* Posteriors are processed in blocks, which bounds the transient memory (frontiers
	 * and fine grids); the pool itself (24 B per node) holds the whole batch. */
	for (int64_t r0 = 0; r0 < R; r0 += saq_block) {
		const int64_t Rb = std::min<int64_t>(saq_block, R - r0);

		/* fine theta grids of this block's interpolants (transient) */
		struct { double* p; } fine{nullptr};      /* null: no fine grids (Clenshaw sums) */
		DBuf<int> fine_bad;
		if (algorithm == Algo::BLI && use_fine) {
			ensure_fine_twiddle();
			saq_fine.alloc((size_t)(2 * Rb) * FineBuildBody::stride(Rmax));
			fine.p = saq_fine.p;
			fine_bad.alloc(2 * Rb);
			dzero(fine_bad.p, 2 * Rb * sizeof(int), st);
			launch_fine_build(2 * Rb, FineBuildBody{P.p, bli_c.p, Rmax, fine_twiddle(), fine.p,
			                                        fine_bad.p, r0, 0, 0, cnt.p}, st, &launches);
			tr.stage("  saq fine grids", st, 2 * Rb);
		}

		/* root intervals (simpson.hpp:243-256) */
		Frontier* cur = &front[0];
		Frontier* nxt = &front[1];
		const bool per_post = algorithm == Algo::BLI && !saq_host_ctl && saq_per_post;
		if (!per_post) {
			/* the widest level holds ~3200 intervals per posterior at rtol 1e-8 */
			cur->reserve(std::max<int64_t>(saq_frontier_per_post * Rb, 64));
			nxt->reserve(std::max<int64_t>(saq_frontier_per_post * Rb, 64));
		}
		sp_post.ensure(3 * Rb); sp_x.ensure(3 * Rb); fv.ensure(3 * Rb);
		launch(3 * Rb, SaqRootPointsBody{P.p, sp_post.p, sp_x.p, r0}, st, &launches);
		pdf_values(3 * Rb, sp_post.p, sp_x.p, fv.p, false);
		std::vector<int64_t> level_first;   /* pool index of each level's first node */

		if (per_post) {
			/* One posterior per persistent CTA (SaqPostBody): a single launch per block. */
			const int n_fine = (8 << Rmax) + 1;
			(void)n_fine;
			/* shared memory: the largest pair of fine grids of the block up to the cap below
			 * (what does not fit is read from global memory through L1) */
			int need = 0;
			for (int64_t r = r0; r < r0 + Rb; ++r) {
				if (h_P[r].status != ST_OK) continue;
				const int g0 = FINE_SIGMA * (8 << std::max(0, h_P[r].level[0])) + 1;
				const int g1 = FINE_SIGMA * (8 << std::max(0, h_P[r].level[1])) + 1;
				need = std::max(need, g0 + g1);
			}
			/* 20 KB: grids of up to 5 refinements are staged; the larger ones are read through
			 * L1 / L2 (staging them made no difference: 4.98 ms at 10, 20 and 37 KB, 5.04 at 1) */
			int smem_cap = (int)((size_t)std::min(224 / RHFQ_SAQP_MINB_HOST, 20) * 1024 / sizeof(double)) - (int)saq_post_header_doubles();
			if (const char* e = std::getenv("RHFQ_SAQP_SMEM_KB")) {
				const int v = std::atoi(e);
				if (v > 0) smem_cap = (int)((size_t)v * 1024 / sizeof(double)) - (int)saq_post_header_doubles();
			}
			const int smem_doubles = use_fine ? std::max(16, std::min(need, smem_cap)) : 16;
			int ctas_per_sm = RHFQ_SAQP_MINB_HOST;
			if (const char* e = std::getenv("RHFQ_SAQP_CTAS")) {
				const int v = std::atoi(e);
				if (v > 0) ctas_per_sm = v;
			}
			int threads = RHFQ_SAQP_THREADS_HOST;
			if (const char* e = std::getenv("RHFQ_SAQP_THREADS")) {
				const int v = std::atoi(e);
				if (v >= 32 && v <= RHFQ_SAQP_THREADS_HOST && v % 32 == 0) threads = v;
			}
			const int n_cta = (int)std::min<int64_t>(Rb, (int64_t)n_sm * ctas_per_sm);
			DBuf<double>* const ff = saq_ff;
			DBuf<unsigned>& fj = saq_fj;
			DBuf<unsigned long long> pool_ptr;
			DBuf<int> ctl2;
			pool_ptr.alloc(2); ctl2.alloc(2);
			if (!saq_lv.p) { saq_lv.alloc((size_t)R * SAQ_POST_LEVELS); saq_lvc.alloc((size_t)R * SAQ_POST_LEVELS); }
			saq_relative = true;
			long long* const lv_base_p = saq_lv.p + r0 * SAQ_POST_LEVELS;
			int* const lv_count_p = saq_lvc.p + r0 * SAQ_POST_LEVELS;
			for (;;) {
				const int64_t cap = saq_post_frontier;
				for (int i = 0; i < 3; ++i) ff[i].ensure((size_t)2 * n_cta * cap);
				fj.ensure((size_t)2 * n_cta * cap);
				tree_reserve(n_nodes + Rb * saq_nodes_per_post);
				const unsigned long long h_pool[2] = {(unsigned long long)n_nodes, 0ull};
				h2d(pool_ptr.p, h_pool, sizeof(h_pool), st);
				dzero(ctl2.p, 2 * sizeof(int), st);
				SaqSplit sp{SaqFrontier{}, SaqFrontier{}, tree(), nullptr, 0, 0, 0, max_levels, min_levels};
				sp.tol = sqrt_eps;
				SaqStepBliBody sb{sp, P.p, bli_f.p, bli_c.p, Rmax, cheb_table(), cnt.p, fine.p, fine_bad.p, r0};
				launch_saq_post(n_cta, threads, SaqPostBody{sb, fv.p, Rb, root_f1.p, root_f3.p, root_id.p, saq_norm.p,
				                                   pool_ptr.p, (int64_t)tree_fmid.n, ff[0].p, ff[1].p, ff[2].p,
				                                   fj.p, cap, ctl2.p, ctl2.p + 1, max_levels, min_levels,
				                                   smem_doubles, lv_base_p, lv_count_p}, st, &launches, saq_timer);
				struct { unsigned long long pool[2]; int ctl[2]; } back;
				d2h(back.pool, pool_ptr.p, 2 * sizeof(unsigned long long), st);
				d2h(back.ctl, ctl2.p, 2 * sizeof(int), st);
				if (back.ctl[1]) {
					saq_post_frontier *= 2;
					saq_nodes_per_post *= 2;
					if (tr.on) std::fprintf(stderr, "[rhfq]   saq block overflow: redo\n");
					continue;
				}
				const int64_t block_nodes = (int64_t)back.pool[0] - n_nodes;
				n_nodes = (int64_t)back.pool[0];
				saq_steps += back.pool[1];
				launch_saq_sum_post(Rb, SaqPostSumBody{tree(), lv_base_p, lv_count_p, root_id.p, saq_norm.p, r0},
				                    st, &launches);
				if (r0 + Rb < R) {
					const int64_t per_post = block_nodes / Rb + 1;
					saq_nodes_per_post = std::max<int64_t>(per_post + per_post / 4, 1024);
					tree_reserve(n_nodes + (R - r0 - Rb) * saq_nodes_per_post);
				}
				break;
			}
			tr.stage("  saq per-posterior CTAs", st, n_nodes);
			continue;          /* sums and norms were done by the CTAs */
		} else if (algorithm == Algo::BLI && !saq_host_ctl) {
			/* Device-controlled levels: every launch reads its frontier length and pool
			 * position from the control block, so all levels of the block are queued without
			 * a read-back.  Frontier and pool are sized from estimates; a claim beyond them
			 * raises the overflow flag and the block is redone with doubled capacities. */
			for (;;) {
				cur = &front[0]; nxt = &front[1];
				tree_reserve(n_nodes + Rb * saq_nodes_per_post);
				SaqCtl h_ctl;
				std::memset(&h_ctl, 0, sizeof(h_ctl));
				h_ctl.count[0] = (unsigned long long)Rb;
				h_ctl.first[0] = (unsigned long long)n_nodes;
				h2d(ctl.p, &h_ctl, sizeof(h_ctl), st);
				launch(Rb, SaqRootBody{P.p, fv.p, cur->view(), tree(), r0, n_nodes, root_f1.p, root_f3.p,
				                       root_id.p}, st, &launches);
				for (int L = 0; L < max_levels; ++L) {
					const int64_t bound = std::min<int64_t>(cur->capacity(), L < 40 ? (Rb << L) : cur->capacity());
					SaqSplit sp{cur->view(), nxt->view(), tree(), nullptr, 0, 0, L + 1, max_levels, min_levels};
					sp.tol = sqrt_eps;
					sp.nx_cap = nxt->capacity();
					sp.pool_cap = (int64_t)tree_fmid.n;
					sp.inv_pow = std::ldexp(1.0, -L);
					launch_saq_ctl(bound, SaqStepBliBody{sp, P.p, bli_f.p, bli_c.p, Rmax, cheb_table(),
					                                     cnt.p, fine.p, fine_bad.p, r0}, ctl.p, L,
					               saq_max_ctas, st, &launches, saq_timer);
					std::swap(cur, nxt);
				}
				d2h(&h_ctl, ctl.p, sizeof(h_ctl), st);
				if (h_ctl.overflow) {
					saq_frontier_per_post *= 2;
					saq_nodes_per_post *= 2;
					front[0].reserve(std::max<int64_t>(saq_frontier_per_post * Rb, 64));
					front[1].reserve(std::max<int64_t>(saq_frontier_per_post * Rb, 64));
					if (tr.on) std::fprintf(stderr, "[rhfq]   saq block overflow: redo\n");
					continue;
				}
				int64_t block_nodes = 0;
				for (int L = 0; L < max_levels && h_ctl.count[L] > 0; ++L) {
					level_first.push_back((int64_t)h_ctl.first[L]);
					saq_steps += h_ctl.count[L];
					block_nodes += (int64_t)h_ctl.count[L];
				}
				n_nodes += block_nodes;
				level_first.push_back(n_nodes);
				/* later blocks: pool sized from what this one needed */
				if (r0 + Rb < R) {
					const int64_t per_post = block_nodes / Rb + 1;
					saq_nodes_per_post = std::max<int64_t>(per_post + per_post / 4, 1024);
					tree_reserve(n_nodes + (R - r0 - Rb) * saq_nodes_per_post);
				}
				break;
			}
		} else {
		cur->n = Rb;
		tree_reserve(n_nodes + Rb);
		launch(Rb, SaqRootBody{P.p, fv.p, cur->view(), tree(), r0, n_nodes, root_f1.p, root_f3.p,
		                       root_id.p}, st, &launches);
		level_first.push_back(n_nodes);
		n_nodes += Rb;

		/* bisection: per tree level one launch over the frontier of the block (several if
		 * the next frontier, which must be sized before its length is known, could overflow) */
		for (int L = 0; L < max_levels + 1 && cur->n > 0; ++L) {
			const int64_t n = cur->n;
			dzero(next_count.p, sizeof(unsigned long long), st);
			int64_t used = 0;
			for (int64_t a = 0; a < n;) {
				int64_t m = std::min<int64_t>(n - a, (nxt->capacity() - used) / 2);
				if (m <= 0) {
					nxt->grow(used, st);
					continue;
				}
				SaqSplit sp{cur->view(), nxt->view(), tree(), next_count.p, n_nodes, a, L + 1,
				            max_levels, min_levels};
				sp.tol = sqrt_eps;
				sp.level_base = level_first.back();
				sp.inv_pow = std::ldexp(1.0, -L);
				if (algorithm == Algo::BLI) {
					launch_timed(m, SaqStepBliBody{sp, P.p, bli_f.p, bli_c.p, Rmax, cheb_table(),
					                               cnt.p, fine.p, fine_bad.p, r0}, st, &launches,
					             saq_timer);
				} else {
					/* explicit pdf: the evaluation is the node kernel over all sample sets */
					sp_post.ensure(2 * (size_t)m); sp_x.ensure(2 * (size_t)m); fv.ensure(2 * (size_t)m);
					launch(2 * m, SaqPointsBody{cur->view(), a, P.p, L, sp_post.p, sp_x.p}, st, &launches);
					pdf_values(2 * m, sp_post.p, sp_x.p, fv.p, false);
					launch(m, SaqDecideBody{sp, P.p, fv.p}, st, &launches);
				}
				unsigned long long h_used = 0;
				d2h(&h_used, next_count.p, sizeof(h_used), st);
				used = (int64_t)h_used;
				a += m;
			}
			saq_steps += (unsigned long long)n;
			level_first.push_back(n_nodes);
			tree_reserve(n_nodes + used);   /* copies the n_nodes in use if it grows */
			n_nodes += used;
			nxt->n = used;
			std::swap(cur, nxt);
		}
		level_first.push_back(n_nodes);
		}
		tr.stage("  saq levels", st, n_nodes);

		/* subtree masses bottom-up (the last level holds leaves only) */
		const int D = (int)level_first.size() - 1;
		for (int L = D - 2; L >= 0; --L)
			launch(level_first[L + 1] - level_first[L], SaqSumBody{tree(), level_first[L]}, st,
			       &launches);
		launch(Rb, SaqNormBody{tree(), root_id.p, saq_norm.p, r0}, st, &launches);
		tr.stage("  saq block", st, n_nodes);
	}

	/* every inner node has two children: leaves = (nodes + roots) / 2 */
	n_leaves = (n_nodes + R) / 2;
	launch(R, PostErrBody{P.p, post_err.p}, st, &launches);
	dsync(st);
	saq_built = true;
	sync_status();
	tr.stage("  saq finalize", st, n_nodes);
}

/*
 * Ragged queries: x_off[R+1] gives posterior r the points x[x_off[r]..x_off[r+1]).
 * mode: 0 cdf, 1 tail, 2 tail quantiles, 4 pdf.
 */
inline void Batch::query(int mode, int64_t n, const int64_t* x_off, const double* x, double* out,
                         bool on_device)
{
	if (n <= 0) return;
	/* posterior of every point from the offsets, on the device (a host loop over the points
	 * and its upload were 6 of the 8 ms of a 2.7 M-quantile resilience query) */
	DBuf<int64_t> doff;
	const int64_t* d_off = x_off;
	int64_t last = 0;
	if (on_device) {
		d2h(&last, x_off + R, sizeof(int64_t), st);
	} else {
		last = x_off[R];
		doff.alloc(R + 1);
		h2d(doff.p, x_off, (R + 1) * sizeof(int64_t), st);
		d_off = doff.p;
	}
	if (last != n) throw std::runtime_error("x_off[R] does not match the number of points.");
	DBuf<int> qpost;
	DBuf<double> qx, qout;
	qpost.alloc(n);
	launch(n, PointPostBody{d_off, R, qpost.p}, st, &launches);
	const double* dx = x;
	double* dout = out;
	if (!on_device) {
		qx.alloc(n); qout.alloc(n);
		h2d(qx.p, x, n * sizeof(double), st);
		dx = qx.p; dout = qout.p;
	}
	dzero(post_err.p, R * sizeof(int), st);
	if (mode == 4) {
		if (algorithm == Algo::SIMPSON) {
			build_saq();
			launch(n, QueryBody{3, P.p, tree(), root_id.p, root_f1.p, root_f3.p, saq_norm.p, qpost.p, dx, dout, post_err.p}, st, &launches);
		} else {
			pdf_values(n, qpost.p, dx, dout, true);
		}
	} else {
		build_saq();
		launch(n, QueryBody{mode, P.p, tree(), root_id.p, root_f1.p, root_f3.p, saq_norm.p, qpost.p, dx, dout, post_err.p},
		       st, &launches);
	}
	h_query_err.resize(R);
	d2h(h_query_err.data(), post_err.p, R * sizeof(int), st);
	if (!on_device)
		d2h(out, dout, n * sizeof(double), st);
}

/* a Workspace is a bundle of (pointer, capacity) pairs: swapping two of them swaps the bytes */
struct WorkspaceBits { unsigned char b[sizeof(Workspace)]; };

/* The large scratch of the most recently destroyed batch is kept per device and adopted by the
 * next batch created there (Workspace): a caller that builds one batch after the other --
 * the bench loop, a sweep over regions -- otherwise returns and re-requests multi-GB blocks
 * every time, and the driver re-maps pool memory at random (stalls of 0.2-0.9 s were measured).
 * All work of a batch has completed when it is destroyed (its stream is synchronised), so the
 * buffers can be handed to a batch on any stream.  rhfq_cache_release() frees them. */
struct WorkspaceCache {
	std::mutex mtx;
	Workspace* slot[16] = {};
	bool enabled = std::getenv("RHFQ_NO_CACHE") == nullptr;
	void adopt(Batch& b, int device) {
		if (!enabled || device < 0 || device >= 16) return;
		std::lock_guard<std::mutex> lock(mtx);
		if (slot[device]) b.exchange(*slot[device]);
	}
	void stash(Batch& b, int device) {
		if (!enabled || device < 0 || device >= 16) return;
		std::lock_guard<std::mutex> lock(mtx);
		if (!slot[device]) slot[device] = new Workspace();
		b.exchange(*slot[device]);      /* the dying batch takes the older buffers with it */
	}
	/* the same for a bare workspace (the resilience loop owns one across its chunks) */
	void adopt(Workspace& w, int device) {
		if (!enabled || device < 0 || device >= 16) return;
		std::lock_guard<std::mutex> lock(mtx);
		if (slot[device]) { std::swap(*reinterpret_cast<WorkspaceBits*>(&w), *reinterpret_cast<WorkspaceBits*>(slot[device])); }
	}
	void stash(Workspace& w, int device) {
		if (!enabled || device < 0 || device >= 16) return;
		std::lock_guard<std::mutex> lock(mtx);
		if (!slot[device]) slot[device] = new Workspace();
		std::swap(*reinterpret_cast<WorkspaceBits*>(&w), *reinterpret_cast<WorkspaceBits*>(slot[device]));
	}
	void release(int device) {
		std::lock_guard<std::mutex> lock(mtx);
		for (int d = 0; d < 16; ++d)
			if (slot[d] && (device < 0 || d == device)) { delete slot[d]; slot[d] = nullptr; }
	}
};
inline WorkspaceCache& workspace_cache()
{
	static WorkspaceCache* c = new WorkspaceCache();      /* never destroyed: CUDA may be gone at exit */
	return *c;
}


} // namespace rhfq
