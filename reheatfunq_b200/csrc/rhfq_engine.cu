/*
 * reheatfunq_b200 — product translation unit: sm_100a kernels + C ABI.
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo (see setup in
 * __graft_entry__.build()).  FP64 CUDA-core work only: nothing on this path is a
 * dense contraction, so there is no tensor-core / TMA code here by design.
 */
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <thrust/iterator/reverse_iterator.h>
#include "rhfq_engine.inl"
#include "rhfq_resilience.inl"
#include "rhfq_gcp.inl"
#include "rhfq_coverings.inl"
#include "rhfq_blitest.inl"

#define RHFQ_API(name) rhfq_##name
#include "rhfq_capi.inl"

/* ------------------------------------------------------------------ *
 *  FP64 peak microbenchmark (roofline denominator)                    *
 * ------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) rhfq_dfma_chain(double* out, int iters, double seed)
{
	/* 8 independent register-resident FMA chains per thread */
	double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
	double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const double m = 0.999999, c = 1e-9;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int u = 0; u < 16; ++u) {
			a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
			a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
		}
	}
	out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int rhfq_measure_fp64_peak(int device, double* tflops, char* err, size_t errlen)
{
	int rc = select_device(device, err, errlen);
	if (rc) return rc;
	return guarded(err, errlen, [&]() -> int {
		cudaDeviceProp prop;
		RHFQ_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
		const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
		double* d = nullptr;
		RHFQ_CUDA_CHECK(cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
		cudaEvent_t e0, e1;
		RHFQ_CUDA_CHECK(cudaEventCreate(&e0));
		RHFQ_CUDA_CHECK(cudaEventCreate(&e1));
		double best = 0.0;
		for (int rep = 0; rep < 5; ++rep) {
			RHFQ_CUDA_CHECK(cudaEventRecord(e0));
			rhfq_dfma_chain<<<blocks, threads>>>(d, iters, 1.0);
			RHFQ_CUDA_CHECK(cudaEventRecord(e1));
			RHFQ_CUDA_CHECK(cudaEventSynchronize(e1));
			float ms = 0;
			RHFQ_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
			const double flop = 2.0 * 8 * 16 * (double)iters * blocks * threads;
			const double tf = flop / (ms * 1e-3) / 1e12;
			if (rep > 0 && tf > best) best = tf;
		}
		cudaEventDestroy(e0);
		cudaEventDestroy(e1);
		cudaFree(d);
		*tflops = best;
		return 0;
	});
}
