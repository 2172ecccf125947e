#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t ld32(const uint32_t *p) { uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p)); return r; }
__device__ __forceinline__ void st64(int2 *p, int2 v) { asm volatile("st.global.L1::no_allocate.v2.s32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory"); }
// per warp-iteration: SPL samples per lane (block of 32*SPL consecutive samples), PF: prefetch next block
template <int SPL, int PF> __global__ void __launch_bounds__(1024) k(const uint32_t *__restrict__ ph, int2 *__restrict__ xy, size_t n) {
	const size_t nblk = n / (32 * SPL), nw = (size_t)gridDim.x * blockDim.x / 32; const unsigned lane = threadIdx.x & 31;
	size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
	uint32_t v[SPL], w[SPL];
	if (PF && b < nblk) {
#pragma unroll
		for (int k = 0; k < SPL; k++) w[k] = ld32(ph + b * (32 * SPL) + (k << 5) + lane);
	}
	for (; b < nblk; b += nw) {
		if (PF) {
#pragma unroll
			for (int k = 0; k < SPL; k++) v[k] = w[k];
			if (b + nw < nblk) {
#pragma unroll
				for (int k = 0; k < SPL; k++) w[k] = ld32(ph + (b + nw) * (32 * SPL) + (k << 5) + lane);
			}
		} else {
#pragma unroll
			for (int k = 0; k < SPL; k++) v[k] = ld32(ph + b * (32 * SPL) + (k << 5) + lane);
		}
#pragma unroll
		for (int k = 0; k < SPL; k++) st64(xy + b * (32 * SPL) + (k << 5) + lane, make_int2((int)v[k], (int)~v[k]));
	}
}
template <int SPL, int PF> void run(const uint32_t *in, int2 *out, size_t n, int threads) {
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int w = 0; w < 3; w++) k<SPL,PF><<<148, threads>>>(in, out, n);
	float best = 1e9f;
	for (int r = 0; r < 7; r++) { cudaEventRecord(e0); k<SPL,PF><<<148, threads>>>(in, out, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
	printf("samples/lane %d prefetch %d threads %4d: %7.3f ms  %7.1f GB/s\n", SPL, PF, threads, best, n * 12.0 / (best * 1e-3) / 1e9);
}
int main() {
	const size_t n = (size_t)1 << 30;
	uint32_t *in; int2 *out; cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 8); cudaMemset(in, 1, n * 4);
	for (int t : {1024, 512}) {
		run<1,0>(in,out,n,t); run<1,1>(in,out,n,t); run<2,0>(in,out,n,t); run<2,1>(in,out,n,t); run<4,0>(in,out,n,t); run<4,1>(in,out,n,t); run<8,1>(in,out,n,t);
	}
	return 0;
}
