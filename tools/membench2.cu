#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int LD> __device__ __forceinline__ uint32_t ld32(const uint32_t *p) { uint32_t r;
  if (LD==0) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  if (LD==1) asm volatile("ld.global.nc.L1::evict_first.u32 %0, [%1];" : "=r"(r) : "l"(p));
  if (LD==2) asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(r) : "l"(p));
  if (LD==3) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r; }
template <int ST> __device__ __forceinline__ void st64(int2 *p, int2 v) {
  if (ST==0) asm volatile("st.global.L1::no_allocate.v2.s32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
  if (ST==1) asm volatile("st.global.cs.v2.s32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
  if (ST==2) asm volatile("st.global.cg.v2.s32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
  if (ST==3) asm volatile("st.global.wt.v2.s32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
  if (ST==4) asm volatile("st.global.v2.s32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
template <int LD, int ST, int PF> __global__ void __launch_bounds__(1024) k(const uint32_t *__restrict__ ph, int2 *__restrict__ xy, size_t n) {
	const size_t nblk = n / 128, nw = (size_t)gridDim.x * blockDim.x / 32; const unsigned lane = threadIdx.x & 31;
	size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
	uint32_t v[4], w[4];
	if (PF) { if (b < nblk) for (int k = 0; k < 4; k++) w[k] = ld32<LD>(ph + (b << 7) + (k << 5) + lane); }
	for (; b < nblk; b += nw) {
		if (PF) {
#pragma unroll
			for (int k = 0; k < 4; k++) v[k] = w[k];
			if (b + nw < nblk) {
#pragma unroll
			for (int k = 0; k < 4; k++) w[k] = ld32<LD>(ph + ((b + nw) << 7) + (k << 5) + lane); }
		} else {
#pragma unroll
			for (int k = 0; k < 4; k++) v[k] = ld32<LD>(ph + (b << 7) + (k << 5) + lane);
		}
#pragma unroll
		for (int k = 0; k < 4; k++) st64<ST>(xy + (b << 7) + (k << 5) + lane, make_int2((int)v[k], (int)~v[k]));
	}
}
template <int LD, int ST, int PF> void run(const char *name, const uint32_t *in, int2 *out, size_t n) {
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int w = 0; w < 3; w++) k<LD,ST,PF><<<148, 1024>>>(in, out, n);
	float best = 1e9f, sum = 0;
	for (int r = 0; r < 7; r++) { cudaEventRecord(e0); k<LD,ST,PF><<<148, 1024>>>(in, out, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; sum += ms; }
	printf("%-40s best %7.3f ms  avg %7.3f ms  %7.1f GB/s\n", name, best, sum / 7, n * 12.0 / (best * 1e-3) / 1e9);
	cudaError_t e = cudaGetLastError(); if (e) printf("  err %s\n", cudaGetErrorString(e));
}
int main() {
	const size_t n = (size_t)1 << 30;
	uint32_t *in; int2 *out; cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 8); cudaMemset(in, 1, n * 4);
	run<0,0,0>("ld nc.noalloc / st noalloc", in, out, n);
	run<0,0,1>("same + 1-block prefetch", in, out, n);
	run<0,1,1>("st.cs (prefetch)", in, out, n);
	run<0,2,1>("st.cg (prefetch)", in, out, n);
	run<0,3,1>("st.wt (prefetch)", in, out, n);
	run<0,4,1>("st default (prefetch)", in, out, n);
	run<1,0,1>("ld L1::evict_first / st noalloc (pf)", in, out, n);
	run<1,2,1>("ld L1::evict_first / st.cg (pf)", in, out, n);
	run<2,1,1>("ld.cs / st.cs (pf)", in, out, n);
	run<3,0,1>("ld L2::256B / st noalloc (pf)", in, out, n);
	return 0;
}
