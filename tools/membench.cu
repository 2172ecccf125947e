// tools/membench.cu -- what HBM delivers for the traffic mixes of the CORDIC kernels (not part of the product):
// copy 1:1 (the MEASURED_PEAKS figure), 4 B read + 8 B written per sample (rotation, constant vector), write-only 8 B
// (NCO), 8 B read + 8 B written (vectoring).  Streaming accesses, no arithmetic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu && ./membench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t ld32(const uint32_t *p) { uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p)); return r; }
__device__ __forceinline__ int4 ld128(const int4 *p) { int4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ void st64(int2 *p, int2 v) { asm volatile("st.global.L1::no_allocate.v2.s32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void st128(int4 *p, int4 v) { asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

// MODE 0: 16 B in -> 16 B out (copy).  1: 16 B in -> 32 B out (1:2).  2: 32 B out (write only).  3: 32 B in -> 32 B out.
// 4: the seeded kernel's access shape: per warp 128 samples, lane reads 4 x 4 B (stride 128 B), writes 4 x 8 B
template <int MODE> __global__ void __launch_bounds__(1024) k(const int4 *__restrict__ in, int4 *__restrict__ out, size_t n16) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	if (MODE == 4) {
		const uint32_t *ph = (const uint32_t *)in; int2 *xy = (int2 *)out;
		const size_t nblk = n16 / 32, nw = stride / 32; const unsigned lane = threadIdx.x & 31;
		for (size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 32; b < nblk; b += nw) {
			uint32_t v[4];
#pragma unroll
			for (int k = 0; k < 4; k++) v[k] = ld32(ph + (b << 7) + (k << 5) + lane);
#pragma unroll
			for (int k = 0; k < 4; k++) st64(xy + (b << 7) + (k << 5) + lane, make_int2((int)v[k], (int)~v[k]));
		}
		return;
	}
	if (MODE == 5) {	// per warp 128 samples: lane reads 2 x 8 B (two samples), writes 2 x 16 B
		const int2 *ph = (const int2 *)in;
		const size_t nblk = n16 / 32, nw = stride / 32; const unsigned lane = threadIdx.x & 31;
		for (size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 32; b < nblk; b += nw) {
			int2 v[2];
#pragma unroll
			for (int k = 0; k < 2; k++) asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(v[k].x), "=r"(v[k].y) : "l"(ph + (b << 6) + (k << 5) + lane));
#pragma unroll
			for (int k = 0; k < 2; k++) st128(out + (b << 6) + (k << 5) + lane, make_int4(v[k].x, ~v[k].x, v[k].y, ~v[k].y));
		}
		return;
	}
	if (MODE == 6) {	// as the kernel's shape, 8 samples per lane per iteration (256-sample blocks)
		const uint32_t *ph = (const uint32_t *)in; int2 *xy = (int2 *)out;
		const size_t nblk = n16 / 64, nw = stride / 32; const unsigned lane = threadIdx.x & 31;
		for (size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 32; b < nblk; b += nw) {
			uint32_t v[8];
#pragma unroll
			for (int k = 0; k < 8; k++) v[k] = ld32(ph + (b << 8) + (k << 5) + lane);
#pragma unroll
			for (int k = 0; k < 8; k++) st64(xy + (b << 8) + (k << 5) + lane, make_int2((int)v[k], (int)~v[k]));
		}
		return;
	}
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
		if (MODE == 0) st128(out + i, ld128(in + i));
		if (MODE == 1) { const int4 v = ld128(in + i); st128(out + 2 * i, make_int4(v.x, ~v.x, v.y, ~v.y)); st128(out + 2 * i + 1, make_int4(v.z, ~v.z, v.w, ~v.w)); }
		if (MODE == 2) { const int t = (int)i; st128(out + 2 * i, make_int4(t, t, t, t)); st128(out + 2 * i + 1, make_int4(t, t, t, t)); }
		if (MODE == 3) { const int4 a = ld128(in + 2 * i), b = ld128(in + 2 * i + 1); st128(out + 2 * i, b); st128(out + 2 * i + 1, a); }
	}
}

template <int MODE> void run(const char *name, const int4 *in, int4 *out, size_t n16, double bytes_per_16, int grid, int block) {
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int w = 0; w < 3; w++) k<MODE><<<grid, block>>>(in, out, n16);
	float best = 1e9f;
	for (int r = 0; r < 5; r++) {
		cudaEventRecord(e0); k<MODE><<<grid, block>>>(in, out, n16); cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
	}
	printf("%-34s grid %5d x %4d  %7.3f ms  %7.1f GB/s\n", name, grid, block, best, n16 * bytes_per_16 / (best * 1e-3) / 1e9);
}

int main() {
	const size_t n16 = (size_t)1 << 28;		// 2^28 16-byte units = 2^30 phases
	int4 *in, *out; cudaMalloc(&in, n16 * 32); cudaMalloc(&out, n16 * 32);
	cudaMemset(in, 1, n16 * 32);
	for (int cfg = 0; cfg < 2; cfg++) {
		const int grid = cfg ? 148 : 148 * 8, block = cfg ? 1024 : 256;
		run<0>("copy 16 B -> 16 B", in, out, n16, 32, grid, block);
		run<1>("4 B in + 8 B out (v4)", in, out, n16, 48, grid, block);
		run<4>("4 B in + 8 B out (kernel's shape)", in, out, n16, 48, grid, block);
		run<5>("4 B in + 8 B out (8 B ld, 16 B st)", in, out, n16, 48, grid, block);
		run<6>("4 B in + 8 B out (shape, 8/lane)", in, out, n16, 48, grid, block);
		run<2>("8 B out only", in, out, n16, 32, grid, block);
		run<3>("8 B in + 8 B out", in, out, n16, 64, grid, block);
	}
	return 0;
}
