// tools/ubench4.cu -- vectoring-mode stage formulations on sm_100a, compute only (not part of the product).
// VERDICT r1 "Next" #6: is an all-FP32 stage (SURVEY App. D varD: fma.rm for the floor shifts, the light FMA pipe) faster
// than the integer stage k_topolar uses?  Every variant runs the complete cfg2 chain (IW16 OW16 WW24 PW24, 21 stages,
// rtl/topolar.v:217-243) on the same pseudo-random in-range vectors, 4 independent samples per thread, and folds
// (mag, phase) into a checksum; variants must agree word for word.  Timed: warp-level chains per second.
//   V_INT   the product's stage: SHF md, IADD3 s, IMAD.MOV -s, 2 SHF, 3 IMAD (8 slots); short form from stage 11 (6 slots)
//   V_FP    all 21 stages in FP32: x kept as X = x + 2^23 (ulp 1), y and the phase delta as plain integer-valued floats
//             Xs = copysign(X, y)                      LOP3      |  w  = fma.rz(Xs, 2^-k, sM) = s*(2^23 + floor(X/2^k))   FFMA
//             sM = copysign(2^23, y)                   LOP3      |  y' = fma(sM, 1 + 2^-k, y - w)                    FADD + FFMA
//             X' = |fma.rm(y, 2^-k, Xs)|               FFMA      |  ph' = fma(sM, a_k/2^23, ph)                           FFMA
//           7 slots, 5 of them on the FMA pipes (x' = x + |floor(y/2^k)| because s*floor(y/2^k) is never negative)
//   V_MIX   FP32 for stages 0-10, the integer short form for stages 11-20
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NST 21
#define TAIL0 11
struct K { int pa[NST]; int na[NST]; float fa[NST]; float ck[NST]; float tk[NST]; unsigned e_phase[4]; };

__device__ __forceinline__ int imad(int a, int b, int c) { int r; asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ int ineg(int a) { int r; asm("neg.s32 %0, %1;" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ float fma_rm(float a, float b, float c) { float r; asm("fma.rm.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fma_rz(float a, float b, float c) { float r; asm("fma.rz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float copysign_bits(float mag, float sgn) {
	return __uint_as_float((__float_as_uint(mag) & 0x7fffffffu) | (__float_as_uint(sgn) & 0x80000000u));	// one LOP3
}

template <int S> __device__ __forceinline__ void int_full(int &x, int &y, unsigned &ph, int pa) {
	const int md = y >> 31, s = md + md + 1, ns = ineg(s);
	const int sy = y >> S, sx = x >> S;
	const int x1 = imad(sy, s, x), y1 = imad(sx, ns, y);
	ph = (unsigned)imad(s, pa, (int)ph); x = x1; y = y1;
}
template <int S> __device__ __forceinline__ void int_tail(int &x, int &y, unsigned &ph, int na) {
	const int md = y >> 31, ns = imad(md, -2, -1), sx = x >> S;
	y = imad(sx, ns, y); x = x - md; ph = (unsigned)imad(ns, na, (int)ph);
}
template <int S> __device__ __forceinline__ void fp_stage(float &X, float &y, float &ph, float tk, float ck, float fa) {
	const float Xs = copysign_bits(X, y), sM = copysign_bits(8388608.0f, y);
	const float Xn = fma_rm(y, tk, Xs);
	const float w = fma_rz(Xs, tk, sM);
	y = __fmaf_rn(sM, ck, y - w);
	ph = __fmaf_rn(sM, fa, ph);
	X = Xn;			// the sign rides along; copysign_bits() ignores it and fma_rz/fma_rm only see Xs
}
template <int K0, int K1> struct FP { static __device__ __forceinline__ void run(float &X, float &y, float &ph, const K &c) {
	fp_stage<K0 + 1>(X, y, ph, c.tk[K0], c.ck[K0], c.fa[K0]); FP<K0 + 1, K1>::run(X, y, ph, c); } };
template <int K1> struct FP<K1, K1> { static __device__ __forceinline__ void run(float &, float &, float &, const K &) {} };
template <int K0, int K1> struct IF { static __device__ __forceinline__ void run(int &x, int &y, unsigned &ph, const K &c) {
	int_full<K0 + 1>(x, y, ph, c.pa[K0]); IF<K0 + 1, K1>::run(x, y, ph, c); } };
template <int K1> struct IF<K1, K1> { static __device__ __forceinline__ void run(int &, int &, unsigned &, const K &) {} };
template <int K0, int K1> struct IT { static __device__ __forceinline__ void run(int &x, int &y, unsigned &ph, const K &c) {
	int_tail<K0 + 1>(x, y, ph, c.na[K0]); IT<K0 + 1, K1>::run(x, y, ph, c); } };
template <int K1> struct IT<K1, K1> { static __device__ __forceinline__ void run(int &, int &, unsigned &, const K &) {} };

enum { V_INT = 0, V_FP = 1, V_MIX = 2 };
template <int V> __global__ void __launch_bounds__(256) kchain(unsigned *out, int iters, const __grid_constant__ K c) {
	unsigned seed = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	unsigned acc = 0;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int s = 0; s < 4; s++) {
			seed = seed * 1664525u + 1013904223u;
			const int ix = (int)(short)(seed >> 16), iy = (int)(short)(seed * 2246822519u >> 13);
			// rtl/topolar.v:83-84,122-152 for IW16/WW24: ex = ix << 6, +-45 degree turn by the input signs
			const int ex = ix << 6, ey = iy << 6;
			const int xn = ex >> 31, yn = ey >> 31, sum = ex + ey, dif = ex - ey;
			int x = yn ? dif : sum, y = yn ? sum : -dif;
			x = xn ? -(yn ? sum : dif) : x;
			y = xn ? (yn ? dif : -sum) : y;
			unsigned ph = c.e_phase[(xn & 2) | (yn & 1)];
			if (V == V_INT) {
				IF<0, TAIL0>::run(x, y, ph, c); IT<TAIL0, NST>::run(x, y, ph, c);
			} else {
				// int -> float through the 1.5*2^23 magic (|v| < 2^22 after the turn for y; x < 2^23 goes in as X = x + 2^23)
				float X = __uint_as_float(0x4B000000u + (unsigned)x);		// 2^23 + x exactly, x in [0, 2^23)
				float yf = __uint_as_float(0x4B400000u + (unsigned)y) - 12582912.0f;
				float pf = 0.0f;
				FP<0, (V == V_FP ? NST : TAIL0)>::run(X, yf, pf, c);
				x = (int)(__float_as_uint(X) & 0x007FFFFFu);
				const int dph = (int)__float_as_uint(pf + 12582912.0f) - 0x4B400000;
				ph += (unsigned)dph << 8;					// left-justified PW24
				if (V == V_MIX) {
					y = (int)__float_as_uint(yf + 12582912.0f) - 0x4B400000;
					IT<TAIL0, NST>::run(x, y, ph, c);
				}
			}
			const int b = (x >> 8) & 1;
			const int mag = (x + 127 + b) >> 8;					// rtl/topolar.v:253-255, D = 8
			acc = acc * 31u + (unsigned)mag * 65599u + (ph >> 8);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

#include <cmath>
#include <vector>
int main() {
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	K c;
	for (int k = 0; k < NST; k++) {		// sw/cordiclib.cpp:157-169, PW = 24
		const unsigned a = (unsigned)(atan2(1.0, pow(2.0, k + 1)) * (4.0 * pow(2.0, 22)) / (2.0 * M_PI));
		c.pa[k] = (int)(a << 8); c.na[k] = -(int)(a << 8);
		c.fa[k] = (float)a / 8388608.0f;			// sM * fa = s * a exactly
		c.tk[k] = (float)ldexp(1.0, -(k + 1)); c.ck[k] = 1.0f + c.tk[k];
	}
	const unsigned E = 1u << 21;
	c.e_phase[0] = (1u * E) << 8; c.e_phase[1] = (7u * E) << 8; c.e_phase[2] = (3u * E) << 8; c.e_phase[3] = (5u * E) << 8;
	const int grid = p.multiProcessorCount * 8, iters = 2048;
	const size_t nthreads = (size_t)grid * 256;
	unsigned *d[3]; std::vector<unsigned> h[3];
	const char *names[3] = {"int (8 slots, 6 from stage 11)", "all FP32 (7 slots)", "FP32 head + int tail"};
	for (int v = 0; v < 3; v++) {
		cudaMalloc(&d[v], nthreads * 4); h[v].resize(nthreads);
		cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
		for (int rep = 0; rep < 2; rep++) {	// second launch is the timed one
			cudaEventRecord(e0);
			if (v == 0) kchain<V_INT><<<grid, 256>>>(d[v], iters, c);
			if (v == 1) kchain<V_FP><<<grid, 256>>>(d[v], iters, c);
			if (v == 2) kchain<V_MIX><<<grid, 256>>>(d[v], iters, c);
			cudaEventRecord(e1); cudaEventSynchronize(e1);
		}
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		cudaMemcpy(h[v].data(), d[v], nthreads * 4, cudaMemcpyDeviceToHost);
		const double samples = (double)nthreads * iters * 4;
		size_t bad = 0;
		for (size_t i = 0; i < nthreads; i++) bad += (h[v][i] != h[0][i]);
		printf("%-32s %8.3f ms  %7.1f Gsamples/s compute-only  %5.2f clk/stage/warp-quad  mismatching threads vs int: %zu\n", names[v], ms,
			samples / (ms * 1e-3) / 1e9, (ms * 1e-3) * (clk * 1e3) * p.multiProcessorCount / (samples / 32 * NST), bad);
	}
	return 0;
}
