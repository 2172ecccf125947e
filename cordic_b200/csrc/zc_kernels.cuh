// zc_kernels.cuh -- device code of the zcordic engine (sm_100a).
//
// One sample per lane; the micro-rotation stages are fully unrolled in registers; the
// arctan table and every other per-configuration constant travel in the kernel parameter
// block, i.e. the constant bank, so stage operands are read as c[0x0][..] immediates.
// The arithmetic is integer shift-add (rtl/cordic.v:253-280, rtl/topolar.v:217-243): there
// is no contraction here for tensor cores to do.
//
// Representation (why the fast kernels are exact):
//   * phase is kept LEFT-justified in 32 bits (phase << (32-PW)), as are the angles, so
//     PW-bit modular arithmetic is plain u32 arithmetic and the "negative phase" test
//     ph[PW-1] (rtl/cordic.v:265) is the sign bit;
//   * x/y are kept right-justified, sign-extended in int32.  The RTL registers are WW bits
//     and wrap; the host only selects a fast kernel when it has proved that no in-range
//     input can exceed WW bits (zc_api.cu: fast_path_is_exact), otherwise the generic
//     kernel below, which wraps after every operation, is used.
#ifndef ZC_KERNELS_CUH
#define ZC_KERNELS_CUH

#include <cstdint>
#include <cuda_runtime.h>

namespace zc {

// SRC_MIX: per-sample (x,y) like SRC_XY, phase from the NCO accumulator like SRC_NCO (a complex mixer)
enum Source { SRC_CONST = 0, SRC_XY = 1, SRC_NCO = 2, SRC_MIX = 3 };

// Per-launch constants (kernel parameter => constant bank).
struct CoreConsts {
	int32_t  na[32];	// -(cordic_angle[k] << (32-PW)) for the stages that are not pass-through
	uint32_t pa[32];	//  (cordic_angle[k] << (32-PW))
	int32_t  cx[4], cy[4];	// SRC_CONST/NCO: extended (x0,y0) after each of the 4 quarter-turn pre-rotations
	uint32_t e_phase[4];	// topolar: pre-rotation phases 1E,3E,5E,7E left-justified, indexed by {xneg,yneg}
	int32_t  pshift;	// 32-PW
	int32_t  in_shl;	// 32-IW: discards the bits above the IW-bit port
	int32_t  in_shr;	// arithmetic shift that sign-extends and leaves the (WW-IW-1|2) zero LSBs
	int32_t  D;		// WW-OW
	int32_t  rc;		// 2^(D-1)-1 when the core rounds (WW > OW+1), else 0
	int32_t  do_round;
	int32_t  neff;		// number of live stages (generic kernel)
	int32_t  wsh;		// 32-WW (generic kernel: wrap)
	uint32_t nco_phase0, nco_step, nco_n0;
};

__device__ __forceinline__ int imad(int a, int b, int c) {
	int r;
	asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
	return r;
}
// Negation kept opaque to the optimiser: left to itself it rewrites -(2m+1) as a second LOP3 on the already
// saturated ALU pipe; as a lone `neg` ptxas emits IMAD.MOV, which runs on the light FMA pipe.
__device__ __forceinline__ int ineg(int a) {
	int r;
	asm("neg.s32 %0, %1;" : "=r"(r) : "r"(a));
	return r;
}

// ---- one micro-rotation, rotation mode (rtl/cordic.v:263-279) ---------------------------
// d = +1 when the residual phase is >= 0, else -1.  x' = x - d*(y>>>k); y' = y + d*(x>>>k);
// p' = p - d*angle.  Both updates read the OLD x and y.
template <int K>
__device__ __forceinline__ void rot_step(int &x, int &y, int &p, const int na) {
	constexpr int S = (K + 1 > 31) ? 31 : (K + 1);
	// Pipe budget per stage (profiles/ubench_r1.txt): the ALU pipe (shifts, LEA/IADD3) and the heavy FMA pipe
	// (IMAD) each retire one warp-instruction per 2 clocks.  3 shifts + (2*md+1) on the ALU pipe, three IMADs
	// on the heavy pipe, and the negation as an IMAD.MOV that ptxas places on the light FMA pipe.
	const int md = p >> 31;
	const int d = md + md + 1;
	const int nd = ineg(d);
	const int sy = y >> S, sx = x >> S;
	const int x1 = imad(sy, nd, x);
	const int y1 = imad(sx, d, y);
	p = imad(d, na, p);
	x = x1;
	y = y1;
}

// ---- one micro-rotation, vectoring mode (rtl/topolar.v:227-243) --------------------------
// s = -1 when y is below the axis (yv[WW-1]), else +1 (y == 0 counts as above).
// x' = x + s*(y>>>k); y' = y - s*(x>>>k); ph' = ph + s*angle.
template <int K>
__device__ __forceinline__ void vec_step(int &x, int &y, uint32_t &ph, const int pa) {
	constexpr int S = (K + 1 > 31) ? 31 : (K + 1);
	const int md = y >> 31;
	const int s = md + md + 1;
	const int ns = ineg(s);
	const int sy = y >> S, sx = x >> S;
	const int x1 = imad(sy, s, x);
	const int y1 = imad(sx, ns, y);
	ph = (uint32_t)imad(s, pa, (int)ph);
	x = x1;
	y = y1;
}

// ---- one micro-rotation, vectoring mode, once y has converged below the shift -----------------------------
// The same assignments as vec_step<K>, for a stage where the host has PROVED -2^S <= y < 2^S for every input
// (zc_api.cu: vec_tail_start; |y_i| <= x_i*2^-i + i by induction over rtl/topolar.v:227-243).  Then y>>>S is
// 0 or -1, i.e. equal to the sign word md, and s*(y>>>S) == -md: x' = x - md.  Six issue slots instead of
// eight: SHF, SHF, IMAD (-s straight from md), IMAD (y'), IADD3 (x'), IMAD (phase, -s * -angle).
template <int K>
__device__ __forceinline__ void vec_step_tail(int &x, int &y, uint32_t &ph, const int na) {
	constexpr int S = (K + 1 > 31) ? 31 : (K + 1);
	const int md = y >> 31;
	const int ns = imad(md, -2, -1);	// as (~md | 1) on the ALU pipe it measured 3 % slower (the ALU pipe is the busy one)
	const int sx = x >> S;
	y = imad(sx, ns, y);
	x = x - md;
	ph = (uint32_t)imad(ns, na, (int)ph);
}

template <int N, int K = 0>
struct Unroll {
	static __device__ __forceinline__ void rot(int &x, int &y, int &p, const CoreConsts &c) {
		rot_step<K>(x, y, p, c.na[K]);
		Unroll<N, K + 1>::rot(x, y, p, c);
	}
	// stages K0.. take the short form (K0 == N: none do)
	template <int K0>
	static __device__ __forceinline__ void vec(int &x, int &y, uint32_t &ph, const CoreConsts &c) {
		if (K >= K0) vec_step_tail<K>(x, y, ph, c.na[K]);
		else vec_step<K>(x, y, ph, (int)c.pa[K]);
		Unroll<N, K + 1>::template vec<K0>(x, y, ph, c);
	}
};
template <int N>
struct Unroll<N, N> {
	static __device__ __forceinline__ void rot(int &, int &, int &, const CoreConsts &) {}
	template <int K0>
	static __device__ __forceinline__ void vec(int &, int &, uint32_t &, const CoreConsts &) {}
};

// Convergent rounding and truncation to OW bits (rtl/cordic.v:290-295,311-312).
__device__ __forceinline__ int round_out(int v, const CoreConsts &c) {
	const int b = (v >> c.D) & c.do_round;
	return (int)((uint32_t)v + (uint32_t)c.rc + (uint32_t)b) >> c.D;	// unsigned add: WW may be 32
}

// Octant pre-rotation of the phase (rtl/cordic.v:131-188): returns the quarter-turn count q
// and leaves the residual phase in [-45,45) degrees, left-justified.
__device__ __forceinline__ int octant(uint32_t P, int &p) {
	const uint32_t t = P + 0x20000000u;
	p = (int)((t & 0x3fffffffu) - 0x20000000u);
	return (int)(t >> 30);
}

__device__ __forceinline__ void quarter_turn(int q, int ex, int ey, int &x, int &y) {
	// q=0:(ex,ey) 1:(-ey,ex) 2:(-ex,-ey) 3:(ey,-ex)
	const int a = (q & 1) ? -ey : ex;
	const int b = (q & 1) ? ex : ey;
	x = (q & 2) ? -a : a;
	y = (q & 2) ? -b : b;
}

__device__ __forceinline__ int pack16(int lo, int hi) { return (int)(((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16)); }

__device__ __forceinline__ int4 ldg_stream(const int4 *p) {
	int4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
		: "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ void stg_stream64(int2 *p, const int2 v) {
	asm volatile("st.global.L1::no_allocate.v2.s32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream(int4 *p, const int4 v) {
	asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
		:: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- rotation mode, fast path: 4 samples per thread, 128-bit loads/stores ------------------
// OUT16: one output word per sample, (int16 o_xval) | (int16 o_yval) << 16 -- the packed ports of a core with OW <= 16
template <int NEFF, int SRC, bool OUT16 = false>
__global__ void __launch_bounds__(256)
k_rotate(const int4 *__restrict__ phase4, const int4 *__restrict__ xyin4,
		int4 *__restrict__ xyout4, size_t ngroups, const __grid_constant__ CoreConsts c) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
		uint32_t P[4];
		int x[4], y[4];
		if (SRC == SRC_NCO || SRC == SRC_MIX) {
			const uint32_t base = c.nco_phase0 + (c.nco_n0 + (uint32_t)(g << 2)) * c.nco_step;
			const uint32_t keep = ~((1u << c.pshift) - 1u);	// i_phase = phase32 >> (32-PW)
#pragma unroll
			for (int s = 0; s < 4; s++)
				P[s] = (base + (uint32_t)s * c.nco_step) & keep;
		} else {
			const int4 pv = ldg_stream(phase4 + g);
			P[0] = (uint32_t)pv.x << c.pshift; P[1] = (uint32_t)pv.y << c.pshift;
			P[2] = (uint32_t)pv.z << c.pshift; P[3] = (uint32_t)pv.w << c.pshift;
		}
		int ex[4], ey[4];
		if (SRC == SRC_XY || SRC == SRC_MIX) {
			const int4 a = ldg_stream(xyin4 + 2 * g), b = ldg_stream(xyin4 + 2 * g + 1);
			const int raw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
			for (int s = 0; s < 4; s++) {
				ex[s] = (raw[2 * s] << c.in_shl) >> c.in_shr;
				ey[s] = (raw[2 * s + 1] << c.in_shl) >> c.in_shr;
			}
		}
		int ox[4], oy[4];
#pragma unroll
		for (int s = 0; s < 4; s++) {
			int p;
			const int q = octant(P[s], p);
			if (SRC == SRC_XY || SRC == SRC_MIX) {
				quarter_turn(q, ex[s], ey[s], x[s], y[s]);
			} else {
				const int xa = (q & 1) ? c.cx[1] : c.cx[0], ya = (q & 1) ? c.cy[1] : c.cy[0];
				const int xb = (q & 1) ? c.cx[3] : c.cx[2], yb = (q & 1) ? c.cy[3] : c.cy[2];
				x[s] = (q & 2) ? xb : xa;
				y[s] = (q & 2) ? yb : ya;
			}
			Unroll<NEFF>::rot(x[s], y[s], p, c);
			ox[s] = round_out(x[s], c);
			oy[s] = round_out(y[s], c);
		}
		if (OUT16) {
			stg_stream(xyout4 + g, make_int4(pack16(ox[0], oy[0]), pack16(ox[1], oy[1]), pack16(ox[2], oy[2]), pack16(ox[3], oy[3])));
		} else {
			stg_stream(xyout4 + 2 * g, make_int4(ox[0], oy[0], ox[1], oy[1]));
			stg_stream(xyout4 + 2 * g + 1, make_int4(ox[2], oy[2], ox[3], oy[3]));
		}
	}
}

// ---- vectoring mode, fast path ---------------------------------------------------------------
// NTAIL: how many of the last stages run in the short form (vec_step_tail)
// IN16: one input word per sample, (int16 i_xval) | (int16 i_yval) << 16 -- the packed ports of a core with IW <= 16
template <int NEFF, int NTAIL, bool IN16 = false>
__global__ void __launch_bounds__(256)
k_topolar(const int4 *__restrict__ xyin4, int4 *__restrict__ mag4, int4 *__restrict__ ph4,
		size_t ngroups, const __grid_constant__ CoreConsts c) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
		int raw[8];
		if (IN16) {
			const int4 a = ldg_stream(xyin4 + g);
			const int w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
			for (int s = 0; s < 4; s++) { raw[2 * s] = w[s]; raw[2 * s + 1] = w[s] >> 16; }	// bits above IW are shifted out below
		} else {
			const int4 a = ldg_stream(xyin4 + 2 * g), b = ldg_stream(xyin4 + 2 * g + 1);
			raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w; raw[4] = b.x; raw[5] = b.y; raw[6] = b.z; raw[7] = b.w;
		}
		int om[4], op[4];
#pragma unroll
		for (int s = 0; s < 4; s++) {
			// rtl/topolar.v:83-84: two sign bits, the input, WW-IW-2 zeros
			const int ex = (raw[2 * s] << c.in_shl) >> c.in_shr;
			const int ey = (raw[2 * s + 1] << c.in_shl) >> c.in_shr;
			// rtl/topolar.v:122-152: a +-45 degree turn selected by the two input signs
			const int xn = ex >> 31, yn = ey >> 31;		// 0 / -1
			const int sum = ex + ey, dif = ex - ey;
			//  {x>=0,y>=0}: ( sum, -dif)  {x>=0,y<0}: ( dif,  sum)
			//  {x<0, y>=0}: (-dif, -sum)  {x<0, y<0}: (-sum,  dif)
			int x = yn ? dif : sum;
			int y = yn ? sum : -dif;
			x = xn ? -((yn) ? sum : dif) : x;
			y = xn ? ((yn) ? dif : -sum) : y;
			uint32_t ph = c.e_phase[(xn & 2) | (yn & 1)];
			Unroll<NEFF>::template vec<NEFF - NTAIL>(x, y, ph, c);
			om[s] = round_out(x, c);
			op[s] = (int)(ph >> c.pshift);
		}
		stg_stream(mag4 + g, make_int4(om[0], om[1], om[2], om[3]));
		stg_stream(ph4 + g, make_int4(op[0], op[1], op[2], op[3]));
	}
}

enum { PROBE_LOCAL = 0, PROBE_SCATTERED = 2 };
// ---- probe: are neighbouring samples' phases neighbours? ------------------------------------------------
// 8 windows of 32 consecutive samples spread over the stream; a pair counts as local when the circular phase
// difference is at most `lim` (one phase LSB for the CORDIC tables: a sweep or a slow NCO, whose word rows are
// conflict-free; one table entry, in 32-bit phase units, for the LUT cores), and the verdict is PROBE_LOCAL when at least
// 7 pairs in 8 are, PROBE_SCATTERED otherwise.
// Called by EVERY thread of a CTA of at least 256 threads (it contains a barrier).  The verdict is a pure function of
// (phase[], n, pshift, lim): the two kernels of an auto-selected call evaluate it independently, CTA by CTA, and always
// agree -- exactly one of them processes the batch, whatever other streams or graph replays are doing.  (Round 1 passed
// the verdict through a reused ring of device words written by a separate probe kernel; a concurrent call that was
// handed the same slot could flip it between the two kernels' reads, and then neither ran.)
__device__ __forceinline__ int probe_local(const uint32_t *__restrict__ phase, size_t n, int pshift, int lim) {
	const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31u;
	int local = 0;
	if (threadIdx.x < 256u && n >= 34) {
		size_t off = ((n / 8) * w) & ~(size_t)31;
		if (off + 33 > n) off = 0;
		const uint32_t a = phase[off + l], b = phase[off + l + 1];
		const int d = (int)((b - a) << pshift) >> pshift;
		local = (d >= -lim && d <= lim);
	}
	const int votes = __syncthreads_count(local);
	return (votes * 8 >= 256 * 7) ? PROBE_LOCAL : PROBE_SCATTERED;
}

} // namespace zc
#endif
