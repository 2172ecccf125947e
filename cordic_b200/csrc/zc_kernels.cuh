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

__device__ __forceinline__ int4 ldg_stream(const int4 *p) {
	int4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
		: "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ void stg_stream(int4 *p, const int4 v) {
	asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
		:: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- rotation mode, fast path: 4 samples per thread, 128-bit loads/stores ------------------
template <int NEFF, int SRC>
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
		stg_stream(xyout4 + 2 * g, make_int4(ox[0], oy[0], ox[1], oy[1]));
		stg_stream(xyout4 + 2 * g + 1, make_int4(ox[2], oy[2], ox[3], oy[3]));
	}
}

// ---- vectoring mode, fast path ---------------------------------------------------------------
// NTAIL: how many of the last stages run in the short form (vec_step_tail)
template <int NEFF, int NTAIL>
__global__ void __launch_bounds__(256)
k_topolar(const int4 *__restrict__ xyin4, int4 *__restrict__ mag4, int4 *__restrict__ ph4,
		size_t ngroups, const __grid_constant__ CoreConsts c) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
		const int4 a = ldg_stream(xyin4 + 2 * g), b = ldg_stream(xyin4 + 2 * g + 1);
		const int raw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
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

// ---- generic kernels: any configuration, any alignment, WW-bit wrap modelled ----------------
__device__ __forceinline__ int wrapw(int v, int wsh) { return (int)((uint32_t)v << wsh) >> wsh; }

template <int SRC>
__global__ void __launch_bounds__(256)
k_rotate_generic(const uint32_t *__restrict__ phase, const int32_t *__restrict__ xyin,
		int32_t *__restrict__ xyout, size_t n, const __grid_constant__ CoreConsts c) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		uint32_t P;
		if (SRC == SRC_NCO || SRC == SRC_MIX) {
			const uint32_t keep = ~((1u << c.pshift) - 1u);
			P = (c.nco_phase0 + (c.nco_n0 + (uint32_t)i) * c.nco_step) & keep;
		} else {
			P = phase[i] << c.pshift;
		}
		int p;
		const int q = octant(P, p);
		int x, y;
		if (SRC == SRC_XY || SRC == SRC_MIX) {
			const int ex = (xyin[2 * i] << c.in_shl) >> c.in_shr;
			const int ey = (xyin[2 * i + 1] << c.in_shl) >> c.in_shr;
			quarter_turn(q, ex, ey, x, y);
			x = wrapw(x, c.wsh); y = wrapw(y, c.wsh);
		} else {
			x = c.cx[q]; y = c.cy[q];
		}
		for (int k = 0; k < c.neff; k++) {
			const int sh = (k + 1 > 31) ? 31 : (k + 1);
			const int sy = y >> sh, sx = x >> sh;
			const uint32_t ux = (uint32_t)x, uy = (uint32_t)y;	// unsigned: the sums may wrap (WW up to 32)
			const uint32_t ak = (k < 32) ? c.pa[k] : 0u;		// sequential cores iterate past the last angle
			if (p < 0) {
				x = wrapw((int)(ux + (uint32_t)sy), c.wsh); y = wrapw((int)(uy - (uint32_t)sx), c.wsh);
				p = (int)((uint32_t)p + ak);
			} else {
				x = wrapw((int)(ux - (uint32_t)sy), c.wsh); y = wrapw((int)(uy + (uint32_t)sx), c.wsh);
				p = (int)((uint32_t)p - ak);
			}
		}
		const int bx = (x >> c.D) & c.do_round, by = (y >> c.D) & c.do_round;
		xyout[2 * i] = wrapw((int)((uint32_t)x + (uint32_t)c.rc + (uint32_t)bx), c.wsh) >> c.D;
		xyout[2 * i + 1] = wrapw((int)((uint32_t)y + (uint32_t)c.rc + (uint32_t)by), c.wsh) >> c.D;
	}
}

__global__ void __launch_bounds__(256)
k_topolar_generic(const int32_t *__restrict__ xyin, int32_t *__restrict__ mag,
		uint32_t *__restrict__ phout, size_t n, const __grid_constant__ CoreConsts c) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const int ex = (xyin[2 * i] << c.in_shl) >> c.in_shr;
		const int ey = (xyin[2 * i + 1] << c.in_shl) >> c.in_shr;
		const int xn = ex < 0, yn = ey < 0;
		int x, y;
		const uint32_t ax = (uint32_t)ex, ay = (uint32_t)ey;
		if (!xn && yn)      { x = (int)(ax - ay);  y = (int)(ax + ay); }
		else if (xn && !yn) { x = (int)(ay - ax);  y = (int)(0u - ax - ay); }
		else if (xn && yn)  { x = (int)(0u - ax - ay); y = (int)(ax - ay); }
		else                { x = (int)(ax + ay);  y = (int)(ay - ax); }
		x = wrapw(x, c.wsh); y = wrapw(y, c.wsh);
		uint32_t ph = c.e_phase[(xn << 1) | yn];
		for (int k = 0; k < c.neff; k++) {
			const int sh = (k + 1 > 31) ? 31 : (k + 1);
			const int sy = y >> sh, sx = x >> sh;
			const uint32_t ux = (uint32_t)x, uy = (uint32_t)y;
			const uint32_t ak = (k < 32) ? c.pa[k] : 0u;
			if (y < 0) {
				x = wrapw((int)(ux - (uint32_t)sy), c.wsh); y = wrapw((int)(uy + (uint32_t)sx), c.wsh); ph -= ak;
			} else {
				x = wrapw((int)(ux + (uint32_t)sy), c.wsh); y = wrapw((int)(uy - (uint32_t)sx), c.wsh); ph += ak;
			}
		}
		const int b = (x >> c.D) & c.do_round;
		mag[i] = wrapw((int)((uint32_t)x + (uint32_t)c.rc + (uint32_t)b), c.wsh) >> c.D;
		phout[i] = ph >> c.pshift;
	}
}

// ---- LUT cores (rtl/sintable.v:71-75, rtl/quarterwav.v:92-109) --------------------------------
struct LutConsts {
	int32_t pshift;		// 32-pw
	int32_t osh;		// 32-ow : sign-extension of the OW-bit table word
	uint32_t lowmask;	// quarterwav: 2^(pw-2)-1
	int32_t pw;
};

// which LUT kernel a probed batch goes to (zc_seeded.cuh: k_seed_probe writes TD_TABLE = 0 for neighbouring phases,
// TD_PACKED = 2 for scattered ones)
enum { LUT_GATE_L2 = 0, LUT_GATE_SMEM = 2 };
#ifndef ZC_LUT_MLP
#define ZC_LUT_MLP 2
#endif
constexpr int LUT_MLP = ZC_LUT_MLP;

template <bool QUARTER>
__device__ __forceinline__ int lut_one(uint32_t phase32, const uint32_t *__restrict__ tbl, const LutConsts &c) {
	const uint32_t ip = phase32 >> c.pshift;
	if (!QUARTER) {
		return (int)(__ldg(tbl + ip) << c.osh) >> c.osh;
	} else {
		const uint32_t fold = 0u - ((ip >> (c.pw - 2)) & 1u);	// all-ones when i_phase[PW-2]
		const uint32_t idx = (ip ^ fold) & c.lowmask;
		const int neg = -(int)((ip >> (c.pw - 1)) & 1u);	// -1 when i_phase[PW-1]
		const int v = (int)__ldg(tbl + idx);
		return (((v ^ neg) - neg) << c.osh) >> c.osh;
	}
}

template <bool QUARTER>
__global__ void __launch_bounds__(256)
k_lut(const int4 *__restrict__ phase4, int4 *__restrict__ out4, const uint32_t *__restrict__ tbl,
		size_t ngroups, const __grid_constant__ LutConsts c, const int *__restrict__ gate) {
	if (gate != nullptr && *gate != LUT_GATE_L2) return;	// a probe kernel chose the shared-memory kernel for this batch
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
		const int4 pv = ldg_stream(phase4 + g);
		int4 o;
		o.x = lut_one<QUARTER>((uint32_t)pv.x, tbl, c);
		o.y = lut_one<QUARTER>((uint32_t)pv.y, tbl, c);
		o.z = lut_one<QUARTER>((uint32_t)pv.z, tbl, c);
		o.w = lut_one<QUARTER>((uint32_t)pv.w, tbl, c);
		stg_stream(out4 + g, o);
	}
}

// ---- LUT cores with the table resident in shared memory ---------------------------------------------------
// k_lut gathers from a table that lives in L2 (512 KB for the shipped sintable): ideal for sweeps (1.00 / 0.97 of the
// HBM copy peak), an L2 gather per sample for scattered phases (286 / 434 Gsamples/s).  Here every CTA first stages a lossless compressed copy of the
// table in shared memory and then looks every sample up there, whatever the phase pattern:
//   sintable   (rtl/sintable.v:71-75)   the second half-wave is the negated first one -- IF the table really is like that;
//              the staging loop checks tbl[i + N/2] == -tbl[i] and 16-bit range entry by entry (the generator's tables
//              pass: C truncation toward zero is odd-symmetric, sw/sintable.cpp:156-168), and stores N/2 int16;
//   quarterwav (rtl/quarterwav.v:92-109) the words are magnitudes below 2^16 (u16) or 2^24 (u16 + u8, HI8).
// The table is the caller's memory and may hold anything: when a check fails the CTA (every CTA reaches the same
// verdict, they all read the whole table) serves its samples from global memory exactly as k_lut does.
template <bool QUARTER, bool HI8>
__global__ void __launch_bounds__(1024, 1)
k_lut_smem(const int4 *__restrict__ phase4, int4 *__restrict__ out4, const uint32_t *__restrict__ tbl,
		size_t ngroups, const __grid_constant__ LutConsts c, const int *__restrict__ gate) {
	if (gate != nullptr && *gate != LUT_GATE_SMEM) return;
	extern __shared__ __align__(16) unsigned char lsm[];
	const uint32_t nent = QUARTER ? (1u << (c.pw - 2)) : (1u << (c.pw - 1));
	unsigned short *const lo = reinterpret_cast<unsigned short *>(lsm);
	unsigned char *const hi = lsm + 2 * (size_t)nent;
	int ok = 1;
	for (uint32_t i = threadIdx.x; i < nent; i += blockDim.x) {
		if (!QUARTER) {
			const int v = (int)(tbl[i] << c.osh) >> c.osh, w = (int)(tbl[i + nent] << c.osh) >> c.osh;
			ok &= (w == -v) & (v >= -32768) & (v <= 32767);
			lo[i] = (unsigned short)v;
		} else {
			const uint32_t v = tbl[i];
			ok &= HI8 ? (v < (1u << 24)) : (v < (1u << 16));
			lo[i] = (unsigned short)v;
			if (HI8) hi[i] = (unsigned char)(v >> 16);
		}
	}
	ok = __syncthreads_and(ok);
	auto one = [&](uint32_t phase32) -> int {
		const uint32_t ip = phase32 >> c.pshift;
		if (!QUARTER) {
			const int neg = -(int)(ip >> (c.pw - 1));			// -1 in the second half-wave
			const int v = (short)lo[ip & (nent - 1u)];
			return (v ^ neg) - neg;
		} else {
			const uint32_t fold = 0u - ((ip >> (c.pw - 2)) & 1u);
			const uint32_t idx = (ip ^ fold) & c.lowmask;
			const int neg = -(int)((ip >> (c.pw - 1)) & 1u);
			const int v = (int)(HI8 ? ((uint32_t)lo[idx] | ((uint32_t)hi[idx] << 16)) : (uint32_t)lo[idx]);
			return (((v ^ neg) - neg) << c.osh) >> c.osh;
		}
	};
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (ok) {
		// LUT_MLP 16-byte loads in flight per thread: one CTA of 1024 threads per SM has to cover HBM's latency alone
		for (; g + (LUT_MLP - 1) * stride < ngroups; g += LUT_MLP * stride) {
			int4 pv[LUT_MLP];
#pragma unroll
			for (int k = 0; k < LUT_MLP; k++) pv[k] = ldg_stream(phase4 + g + k * stride);
#pragma unroll
			for (int k = 0; k < LUT_MLP; k++)
				stg_stream(out4 + g + k * stride, make_int4(one((uint32_t)pv[k].x), one((uint32_t)pv[k].y),
					one((uint32_t)pv[k].z), one((uint32_t)pv[k].w)));
		}
		for (; g < ngroups; g += stride) {
			const int4 pv = ldg_stream(phase4 + g);
			stg_stream(out4 + g, make_int4(one((uint32_t)pv.x), one((uint32_t)pv.y), one((uint32_t)pv.z), one((uint32_t)pv.w)));
		}
	} else {
		for (; g < ngroups; g += stride) {
			const int4 pv = ldg_stream(phase4 + g);
			int4 o;
			o.x = lut_one<QUARTER>((uint32_t)pv.x, tbl, c);
			o.y = lut_one<QUARTER>((uint32_t)pv.y, tbl, c);
			o.z = lut_one<QUARTER>((uint32_t)pv.z, tbl, c);
			o.w = lut_one<QUARTER>((uint32_t)pv.w, tbl, c);
			stg_stream(out4 + g, o);
		}
	}
}

template <bool QUARTER>
__global__ void __launch_bounds__(256)
k_lut_scalar(const uint32_t *__restrict__ phase, int32_t *__restrict__ out,
		const uint32_t *__restrict__ tbl, size_t n, const __grid_constant__ LutConsts c) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
		out[i] = lut_one<QUARTER>(phase[i], tbl, c);
}

} // namespace zc
#endif
