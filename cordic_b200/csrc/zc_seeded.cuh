// zc_seeded.cuh -- rotation mode with a constant input vector (the sin/cos generator of
// bench/cpp/cordic_tb.cpp:61-80 and the NCO): table-seeded prefix + register suffix.
//
// Why this is still bit-exact rtl/cordic.v.  In rotation mode the direction d_k of stage k depends
// only on the phase (rtl/cordic.v:265), never on x/y.  With (i_xval,i_yval) fixed, the x/y registers
// after the first M stages are therefore a function of (quarter turn q, d_0..d_{M-1}) only.  The map
// phase -> (d_0..d_{M-1}) is a monotone step function of the residual phase whose steps sit at the
// partial sums of +-angle[k]; we enumerate its R <= 2^M intervals on the host (phase arithmetic
// only), let a setup KERNEL run the real stage arithmetic once per (q, interval) to fill the x/y
// table, and at run time find a sample's interval with one bucket lookup: the reduced phase's top
// bits select a bucket that contains at most one step (the plan refuses any geometry where that is
// not true), and `entry + low_bits` carries into the interval number exactly when the sample lies
// at or above that step.  The remaining NS = N-M stages run in registers exactly like the plain
// kernel, except that their directions come from a second small table indexed by the residual
// phase (again a function of the phase alone), which removes the phase recursion from the ALU.
//
// Shared-memory layout (one CTA of 1024 threads per SM, tables copied in by bulk-TMA):
//   T1[2^LB]      u32   bucket -> 16*((interval_at_bucket_start << lgW) + (W - offset_of_step_in_bucket) - (bucket << lgW)),
//                         so that (T1[bucket] + 16*u) >> (lgW+4) is the interval of reduced phase u
//   TS[R]         i32   interval -> 16*(partial angle sum + 2^(PW-3) + rmin), so that 16*u - TS = byte offset of the TD row
//   T2[R][4]      int2  (interval, q) -> (x, y) after M stages
//   TD[NSP/4][nres] int4 residual -> d_M .. d_{M+NS-1} as +1/-1 words, in planes of four stages so that
//                         consecutive residuals (a phase sweep) read consecutive 16-byte slots: no bank conflicts
// Sample-to-lane mapping: a warp owns 128 consecutive samples per iteration and lane l takes samples
// l, l+32, l+64, l+96 of them, so that for a phase sweep neighbouring lanes read neighbouring table rows
// (or the same row: a broadcast) and every global access is a fully coalesced 128/256-byte row.
#ifndef ZC_SEEDED_CUH
#define ZC_SEEDED_CUH

#include "zc_internal.h"
#include "zc_kernels.cuh"

#include <cstring>
#include <memory>
#include <mutex>
#include <utility>
#include <vector>

namespace zc {

// Suffix stages whose -d is formed as d ^ ~1 (LOP3, ALU pipe) instead of a negation (IMAD.MOV/IADD3): one stage in
// four balances the heavy FMA pipe (81 % busy with every negation on it) against the ALU pipe (70 %); measured on
// the cfg1 sweep, 20 steps: none 456-457, 0x22 463-464, 0x88 465-467, 0xAA 461, 0xFF 460 Gsamples/s.
#ifndef ZC_XNEG_MASK
#define ZC_XNEG_MASK 0x8888
#endif
constexpr int SEED_MAX_NS = 16;
constexpr size_t SEED_MAX_BLOCKS = 0xF0000000u;	// the table kernels count 128-sample blocks in 32 bits (2^38.9 samples)
constexpr size_t SEED_SMEM_LIMIT = 227 * 1024 - 64;	// opt-in maximum per CTA minus the mbarrier slot

struct SeedConsts {
	int32_t  M;		// stages folded into the table
	uint32_t mul_q;		// 2^(32-PW): phase*mul_q + 2^29 puts the quarter turn in bits 31:30
	uint32_t mul_u;		// 2^(34-PW): phase*mul_u + 2^31 left-justifies the reduced phase u (PW-2 bits)
	int32_t  bsh;		// u_left >> bsh = bucket number (32-LB)
	int32_t  ush;		// u_left >> ush = u << lgrow, u = the reduced phase in LSBs, offset binary
	int32_t  rsh;		// (T1 entry + (u << lgrow)) >> rsh = interval number (lgW+lgrow)
	int32_t  lgw;
	uint32_t mul_r;		// 2^(32-PW-lgrow): TD byte offset * mul_r + res_bias = residual phase, left-justified
	int32_t  lgrow;		// log2(bytes per TD row slot): every T1/TS entry is scaled by it
	int32_t  res_bias;	// rmin << (32-PW)
	float    rscale, rbias;	// float rounding: fma(2^23*1.5 + v, 2^-D, 2^23*1.5*(1-2^-D)) rounds v/2^D to nearest even
	uint32_t off_ts, off_t2, off_td;	// byte offsets of the tables in shared memory
	int32_t  td_plane;	// bytes per TD plane (nres*16)
	uint32_t total_bytes;	// multiple of 16
	int32_t  sh[SEED_MAX_NS];	// arithmetic shift of suffix stage j: min(M+j+1, 31)
	uint32_t R;
};

// ---- setup kernel: the x/y table, by running the real stages --------------------------------------
__global__ void k_seed_fill_xy(const uint32_t *__restrict__ rep_phase /* left-justified, one per interval */,
		int2 *__restrict__ t2, uint32_t R, int M, const __grid_constant__ CoreConsts c) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 4u * R) return;
	const uint32_t rank = i >> 2, q = i & 3u;
	int p = (int)rep_phase[rank];
	int x = c.cx[q], y = c.cy[q];
	for (int k = 0; k < M; k++) {
		const int sh = (k + 1 > 31) ? 31 : (k + 1);
		const int sy = y >> sh, sx = x >> sh;
		if (p < 0) { x = x + sy; y = y - sx; p += (int)c.pa[k]; }
		else       { x = x - sy; y = y + sx; p -= (int)c.pa[k]; }
	}
	t2[i] = make_int2(x, y);
}

enum { MA_NEG = 0, MA_XNEG = 1, MA_DP2A = 2 };
__device__ __forceinline__ int dp2a_lo(int a, int b, int c) {
	int r;
	asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
	return r;
}
__device__ __forceinline__ int dp2a_hi(int a, int b, int c) {
	int r;
	asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
	return r;
}

template <int NS, int J = 0>
struct Suffix {
	// directions from the table: 2 shifts + 2 multiply-adds by +-1 + one negation per stage
	// MA_NEG: -d by negation; MA_XNEG: as MA_NEG with ZC_XNEG_MASK (word-table flavour; the byte flavour already spends a
	// PRMT per stage on the ALU pipe); MA_DP2A: d[J] is the word {0, d, 0, -d} (bytes 3..0) and both updates are IDP.2A:
	//   dp2a.lo(sy, w, x) = x + sy.h0 * (-d) + sy.h1 * 0      dp2a.hi(sx, w, y) = y + sx.h0 * d + sx.h1 * 0
	// which is the stage exactly as long as y>>sh and x>>sh fit 16 bits (the host checks sh >= WW-16): no negation at all,
	// 4 issue slots per stage instead of 5 (tools/ubench3.cu: IDP.2A issues at IMAD's rate, on IMAD's pipe).
	template <int MA>
	static __device__ __forceinline__ void run(int &x, int &y, const int (&d)[SEED_MAX_NS], const SeedConsts &s) {
		const int sy = y >> s.sh[J], sx = x >> s.sh[J];
		int x1, y1;
		if (MA == MA_DP2A) {
			x1 = dp2a_lo(sy, d[J], x);
			y1 = dp2a_hi(sx, d[J], y);
		} else {
			// -d for d = +-1: as a negation (IMAD.MOV / IADD3, ptxas' choice) or as d ^ ~1 (LOP3, ALU pipe), per stage
			const int nd = (MA == MA_XNEG && ((ZC_XNEG_MASK >> J) & 1)) ? (d[J] ^ -2) : -d[J];
			x1 = imad(sy, nd, x);
			y1 = imad(sx, d[J], y);
		}
		x = x1; y = y1;
		Suffix<NS, J + 1>::template run<MA>(x, y, d, s);
	}
	// directions from the phase recursion in registers (rtl/cordic.v:263-279), as in k_rotate
	static __device__ __forceinline__ void run_reg(int &x, int &y, int &p, const CoreConsts &c, const SeedConsts &s) {
		const int md = p >> 31;
		const int d = md + md + 1, nd = ineg(d);
		const int sy = y >> s.sh[J], sx = x >> s.sh[J];
		const int x1 = imad(sy, nd, x);
		const int y1 = imad(sx, d, y);
		p = imad(d, c.na[s.M + J], p);
		x = x1; y = y1;
		Suffix<NS, J + 1>::run_reg(x, y, p, c, s);
	}
};
template <int NS>
struct Suffix<NS, NS> {
	template <int MA>
	static __device__ __forceinline__ void run(int &, int &, const int (&)[SEED_MAX_NS], const SeedConsts &) {}
	static __device__ __forceinline__ void run_reg(int &, int &, int &, const CoreConsts &, const SeedConsts &) {}
};

enum { TD_TABLE = 0, TD_REGS = 1, TD_PACKED = 2, TD_TABLE_DP = 3 };
// plan flavours: word TD + x/y table, byte TD + x/y table, byte TD + per-interval prefix directions (no x/y table), and
// and the IDP.2A form of the first: word TD holding the multiplier words {0, d, 0, -d}.  (The byte flavours have no
// IDP.2A form: storing the pair (-d, d) per stage and building the multiplier word with one PRMT saves the negation but
// doubles the rows, and their scattered reads cost more than that: 320 vs 374 Gsamples/s for the constant-vector kernel
// on random phases, 181 vs 195 for per-sample vectors -- measured, dropped.)
enum { FL_WORDS = 0, FL_PACKED = 1, FL_DIRS = 2, FL_WORDS_DP = 3 };
static inline bool fl_packed(int flavour) { return flavour == FL_PACKED || flavour == FL_DIRS; }
static inline bool fl_dirs(int flavour) { return flavour == FL_DIRS; }
constexpr int DIRS_M = 12;	// prefix depth of the FL_DIRS flavour (its kernel unrolls the byte indices)

// Sign-extends byte `b` of w with one PRMT (selector nibble 8|b replicates that byte's sign bit;
// __byte_perm() masks the replicate bit off, hence the PTX).
__device__ __forceinline__ int sext_byte(uint32_t w, int b) {
	int r;
	asm("prmt.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(w), "r"(0x8880 + 0x1111 * b));
	return r;
}

__device__ __forceinline__ uint32_t ldg_stream32(const uint32_t *p) {
	uint32_t r;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
	return r;
}
__device__ __forceinline__ void stg_stream64(int2 *p, const int2 v) {
	asm volatile("st.global.L1::no_allocate.v2.s32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// Convergent rounding on the FMA pipe (the ALU pipe is this kernel's bottleneck).  For |v| < 2^22 the word
// 0x4B400000+v IS the float 1.5*2^23+v; one fused multiply-add forms 1.5*2^23 + v/2^D exactly and rounds it
// to the nearest integer, ties to even -- which is precisely rtl/cordic.v:290-295 (add 2^(D-1) when bit D is
// set, 2^(D-1)-1 when it is clear, then drop D bits).  Only used when WW <= 23 and the core rounds (D >= 2).
__device__ __forceinline__ int round_out_fma(int v, const SeedConsts &s) {
	const float r = __fmaf_rn(__int_as_float(v + 0x4B400000), s.rscale, s.rbias);
	return __float_as_int(r) - 0x4B400000;
}

__device__ __forceinline__ int4 lds128(uint32_t addr) {
	int4 r;
	asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
	return r;
}

// `nblocks` blocks of 128 consecutive samples; warp w of the grid takes blocks w, w+W, w+2W, ...
// TDM: where the suffix directions come from.  TD_TABLE: one 32-bit word per stage from the TD table (no unpacking;
// fastest when neighbouring lanes read neighbouring rows -- sweeps, slow NCOs -- but two 16-byte reads per sample
// make it bank-conflict bound for scattered phases).  TD_PACKED: one signed byte per stage (a quarter of the
// shared-memory traffic, one PRMT per stage to unpack): the better trade for scattered phases.  TD_REGS: the phase
// recursion in registers (3 more issue slots per stage, no table traffic).
template <int NS, int SRC, bool RF, int TDM>
__global__ void __launch_bounds__(1024, 1)
k_rotate_seeded(const uint32_t *__restrict__ phase, int2 *__restrict__ xyout, size_t nblocks,
		const __grid_constant__ CoreConsts c, const __grid_constant__ SeedConsts s,
		const uint4 *__restrict__ tables, const int *__restrict__ gate) {
	extern __shared__ __align__(128) unsigned char smem[];
	// Auto-selection: both table flavours are enqueued behind a probe kernel that writes which one suits the
	// data; the other returns here, before touching shared memory.
	if (gate != nullptr && *gate != (TDM == TD_TABLE_DP ? TD_TABLE : TDM)) return;
	const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
	const uint32_t mbar = sbase + s.total_bytes;		// 8-byte slot after the tables

	// ---- stage the tables: one thread issues bulk-TMA copies that complete on an mbarrier --------
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(s.total_bytes) : "memory");
		const char *src = reinterpret_cast<const char *>(tables);
		for (uint32_t off = 0; off < s.total_bytes; off += 32768u) {
			const uint32_t len = (s.total_bytes - off < 32768u) ? (s.total_bytes - off) : 32768u;
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				:: "r"(sbase + off), "l"(src + off), "r"(len), "r"(mbar) : "memory");
		}
	}
	__syncthreads();		// the mbarrier is initialised before anyone waits on it
	{
		uint32_t done = 0;
		while (!done) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
				"selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar) : "memory");
		}
	}

	const uint32_t *const T1 = reinterpret_cast<const uint32_t *>(smem);
	const int32_t *const TS = reinterpret_cast<const int32_t *>(smem + s.off_ts);
	const int2 *const T2 = reinterpret_cast<const int2 *>(smem + s.off_t2);
	const unsigned char *const TD = smem + s.off_td;
	const uint32_t lane = threadIdx.x & 31u;
	// Shared-window address of each TD plane, kept opaque so that ptxas holds one uniform register per plane and reads
	// [row + UR] instead of re-adding the plane stride per sample.
	uint32_t tdbase[SEED_MAX_NS / 4];
#pragma unroll
	for (int j = 0; j < SEED_MAX_NS / 4; j++)
		asm("mov.b32 %0, %1;" : "=r"(tdbase[j]) : "r"(sbase + s.off_td + (uint32_t)(j * s.td_plane)));
	// 32-bit block counters (the host keeps nblocks + nwarps below 2^32): one IADD/ISETP per iteration, not pairs
	const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
	const uint32_t nblk = (uint32_t)nblocks;
	uint32_t blk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	// One block of 128 samples.  `tin` holds the block's four phases per lane; they are folded first and then, when
	// `reload` names a block, the registers are refilled at once with that block's phases (software prefetch, one block
	// ahead; two register sets refilled two blocks ahead measured the same in short runs and 3 % slower under the power
	// cap).
	auto body = [&](uint32_t (&tin)[4], const uint32_t blk, const uint32_t reload) {
		uint32_t tq[4], tu[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			// octant fold (rtl/cordic.v:131-188): phase + 45 degrees; bits above PW fall off the top
			tq[k] = (uint32_t)imad((int)tin[k], (int)s.mul_q, 0x20000000);		// [q:2][u:PW-2][0...]
			tu[k] = (uint32_t)imad((int)tin[k], (int)s.mul_u, (int)0x80000000u);	// [u:PW-2][0...]
		}
		if (SRC == SRC_CONST && reload < nblk) {
#pragma unroll
			for (int k = 0; k < 4; k++) tin[k] = ldg_stream32(phase + ((size_t)reload << 7) + (k << 5) + lane);
		}
		int2 *const dst = xyout + ((size_t)blk << 7) + lane;
		int x[4], y[4];
		uint32_t row16[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const uint32_t u16 = tu[k] >> s.ush;				// 16 * reduced phase
			const uint32_t rank = (T1[tu[k] >> s.bsh] + u16) >> s.rsh;	// carries past the step, if any
			const int2 xy = T2[__funnelshift_l(tq[k], rank, 2)];		// row rank*4 + quarter turn
			x[k] = xy.x; y[k] = xy.y;
			row16[k] = u16 - (uint32_t)TS[rank];		// residual after M stages, as the byte offset of its TD row
		}
#pragma unroll
		for (int k = 0; k < 4; k++) {
			if (NS > 0) {
				if (TDM == TD_TABLE || TDM == TD_TABLE_DP) {
					int d[SEED_MAX_NS];
#pragma unroll
					for (int j = 0; j < NS; j += 4) {
						// one uniform shared-window base per plane, so every plane is read as [row + uniform register]
						const int4 dv = lds128(row16[k] + tdbase[j >> 2]);
						d[j] = dv.x;
						if (j + 1 < SEED_MAX_NS) d[j + 1] = dv.y;
						if (j + 2 < SEED_MAX_NS) d[j + 2] = dv.z;
						if (j + 3 < SEED_MAX_NS) d[j + 3] = dv.w;
					}
					Suffix<NS>::template run<TDM == TD_TABLE_DP ? MA_DP2A : MA_XNEG>(x[k], y[k], d, s);
				} else if (TDM == TD_PACKED) {
					const unsigned char *row = TD + (int)row16[k];	// 8-byte (NS<=8) or 16-byte rows
					uint32_t w[4] = {0, 0, 0, 0};
					if (NS <= 8) {
						const int2 v = *reinterpret_cast<const int2 *>(row);
						w[0] = (uint32_t)v.x; w[1] = (uint32_t)v.y;
					} else {
						const int4 v = *reinterpret_cast<const int4 *>(row);
						w[0] = (uint32_t)v.x; w[1] = (uint32_t)v.y; w[2] = (uint32_t)v.z; w[3] = (uint32_t)v.w;
					}
					int d[SEED_MAX_NS];
#pragma unroll
					for (int j = 0; j < NS; j++)		// sign-extend byte j&3 of word j>>2
						d[j] = sext_byte(w[j >> 2], j & 3);
					Suffix<NS>::template run<MA_NEG>(x[k], y[k], d, s);
				} else {
					int p = imad((int)row16[k], (int)s.mul_r, s.res_bias);	// residual phase, left-justified
					Suffix<NS>::run_reg(x[k], y[k], p, c, s);
				}
			}
			const int ox = RF ? round_out_fma(x[k], s) : round_out(x[k], c);
			const int oy = RF ? round_out_fma(y[k], s) : round_out(y[k], c);
			stg_stream64(dst + (k << 5), make_int2(ox, oy));
		}
	};
	if (SRC == SRC_NCO) {
		for (; blk < nblk; blk += nwarps) {
			uint32_t tin[4];
			const uint32_t base = c.nco_phase0 + (c.nco_n0 + (blk << 7) + lane) * c.nco_step;
#pragma unroll
			for (int k = 0; k < 4; k++) tin[k] = (base + (uint32_t)(k << 5) * c.nco_step) >> c.pshift;
			body(tin, blk, nblk);
		}
	} else {
		uint32_t pin[4] = {0, 0, 0, 0};
		if (blk < nblk) {
#pragma unroll
			for (int k = 0; k < 4; k++) pin[k] = ldg_stream32(phase + ((size_t)blk << 7) + (k << 5) + lane);
		}
		for (; blk < nblk; blk += nwarps) body(pin, blk, blk + nwarps);
	}
}

// ---- probe: are neighbouring samples' phases neighbours? ------------------------------------------------
// 8 windows of 32 consecutive samples spread over the stream; a pair counts as local when the circular phase
// difference is at most one LSB (a sweep or a slow NCO: word rows are conflict-free), and the word flavour is
// chosen when at least 7 pairs in 8 are.  Scattered phases get the byte-packed flavour.
// `lim`: the largest circular difference that still counts as local (1 phase LSB for the CORDIC tables; one table entry,
// in 32-bit phase units, for the LUT cores).
__global__ void k_seed_probe(const uint32_t *__restrict__ phase, size_t n, int pshift, int lim, int *gate) {
	const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31u;
	size_t off = ((n / 8) * w) & ~(size_t)31;
	if (off + 33 > n) off = 0;
	int local = 0;
	if (n >= 34) {
		const uint32_t a = phase[off + l], b = phase[off + l + 1];
		const int d = (int)((b - a) << pshift) >> pshift;
		local = (d >= -lim && d <= lim);
	}
	const int votes = __syncthreads_count(local);
	if (threadIdx.x == 0) *gate = (votes * 8 >= (int)blockDim.x * 7) ? TD_TABLE : TD_PACKED;
}

// ---- host: plan construction and cache -----------------------------------------------------------
struct SeedPlan {
	zc_params p;
	int32_t x0c[4], y0c[4];		// the pre-rotated constant vector (identifies x0,y0 modulo IW)
	int device = -1;
	int NS = 0;
	int flavour = FL_WORDS;
	SeedConsts s;
	void *dev = nullptr;		// tables, laid out as in shared memory
	std::shared_ptr<void> hold;	// owns `dev`: a copy of the plan keeps the tables alive across a cache eviction
	bool usable = false;		// false: geometry does not fit; cached so we do not retry
	uint64_t stamp = 0;
};

struct Interval { int64_t lo, hi, S; uint32_t neg; };	// neg: bit k set when stage k rotates clockwise (d_k = -1)

// Enumerates the intervals of constant (d_0..d_{M-1}) over the reduced phase range
// [-2^(PW-3), 2^(PW-3)), in ascending order.  Phase arithmetic only (rtl/cordic.v:265-279).
static void seed_intervals(const zc_params *p, int M, std::vector<Interval> &iv) {
	const int64_t half = (int64_t)1 << (p->pw - 3);
	iv.assign(1, Interval{-half, half, 0, 0u});
	std::vector<Interval> next;
	for (int k = 0; k < M; k++) {
		next.clear();
		const int64_t a = p->angle[k];
		for (const Interval &it : iv) {
			// residual = phase - S ; negative residual -> rotate clockwise, S' = S - angle
			if (it.lo < it.S) next.push_back(Interval{it.lo, it.hi < it.S ? it.hi : it.S, it.S - a, it.neg | (1u << k)});
			if (it.hi > it.S) next.push_back(Interval{it.lo > it.S ? it.lo : it.S, it.hi, it.S + a, it.neg});
		}
		iv.swap(next);
	}
}

static bool seed_geometry(const zc_params *p, int neff, int M, int flavour, std::vector<Interval> &iv, SeedConsts &s,
		int &NS, int64_t &rmin, int64_t &rmax) {
	const bool packed = fl_packed(flavour);
	NS = neff - M;
	if (NS < 0 || NS > SEED_MAX_NS) return false;
	seed_intervals(p, M, iv);
	int64_t wmin = INT64_MAX;
	rmin = INT64_MAX; rmax = INT64_MIN;
	for (const Interval &it : iv) {
		if (it.hi - it.lo < wmin) wmin = it.hi - it.lo;
		if (it.lo - it.S < rmin) rmin = it.lo - it.S;
		if (it.hi - 1 - it.S > rmax) rmax = it.hi - 1 - it.S;
	}
	int lgw = 0;
	while (((int64_t)2 << lgw) <= wmin) lgw++;		// largest W = 2^lgw <= wmin: at most one step per bucket
	if (lgw < 2) return false;
	if (lgw > 16) lgw = 16;
	const int LB = p->pw - 2 - lgw;				// log2(number of buckets)
	if (LB < 0 || LB > 15) return false;
	const size_t R = iv.size();
	const int nsp = (NS + 3) & ~3;
	const size_t nres = (size_t)(rmax - rmin + 1);
	// bytes per TD row slot: 16 (one int4 plane entry) or, packed, one signed byte per stage
	const int lgrow = (packed && NS <= 8) ? 3 : 4;
	const size_t b_t1 = (size_t)4 << LB, b_ts = (R * 4 + 15) & ~(size_t)15, b_t2 = R * (flavour == FL_DIRS ? 16 : 32),
		     b_td = packed ? ((nres << lgrow) + 15) & ~(size_t)15 : nres * (size_t)nsp * 4;
	const size_t total = b_t1 + b_ts + b_t2 + b_td;
	if (total + 16 > SEED_SMEM_LIMIT) return false;
	if ((R << lgw) >= ((uint64_t)1 << 32)) return false;
	std::memset(&s, 0, sizeof(s));
	s.M = M; s.lgw = lgw; s.R = (uint32_t)R;
	s.mul_q = (uint32_t)1 << (32 - p->pw);
	s.mul_u = (uint32_t)1 << (34 - p->pw);
	s.bsh = 32 - LB;
	s.ush = 34 - p->pw - lgrow;
	s.rsh = lgw + lgrow;
	s.lgrow = lgrow;
	if (LB < 1 || s.ush < 0 || ((uint64_t)R << (lgw + lgrow)) >= ((uint64_t)1 << 31)) return false;
	if (p->pw > 28 || 32 - p->pw - lgrow < 0) return false;	// residual reconstruction needs 2^(32-PW-lgrow)
	s.mul_r = (uint32_t)1 << (32 - p->pw - lgrow);
	s.res_bias = (int32_t)((uint64_t)rmin << (32 - p->pw));
	{
		const int D = p->ww - p->ow;
		s.rscale = 1.0f / (float)((uint32_t)1 << D);
		s.rbias = 12582912.0f - 12582912.0f / (float)((uint32_t)1 << D);
	}
	s.off_ts = (uint32_t)b_t1;
	s.off_t2 = (uint32_t)(b_t1 + b_ts);
	s.off_td = (uint32_t)(b_t1 + b_ts + b_t2);
	s.td_plane = (int32_t)(nres * 16);
	s.total_bytes = (uint32_t)((total + 15) & ~(size_t)15);
	for (int j = 0; j < SEED_MAX_NS; j++) s.sh[j] = (M + j + 1 > 31) ? 31 : (M + j + 1);
	return true;
}

static std::mutex g_seed_mu;
static std::vector<SeedPlan> g_seed_cache;
static uint64_t g_seed_clock = 0;

// cudaFree waits for the device to go idle, so kernels already enqueued on the tables finish first.
struct DevFree { void operator()(void *ptr) const { if (ptr) cudaFree(ptr); } };

// Builds (or finds) the plan for (p, constant vector, device).  Called with the device current.
static int seed_plan_get(const zc_params *p, const CoreConsts &c, int device, int flavour, cudaStream_t st,
		SeedPlan &out) {
	const bool packed = fl_packed(flavour);
	std::lock_guard<std::mutex> lk(g_seed_mu);
	for (SeedPlan &pl : g_seed_cache) {
		if (pl.device == device && pl.flavour == flavour && std::memcmp(&pl.p, p, sizeof(*p)) == 0 &&
		    std::memcmp(pl.x0c, c.cx, sizeof(pl.x0c)) == 0 && std::memcmp(pl.y0c, c.cy, sizeof(pl.y0c)) == 0) {
			pl.stamp = ++g_seed_clock;
			out = pl;
			return ZC_OK;
		}
	}
	SeedPlan pl;
	pl.p = *p; pl.device = device; pl.flavour = flavour; pl.stamp = ++g_seed_clock;
	std::memcpy(pl.x0c, c.cx, sizeof(pl.x0c));
	std::memcpy(pl.y0c, c.cy, sizeof(pl.y0c));
	std::vector<Interval> iv;
	int64_t rmin = 0, rmax = 0;
	bool ok = false;
	const int neff = c.neff;
	if (fl_dirs(flavour)) {
		if (neff >= DIRS_M) ok = seed_geometry(p, neff, DIRS_M, flavour, iv, pl.s, pl.NS, rmin, rmax);
	} else {
		for (int M = (neff < 13 ? neff : 13); M >= 6 && !ok; M--)
			ok = seed_geometry(p, neff, M, flavour, iv, pl.s, pl.NS, rmin, rmax);
		// IDP.2A multiplies the low 16 bits of x>>sh, y>>sh: exact when |x|,|y| < 2^(WW-1) and sh >= WW-16
		if (ok && flavour == FL_WORDS_DP && pl.s.M + 1 < p->ww - 16) ok = false;
	}
	if (ok) {
		const SeedConsts &s = pl.s;
		const int pshift = c.pshift;
		const size_t R = iv.size();
		std::vector<uint32_t> host(s.total_bytes / 4 + R, 0u);		// tables + representative phases
		uint32_t *t1 = host.data(), *ts = host.data() + s.off_ts / 4, *td = host.data() + s.off_td / 4;
		uint32_t *rep = host.data() + s.total_bytes / 4;
		const int64_t half = (int64_t)1 << (p->pw - 3), W = (int64_t)1 << s.lgw;
		const size_t nb = (size_t)1 << (p->pw - 2 - s.lgw);
		size_t r = 0;
		for (size_t b = 0; b < nb && ok; b++) {
			const int64_t b0 = -half + (int64_t)b * W;
			while (r + 1 < R && iv[r + 1].lo <= b0) r++;
			int64_t off = W;					// no step inside this bucket
			if (r + 1 < R && iv[r + 1].lo < b0 + W) {
				off = iv[r + 1].lo - b0;			// in [1, W-1]
				if (r + 2 < R && iv[r + 2].lo < b0 + W) ok = false;	// two steps: refuse
			}
			t1[b] = (uint32_t)(((uint64_t)r << s.lgw) + (uint64_t)(W - off) - ((uint64_t)b << s.lgw)) << s.lgrow;
		}
		if (flavour == FL_DIRS) {		// per-interval prefix directions, one signed byte per stage, 16-byte rows
			unsigned char *tp = reinterpret_cast<unsigned char *>(host.data()) + s.off_t2;
			for (size_t k = 0; k < R; k++)
				for (int j = 0; j < DIRS_M; j++)
					tp[k * 16 + j] = (unsigned char)(((iv[k].neg >> j) & 1u) ? 0xff : 0x01);
		}
		for (size_t k = 0; k < R && ok; k++) {
			ts[k] = (uint32_t)(int32_t)((iv[k].S + half + rmin) * ((int64_t)1 << s.lgrow));
			rep[k] = (uint32_t)((uint64_t)iv[k].lo << pshift);
		}
		const size_t nres = (size_t)(rmax - rmin + 1);
		for (int64_t res = rmin; res <= rmax && ok; res++) {
			int64_t ph = res;
			for (int j = 0; j < pl.NS; j++) {			// rtl/cordic.v:265-279, phase only
				const bool neg = ph < 0;
				if (packed)
					reinterpret_cast<unsigned char *>(td)[((size_t)(res - rmin) << s.lgrow) + j] = (unsigned char)(neg ? 0xff : 0x01);
				else if (flavour == FL_WORDS_DP)	// bytes 3..0 = {0, d, 0, -d}
					td[((size_t)(j >> 2) * nres + (size_t)(res - rmin)) * 4 + (j & 3)] = neg ? 0x00FF0001u : 0x000100FFu;
				else
					td[((size_t)(j >> 2) * nres + (size_t)(res - rmin)) * 4 + (j & 3)] = (uint32_t)(neg ? -1 : 1);
				ph += neg ? (int64_t)p->angle[s.M + j] : -(int64_t)p->angle[s.M + j];
			}
		}
		if (ok) {
			cudaError_t e = cudaMalloc(&pl.dev, host.size() * 4);
			if (e == cudaSuccess) pl.hold = std::shared_ptr<void>(pl.dev, DevFree());
			if (e == cudaSuccess) e = cudaMemcpyAsync(pl.dev, host.data(), host.size() * 4, cudaMemcpyHostToDevice, st);
			if (e == cudaSuccess && !fl_dirs(flavour)) {
				const uint32_t nthreads = 4u * (uint32_t)R;
				k_seed_fill_xy<<<(nthreads + 255) / 256, 256, 0, st>>>(
					reinterpret_cast<const uint32_t *>(pl.dev) + s.total_bytes / 4,
					reinterpret_cast<int2 *>(reinterpret_cast<char *>(pl.dev) + s.off_t2), (uint32_t)R, s.M, c);
				e = cudaGetLastError();
			}
			if (e == cudaSuccess) e = cudaStreamSynchronize(st);	// tables complete before any stream uses them
			if (e != cudaSuccess) {
				cudaGetLastError();
				return set_error(ZC_ECUDA, "seed table setup failed: %s", cudaGetErrorString(e));
			}
		}
	}
	pl.usable = ok;
	if (g_seed_cache.size() >= 16) {		// evict the least recently used plan
		size_t victim = 0;
		for (size_t k = 1; k < g_seed_cache.size(); k++)
			if (g_seed_cache[k].stamp < g_seed_cache[victim].stamp) victim = k;
		g_seed_cache.erase(g_seed_cache.begin() + victim);	// the tables go when the last user's copy does
	}
	g_seed_cache.push_back(pl);
	out = pl;
	return ZC_OK;
}

// cudaFuncSetAttribute once per (device, kernel, size): launches of a configured kernel then consist of the
// launch alone, which keeps them legal inside a stream capture.
static cudaError_t ensure_dynamic_smem(const void *kern, size_t smem) {
	static std::mutex mu;
	static std::vector<std::pair<std::pair<int, const void *>, size_t>> seen;
	int device = 0;
	cudaError_t e = cudaGetDevice(&device);
	if (e != cudaSuccess) return e;
	std::lock_guard<std::mutex> lk(mu);
	for (auto &it : seen)
		if (it.first.first == device && it.first.second == kern) {
			if (it.second >= smem) return cudaSuccess;
			e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if (e == cudaSuccess) it.second = smem;
			return e;
		}
	e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e == cudaSuccess) seen.push_back({{device, kern}, smem});
	return e;
}

// Drops the cached plans of `device` (all devices when negative).  The gate ring stays: 4 KB per device, and a
// launch already enqueued may still read its slot.
static void seed_trim(int device) {
	std::lock_guard<std::mutex> lk(g_seed_mu);
	for (size_t k = 0; k < g_seed_cache.size();) {
		if (device < 0 || g_seed_cache[k].device == device) g_seed_cache.erase(g_seed_cache.begin() + k);
		else k++;
	}
}

template <int SRC, int NS>
struct SeedTable {
	static cudaError_t launch(int ns, int tdm, int grid, size_t smem, cudaStream_t st, const uint32_t *ph, int2 *out, size_t nblocks,
			const CoreConsts &c, const SeedConsts &s, const uint4 *tables, const int *gate) {
		if (ns == NS) {
			// float rounding needs every register value to fit 1.5*2^23 +- 2^22 and a rounding core (D >= 2)
			const bool rf = c.do_round && c.wsh >= 9;
			typedef void (*kern_t)(const uint32_t *, int2 *, size_t, const CoreConsts, const SeedConsts, const uint4 *, const int *);
			kern_t kern;
			if (tdm == TD_TABLE) kern = rf ? k_rotate_seeded<NS, SRC, true, TD_TABLE> : k_rotate_seeded<NS, SRC, false, TD_TABLE>;
			else if (tdm == TD_TABLE_DP) kern = rf ? k_rotate_seeded<NS, SRC, true, TD_TABLE_DP> : k_rotate_seeded<NS, SRC, false, TD_TABLE_DP>;
			else if (tdm == TD_REGS) kern = rf ? k_rotate_seeded<NS, SRC, true, TD_REGS> : k_rotate_seeded<NS, SRC, false, TD_REGS>;
			else kern = rf ? k_rotate_seeded<NS, SRC, true, TD_PACKED> : k_rotate_seeded<NS, SRC, false, TD_PACKED>;
			cudaError_t e = ensure_dynamic_smem((const void *)kern, smem);
			if (e != cudaSuccess) return e;
			kern<<<grid, 1024, smem, st>>>(ph, out, nblocks, c, s, tables, gate);
			return cudaGetLastError();
		}
		return SeedTable<SRC, NS - 1>::launch(ns, tdm, grid, smem, st, ph, out, nblocks, c, s, tables, gate);
	}
};
template <int SRC>
struct SeedTable<SRC, -1> {
	static cudaError_t launch(int, int, int, size_t, cudaStream_t, const uint32_t *, int2 *, size_t, const CoreConsts &,
			const SeedConsts &, const uint4 *, const int *) { return cudaErrorInvalidValue; }
};

// ---- per-sample input vectors: every stage in registers, every direction from a table ----------------------
// With (x, y) varying per sample the x/y table is gone, but the directions still depend on the phase alone: the
// DIRS_M prefix directions come from the sample's interval (16-byte row of signed bytes), the rest from the
// residual row as in the byte flavour above.  A stage is then PRMT + IMAD.MOV + 2 SHF + 2 IMAD (6 issue slots
// instead of 8), and the phase recursion is gone altogether.
template <int NS, int J = 0>
struct DirStages {
	static __device__ __forceinline__ void run(int &x, int &y, const uint32_t (&tp)[4], const uint32_t (&td)[4]) {
		constexpr int S = (J + 1 > 31) ? 31 : (J + 1);
		const int d = (J < DIRS_M) ? sext_byte(tp[J >> 2], J & 3) : sext_byte(td[(J - DIRS_M) >> 2], (J - DIRS_M) & 3);
		const int nd = ineg(d);
		const int sy = y >> S, sx = x >> S;
		const int x1 = imad(sy, nd, x);
		const int y1 = imad(sx, d, y);
		x = x1; y = y1;
		DirStages<NS, J + 1>::run(x, y, tp, td);
	}
};
template <int NS>
struct DirStages<NS, DIRS_M + NS> {
	static __device__ __forceinline__ void run(int &, int &, const uint32_t (&)[4], const uint32_t (&)[4]) {}
};

__device__ __forceinline__ int2 ldg_stream64(const int2 *p) {
	int2 r;
	asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
	return r;
}

template <int NS, int SRC, bool RF>
__global__ void __launch_bounds__(1024, 1)
k_rotate_dirs(const uint32_t *__restrict__ phase, const int2 *__restrict__ xyin, int2 *__restrict__ xyout,
		size_t nblocks, const __grid_constant__ CoreConsts c, const __grid_constant__ SeedConsts s,
		const uint4 *__restrict__ tables) {
	extern __shared__ __align__(128) unsigned char smem[];
	const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
	const uint32_t mbar = sbase + s.total_bytes;
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(s.total_bytes) : "memory");
		const char *src = reinterpret_cast<const char *>(tables);
		for (uint32_t off = 0; off < s.total_bytes; off += 32768u) {
			const uint32_t len = (s.total_bytes - off < 32768u) ? (s.total_bytes - off) : 32768u;
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				:: "r"(sbase + off), "l"(src + off), "r"(len), "r"(mbar) : "memory");
		}
	}
	__syncthreads();
	{
		uint32_t done = 0;
		while (!done) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
				"selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar) : "memory");
		}
	}
	const uint32_t *const T1 = reinterpret_cast<const uint32_t *>(smem);
	const int32_t *const TS = reinterpret_cast<const int32_t *>(smem + s.off_ts);
	const uint4 *const TP = reinterpret_cast<const uint4 *>(smem + s.off_t2);
	const unsigned char *const TD = smem + s.off_td;
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t nwarps = gridDim.x * (blockDim.x >> 5), nblk = (uint32_t)nblocks;	// see k_rotate_seeded
	// software prefetch of the next block's inputs (12 registers), as in k_rotate_seeded: the loads of block b+W are in
	// flight while block b runs its 20 stages
	uint32_t blk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	uint32_t pph[4] = {0, 0, 0, 0};
	int2 pv[4] = {make_int2(0, 0), make_int2(0, 0), make_int2(0, 0), make_int2(0, 0)};
	if (blk < nblk) {
		const size_t b0 = ((size_t)blk << 7) + lane;
#pragma unroll
		for (int k = 0; k < 4; k++) pv[k] = ldg_stream64(xyin + b0 + (k << 5));
		if (SRC != SRC_MIX) {
#pragma unroll
			for (int k = 0; k < 4; k++) pph[k] = ldg_stream32(phase + b0 + (k << 5));
		}
	}
	for (; blk < nblk; blk += nwarps) {
		const size_t base = ((size_t)blk << 7) + lane;
		uint32_t ph[4];
		int2 v[4];
#pragma unroll
		for (int k = 0; k < 4; k++) v[k] = pv[k];
		if (SRC == SRC_MIX) {
			const uint32_t p0 = c.nco_phase0 + (c.nco_n0 + (uint32_t)base) * c.nco_step;
#pragma unroll
			for (int k = 0; k < 4; k++) ph[k] = (p0 + (uint32_t)(k << 5) * c.nco_step) >> c.pshift;
		} else {
#pragma unroll
			for (int k = 0; k < 4; k++) ph[k] = pph[k];
		}
		if (blk + nwarps < nblk) {
			const size_t nb = ((size_t)(blk + nwarps) << 7) + lane;
#pragma unroll
			for (int k = 0; k < 4; k++) pv[k] = ldg_stream64(xyin + nb + (k << 5));
			if (SRC != SRC_MIX) {
#pragma unroll
				for (int k = 0; k < 4; k++) pph[k] = ldg_stream32(phase + nb + (k << 5));
			}
		}
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const uint32_t tq = (uint32_t)imad((int)ph[k], (int)s.mul_q, 0x20000000);
			const uint32_t tu = (uint32_t)imad((int)ph[k], (int)s.mul_u, (int)0x80000000u);
			const uint32_t ur = tu >> s.ush;
			const uint32_t rank = (T1[tu >> s.bsh] + ur) >> s.rsh;
			const unsigned char *row = TD + (int)(ur - (uint32_t)TS[rank]);
			// rtl/cordic.v:85-86 (extend) and :131-188 (quarter turn selected by the octant)
			const int ex = (v[k].x << c.in_shl) >> c.in_shr, ey = (v[k].y << c.in_shl) >> c.in_shr;
			int x, y;
			quarter_turn((int)(tq >> 30), ex, ey, x, y);
			const uint4 tpv = TP[rank];
			const uint32_t tp[4] = {tpv.x, tpv.y, tpv.z, tpv.w};
			uint32_t td[4] = {0, 0, 0, 0};
			if (NS > 0 && NS <= 8) {
				const int2 w = *reinterpret_cast<const int2 *>(row);
				td[0] = (uint32_t)w.x; td[1] = (uint32_t)w.y;
			} else if (NS > 8) {
				const int4 w = *reinterpret_cast<const int4 *>(row);
				td[0] = (uint32_t)w.x; td[1] = (uint32_t)w.y; td[2] = (uint32_t)w.z; td[3] = (uint32_t)w.w;
			}
			DirStages<NS>::run(x, y, tp, td);
			const int ox = RF ? round_out_fma(x, s) : round_out(x, c);
			const int oy = RF ? round_out_fma(y, s) : round_out(y, c);
			stg_stream64(xyout + base + (k << 5), make_int2(ox, oy));
		}
	}
}

template <int SRC, int NS>
struct DirsTable {
	static cudaError_t launch(int ns, int grid, size_t smem, cudaStream_t st, const uint32_t *ph, const int2 *xin, int2 *out,
			size_t nblocks, const CoreConsts &c, const SeedConsts &s, const uint4 *tables) {
		if (ns == NS) {
			const bool rf = c.do_round && c.wsh >= 9;
			typedef void (*kern_t)(const uint32_t *, const int2 *, int2 *, size_t, const CoreConsts, const SeedConsts, const uint4 *);
			kern_t kern = rf ? (kern_t)k_rotate_dirs<NS, SRC, true> : (kern_t)k_rotate_dirs<NS, SRC, false>;
			cudaError_t e = ensure_dynamic_smem((const void *)kern, smem);
			if (e != cudaSuccess) return e;
			kern<<<grid, 1024, smem, st>>>(ph, xin, out, nblocks, c, s, tables);
			return cudaGetLastError();
		}
		return DirsTable<SRC, NS - 1>::launch(ns, grid, smem, st, ph, xin, out, nblocks, c, s, tables);
	}
};
template <int SRC>
struct DirsTable<SRC, -1> {
	static cudaError_t launch(int, int, size_t, cudaStream_t, const uint32_t *, const int2 *, int2 *, size_t, const CoreConsts &,
			const SeedConsts &, const uint4 *) { return cudaErrorInvalidValue; }
};

// Per-sample (x,y): tries the table-directed kernel on the first floor(n/128)*128 samples.
template <int SRC>
static int dirs_rotate_try(const zc_params *p, const CoreConsts &c, const uint32_t *phase, const int32_t *xy_in,
		int32_t *xy_out, size_t n, int device, int sms, cudaStream_t st, uint32_t flags, size_t &done, int &launches) {
	done = 0; launches = 0;
	const size_t nblocks = n >> 7;
	if (nblocks == 0 || nblocks > SEED_MAX_BLOCKS) return ZC_OK;
	if (!(flags & ZC_F_FORCE_SEED) && n < ((size_t)1 << 20)) return ZC_OK;
	if (c.neff < DIRS_M || p->pw < 12) return ZC_OK;
	CoreConsts key = c;		// the plan does not depend on the input vector
	for (int q = 0; q < 4; q++) key.cx[q] = key.cy[q] = 0;
	SeedPlan pl;
	int rc = seed_plan_get(p, key, device, FL_DIRS, st, pl);
	if (rc != ZC_OK) return rc;
	if (!pl.usable) return ZC_OK;
	cudaError_t e = DirsTable<SRC, SEED_MAX_NS>::launch(pl.NS, sms, pl.s.total_bytes + 16, st, phase, (const int2 *)xy_in,
		(int2 *)xy_out, nblocks, c, pl.s, (const uint4 *)pl.dev);
	if (e != cudaSuccess)
		return set_error(ZC_ECUDA, "launch of k_rotate_dirs failed: %s", cudaGetErrorString(e));
	launches = 1;
	done = nblocks << 7;
	return ZC_OK;
}

// A small ring of device-side gate words per device for the auto-selected launches (stream-ordered use; a slot
// is reused 1024 seeded calls later).
static std::mutex g_gate_mu;
static int *g_gate_ring[64] = {};
static unsigned g_gate_next[64] = {};
constexpr unsigned GATE_RING = 1024;

static int gate_slot(int device, int **slot) {
	std::lock_guard<std::mutex> lk(g_gate_mu);
	if (!g_gate_ring[device]) {
		cudaError_t e = cudaMalloc((void **)&g_gate_ring[device], GATE_RING * sizeof(int));
		if (e != cudaSuccess) return set_error(ZC_ECUDA, "cudaMalloc(gate ring): %s", cudaGetErrorString(e));
	}
	*slot = g_gate_ring[device] + (g_gate_next[device]++ % GATE_RING);
	return ZC_OK;
}

// Tries the seeded path on the first floor(n/128)*128 samples.  done=0 means "not applicable here"; launches
// reports how many kernels were enqueued.
template <int SRC>
static int seeded_rotate_try(const zc_params *p, const CoreConsts &c, const uint32_t *phase, int32_t *xy_out,
		size_t n, int device, int sms, cudaStream_t st, uint32_t flags, size_t &done, int &launches) {
	done = 0; launches = 0;
	const size_t nblocks = n >> 7;
	if (nblocks == 0 || nblocks > SEED_MAX_BLOCKS) return ZC_OK;
	if (!(flags & ZC_F_FORCE_SEED) && n < ((size_t)1 << 20)) return ZC_OK;	// not worth the table load
	if (c.neff < 6 || p->pw < 12) return ZC_OK;
	// Which flavour of direction table.  The caller's flag wins.  For the NCO the host knows the pattern: byte rows
	// when neighbouring lanes land more than one table row apart (|step| >= 2 phase LSBs).  For a phase stream of
	// 4 Mi samples or more, a probe kernel decides on the device and both flavours are enqueued behind it.
	const bool forced = (flags & (ZC_F_SEED_REGS | ZC_F_SEED_PACKED | ZC_F_SEED_WORDS)) != 0;
	int tdm = (flags & ZC_F_SEED_REGS) ? TD_REGS : (flags & ZC_F_SEED_PACKED) ? TD_PACKED : TD_TABLE;
	bool probe = false;
	if (!forced) {
		if (SRC == SRC_NCO) {
			const int32_t sstep = (int32_t)c.nco_step;
			const uint32_t mag = (uint32_t)(sstep < 0 ? -(int64_t)sstep : (int64_t)sstep);
			if ((mag >> c.pshift) >= 2u) tdm = TD_PACKED;
		} else if (n >= ((size_t)1 << 22)) {
			probe = true;
		}
	}
	SeedPlan pl, pl2;
	int rc = ZC_OK;
	if (tdm == TD_TABLE && !(flags & ZC_F_NO_DP2A)) {	// the word table as IDP.2A multipliers, when the shifts allow
		if ((rc = seed_plan_get(p, c, device, FL_WORDS_DP, st, pl)) != ZC_OK) return rc;
		if (pl.usable) tdm = TD_TABLE_DP;
	}
	if (tdm != TD_TABLE_DP) {
		if ((rc = seed_plan_get(p, c, device, tdm == TD_PACKED ? FL_PACKED : FL_WORDS, st, pl)) != ZC_OK) return rc;
		if (!pl.usable) return ZC_OK;
	}
	int *gate = nullptr;
	if (probe) {
		if ((rc = seed_plan_get(p, c, device, FL_PACKED, st, pl2)) != ZC_OK) return rc;
		if (!pl2.usable) probe = false;
	}
	if (probe) {
		if ((rc = gate_slot(device, &gate)) != ZC_OK) return rc;
		k_seed_probe<<<1, 256, 0, st>>>(phase, n, c.pshift, 1, gate);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) return set_error(ZC_ECUDA, "launch of k_seed_probe failed: %s", cudaGetErrorString(e));
		launches++;
	}
	cudaError_t e = SeedTable<SRC, SEED_MAX_NS>::launch(pl.NS, tdm, sms, pl.s.total_bytes + 16, st, phase, (int2 *)xy_out,
		nblocks, c, pl.s, (const uint4 *)pl.dev, gate);
	if (e == cudaSuccess) launches++;
	if (e == cudaSuccess && probe) {
		e = SeedTable<SRC, SEED_MAX_NS>::launch(pl2.NS, TD_PACKED, sms, pl2.s.total_bytes + 16, st, phase, (int2 *)xy_out,
			nblocks, c, pl2.s, (const uint4 *)pl2.dev, gate);
		if (e == cudaSuccess) launches++;
	}
	if (e != cudaSuccess)
		return set_error(ZC_ECUDA, "launch of k_rotate_seeded failed: %s", cudaGetErrorString(e));
	done = nblocks << 7;
	return ZC_OK;
}

} // namespace zc
#endif
