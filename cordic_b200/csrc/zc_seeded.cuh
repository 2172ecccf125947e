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
//   TD[NSP/4][nres] int4 residual -> d_M .. d_{M+NS-1} as +1/-1 words (or IDP.2A multiplier words {0, d, 0, -d}), in planes
//                         of four stages so that consecutive residuals (a phase sweep) read consecutive 16-byte slots: no
//                         bank conflicts.  Byte flavours: one signed byte per stage, 8- or 16-byte rows.
//   Merged byte flavour (FL_PACKED_M): the T2 words are (x << (32-WW) | low byte of TS), (y << (32-WW) | high byte of TS) with
//                         TS taken modulo 2^16, and the TS table is not read: three dependent lookups instead of four.
//   Per-sample-vector flavours (FL_DIRS*): T2 holds one 16-byte row per interval instead -- 12 prefix directions as signed
//                         bytes and TS in the last word.
// Sample-to-lane mapping (MAP_BLOCK): a warp owns 128 consecutive samples per iteration and lane l takes samples
// l, l+32, l+64, l+96 of them, so that for a phase sweep neighbouring lanes read neighbouring table rows
// (or the same row: a broadcast) and every global access is a fully coalesced 128/256-byte row.  MAP_COMB (NCO with a
// near-period K): see k_rotate_seeded.
// What bounds these kernels is the LSU data pipe: wavefronts of 128 bytes of distinct data per access, identical
// addresses served once, 128-bit accesses split by quarter-warp ("The LSU budget").
#ifndef ZC_SEEDED_CUH
#define ZC_SEEDED_CUH

#include "zc_seedplan.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <utility>
#include <vector>

namespace zc {

// Suffix stages whose -d is formed as d ^ ~1 (LOP3, ALU pipe) instead of a negation (IMAD.MOV/IADD3): one stage in
// four balances the heavy FMA pipe (81 % busy with every negation on it) against the ALU pipe (70 %); measured on
// the cfg1 sweep, 20 steps: none 456-457, 0x22 463-464, 0x88 465-467, 0xAA 461, 0xFF 460 Gsamples/s.
#ifndef ZC_XNEG_MASK
#define ZC_XNEG_MASK 0x8888
#endif
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

// TD_PACKED_M: TD_PACKED on a FL_PACKED_M plan (the TS lookup folded into the (x, y) record)
enum { TD_TABLE = PROBE_LOCAL, TD_REGS = 1, TD_PACKED = PROBE_SCATTERED, TD_TABLE_DP = 3, TD_PACKED_M = 4 };
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
	return r;
}
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

// Convergent rounding on the FMA pipe (the ALU pipe is this kernel's bottleneck).  For |v| < 2^22 the word
// 0x4B400000+v IS the float 1.5*2^23+v; one fused multiply-add forms 1.5*2^23 + v/2^D exactly and rounds it
// to the nearest integer, ties to even -- which is precisely rtl/cordic.v:290-295 (add 2^(D-1) when bit D is
// set, 2^(D-1)-1 when it is clear, then drop D bits).  Only used when WW <= 23 and the core rounds (D >= 2).
__device__ __forceinline__ int round_out_fma(int v, const SeedConsts &s) {
	const float r = __fmaf_rn(__int_as_float(v + 0x4B400000), s.rscale, s.rbias);
	return __float_as_int(r) - 0x4B400000;
}

// 256-bit store (sm_100: STG.256), 32-byte aligned
__device__ __forceinline__ void stg256(int2 *p, int a, int b, int c, int d, int e, int f, int g, int h) {
	asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d),
		"r"(e), "r"(f), "r"(g), "r"(h) : "memory");
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
//
// MAP: which samples a lane takes.
//   MAP_BLOCK  a warp owns 128 consecutive samples per iteration, lane l takes l, l+32, l+64, l+96 (see the file header).
//   MAP_COMB   (NCO only) for a phase accumulator whose step scatters consecutive samples over the circle.  The host
//              finds a run length K with K*step = delta (mod 2^32), |delta| a fraction of a phase LSB (comb_search).
//              The stream is cut into tiles of 8K samples; lane l = (a = l & 7, b = l >> 3) works in run a of the tile,
//              [a*K, (a+1)*K), and per iteration the warp takes one 16-sample chunk of each of the 8 runs: lane (a, b)
//              computes samples 16m + 4b + {0, 1, 2, 3} of run a.  At every instruction the 8 lanes of a quarter-warp (the
//              unit in which a 128-bit shared-memory read is served) then hold phases delta apart -- mostly the SAME
//              table rows, which the LSU serves once.  Before storing, the results are transposed with warp shuffles so
//              that a quarter-warp holds two whole 128-byte lines (runs 2h and 2h+1) and writes them with one 256-bit
//              store per lane.  Round 1's comb gave every LANE its own run and died on scattered 8/16-byte stores.
// OUT16: outputs packed as (int16 o_xval, int16 o_yval) in one word (cores with OW <= 16; zc_rotate_const_o16).
enum { MAP_BLOCK = 0, MAP_COMB = 1 };
struct CombConsts {
	uint32_t K;		// run length, a multiple of 4: every run starts on a 32-byte boundary (one 256-bit store per lane)
	uint32_t cpr;		// 16-sample chunks per run: ceil(K/16)
	uint32_t nunits;	// tiles * cpr
	uint32_t tile;		// 8*K samples
	uint32_t tile_step;	// 8*K*step (mod 2^32)
	uint32_t chunk_step;	// 16*step
	uint32_t dt, dm;	// (number of warps in the grid) / cpr and % cpr: how (tile, chunk) advances per iteration
};

template <int NS, int SRC, bool RF, int TDM, int MAP, bool OUT16>
__global__ void __launch_bounds__(1024, 1)
k_rotate_seeded(const uint32_t *__restrict__ phase, int2 *__restrict__ xyout, size_t nblocks,
		const __grid_constant__ CoreConsts c, const __grid_constant__ SeedConsts s,
		const uint4 *__restrict__ tables, const int probe_lim, const __grid_constant__ CombConsts cb) {
	extern __shared__ __align__(128) unsigned char smem[];
	// Auto-selection: both table flavours are enqueued back to back and every CTA of both evaluates the same probe of
	// the same input (probe_local: a pure function of the phase stream); the flavour the verdict does not name returns
	// here, before touching shared memory.  No device-side state is shared between calls, streams or graph replays.
	if (probe_lim >= 0 && probe_local(phase, nblocks << 7, c.pshift, probe_lim) !=
			(TDM == TD_TABLE_DP ? TD_TABLE : TDM == TD_PACKED_M ? TD_PACKED : TDM)) return;
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
	auto body = [&](uint32_t (&tin)[4], const uint32_t reload, auto &&store) {
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
		int x[4], y[4];
		uint32_t row16[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const uint32_t u16 = tu[k] >> s.ush;				// 16 * reduced phase
			const uint32_t rank = (T1[tu[k] >> s.bsh] + u16) >> s.rsh;	// carries past the step, if any
			const int2 xy = T2[__funnelshift_l(tq[k], rank, 2)];		// row rank*4 + quarter turn
			if (TDM == TD_PACKED_M) {
				// record words: (x << recsh | low byte of TS), (y << recsh | high byte of TS), TS taken modulo 2^16 -- the
				// TD table is smaller than that, so the difference below is exact after the mask; one lookup fewer
				x[k] = xy.x >> s.recsh; y[k] = xy.y >> s.recsh;
				row16[k] = (u16 - prmt((uint32_t)xy.x, (uint32_t)xy.y, 0x7640u)) & 0xffffu;
			} else {
				x[k] = xy.x; y[k] = xy.y;
				row16[k] = u16 - (uint32_t)TS[rank];	// residual after M stages, as the byte offset of its TD row
			}
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
				} else if (TDM == TD_PACKED || TDM == TD_PACKED_M) {
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
			store(k, ox, oy);
		}
	};
	// MAP_BLOCK stores: sample k of the lane goes to block*128 + 32k + lane (a coalesced 256-byte row per instruction;
	// 128 bytes with packed outputs)
	auto store_block = [&](const uint32_t blk) {
		return [=](const int k, const int ox, const int oy) {
			if (OUT16) {
				uint32_t *const dst = reinterpret_cast<uint32_t *>(xyout) + ((size_t)blk << 7) + lane;
				asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(dst + (k << 5)),
					"r"(((uint32_t)ox & 0xffffu) | ((uint32_t)oy << 16)) : "memory");
			} else {
				stg_stream64(xyout + ((size_t)blk << 7) + lane + (k << 5), make_int2(ox, oy));
			}
		};
	};
	if (SRC == SRC_NCO && MAP == MAP_COMB) {
		// compute layout: lane (a = l & 7, b = l >> 3) takes samples 16m + 4b + {0,1,2,3} of run a -- the 8 lanes of a
		// quarter-warp sit in 8 different runs, delta apart in phase.  store layout: lane (a' = l >> 2, b' = l & 3) holds the
		// 32 bytes of run a', samples 16m + 4b' .. +3 -- a quarter-warp writes two full 128-byte lines with one 256-bit
		// store per lane.  Eight shuffles per 128 samples move the results from one layout to the other (stored straight
		// from the compute layout, every lane of a quarter-warp hit a different line: 32 LSU wavefronts per store
		// instruction, profiles/r2_ncu_comb_untransposed.md -- the data pipe was 95 % busy, 54 % of it those stores).
		const uint32_t a = lane & 7u, b = lane >> 3;
		const uint32_t lane_off = a * cb.K + 4u * b;		// the lane's first sample within (tile, chunk 0)
		const uint32_t lane_phase = c.nco_phase0 + (c.nco_n0 + lane_off) * c.nco_step;
		const uint32_t s1 = c.nco_step, s2 = 2u * s1, s3 = 3u * s1;
		const uint32_t src_lane = (lane >> 2) + 8u * (lane & 3u);	// who computed what this lane stores
		const uint32_t st_off = (lane >> 2) * cb.K + 4u * (lane & 3u);
		uint32_t t = blk / cb.cpr, m = blk - t * cb.cpr;	// unit = (tile t, chunk m); warp w takes units w, w+W, ...
		for (; blk < cb.nunits; blk += nwarps) {
			const uint32_t base = lane_phase + t * cb.tile_step + m * cb.chunk_step;
			uint32_t tin[4] = {base >> c.pshift, (base + s1) >> c.pshift, (base + s2) >> c.pshift, (base + s3) >> c.pshift};
			int r[8];
			body(tin, 0u, [&](const int k, const int ox, const int oy) { r[2 * k] = ox; r[2 * k + 1] = oy; });
#pragma unroll
			for (int i = 0; i < 8; i++) r[i] = __shfl_sync(0xffffffffu, r[i], (int)src_lane);
			const uint32_t j0 = 16u * m + 4u * (lane & 3u);
			int2 *const dst = xyout + ((size_t)t * cb.tile + st_off + 16u * m);
			if (j0 < cb.K) stg256(dst, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);	// K % 4 == 0: all four or none, 32-byte aligned
			m += cb.dm; t += cb.dt;
			if (m >= cb.cpr) { m -= cb.cpr; t++; }
		}
	} else if (SRC == SRC_NCO) {
		for (; blk < nblk; blk += nwarps) {
			uint32_t tin[4];
			const uint32_t base = c.nco_phase0 + (c.nco_n0 + (blk << 7) + lane) * c.nco_step;
#pragma unroll
			for (int k = 0; k < 4; k++) tin[k] = (base + (uint32_t)(k << 5) * c.nco_step) >> c.pshift;
			body(tin, nblk, store_block(blk));
		}
	} else {
		uint32_t pin[4] = {0, 0, 0, 0};
		if (blk < nblk) {
#pragma unroll
			for (int k = 0; k < 4; k++) pin[k] = ldg_stream32(phase + ((size_t)blk << 7) + (k << 5) + lane);
		}
		for (; blk < nblk; blk += nwarps) body(pin, blk + nwarps, store_block(blk));
	}
}

struct SeedLaunch {		// what a launch needs besides the template selectors
	int grid; size_t smem; cudaStream_t st; const uint32_t *ph; void *out; size_t nblocks;
	const CoreConsts *c; const SeedConsts *s; const uint4 *tables; int probe_lim; const CombConsts *cb;
};

template <int SRC, int NS>
struct SeedTable {
	// Instantiated combinations: every TDM for the block mapping with 32-bit outputs; the comb mapping only exists for
	// the word tables, packed outputs for everything but the register-recursion flavour (an A/B switch).
	template <int MAP, bool OUT16>
	static cudaError_t launch(int ns, int tdm, const SeedLaunch &L) {
		if (ns == NS) {
			// float rounding needs every register value to fit 1.5*2^23 +- 2^22 and a rounding core (D >= 2)
			const bool rf = L.c->do_round && L.c->wsh >= 9;
			typedef void (*kern_t)(const uint32_t *, int2 *, size_t, const CoreConsts, const SeedConsts, const uint4 *, int,
				const CombConsts);
			kern_t kern = nullptr;
#define ZC_PICK(T) (rf ? (kern_t)k_rotate_seeded<NS, SRC, true, T, MAP, OUT16> : (kern_t)k_rotate_seeded<NS, SRC, false, T, MAP, OUT16>)
			if (tdm == TD_TABLE) kern = ZC_PICK(TD_TABLE);
			else if (tdm == TD_TABLE_DP) kern = ZC_PICK(TD_TABLE_DP);
			if constexpr (MAP == MAP_BLOCK) {
				if (tdm == TD_PACKED) kern = ZC_PICK(TD_PACKED);
				if (tdm == TD_PACKED_M) kern = ZC_PICK(TD_PACKED_M);
				if constexpr (!OUT16) { if (tdm == TD_REGS) kern = ZC_PICK(TD_REGS); }
			}
#undef ZC_PICK
			if (!kern) return cudaErrorInvalidValue;
			cudaError_t e = ensure_dynamic_smem((const void *)kern, L.smem);
			if (e != cudaSuccess) return e;
			kern<<<L.grid, 1024, L.smem, L.st>>>(L.ph, (int2 *)L.out, L.nblocks, *L.c, *L.s, L.tables, L.probe_lim,
				L.cb ? *L.cb : CombConsts{});
			return cudaGetLastError();
		}
		return SeedTable<SRC, NS - 1>::template launch<MAP, OUT16>(ns, tdm, L);
	}
};
template <int SRC>
struct SeedTable<SRC, -1> {
	template <int MAP, bool OUT16>
	static cudaError_t launch(int, int, const SeedLaunch &) { return cudaErrorInvalidValue; }
};

// ---- per-sample input vectors: every stage in registers, every direction from a table ----------------------
// With (x, y) varying per sample the x/y table is gone, but the directions still depend on the phase alone: the
// DIRS_M prefix directions come from the sample's interval (16-byte row of signed bytes), the rest from the
// residual row as in the byte flavour above.  A stage is then PRMT + IMAD.MOV + 2 SHF + 2 IMAD (6 issue slots
// instead of 8), and the phase recursion is gone altogether.
// TDW: the suffix stages (J >= DIRS_M) take their directions as IDP.2A multiplier words {0, d, 0, -d} from word planes
// (as k_rotate_seeded's TD_TABLE_DP: 4 issue slots instead of 6; conflict-free when neighbouring samples have
// neighbouring phases, which a probe of the phase stream establishes per call).
template <int NS, bool TDW, int J = 0>
struct DirStages {
	static __device__ __forceinline__ void run(int &x, int &y, const uint32_t (&tp)[4], const uint32_t (&td)[4],
			const int (&wd)[SEED_MAX_NS]) {
		constexpr int S = (J + 1 > 31) ? 31 : (J + 1);
		const int sy = y >> S, sx = x >> S;
		int x1, y1;
		if (TDW && J >= DIRS_M) {
			x1 = dp2a_lo(sy, wd[J - DIRS_M], x);
			y1 = dp2a_hi(sx, wd[J - DIRS_M], y);
		} else {
			const int d = (J < DIRS_M) ? sext_byte(tp[J >> 2], J & 3) : sext_byte(td[(J - DIRS_M) >> 2], (J - DIRS_M) & 3);
			const int nd = ineg(d);
			x1 = imad(sy, nd, x);
			y1 = imad(sx, d, y);
		}
		x = x1; y = y1;
		DirStages<NS, TDW, J + 1>::run(x, y, tp, td, wd);
	}
};
template <int NS, bool TDW>
struct DirStages<NS, TDW, DIRS_M + NS> {
	static __device__ __forceinline__ void run(int &, int &, const uint32_t (&)[4], const uint32_t (&)[4], const int (&)[SEED_MAX_NS]) {}
};

__device__ __forceinline__ int2 ldg_stream64(const int2 *p) {
	int2 r;
	asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
	return r;
}

template <int NS, int SRC, bool RF, bool TDW>
__global__ void __launch_bounds__(1024, 1)
k_rotate_dirs(const uint32_t *__restrict__ phase, const int2 *__restrict__ xyin, int2 *__restrict__ xyout,
		size_t nblocks, const __grid_constant__ CoreConsts c, const __grid_constant__ SeedConsts s,
		const uint4 *__restrict__ tables, const int probe_lim) {
	extern __shared__ __align__(128) unsigned char smem[];
	// auto-selection as in k_rotate_seeded: the word-suffix and the byte-suffix kernels are both enqueued, both evaluate
	// the same probe of the same phase stream, one proceeds
	if (probe_lim >= 0 && probe_local(phase, nblocks << 7, c.pshift, probe_lim) != (TDW ? PROBE_LOCAL : PROBE_SCATTERED)) return;
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
	const uint4 *const TP = reinterpret_cast<const uint4 *>(smem + s.off_t2);
	const unsigned char *const TD = smem + s.off_td;
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t tdbase[SEED_MAX_NS / 4];		// one uniform shared-window base per word plane (see k_rotate_seeded)
#pragma unroll
	for (int j = 0; j < SEED_MAX_NS / 4; j++)
		asm("mov.b32 %0, %1;" : "=r"(tdbase[j]) : "r"(sbase + s.off_td + (uint32_t)(j * s.td_plane)));
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
			const uint4 tpv = TP[rank];		// 12 prefix directions (signed bytes) + the interval's row offset in the last word
			const uint32_t row16 = ur - tpv.w;	// byte offset of the residual's row (or 16-byte plane slot)
			const unsigned char *row = TD + (int)row16;
			// rtl/cordic.v:85-86 (extend) and :131-188 (quarter turn selected by the octant)
			const int ex = (v[k].x << c.in_shl) >> c.in_shr, ey = (v[k].y << c.in_shl) >> c.in_shr;
			int x, y;
			quarter_turn((int)(tq >> 30), ex, ey, x, y);
			const uint32_t tp[4] = {tpv.x, tpv.y, tpv.z, tpv.w};
			uint32_t td[4] = {0, 0, 0, 0};
			int wd[SEED_MAX_NS];
			if (TDW) {
#pragma unroll
				for (int j = 0; j < NS; j += 4) {
					const int4 dv = lds128(row16 + tdbase[j >> 2]);
					wd[j] = dv.x;
					if (j + 1 < SEED_MAX_NS) wd[j + 1] = dv.y;
					if (j + 2 < SEED_MAX_NS) wd[j + 2] = dv.z;
					if (j + 3 < SEED_MAX_NS) wd[j + 3] = dv.w;
				}
			} else if (NS > 0 && NS <= 8) {
				const int2 w = *reinterpret_cast<const int2 *>(row);
				td[0] = (uint32_t)w.x; td[1] = (uint32_t)w.y;
			} else if (NS > 8) {
				const int4 w = *reinterpret_cast<const int4 *>(row);
				td[0] = (uint32_t)w.x; td[1] = (uint32_t)w.y; td[2] = (uint32_t)w.z; td[3] = (uint32_t)w.w;
			}
			DirStages<NS, TDW>::run(x, y, tp, td, wd);
			const int ox = RF ? round_out_fma(x, s) : round_out(x, c);
			const int oy = RF ? round_out_fma(y, s) : round_out(y, c);
			stg_stream64(xyout + base + (k << 5), make_int2(ox, oy));
		}
	}
}

template <int SRC, int NS>
struct DirsTable {
	static cudaError_t launch(int ns, bool tdw, int grid, size_t smem, cudaStream_t st, const uint32_t *ph, const int2 *xin, int2 *out,
			size_t nblocks, const CoreConsts &c, const SeedConsts &s, const uint4 *tables, int probe_lim) {
		if (ns == NS) {
			const bool rf = c.do_round && c.wsh >= 9;
			typedef void (*kern_t)(const uint32_t *, const int2 *, int2 *, size_t, const CoreConsts, const SeedConsts, const uint4 *, int);
			kern_t kern = tdw ? (rf ? (kern_t)k_rotate_dirs<NS, SRC, true, true> : (kern_t)k_rotate_dirs<NS, SRC, false, true>)
					  : (rf ? (kern_t)k_rotate_dirs<NS, SRC, true, false> : (kern_t)k_rotate_dirs<NS, SRC, false, false>);
			cudaError_t e = ensure_dynamic_smem((const void *)kern, smem);
			if (e != cudaSuccess) return e;
			kern<<<grid, 1024, smem, st>>>(ph, xin, out, nblocks, c, s, tables, probe_lim);
			return cudaGetLastError();
		}
		return DirsTable<SRC, NS - 1>::launch(ns, tdw, grid, smem, st, ph, xin, out, nblocks, c, s, tables, probe_lim);
	}
};
template <int SRC>
struct DirsTable<SRC, -1> {
	static cudaError_t launch(int, bool, int, size_t, cudaStream_t, const uint32_t *, const int2 *, int2 *, size_t, const CoreConsts &,
			const SeedConsts &, const uint4 *, int) { return cudaErrorInvalidValue; }
};

// Per-sample (x,y): tries the table-directed kernel on the first floor(n/128)*128 samples.  Suffix directions: IDP.2A
// word planes when neighbouring samples have neighbouring phases (the NCO mixer: the host knows from the step; a phase
// stream of 4 Mi samples or more: both kernels are enqueued and probe; ZC_F_SEED_WORDS / ZC_F_SEED_PACKED force), byte
// rows otherwise.
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
	bool words = (flags & ZC_F_SEED_WORDS) != 0, probe = false;
	if (!(flags & (ZC_F_SEED_WORDS | ZC_F_SEED_PACKED)) && !(flags & ZC_F_NO_DP2A)) {
		if (SRC == SRC_MIX) {
			const int32_t sstep = (int32_t)c.nco_step;
			const uint32_t mag = (uint32_t)(sstep < 0 ? -(int64_t)sstep : (int64_t)sstep);
			words = (mag >> c.pshift) < 2u;
		} else if (n >= ((size_t)1 << 22)) {
			probe = true;
		}
	}
	SeedPlan pl, plw;
	int rc = ZC_OK;
	if (words || probe) {
		if ((rc = seed_plan_get(p, key, device, FL_DIRS_DP, st, plw)) != ZC_OK) return rc;
		if (!plw.usable) { words = false; probe = false; }
	}
	if (!words) {
		if ((rc = seed_plan_get(p, key, device, FL_DIRS, st, pl)) != ZC_OK) return rc;
		if (!pl.usable) return ZC_OK;
	}
	const int probe_lim = probe ? 1 : -1;
	cudaError_t e = cudaSuccess;
	if (words || probe) {
		e = DirsTable<SRC, SEED_MAX_NS>::launch(plw.NS, true, sms, plw.s.total_bytes + 16, st, phase, (const int2 *)xy_in,
			(int2 *)xy_out, nblocks, c, plw.s, (const uint4 *)plw.dev, probe_lim);
		if (e == cudaSuccess) launches++;
	}
	if (e == cudaSuccess && !words) {
		e = DirsTable<SRC, SEED_MAX_NS>::launch(pl.NS, false, sms, pl.s.total_bytes + 16, st, phase, (const int2 *)xy_in,
			(int2 *)xy_out, nblocks, c, pl.s, (const uint4 *)pl.dev, probe_lim);
		if (e == cudaSuccess) launches++;
	}
	if (e != cudaSuccess)
		return set_error(ZC_ECUDA, "launch of k_rotate_dirs failed: %s", cudaGetErrorString(e));
	done = nblocks << 7;
	return ZC_OK;
}

// ---- NCO comb mapping: the host side -----------------------------------------------------------------------
// What the mapping buys is ROW SHARING: with the 8 lanes of a quarter-warp `dp` phase LSBs apart, a 128-bit read of the
// direction table touches about 1 + 7|dp| distinct 16-byte rows per quarter-warp instead of 8, and the LSU serves
// identical rows once.  Measured on B200 (profiles/r2_comb_ab.txt): |dp| = 0 .. 0.3 -> 460-510 Gsamples/s; |dp| = 0.95
// (eight distinct rows again, plus 8 shuffles and the transposed stores) -> 245-280, below the 380 of the byte table
// under the block mapping.  Hence the acceptance bound.
constexpr double COMB_MAX_DP = 0.35;

// Finds a run length K (a multiple of 4, 64 <= K, 8K <= n) with K*step = delta (mod 2^32) and |delta| at most
// COMB_MAX_DP phase LSBs, so that 8 lanes K samples apart mostly read the SAME table rows.  Candidates: the denominators of the continued-fraction convergents of
// step / 2^32 -- the K with record-small |K*step mod 2^32| -- and their small multiples.  Returns 0 when none qualifies.
static uint32_t comb_search(uint32_t step, int pshift, size_t n) {
	if (step == 0 || n < 1024) return 0;
	if (const char *force = std::getenv("ZCORDIC_COMB_K")) {	// experiments: any even K is correct, only the bank conflicts change
		const uint64_t K = std::strtoull(force, nullptr, 0);
		return (K >= 16 && !(K & 3) && 8 * K <= n) ? (uint32_t)K : 0;
	}
	const uint64_t kmax = n / 8 > 0x08000000ull ? 0x08000000ull : n / 8;		// tile = 8K <= 2^30 samples
	uint64_t q[48];
	int nq = 0;
	{	// Euclid on (2^32, step): q_{i+1} = a_i * q_i + q_{i-1}
		uint64_t r0 = (uint64_t)1 << 32, r1 = step, q0 = 0, q1 = 1;
		while (r1 != 0 && nq < 48 && q1 <= kmax) {
			q[nq++] = q1;
			const uint64_t a = r0 / r1, r2 = r0 - a * r1, q2 = a * q1 + q0;
			r0 = r1; r1 = r2; q0 = q1; q1 = q2;
		}
		if (r1 == 0 && nq < 48 && q1 <= kmax) q[nq++] = q1;	// the exact period: q1 * step == 0 (mod 2^32)
	}
	const double lsb = (double)((uint64_t)1 << pshift);
	double best = 1e30;
	uint32_t bestK = 0;
	for (int i = 0; i < nq; i++)
		for (uint64_t m = 1; m <= 64; m++) {
			const uint64_t K = q[i] * m;
			if (K > kmax) break;
			if (K < 64 || (K & 3)) continue;
			const double dp = (double)(int32_t)((uint32_t)K * step) / lsb;		// lanes of a quarter-warp: dp LSBs apart
			if (dp > COMB_MAX_DP || dp < -COMB_MAX_DP) continue;
			const double cost = 1.0 + 7.0 * std::fabs(dp);				// distinct rows per quarter-warp read
			const uint64_t cpr = (K + 15) / 16;
			const double waste = (double)(cpr * 16 - K) / (double)(cpr * 16);	// idle lanes in the last chunk of a run
			const uint64_t covered = (n / (8 * K)) * 8 * K;
			const double left = (double)(n - covered) / (double)n;			// what a second pass has to take
			const double score = cost * (1.0 + waste) * (1.0 + 0.5 * left);
			if (score < best - 1e-9) { best = score; bestK = (uint32_t)K; }
		}
	return bestK;
}

// Tries the seeded path on a prefix of the n samples (a multiple of 128).  done=0 means "not applicable here"; launches
// reports how many kernels were enqueued.  OUT16: xy_out receives (int16 x, int16 y) words.
template <int SRC, bool OUT16>
static int seeded_rotate_try(const zc_params *p, const CoreConsts &c, const uint32_t *phase, void *xy_out,
		size_t n, int device, int sms, cudaStream_t st, uint32_t flags, size_t &done, int &launches) {
	done = 0; launches = 0;
	const size_t nblocks = n >> 7;
	if (nblocks == 0 || nblocks > SEED_MAX_BLOCKS) return ZC_OK;
	if (!(flags & ZC_F_FORCE_SEED) && n < ((size_t)1 << 20)) return ZC_OK;	// not worth the table load
	if (c.neff < 6 || p->pw < 12) return ZC_OK;
	// Which flavour of direction table.  The caller's flag wins.  For the NCO the host knows the pattern: a step of
	// 2 phase LSBs or more scatters neighbouring samples, and then either a comb mapping exists that makes the lanes of
	// a quarter-warp neighbours again (word tables, IDP.2A stages) or the byte rows take over.  For a phase stream of
	// 4 Mi samples or more both flavours are enqueued and a probe inside the kernels decides.
	const bool forced = (flags & (ZC_F_SEED_REGS | ZC_F_SEED_PACKED | ZC_F_SEED_WORDS)) != 0;
	int tdm = (flags & ZC_F_SEED_REGS) ? TD_REGS : (flags & ZC_F_SEED_PACKED) ? TD_PACKED : TD_TABLE;
	if (OUT16 && tdm == TD_REGS) tdm = TD_TABLE;
	bool probe = false, scattered_nco = false;
	if (!forced) {
		if (SRC == SRC_NCO) {
			const int32_t sstep = (int32_t)c.nco_step;
			const uint32_t mag = (uint32_t)(sstep < 0 ? -(int64_t)sstep : (int64_t)sstep);
			if ((mag >> c.pshift) >= 2u) scattered_nco = true;
		} else if (n >= ((size_t)1 << 22)) {
			probe = true;
		}
	}
	SeedPlan pl, pl2;
	int rc = ZC_OK;
	bool have_words = false;
	if (tdm == TD_TABLE && !(flags & ZC_F_NO_DP2A)) {	// the word table as IDP.2A multipliers, when the shifts allow
		if ((rc = seed_plan_get(p, c, device, FL_WORDS_DP, st, pl)) != ZC_OK) return rc;
		if (pl.usable) { tdm = TD_TABLE_DP; have_words = true; }
	}
	if (tdm == TD_TABLE) {
		if ((rc = seed_plan_get(p, c, device, FL_WORDS, st, pl)) != ZC_OK) return rc;
		have_words = pl.usable;
	}
	const int grid = sms;
	SeedLaunch L{grid, 0, st, phase, xy_out, 0, &c, nullptr, nullptr, -1, nullptr};
	cudaError_t e = cudaSuccess;
	size_t off = 0;
	if constexpr (SRC == SRC_NCO && !OUT16) {
		// ---- scattered NCO: comb passes over the largest tile-aligned prefix, again over what is left, ... -------------
		const bool aligned = (reinterpret_cast<uintptr_t>(xy_out) & 31u) == 0;	// 256-bit stores
		while (scattered_nco && have_words && aligned && !(flags & ZC_F_NO_COMB)) {
			const size_t rem = n - off;
			if (rem < ((size_t)1 << 20) && !((flags & ZC_F_FORCE_SEED) && off == 0)) break;
			const uint32_t K = comb_search(c.nco_step, c.pshift, rem);
			if (!K) break;
			const uint32_t nwarps = (uint32_t)grid * 32u;
			CombConsts cb;
			cb.K = K; cb.cpr = (K + 15u) / 16u; cb.tile = 8u * K;
			const uint64_t tiles = rem / cb.tile, units = tiles * cb.cpr;
			if (tiles == 0 || units + nwarps >= ((uint64_t)1 << 32)) break;
			cb.nunits = (uint32_t)units;
			cb.tile_step = cb.tile * c.nco_step; cb.chunk_step = 16u * c.nco_step;
			cb.dt = nwarps / cb.cpr; cb.dm = nwarps % cb.cpr;
			CoreConsts cc = c;
			cc.nco_n0 = c.nco_n0 + (uint32_t)off;		// arithmetic is mod 2^32
			L.smem = pl.s.total_bytes + 16; L.out = (int32_t *)xy_out + 2 * off; L.nblocks = 0; L.c = &cc; L.s = &pl.s;
			L.tables = (const uint4 *)pl.dev; L.cb = &cb;
			e = SeedTable<SRC, SEED_MAX_NS>::template launch<MAP_COMB, false>(pl.NS, tdm, L);
			if (e != cudaSuccess) return set_error(ZC_ECUDA, "launch of k_rotate_seeded (comb) failed: %s", cudaGetErrorString(e));
			launches++;
			off += (size_t)tiles * cb.tile;
		}
		L.cb = nullptr; L.c = &c;
	}
	const size_t rest_blocks = (n - off) >> 7;
	if (off && (rest_blocks << 7) < ((size_t)1 << 20)) { done = off; return ZC_OK; }	// the plain kernels take the tail
	if (scattered_nco) { tdm = TD_PACKED; have_words = false; }
	// the byte table comes with the TS lookup folded into the (x, y) records when the core allows (ZC_F_NO_MERGE: A/B)
	auto packed_plan = [&](SeedPlan &out, int &tdm_out) -> int {
		tdm_out = TD_PACKED;
		if (!(flags & ZC_F_NO_MERGE)) {
			const int prc = seed_plan_get(p, c, device, FL_PACKED_M, st, out);
			if (prc != ZC_OK) return prc;
			if (out.usable) { tdm_out = TD_PACKED_M; return ZC_OK; }
		}
		return seed_plan_get(p, c, device, FL_PACKED, st, out);
	};
	int tdm2 = TD_PACKED;
	if (tdm == TD_PACKED) {
		if ((rc = packed_plan(pl, tdm)) != ZC_OK) return rc;
		if (!pl.usable) { done = off; return ZC_OK; }
	} else if (tdm == TD_REGS) {
		if ((rc = seed_plan_get(p, c, device, FL_WORDS, st, pl)) != ZC_OK) return rc;
		if (!pl.usable) { done = off; return ZC_OK; }
	} else if (!have_words) {
		done = off; return ZC_OK;
	}
	if (probe) {
		if ((rc = packed_plan(pl2, tdm2)) != ZC_OK) return rc;
		if (!pl2.usable) probe = false;
	}
	CoreConsts cc = c;
	cc.nco_n0 = c.nco_n0 + (uint32_t)off;
	// probe: both flavours are enqueued; each evaluates probe_local() on the same input and exactly one proceeds
	L.probe_lim = probe ? 1 : -1;
	L.c = &cc; L.nblocks = rest_blocks;
	L.out = OUT16 ? (void *)((int32_t *)xy_out + off) : (void *)((int32_t *)xy_out + 2 * off);
	L.ph = phase ? phase + off : nullptr;
	L.smem = pl.s.total_bytes + 16; L.s = &pl.s; L.tables = (const uint4 *)pl.dev;
	e = SeedTable<SRC, SEED_MAX_NS>::template launch<MAP_BLOCK, OUT16>(pl.NS, tdm, L);
	if (e == cudaSuccess) launches++;
	if (e == cudaSuccess && probe) {
		L.smem = pl2.s.total_bytes + 16; L.s = &pl2.s; L.tables = (const uint4 *)pl2.dev;
		e = SeedTable<SRC, SEED_MAX_NS>::template launch<MAP_BLOCK, OUT16>(pl2.NS, tdm2, L);
		if (e == cudaSuccess) launches++;
	}
	if (e != cudaSuccess)
		return set_error(ZC_ECUDA, "launch of k_rotate_seeded failed: %s", cudaGetErrorString(e));
	done = off + (rest_blocks << 7);
	return ZC_OK;
}

} // namespace zc
#endif
