// zc_seedplan.h -- the table plans of the table-seeded rotation kernels (zc_seeded.cuh): constants, host-side plan
// structure and the plan cache's interface.  The cache and the setup kernel live in zc_seedplan.cu; the kernels that
// consume the plans are compiled in separate translation units (zc_rot_const.cu, zc_rot_nco.cu, zc_rot_dirs.cu) so that
// the library builds in parallel.
#ifndef ZC_SEEDPLAN_H
#define ZC_SEEDPLAN_H

#include "zc_internal.h"
#include "zc_kernels.cuh"

#include <memory>

namespace zc {

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
	int32_t  recsh;		// FL_PACKED_M: 32-WW, the shift that brings x / y down from the top of their record words
	uint32_t R;
};

// plan flavours: word TD + x/y table, byte TD + x/y table, byte TD + per-interval prefix directions (no x/y table), and
// and the IDP.2A form of the first: word TD holding the multiplier words {0, d, 0, -d}.  (The byte flavours have no
// IDP.2A form: storing the pair (-d, d) per stage and building the multiplier word with one PRMT saves the negation but
// doubles the rows, and their scattered reads cost more than that: 320 vs 374 Gsamples/s for the constant-vector kernel
// on random phases, 181 vs 195 for per-sample vectors -- measured, dropped.)
// FL_DIRS_DP: per-interval prefix directions as in FL_DIRS, suffix directions as IDP.2A multiplier words as in FL_WORDS_DP
// FL_PACKED_M: FL_PACKED with the interval's TD-row offset merged into its (x, y) record -- (x << recsh | low byte,
// y << recsh | high byte of TS mod 2^16) -- so that a sample costs three shared-memory lookups (bucket, record, direction
// row) instead of four; needs WW <= 24 (eight free bits under each of x and y) and a direction table below 64 KB.
// (The same merge under the IDP.2A word table measured 1 % slower on sweeps and 4 % slower on slow NCOs -- neighbouring
// lanes share the TS word there, so the lookup it saves was nearly free: profiles/r2_merged_records_ab.txt -- dropped.)
enum { FL_WORDS = 0, FL_PACKED = 1, FL_DIRS = 2, FL_WORDS_DP = 3, FL_DIRS_DP = 4, FL_PACKED_M = 5 };
static inline bool fl_packed(int flavour) { return flavour == FL_PACKED || flavour == FL_DIRS || flavour == FL_PACKED_M; }
static inline bool fl_merged(int flavour) { return flavour == FL_PACKED_M; }
static inline bool fl_dirs(int flavour) { return flavour == FL_DIRS || flavour == FL_DIRS_DP; }
static inline bool fl_dp(int flavour) { return flavour == FL_WORDS_DP || flavour == FL_DIRS_DP; }
constexpr int DIRS_M = 12;	// prefix depth of the FL_DIRS flavour (its kernel unrolls the byte indices)


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

// cudaFree waits for the device to go idle, so kernels already enqueued on the tables finish first.
struct DevFree { void operator()(void *ptr) const { if (ptr) cudaFree(ptr); } };

// Builds (or finds) the plan for (p, constant vector, device).  Called with the device current.
int seed_plan_get(const zc_params *p, const CoreConsts &c, int device, int flavour, cudaStream_t st, SeedPlan &out);
// cudaFuncSetAttribute once per (device, kernel, size): launches of a configured kernel then consist of the
// launch alone, which keeps them legal inside a stream capture.
cudaError_t ensure_dynamic_smem(const void *kern, size_t smem);
// Drops the cached plans of `device` (all devices when negative).
void seed_trim(int device);


// ---- kernel families compiled in their own translation units -----------------------------------------------------
// Table-seeded / table-directed rotation over a prefix of the n samples: `done` (a multiple of 128, 0 = not applicable)
// says how far they got, `launches` how many kernels were enqueued; the caller finishes with the plain kernels.
int seeded_rotate_const(const zc_params *p, const CoreConsts &c, const uint32_t *phase, void *xy_out, bool out16, size_t n,
		int device, int sms, cudaStream_t st, uint32_t flags, size_t &done, int &launches);
int seeded_rotate_nco(const zc_params *p, const CoreConsts &c, void *xy_out, size_t n, int device, int sms,
		cudaStream_t st, uint32_t flags, size_t &done, int &launches);
long long nco_comb_run(const zc_params *p, uint32_t step, size_t n);
int dirs_rotate_xy(const zc_params *p, const CoreConsts &c, const uint32_t *phase, const int32_t *xy_in, int32_t *xy_out,
		size_t n, int device, int sms, cudaStream_t st, uint32_t flags, size_t &done, int &launches);
int dirs_rotate_mix(const zc_params *p, const CoreConsts &c, const int32_t *xy_in, int32_t *xy_out, size_t n, int device,
		int sms, cudaStream_t st, uint32_t flags, size_t &done, int &launches);
// Fully unrolled plain kernels, 4 samples per thread (groups = samples / 4, 16-byte aligned buffers).
void launch_rotate_plain(int src, bool out16, int neff, int grid, cudaStream_t st, const int4 *ph, const int4 *xin,
		int4 *out, size_t groups, const CoreConsts &c);
void launch_topolar_plain(bool in16, int neff, int ntail, int grid, cudaStream_t st, const int4 *xin, int4 *mag, int4 *ph,
		size_t groups, const CoreConsts &c);

} // namespace zc
#endif
