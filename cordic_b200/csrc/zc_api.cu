// zc_api.cu -- the C ABI of libzcordic (include/zcordic.h): argument checking, constant
// preparation, kernel selection and launch, and the host-buffer pipelines.
#include "zc_internal.h"
#include "zc_kernels.cuh"
#include "zc_generic.cuh"
#include "zc_seedplan.h"
#include "zc_quadtbl.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

namespace zc {

// ---- errors -----------------------------------------------------------------------------
static thread_local char g_errbuf[512] = "";

int set_error(int code, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_errbuf, sizeof(g_errbuf), fmt, ap);
	va_end(ap);
	return code;
}

#define ZC_CUDA(call)                                                                      \
	do {                                                                               \
		cudaError_t e_ = (call);                                                   \
		if (e_ != cudaSuccess)                                                     \
			return set_error(ZC_ECUDA, "%s failed: %s (%s:%d)", #call,         \
				cudaGetErrorString(e_), __FILE__, __LINE__);                \
	} while (0)

static std::atomic<uint64_t> g_launches{0};

// ---- device bookkeeping -------------------------------------------------------------------
struct DeviceInfo { int sms; bool ok; };
static DeviceInfo g_dev[64];
static std::mutex g_dev_mu;

static int device_info(int device, DeviceInfo &out) {
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count <= 0) {
		cudaGetLastError();
		return set_error(ZC_ENODEV, "no usable CUDA device (%s); libzcordic has no CPU path",
			e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
	}
	if (device < 0 || device >= count || device >= 64)
		return set_error(ZC_ENODEV, "device ordinal %d out of range [0,%d)", device, count);
	std::lock_guard<std::mutex> lk(g_dev_mu);
	if (!g_dev[device].ok) {
		cudaDeviceProp prop;
		ZC_CUDA(cudaGetDeviceProperties(&prop, device));
		// the library is an sm_100a cubin (architecture-specific, no PTX): any other device would fail every launch
		// later with "no kernel image"
		if (prop.major != 10 || prop.minor != 0)
			return set_error(ZC_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
				device, prop.major, prop.minor);
		g_dev[device].sms = prop.multiProcessorCount;
		g_dev[device].ok = true;
	}
	out = g_dev[device];
	return ZC_OK;
}

// Makes `device` current for the scope and restores the caller's device afterwards.
struct DeviceScope {
	int prev = -1;
	bool changed = false;
	int enter(int device) {
		ZC_CUDA(cudaGetDevice(&prev));
		if (prev != device) {
			ZC_CUDA(cudaSetDevice(device));
			changed = true;
		}
		return ZC_OK;
	}
	~DeviceScope() { if (changed) cudaSetDevice(prev); }
};

// ---- parameter checks -----------------------------------------------------------------------
int check_params(const zc_params *p, int want_mode) {
	if (!p) return set_error(ZC_EINVAL, "NULL zc_params");
	if (p->mode != want_mode)
		return set_error(ZC_EINVAL, "zc_params.mode=%d but this entry point needs %s", p->mode,
			want_mode == ZC_MODE_P2R ? "ZC_MODE_P2R (zc_derive_p2r)" : "ZC_MODE_R2P (zc_derive_r2p)");
	if (p->pw < 3 || p->pw > 32 || p->ww < 2 || p->ww > 32 || p->iw < 1 || p->ow < 1 ||
	    p->iw > p->ww || p->ow >= p->ww || p->nstages < 0 || p->nstages > ZC_MAX_STAGES)
		return set_error(ZC_ERANGE, "unsupported configuration IW=%d OW=%d WW=%d PW=%d NSTAGES=%d",
			p->iw, p->ow, p->ww, p->pw, p->nstages);
	if (want_mode == ZC_MODE_P2R ? (p->ww - p->iw < 1) : (p->ww - p->iw < 2))
		return set_error(ZC_ERANGE, "working width %d too small for IW=%d", p->ww, p->iw);
	if (p->seq != 0 && p->seq != 1)
		return set_error(ZC_EINVAL, "zc_params.seq=%d (0: pipelined core, 1: sequential core)", p->seq);
	if (p->seq && (want_mode == ZC_MODE_P2R ? p->nstages < 3 : (p->nstages < 1 || ((p->nstages + 1) & p->nstages) == 0)))
		return set_error(ZC_ERANGE, "sequential core with NSTAGES=%d does not exist in the reference", p->nstages);
	return ZC_OK;
}

// Stage updates that reach the output.  Pipelined cores: NSTAGES (of which some may be pass-through).  Sequential
// cores: rtl/seqcordic.v:319-324 registers its outputs two iterations early; rtl/seqpolar.v runs all NSTAGES.
int iterations(const zc_params *p) {
	if (!p->seq) return p->nstages;
	return p->mode == ZC_MODE_P2R ? p->nstages - 2 : p->nstages;
}

// Stages i with cordic_angle[i]==0 or i>=WW are pass-through (rtl/cordic.v:253).  The angle
// table is non-increasing, so the live stages are a prefix; its length is what we unroll.
// The sequential machines have no such test (rtl/seqcordic.v:281-299): every iteration counts there.
static int live_stages(const zc_params *p) {
	if (p->seq) return iterations(p);
	int n = 0;
	while (n < p->nstages && n < p->ww && p->angle[n] != 0) n++;
	return n;
}

// The table kernels derive the directions from strictly positive angles; a sequential core may run zero ones.
static bool angles_all_live(const zc_params *p, int neff) {
	for (int k = 0; k < neff; k++)
		if (k >= ZC_MAX_STAGES || p->angle[k] == 0) return false;
	return true;
}

// True when 32-bit non-wrapping arithmetic provably equals the RTL's WW-bit wrapping
// arithmetic for EVERY in-range input: the vector norm entering stage 0 is at most
// sqrt(2)*2^(WW-2) (p2r: |e| <= 2^(WW-2) per axis; r2p: |e| <= 2^(WW-3) per axis, then the
// 45-degree turn), each stage scales it by at most sqrt(1+4^-k) (total < 1.1645) and adds at
// most sqrt(2) of truncation error, and the rounding add contributes 2^(D-1).
static bool fast_path_is_exact(const zc_params *p) {
	if (p->ww == 32) return true;	// int32 arithmetic *is* the WW-bit arithmetic
	const double limit = (double)(1ull << (p->ww - 1)) - 1.0;
	const double r0 = (p->mode == ZC_MODE_P2R ? 1.41421356237309515 : 1.0) * (double)(1ull << (p->ww - 2));
	const int D = p->ww - p->ow;
	const double bound = r0 * 1.1645 + 1.7 * (double)p->nstages + 2.0 + (double)(1ull << (D - 1));
	return bound <= limit;
}

static void fill_consts(const zc_params *p, CoreConsts &c) {
	std::memset(&c, 0, sizeof(c));
	const int neff = live_stages(p);
	c.neff = neff;
	c.pshift = 32 - p->pw;
	for (int k = 0; k < neff && k < 32; k++) {
		c.pa[k] = p->angle[k] << c.pshift;
		c.na[k] = (int32_t)(0u - c.pa[k]);
	}
	const int lsh = (p->mode == ZC_MODE_P2R) ? (p->ww - p->iw - 1) : (p->ww - p->iw - 2);
	c.in_shl = 32 - p->iw;
	c.in_shr = 32 - p->iw - lsh;
	c.D = p->ww - p->ow;
	c.do_round = (p->ww > p->ow + 1) ? 1 : 0;
	c.rc = c.do_round ? (int32_t)((1u << (c.D - 1)) - 1u) : 0;
	c.wsh = 32 - p->ww;
	if (p->mode == ZC_MODE_R2P) {
		// rtl/topolar.v:122-152; index = {i_xval[IW-1], i_yval[IW-1]}
		const uint32_t E = 1u << (p->pw - 3);
		c.e_phase[0] = (1u * E) << c.pshift;	// 2'b00 (default)
		c.e_phase[1] = (7u * E) << c.pshift;	// 2'b01
		c.e_phase[2] = (3u * E) << c.pshift;	// 2'b10
		c.e_phase[3] = (5u * E) << c.pshift;	// 2'b11
	}
}

static inline int wrap_to(int64_t v, int w) {
	const int sh = 64 - w;
	return (int)((int64_t)((uint64_t)v << sh) >> sh);
}

// The four quarter-turn pre-rotations of the extended constant input (rtl/cordic.v:85-86,131-188).
static void fill_const_xy(const zc_params *p, int32_t x0, int32_t y0, CoreConsts &c) {
	const int lsh = p->ww - p->iw - 1;
	const int64_t ex = (int64_t)wrap_to(x0, p->iw) * ((int64_t)1 << lsh);
	const int64_t ey = (int64_t)wrap_to(y0, p->iw) * ((int64_t)1 << lsh);
	const int64_t xs[4] = {ex, -ey, -ex, ey}, ys[4] = {ey, ex, -ey, -ex};
	for (int q = 0; q < 4; q++) {
		c.cx[q] = wrap_to(xs[q], p->ww);
		c.cy[q] = wrap_to(ys[q], p->ww);
	}
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool aligned8(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

static int grid_for(size_t work, const DeviceInfo &di, int per_sm) {
	size_t blocks = (work + 255) / 256;
	const size_t cap = (size_t)di.sms * (size_t)per_sm;
	if (blocks > cap) blocks = cap;
	if (blocks < 1) blocks = 1;
	return (int)blocks;
}

static int post_launch(const char *what) {
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		return set_error(ZC_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
	g_launches.fetch_add(1, std::memory_order_relaxed);
	return ZC_OK;
}

// The fully unrolled plain kernels are instantiated in zc_rot_plain.cu (k_rotate<N,SRC,OUT16>) and zc_topolar.cu
// (k_topolar<N,NTAIL,IN16>), the table-seeded ones in zc_rot_const.cu / zc_rot_nco.cu / zc_rot_dirs.cu: see zc_seedplan.h.

// First vectoring stage from which y>>>(i+1) is provably 0 or -1 for every input, so that the short stage form
// (zc_kernels.cuh: vec_step_tail) is the same function.  After the +-45 degree turn of rtl/topolar.v:122-152,
// x_0 >= 0 and |y_0| <= x_0.  A stage (rtl/topolar.v:227-243, shift s = i+1) never decreases x, and
// |y_i| <= x_i*2^-i + i follows by induction (x>>>s lies in (x/2^s - 1, x/2^s]).  With X an upper bound of x
// over all stages (the one fast_path_is_exact() uses), -2^(i+1) <= y_i < 2^(i+1) holds once
// X < 2^(2i+1) - i*2^i.  Only meaningful when the fast path is exact (no WW-bit wrap), hence never for WW = 32.
static int vec_tail_start(const zc_params *p, int neff) {
	if (p->ww >= 32 || !fast_path_is_exact(p)) return neff;
	const double xb = (double)(1ull << (p->ww - 2)) * 1.1645 + 1.7 * (double)p->nstages + 2.0;
	for (int i = 1; i < neff && i < 31; i++) {
		const double lim = (double)(1ull << (2 * i + 1 > 62 ? 62 : 2 * i + 1)) - (double)i * (double)(1ull << i);
		if (xb < lim) return i;
	}
	return neff;
}

// out16: xy_out receives one word per sample, (int16 o_xval) | (int16 o_yval) << 16 (zc_rotate_const_o16)
template <int SRC>
static int launch_rotate(const zc_params *p, CoreConsts &c, const uint32_t *phase, const int32_t *xy_in,
		int32_t *xy_out, size_t n, int device, void *stream, uint32_t flags, bool out16 = false) {
	DeviceInfo di;
	int rc = device_info(device, di);
	if (rc != ZC_OK) return rc;
	if (n == 0) return ZC_OK;
	DeviceScope scope;
	if ((rc = scope.enter(device)) != ZC_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;

	constexpr bool has_phase = (SRC == SRC_CONST || SRC == SRC_XY), has_xy = (SRC == SRC_XY || SRC == SRC_MIX);
	// non-wrapping arithmetic must be provably exact for this configuration, else the generic kernel does it all
	const bool math_ok = !(flags & ZC_F_FORCE_GENERIC) && fast_path_is_exact(p) && c.neff >= 1 && c.neff <= 32;
	// the table kernels move 4-byte phases and 8-byte (x,y) pairs; the plain fast kernel moves 16-byte vectors
	const bool pair_ok = (out16 || aligned8(xy_out)) && (!has_xy || aligned8(xy_in));
	const size_t ow_ = out16 ? 1 : 2;		// output words per sample
	const bool vec_ok = aligned16(xy_out) && (!has_phase || aligned16(phase)) && (!has_xy || aligned16(xy_in));
	size_t done = 0;
	if (math_ok && pair_ok && !(flags & ZC_F_NO_SEED) && (!p->seq || angles_all_live(p, c.neff))) {
		int launched = 0;
		if constexpr (SRC == SRC_XY) rc = dirs_rotate_xy(p, c, phase, xy_in, xy_out, n, device, di.sms, st, flags, done, launched);
		else if constexpr (SRC == SRC_MIX) rc = dirs_rotate_mix(p, c, xy_in, xy_out, n, device, di.sms, st, flags, done, launched);
		else if constexpr (SRC == SRC_NCO) rc = seeded_rotate_nco(p, c, xy_out, n, device, di.sms, st, flags, done, launched);
		else rc = seeded_rotate_const(p, c, phase, xy_out, out16, n, device, di.sms, st, flags, done, launched);
		if (rc != ZC_OK) return rc;
		g_launches.fetch_add((uint64_t)launched, std::memory_order_relaxed);
	}
	if (math_ok && vec_ok) {		// `done` is a multiple of 128 samples: the remainder keeps the base alignment
		const size_t groups = (n - done) / 4;
		if (groups) {
			CoreConsts t = c;
			t.nco_n0 = c.nco_n0 + (uint32_t)done;
			launch_rotate_plain(SRC, out16, c.neff, grid_for(groups, di, 16), st,
				(const int4 *)(phase ? phase + done : nullptr), (const int4 *)(xy_in ? xy_in + 2 * done : nullptr),
				(int4 *)(xy_out + ow_ * done), groups, t);
			if ((rc = post_launch("k_rotate")) != ZC_OK) return rc;
			done += groups * 4;
		}
	}
	if (done < n) {
		CoreConsts t = c;
		t.nco_n0 = c.nco_n0 + (uint32_t)done;
		const size_t rest = n - done;
		k_rotate_generic<SRC><<<grid_for(rest, di, 16), 256, 0, st>>>(
			phase ? phase + done : nullptr, xy_in ? xy_in + 2 * done : nullptr,
			xy_out + ow_ * done, rest, t, out16 ? 1 : 0);
		if ((rc = post_launch("k_rotate_generic")) != ZC_OK) return rc;
	}
	return ZC_OK;
}

// in16: xy_in holds one word per sample, (int16 i_xval) | (int16 i_yval) << 16 (zc_topolar_i16)
static int launch_topolar(const zc_params *p, const int32_t *xy_in, int32_t *mag, uint32_t *phase,
		size_t n, int device, void *stream, uint32_t flags, bool in16 = false) {
	int rc = check_params(p, ZC_MODE_R2P);
	if (rc != ZC_OK) return rc;
	if (n && (!xy_in || !mag || !phase)) return set_error(ZC_EINVAL, "NULL buffer");
	if (in16 && p->iw > 16) return set_error(ZC_ERANGE, "packed int16 inputs need IW <= 16 (IW=%d)", p->iw);
	DeviceInfo di;
	if ((rc = device_info(device, di)) != ZC_OK) return rc;
	if (n == 0) return ZC_OK;
	DeviceScope scope;
	if ((rc = scope.enter(device)) != ZC_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	CoreConsts c;
	fill_consts(p, c);
	const bool fast = !(flags & ZC_F_FORCE_GENERIC) && fast_path_is_exact(p) && aligned16(xy_in) &&
		aligned16(mag) && aligned16(phase) && c.neff >= 1 && c.neff <= 32;
	size_t done = 0;
	if (fast) {
		const size_t groups = n / 4;
		if (groups) {
			const int ntail = (flags & ZC_F_NO_TAIL) ? 0 : c.neff - vec_tail_start(p, c.neff);
			launch_topolar_plain(in16, c.neff, ntail, grid_for(groups, di, 16), st, (const int4 *)xy_in, (int4 *)mag,
				(int4 *)phase, groups, c);
			if ((rc = post_launch("k_topolar")) != ZC_OK) return rc;
			done = groups * 4;
		}
	}
	if (done < n) {
		const size_t rest = n - done;
		k_topolar_generic<<<grid_for(rest, di, 16), 256, 0, st>>>(xy_in + (in16 ? 1 : 2) * done, mag + done,
			phase + done, rest, c, in16 ? 1 : 0);
		if ((rc = post_launch("k_topolar_generic")) != ZC_OK) return rc;
	}
	return ZC_OK;
}

// out16: `out` receives int16 words (tables with OW <= 16).  nco: no phase stream, the 32-bit accumulator
// nco[0] + (nco[2] + i) * nco[1] is generated in registers (zc_nco_lut_*).
template <bool QUARTER>
static int launch_lut(int pw, int ow, const uint32_t *tbl, const uint32_t *phase32, void *out,
		size_t n, int device, void *stream, bool out16 = false, const uint32_t *nco = nullptr) {
	int rc = check_lut(QUARTER, pw, ow);
	if (rc != ZC_OK) return rc;
	if (out16 && ow > 16) return set_error(ZC_ERANGE, "packed int16 outputs need OW <= 16 (OW=%d)", ow);
	if (!tbl || (n && ((!nco && !phase32) || !out))) return set_error(ZC_EINVAL, "NULL buffer");
	DeviceInfo di;
	if ((rc = device_info(device, di)) != ZC_OK) return rc;
	if (n == 0) return ZC_OK;
	DeviceScope scope;
	if ((rc = scope.enter(device)) != ZC_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	LutConsts c;
	c.pshift = 32 - pw; c.osh = 32 - ow; c.pw = pw;
	c.lowmask = QUARTER ? ((1u << (pw - 2)) - 1u) : 0u;
	c.nco = nco ? 1 : 0;
	c.nco_phase0 = nco ? nco[0] : 0u; c.nco_step = nco ? nco[1] : 0u; c.nco_n0 = nco ? nco[2] : 0u;
	size_t done = 0;
	if ((nco || aligned16(phase32)) && (out16 ? aligned8(out) : aligned16(out)) && n >= 4) {
		const size_t groups = n / 4;
		// Large batches of a table that fits shared memory once compressed (int16 half-wave / u16[+u8] magnitudes) go
		// through the kernel that keeps it there: indifferent to the phase pattern.  ZCORDIC_LUT_SMEM=0 keeps the L2 path.
		const size_t nent = QUARTER ? ((size_t)1 << (pw - 2)) : ((size_t)1 << (pw - 1));
		const bool hi8 = QUARTER && ow > 17;
		const size_t smem = nent * (hi8 ? 3 : 2);
		// ZCORDIC_LUT_SMEM=0: never; =2: always (A/B); default: a probe of the phases decides on the device -- neighbouring
		// phases (a sweep) keep the L2 kernel, which then streams at the HBM copy peak; scattered ones take shared memory.
		static const int smem_mode = std::getenv("ZCORDIC_LUT_SMEM") ? std::atoi(std::getenv("ZCORDIC_LUT_SMEM")) : 1;
		const bool fits = n >= ((size_t)1 << 22) && smem <= 200 * 1024 && (QUARTER ? ow <= 25 : ow <= 16);
		// both kernels evaluate the same probe of the same phases (zc_kernels.cuh: probe_local) and exactly one proceeds
		const int entry = (int)(1u << (pw < 2 ? 30 : (32 - pw > 30 ? 30 : 32 - pw)));	// one table entry, in 32-bit phase units
		int probe_lim = (smem_mode == 1 && fits) ? entry : -1;
		bool use_l2 = (smem_mode != 2 || !fits), use_smem = (smem_mode != 0 && fits);
		if (nco) {		// the host knows the pattern: neighbouring samples within one table entry keep the L2 kernel
			const int32_t sstep = (int32_t)nco[1];
			const bool local = sstep >= -entry && sstep <= entry;
			probe_lim = -1;
			if (smem_mode == 1 && fits) { use_l2 = local; use_smem = !local; }
		}
		if (use_l2) {
			if (out16) k_lut<QUARTER, true><<<grid_for(groups, di, 32), 256, 0, st>>>((const int4 *)phase32, out, tbl, groups, c, probe_lim);
			else k_lut<QUARTER, false><<<grid_for(groups, di, 32), 256, 0, st>>>((const int4 *)phase32, out, tbl, groups, c, probe_lim);
			if ((rc = post_launch("k_lut")) != ZC_OK) return rc;
		}
		if (use_smem) {
			typedef void (*kern_t)(const int4 *, void *, const uint32_t *, size_t, const LutConsts, int);
			kern_t kern = out16 ? (hi8 ? (kern_t)k_lut_smem<QUARTER, true, true> : (kern_t)k_lut_smem<QUARTER, false, true>)
					    : (hi8 ? (kern_t)k_lut_smem<QUARTER, true, false> : (kern_t)k_lut_smem<QUARTER, false, false>);
			cudaError_t e = ensure_dynamic_smem((const void *)kern, smem);
			if (e != cudaSuccess) return set_error(ZC_ECUDA, "k_lut_smem shared memory: %s", cudaGetErrorString(e));
			kern<<<di.sms, 1024, smem, st>>>((const int4 *)phase32, out, tbl, groups, c, probe_lim);
			if ((rc = post_launch("k_lut_smem")) != ZC_OK) return rc;
		}
		done = groups * 4;
	}
	if (done < n) {
		const size_t rest = n - done;
		void *tail = out16 ? (void *)((int16_t *)out + done) : (void *)((int32_t *)out + done);
		c.nco_n0 += (uint32_t)done;
		k_lut_scalar<QUARTER><<<grid_for(rest, di, 32), 256, 0, st>>>(phase32 ? phase32 + done : nullptr, tail, tbl, rest, c, out16 ? 1 : 0);
		if ((rc = post_launch("k_lut_scalar")) != ZC_OK) return rc;
	}
	return ZC_OK;
}

// ---- quadtbl ------------------------------------------------------------------------------------
// Device copies of coefficient tables, keyed by content (the zc_quadtbl is caller memory).
// The key is the whole configuration the cached rows and the cached `nowrap` verdict depend on: the widths that decide
// the sign extension (CBITS/LBITS/QBITS), the ones the wrap proof walks (PW via DXBITS, LBITS, CBITS, WW) and the words.
struct QtKey {
	int32_t pw, ww, lgtbl, dxbits, cbits, lbits, qbits;
	std::vector<uint32_t> words;		// ctbl | ltbl | qtbl, 2^LGTBL each
	bool operator==(const QtKey &o) const {
		return pw == o.pw && ww == o.ww && lgtbl == o.lgtbl && dxbits == o.dxbits && cbits == o.cbits &&
			lbits == o.lbits && qbits == o.qbits && words == o.words;
	}
};
struct QtDevTable { int device; uint64_t hash; QtKey key; std::shared_ptr<void> dev; bool nowrap; uint64_t stamp; };
static std::mutex g_qt_mu;
static std::vector<QtDevTable> g_qt_cache;
static uint64_t g_qt_clock = 0;

static uint64_t qt_hash(const zc_quadtbl *q) {
	uint64_t h = 1469598103934665603ull;
	auto mix = [&](uint32_t w) { h = (h ^ w) * 1099511628211ull; };
	const int n = 1 << q->lgtbl;
	mix((uint32_t)q->lgtbl); mix((uint32_t)q->pw); mix((uint32_t)q->ww); mix((uint32_t)q->dxbits);
	mix((uint32_t)q->cbits); mix((uint32_t)q->lbits); mix((uint32_t)q->qbits);
	for (int k = 0; k < n; k++) { mix(q->ctbl[k]); mix(q->ltbl[k]); mix(q->qtbl[k]); }
	return h;
}

static QtKey qt_key(const zc_quadtbl *q) {
	QtKey k{q->pw, q->ww, q->lgtbl, q->dxbits, q->cbits, q->lbits, q->qbits, {}};
	const size_t n = (size_t)1 << q->lgtbl;
	k.words.reserve(3 * n);
	k.words.insert(k.words.end(), q->ctbl, q->ctbl + n);
	k.words.insert(k.words.end(), q->ltbl, q->ltbl + n);
	k.words.insert(k.words.end(), q->qtbl, q->qtbl + n);
	return k;
}

static inline int32_t sext32(uint32_t w, int bits) { return (int32_t)(w << (32 - bits)) >> (32 - bits); }

// Uploads (once) the coefficient tables sign-extended, and decides whether the register wraps can be skipped:
// |lsum| <= |l| + |q| (dx < 2^(DXBITS-1), so the renormalised product is at most |q|) must fit LBITS, and
// |r| <= |c| + |l| + |q| must fit CBITS, entry by entry.
static int qt_device_tables(const zc_quadtbl *q, int device, cudaStream_t st, std::shared_ptr<void> *out, bool *nowrap) {
	const uint64_t h = qt_hash(q);
	const int n = 1 << q->lgtbl;
	QtKey key = qt_key(q);
	std::lock_guard<std::mutex> lk(g_qt_mu);
	for (QtDevTable &e : g_qt_cache)
		if (e.device == device && e.hash == h && e.key == key) {	// a hash hit is confirmed word for word
			e.stamp = ++g_qt_clock; *out = e.dev; *nowrap = e.nowrap;
			return ZC_OK;
		}
	std::vector<int32_t> host(4 * (size_t)n);		// one row {c, l, q, 0} per index
	bool safe = true;
	// Can the LBITS-bit lsum or the CBITS-bit r_value register wrap for SOME phase?  Tables of up to 2^24 phases are
	// simply walked: every (entry, dx) pair through the exact arithmetic of rtl/quadtbl.v:196-260 (a peak entry has
	// |c| at full scale, so the triangle inequality alone would condemn every real table); larger ones get the bound.
	const int dxs = q->dxbits - 1;
	const bool walk = (q->lgtbl + dxs) <= 24;
	const int rbits = q->cbits < q->ww ? q->cbits : q->ww;		// r must also fit OW+XTRA bits for the rounding not to wrap
	const int64_t llim = (int64_t)1 << (q->lbits - 1), rlim = (int64_t)1 << (rbits - 1);
	for (int k = 0; k < n; k++) {
		const int64_t cv = sext32(q->ctbl[k], q->cbits), lv = sext32(q->ltbl[k], q->lbits), qv = sext32(q->qtbl[k], q->qbits);
		host[4 * k] = (int32_t)cv; host[4 * k + 1] = (int32_t)lv; host[4 * k + 2] = (int32_t)qv; host[4 * k + 3] = 0;
		if (!safe) continue;
		if (walk) {
			for (int64_t dx = 0; dx < ((int64_t)1 << dxs) && safe; dx++) {
				const int64_t lsum = ((qv * dx) >> dxs) + lv;
				const int64_t r = ((lsum * dx) >> dxs) + cv;
				if (lsum < -llim || lsum >= llim || r < -rlim || r >= rlim) safe = false;
			}
		} else {
			const int64_t al = (lv < 0 ? -lv : lv) + (qv < 0 ? -qv : qv) + 1, ac = (cv < 0 ? -cv : cv) + al + 1;
			if (al >= llim || ac >= rlim) safe = false;
		}
	}
	int32_t *dev = nullptr;
	ZC_CUDA(cudaMalloc((void **)&dev, host.size() * 4));
	cudaError_t e = cudaMemcpyAsync(dev, host.data(), host.size() * 4, cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) { cudaFree(dev); return set_error(ZC_ECUDA, "quadtbl table upload failed: %s", cudaGetErrorString(e)); }
	if (g_qt_cache.size() >= 16) {		// evict the least recently used; its memory goes with the last user's copy
		size_t victim = 0;
		for (size_t k = 1; k < g_qt_cache.size(); k++) if (g_qt_cache[k].stamp < g_qt_cache[victim].stamp) victim = k;
		g_qt_cache.erase(g_qt_cache.begin() + victim);
	}
	g_qt_cache.push_back(QtDevTable{device, h, std::move(key), std::shared_ptr<void>(dev, DevFree()), safe, ++g_qt_clock});
	*out = g_qt_cache.back().dev; *nowrap = safe;
	return ZC_OK;
}

static int launch_quadtbl(const zc_quadtbl *q, const uint32_t *phase32, int32_t *out, size_t n, int device, void *stream) {
	int rc = check_qtbl(q);
	if (rc != ZC_OK) return rc;
	if (n && (!phase32 || !out)) return set_error(ZC_EINVAL, "NULL buffer");
	DeviceInfo di;
	if ((rc = device_info(device, di)) != ZC_OK) return rc;
	if (n == 0) return ZC_OK;
	DeviceScope scope;
	if ((rc = scope.enter(device)) != ZC_OK) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	std::shared_ptr<void> hold;
	bool nowrap = false;
	if ((rc = qt_device_tables(q, device, st, &hold, &nowrap)) != ZC_OK) return rc;
	const int32_t *tables = static_cast<const int32_t *>(hold.get());
	QtConsts c;
	c.pshift = 32 - q->pw; c.dxs = q->dxbits - 1; c.dxmask = (1u << (q->dxbits - 1)) - 1u;
	c.qsh = 32 - q->qbits; c.lsh = 32 - q->lbits; c.csh = 32 - q->cbits;
	c.xtra = q->nextra; c.rc = (1 << (q->nextra - 1)) - 1;
	c.keep_hi = (1 << (q->ow - 1)) - 1; c.keep_lo = -(1 << (q->ow - 2));
	c.osh = 32 - q->ow; c.ntbl = 1 << q->lgtbl;
	const bool vec = aligned16(phase32) && aligned16(out);
	const size_t groups = vec ? n / 4 : 0, tail = n - groups * 4;
	const bool wide = (q->qbits + q->dxbits > 31) || (q->lbits + q->dxbits > 31);
	const int grid = grid_for(groups ? groups : tail, di, 8);
	const size_t smem = (size_t)c.ntbl * 16;		// up to 64 KB (LGTBL = 12): beyond the 48 KB default, opt in
#define ZC_QT_LAUNCH(W, NW)                                                                           \
	do {                                                                                              \
		cudaError_t qe = ensure_dynamic_smem((const void *)k_quadtbl<W, NW>, smem);                    \
		if (qe != cudaSuccess) return set_error(ZC_ECUDA, "k_quadtbl shared memory: %s", cudaGetErrorString(qe)); \
		k_quadtbl<W, NW><<<grid, 256, smem, st>>>((const int4 *)phase32, (int4 *)out, tables, groups,   \
			phase32 + groups * 4, out + groups * 4, (int)tail, c);                                   \
	} while (0)
	if (wide) { if (nowrap) ZC_QT_LAUNCH(true, true); else ZC_QT_LAUNCH(true, false); }
	else      { if (nowrap) ZC_QT_LAUNCH(false, true); else ZC_QT_LAUNCH(false, false); }
#undef ZC_QT_LAUNCH
	return post_launch("k_quadtbl");
}

// ---- host-buffer pipeline ---------------------------------------------------------------------
// Streams chunks H2D -> kernel -> D2H through NBUF device buffers on three streams so copies in
// both directions overlap compute.  `launch(chunk_index_offset, count, din0, din1, dout0, dout1,
// stream)` enqueues the device entry point for one chunk.
struct Lane { size_t bytes_per_sample; const char *host_in; char *host_out; };

// Device staging buffers, streams and events are kept between calls (at most two sets per device):
// a test bench that sends batch after batch should not pay cudaMalloc + stream creation every time.
constexpr int PIPE_NBUF = 3;
struct PipeCtx {
	int device = -1;
	char *pool = nullptr;
	size_t pool_bytes = 0;
	cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
	cudaEvent_t ev_in[PIPE_NBUF] = {}, ev_k[PIPE_NBUF] = {}, ev_out[PIPE_NBUF] = {};
};
static std::mutex g_pipe_mu;
static std::vector<PipeCtx *> g_pipe_free;

// Called with ctx->device current.
static void pipe_destroy(PipeCtx *ctx) {
	for (int b = 0; b < PIPE_NBUF; b++) {
		if (ctx->ev_in[b]) cudaEventDestroy(ctx->ev_in[b]);
		if (ctx->ev_k[b]) cudaEventDestroy(ctx->ev_k[b]);
		if (ctx->ev_out[b]) cudaEventDestroy(ctx->ev_out[b]);
	}
	if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
	if (ctx->s_k) cudaStreamDestroy(ctx->s_k);
	if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
	if (ctx->pool) cudaFree(ctx->pool);
	delete ctx;
}

// Called with `device` current.  Returns a context whose pool holds at least `bytes`.
static int pipe_acquire(int device, size_t bytes, PipeCtx **out) {
	PipeCtx *ctx = nullptr;
	{
		std::lock_guard<std::mutex> lk(g_pipe_mu);
		size_t pick = g_pipe_free.size();
		for (size_t k = 0; k < g_pipe_free.size(); k++)
			if (g_pipe_free[k]->device == device &&
			    (pick == g_pipe_free.size() || g_pipe_free[k]->pool_bytes > g_pipe_free[pick]->pool_bytes)) pick = k;
		if (pick < g_pipe_free.size()) {
			ctx = g_pipe_free[pick];
			g_pipe_free.erase(g_pipe_free.begin() + pick);
		}
	}
	cudaError_t e = cudaSuccess;
	if (!ctx) {
		ctx = new PipeCtx();
		ctx->device = device;
		e = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_k, cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking);
		for (int b = 0; b < PIPE_NBUF && e == cudaSuccess; b++) {
			e = cudaEventCreateWithFlags(&ctx->ev_in[b], cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_k[b], cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_out[b], cudaEventDisableTiming);
		}
	}
	if (e == cudaSuccess && ctx->pool_bytes < bytes) {
		if (ctx->pool) cudaFree(ctx->pool);
		ctx->pool = nullptr; ctx->pool_bytes = 0;
		e = cudaMalloc((void **)&ctx->pool, bytes);
		if (e == cudaSuccess) ctx->pool_bytes = bytes;
	}
	if (e != cudaSuccess) {
		cudaGetLastError();
		pipe_destroy(ctx);
		return set_error(e == cudaErrorMemoryAllocation ? ZC_ENOMEM : ZC_ECUDA,
			"host pipeline setup (%zu staging bytes) failed: %s", bytes, cudaGetErrorString(e));
	}
	*out = ctx;
	return ZC_OK;
}

// Called with ctx->device current and all of its streams idle.
static void pipe_release(PipeCtx *ctx) {
	{
		std::lock_guard<std::mutex> lk(g_pipe_mu);
		int kept = 0;
		for (PipeCtx *f : g_pipe_free) kept += (f->device == ctx->device);
		if (kept < 2) { g_pipe_free.push_back(ctx); return; }
	}
	pipe_destroy(ctx);
}

template <class Launch>
static int host_pipeline(int device, size_t n, const Lane in[2], const Lane out[2], Launch launch) {
	DeviceInfo di;
	int rc = device_info(device, di);
	if (rc != ZC_OK) return rc;
	if (n == 0) return ZC_OK;
	DeviceScope scope;
	if ((rc = scope.enter(device)) != ZC_OK) return rc;
	constexpr int NBUF = PIPE_NBUF;
	// samples per pipeline chunk: 4 Mi by default; ZCORDIC_HOST_CHUNK_LG2 (20..28) for sweeps of it
	static const size_t chunk_max = [] {
		const char *e = std::getenv("ZCORDIC_HOST_CHUNK_LG2");
		const int lg = e ? std::atoi(e) : 22;
		return (size_t)1 << (lg < 20 ? 20 : lg > 28 ? 28 : lg);
	}();
	const size_t chunk = (n < chunk_max) ? ((n + 3) & ~(size_t)3) : chunk_max;
	const size_t sizes[4] = {in[0].bytes_per_sample * chunk, in[1].bytes_per_sample * chunk,
		out[0].bytes_per_sample * chunk, out[1].bytes_per_sample * chunk};
	size_t per_buf = 0;
	for (int k = 0; k < 4; k++) per_buf += (sizes[k] + 255) & ~(size_t)255;	// every sub-buffer 256-byte aligned
	PipeCtx *ctx = nullptr;
	if ((rc = pipe_acquire(device, per_buf * NBUF, &ctx)) != ZC_OK) return rc;
	const cudaStream_t s_in = ctx->s_in, s_k = ctx->s_k, s_out = ctx->s_out;
#define ZC_PIPE(call)                                                                          \
	do {                                                                                   \
		cudaError_t e_ = (call);                                                       \
		if (e_ != cudaSuccess) {                                                       \
			cudaDeviceSynchronize();                                               \
			cudaGetLastError();                                                    \
			pipe_destroy(ctx);                                                     \
			return set_error(ZC_ECUDA, "%s failed: %s (%s:%d)", #call,             \
				cudaGetErrorString(e_), __FILE__, __LINE__);                    \
		}                                                                              \
	} while (0)
	auto sub = [&](int b, int which) -> char * {	// which: 0,1 = in ; 2,3 = out
		size_t off = 0;
		for (int k = 0; k < which; k++) off += (sizes[k] + 255) & ~(size_t)255;
		return ctx->pool + per_buf * b + off;
	};
	size_t ci = 0;
	for (size_t off = 0; off < n; off += chunk, ci++) {
		const int b = (int)(ci % NBUF);
		const size_t cnt = (n - off < chunk) ? (n - off) : chunk;
		if (ci >= NBUF) {	// the buffer's previous contents must have left the device
			ZC_PIPE(cudaStreamWaitEvent(s_in, ctx->ev_out[b], 0));
		}
		for (int k = 0; k < 2; k++)
			if (in[k].bytes_per_sample)
				ZC_PIPE(cudaMemcpyAsync(sub(b, k), in[k].host_in + off * in[k].bytes_per_sample,
					cnt * in[k].bytes_per_sample, cudaMemcpyHostToDevice, s_in));
		ZC_PIPE(cudaEventRecord(ctx->ev_in[b], s_in));
		ZC_PIPE(cudaStreamWaitEvent(s_k, ctx->ev_in[b], 0));
		if (ci >= NBUF) ZC_PIPE(cudaStreamWaitEvent(s_k, ctx->ev_out[b], 0));
		rc = launch(off, cnt, sub(b, 0), sub(b, 1), sub(b, 2), sub(b, 3), s_k);
		if (rc != ZC_OK) {
			cudaDeviceSynchronize();
			pipe_destroy(ctx);
			return rc;
		}
		ZC_PIPE(cudaEventRecord(ctx->ev_k[b], s_k));
		ZC_PIPE(cudaStreamWaitEvent(s_out, ctx->ev_k[b], 0));
		for (int k = 0; k < 2; k++)
			if (out[k].bytes_per_sample)
				ZC_PIPE(cudaMemcpyAsync(out[k].host_out + off * out[k].bytes_per_sample, sub(b, 2 + k),
					cnt * out[k].bytes_per_sample, cudaMemcpyDeviceToHost, s_out));
		ZC_PIPE(cudaEventRecord(ctx->ev_out[b], s_out));
	}
	ZC_PIPE(cudaStreamSynchronize(s_out));
	ZC_PIPE(cudaStreamSynchronize(s_k));
	ZC_PIPE(cudaStreamSynchronize(s_in));
#undef ZC_PIPE
	pipe_release(ctx);
	return ZC_OK;
}

// Frees the cached device allocations of `device` (all devices when negative): staging pools, seed tables,
// quadtbl tables.  Safe at any time: cudaFree lets in-flight work finish first.
static int trim_caches(int device) {
	std::vector<PipeCtx *> victims;
	{
		std::lock_guard<std::mutex> lk(g_pipe_mu);
		for (size_t k = 0; k < g_pipe_free.size();) {
			if (device < 0 || g_pipe_free[k]->device == device) {
				victims.push_back(g_pipe_free[k]);
				g_pipe_free.erase(g_pipe_free.begin() + k);
			} else k++;
		}
	}
	for (PipeCtx *ctx : victims) {
		DeviceScope scope;
		if (scope.enter(ctx->device) == ZC_OK) pipe_destroy(ctx);
	}
	{
		std::lock_guard<std::mutex> lk(g_qt_mu);
		for (size_t k = 0; k < g_qt_cache.size();) {
			if (device < 0 || g_qt_cache[k].device == device) g_qt_cache.erase(g_qt_cache.begin() + k);
			else k++;
		}
	}
	seed_trim(device);
	return ZC_OK;
}

} // namespace zc

using namespace zc;

// =============================== extern "C" ===================================================
extern "C" {

int zc_version(void) { return ZC_VERSION_MAJOR * 1000 + ZC_VERSION_MINOR; }

const char *zc_strerror(int status) {
	switch (status) {
	case ZC_OK: return "ok";
	case ZC_EINVAL: return "invalid argument";
	case ZC_ERANGE: return "configuration out of range";
	case ZC_ECUDA: return "CUDA error";
	case ZC_ENODEV: return "no usable CUDA device";
	case ZC_ENOMEM: return "out of memory";
	default: return "unknown status";
	}
}

const char *zc_last_error(void) { return g_errbuf; }

int zc_device_count(void) {
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return set_error(ZC_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
	}
	return count;
}

uint64_t zc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int zc_trim(int device) { return trim_caches(device); }

int zc_derive_p2r(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *out) {
	return derive_p2r(iw, ow, xtra_user, pw, nstages, out);
}
int zc_derive_r2p(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *out) {
	return derive_r2p(iw, ow, xtra_user, pw, nstages, out);
}
int zc_derive_sp2r(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *out) {
	return derive_sp2r(iw, ow, xtra_user, pw, nstages, out);
}
int zc_derive_sr2p(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *out) {
	return derive_sr2p(iw, ow, xtra_user, pw, nstages, out);
}
int zc_iterations(const zc_params *p) {
	if (!p) return set_error(ZC_EINVAL, "NULL zc_params");
	return iterations(p);
}
int zc_topolar_tail_stages(const zc_params *p) {
	if (!p) return set_error(ZC_EINVAL, "NULL zc_params");
	if (p->mode != ZC_MODE_R2P) return 0;
	const int neff = live_stages(p);
	if (neff < 1 || neff > 32) return 0;
	int t = neff - vec_tail_start(p, neff);
	t &= ~1;
	return t > 16 ? 16 : t;
}
long long zc_nco_comb_run(const zc_params *p, uint32_t step, size_t n) {
	int rc = check_params(p, ZC_MODE_P2R);
	if (rc != ZC_OK) return rc;
	return nco_comb_run(p, step, n);
}
int zc_clocks_per_output(const zc_params *p) {
	if (!p) return set_error(ZC_EINVAL, "NULL zc_params");
	if (!p->seq) return 1;
	return p->mode == ZC_MODE_P2R ? p->nstages + 1 : p->nstages + 3;	// sw/seqcordic.cpp:459, sw/seqpolar.cpp:396
}
int zc_derive_tbl(int iw, int pw, int ow, int *pw_out, int *ow_out) {
	return derive_lut(false, iw, pw, ow, pw_out, ow_out);
}
int zc_derive_qtr(int iw, int pw, int ow, int *pw_out, int *ow_out) {
	return derive_lut(true, iw, pw, ow, pw_out, ow_out);
}
int zc_derive_qtbl(int iw, int ow, int xtra_user, int pw, zc_quadtbl *out) {
	return derive_qtbl(iw, ow, xtra_user, pw, out);
}
int zc_lut_build_sintable(int pw, int ow, uint32_t *tbl) { return build_sintable(pw, ow, tbl); }
int zc_lut_build_quarterwav(int pw, int ow, uint32_t *tbl) { return build_quarterwav(pw, ow, tbl); }

int zc_rotate_const_ex(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase, int32_t *xy,
		size_t n, int device, void *stream, uint32_t flags) {
	int rc = check_params(p, ZC_MODE_P2R);
	if (rc != ZC_OK) return rc;
	if (n && (!phase || !xy)) return set_error(ZC_EINVAL, "NULL buffer");
	CoreConsts c;
	fill_consts(p, c);
	fill_const_xy(p, x0, y0, c);
	return launch_rotate<SRC_CONST>(p, c, phase, nullptr, xy, n, device, stream, flags);
}
int zc_rotate_const(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase, int32_t *xy,
		size_t n, int device, void *stream) {
	return zc_rotate_const_ex(p, x0, y0, phase, xy, n, device, stream, ZC_F_DEFAULT);
}

int zc_rotate_ex(const zc_params *p, const int32_t *xy_in, const uint32_t *phase, int32_t *xy_out,
		size_t n, int device, void *stream, uint32_t flags) {
	int rc = check_params(p, ZC_MODE_P2R);
	if (rc != ZC_OK) return rc;
	if (n && (!phase || !xy_in || !xy_out)) return set_error(ZC_EINVAL, "NULL buffer");
	CoreConsts c;
	fill_consts(p, c);
	return launch_rotate<SRC_XY>(p, c, phase, xy_in, xy_out, n, device, stream, flags);
}
int zc_rotate(const zc_params *p, const int32_t *xy_in, const uint32_t *phase, int32_t *xy_out,
		size_t n, int device, void *stream) {
	return zc_rotate_ex(p, xy_in, phase, xy_out, n, device, stream, ZC_F_DEFAULT);
}

int zc_topolar_ex(const zc_params *p, const int32_t *xy_in, int32_t *mag, uint32_t *phase, size_t n,
		int device, void *stream, uint32_t flags) {
	return launch_topolar(p, xy_in, mag, phase, n, device, stream, flags);
}
int zc_topolar(const zc_params *p, const int32_t *xy_in, int32_t *mag, uint32_t *phase, size_t n,
		int device, void *stream) {
	return launch_topolar(p, xy_in, mag, phase, n, device, stream, ZC_F_DEFAULT);
}

int zc_rotate_const_o16(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase, int16_t *xy16,
		size_t n, int device, void *stream) {
	int rc = check_params(p, ZC_MODE_P2R);
	if (rc != ZC_OK) return rc;
	if (p->ow > 16) return set_error(ZC_ERANGE, "packed int16 outputs need OW <= 16 (OW=%d)", p->ow);
	if (n && (!phase || !xy16)) return set_error(ZC_EINVAL, "NULL buffer");
	if (reinterpret_cast<uintptr_t>(xy16) & 3u) return set_error(ZC_EINVAL, "xy16 must be 4-byte aligned");
	CoreConsts c;
	fill_consts(p, c);
	fill_const_xy(p, x0, y0, c);
	return launch_rotate<SRC_CONST>(p, c, phase, nullptr, reinterpret_cast<int32_t *>(xy16), n, device, stream, ZC_F_DEFAULT, true);
}

int zc_topolar_i16(const zc_params *p, const int16_t *xy16_in, int32_t *mag, uint32_t *phase, size_t n,
		int device, void *stream) {
	if (reinterpret_cast<uintptr_t>(xy16_in) & 3u) return set_error(ZC_EINVAL, "xy16_in must be 4-byte aligned");
	return launch_topolar(p, reinterpret_cast<const int32_t *>(xy16_in), mag, phase, n, device, stream, ZC_F_DEFAULT, true);
}

int zc_nco_rotate_ex(const zc_params *p, int32_t x0, int32_t y0, uint32_t phase0, uint32_t step,
		uint64_t n0, int32_t *xy, size_t n, int device, void *stream, uint32_t flags) {
	int rc = check_params(p, ZC_MODE_P2R);
	if (rc != ZC_OK) return rc;
	if (n && !xy) return set_error(ZC_EINVAL, "NULL buffer");
	CoreConsts c;
	fill_consts(p, c);
	fill_const_xy(p, x0, y0, c);
	c.nco_phase0 = phase0; c.nco_step = step; c.nco_n0 = (uint32_t)n0;	// arithmetic is mod 2^32
	return launch_rotate<SRC_NCO>(p, c, nullptr, nullptr, xy, n, device, stream, flags);
}
int zc_nco_rotate(const zc_params *p, int32_t x0, int32_t y0, uint32_t phase0, uint32_t step,
		uint64_t n0, int32_t *xy, size_t n, int device, void *stream) {
	return zc_nco_rotate_ex(p, x0, y0, phase0, step, n0, xy, n, device, stream, ZC_F_DEFAULT);
}

int zc_nco_mix(const zc_params *p, const int32_t *xy_in, uint32_t phase0, uint32_t step, uint64_t n0,
		int32_t *xy_out, size_t n, int device, void *stream) {
	int rc = check_params(p, ZC_MODE_P2R);
	if (rc != ZC_OK) return rc;
	if (n && (!xy_in || !xy_out)) return set_error(ZC_EINVAL, "NULL buffer");
	CoreConsts c;
	fill_consts(p, c);
	c.nco_phase0 = phase0; c.nco_step = step; c.nco_n0 = (uint32_t)n0;
	return launch_rotate<SRC_MIX>(p, c, nullptr, xy_in, xy_out, n, device, stream, ZC_F_DEFAULT);
}

int zc_lut_sin(int pw, int ow, const uint32_t *tbl_dev, const uint32_t *phase32, int32_t *out, size_t n,
		int device, void *stream) {
	return launch_lut<false>(pw, ow, tbl_dev, phase32, out, n, device, stream);
}
int zc_lut_qwav(int pw, int ow, const uint32_t *tbl_dev, const uint32_t *phase32, int32_t *out, size_t n,
		int device, void *stream) {
	return launch_lut<true>(pw, ow, tbl_dev, phase32, out, n, device, stream);
}

int zc_lut_sin_o16(int pw, int ow, const uint32_t *tbl_dev, const uint32_t *phase32, int16_t *out, size_t n,
		int device, void *stream) {
	return launch_lut<false>(pw, ow, tbl_dev, phase32, out, n, device, stream, true);
}
int zc_lut_qwav_o16(int pw, int ow, const uint32_t *tbl_dev, const uint32_t *phase32, int16_t *out, size_t n,
		int device, void *stream) {
	return launch_lut<true>(pw, ow, tbl_dev, phase32, out, n, device, stream, true);
}

int zc_nco_lut_sin(int pw, int ow, const uint32_t *tbl_dev, uint32_t phase0, uint32_t step, uint64_t n0, int32_t *out,
		size_t n, int device, void *stream) {
	const uint32_t nco[3] = {phase0, step, (uint32_t)n0};		// arithmetic is mod 2^32
	return launch_lut<false>(pw, ow, tbl_dev, nullptr, out, n, device, stream, false, nco);
}
int zc_nco_lut_qwav(int pw, int ow, const uint32_t *tbl_dev, uint32_t phase0, uint32_t step, uint64_t n0, int32_t *out,
		size_t n, int device, void *stream) {
	const uint32_t nco[3] = {phase0, step, (uint32_t)n0};
	return launch_lut<true>(pw, ow, tbl_dev, nullptr, out, n, device, stream, false, nco);
}

int zc_quadtbl_sin(const zc_quadtbl *q, const uint32_t *phase32, int32_t *out, size_t n, int device, void *stream) {
	return launch_quadtbl(q, phase32, out, n, device, stream);
}

// ---- host buffers --------------------------------------------------------------------------
int zc_rotate_const_host(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase, int32_t *xy,
		size_t n, int device) {
	int rc = check_params(p, ZC_MODE_P2R);
	if (rc != ZC_OK) return rc;
	if (n && (!phase || !xy)) return set_error(ZC_EINVAL, "NULL buffer");
	const Lane in[2] = {{4, (const char *)phase, nullptr}, {0, nullptr, nullptr}};
	const Lane out[2] = {{8, nullptr, (char *)xy}, {0, nullptr, nullptr}};
	return host_pipeline(device, n, in, out,
		[&](size_t, size_t cnt, char *i0, char *, char *o0, char *, cudaStream_t st) {
			return zc_rotate_const_ex(p, x0, y0, (const uint32_t *)i0, (int32_t *)o0, cnt, device, st, ZC_F_DEFAULT);
		});
}

int zc_rotate_const_o16_host(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase, int16_t *xy16,
		size_t n, int device) {
	int rc = check_params(p, ZC_MODE_P2R);
	if (rc != ZC_OK) return rc;
	if (p->ow > 16) return set_error(ZC_ERANGE, "packed int16 outputs need OW <= 16 (OW=%d)", p->ow);
	if (n && (!phase || !xy16)) return set_error(ZC_EINVAL, "NULL buffer");
	const Lane in[2] = {{4, (const char *)phase, nullptr}, {0, nullptr, nullptr}};
	const Lane out[2] = {{4, nullptr, (char *)xy16}, {0, nullptr, nullptr}};
	return host_pipeline(device, n, in, out,
		[&](size_t, size_t cnt, char *i0, char *, char *o0, char *, cudaStream_t st) {
			return zc_rotate_const_o16(p, x0, y0, (const uint32_t *)i0, (int16_t *)o0, cnt, device, st);
		});
}

int zc_topolar_i16_host(const zc_params *p, const int16_t *xy16_in, int32_t *mag, uint32_t *phase, size_t n,
		int device) {
	int rc = check_params(p, ZC_MODE_R2P);
	if (rc != ZC_OK) return rc;
	if (p->iw > 16) return set_error(ZC_ERANGE, "packed int16 inputs need IW <= 16 (IW=%d)", p->iw);
	if (n && (!xy16_in || !mag || !phase)) return set_error(ZC_EINVAL, "NULL buffer");
	const Lane in[2] = {{4, (const char *)xy16_in, nullptr}, {0, nullptr, nullptr}};
	const Lane out[2] = {{4, nullptr, (char *)mag}, {4, nullptr, (char *)phase}};
	return host_pipeline(device, n, in, out,
		[&](size_t, size_t cnt, char *i0, char *, char *o0, char *o1, cudaStream_t st) {
			return zc_topolar_i16(p, (const int16_t *)i0, (int32_t *)o0, (uint32_t *)o1, cnt, device, st);
		});
}

int zc_rotate_host(const zc_params *p, const int32_t *xy_in, const uint32_t *phase, int32_t *xy_out,
		size_t n, int device) {
	int rc = check_params(p, ZC_MODE_P2R);
	if (rc != ZC_OK) return rc;
	if (n && (!phase || !xy_in || !xy_out)) return set_error(ZC_EINVAL, "NULL buffer");
	const Lane in[2] = {{4, (const char *)phase, nullptr}, {8, (const char *)xy_in, nullptr}};
	const Lane out[2] = {{8, nullptr, (char *)xy_out}, {0, nullptr, nullptr}};
	return host_pipeline(device, n, in, out,
		[&](size_t, size_t cnt, char *i0, char *i1, char *o0, char *, cudaStream_t st) {
			return zc_rotate_ex(p, (const int32_t *)i1, (const uint32_t *)i0, (int32_t *)o0, cnt, device, st, ZC_F_DEFAULT);
		});
}

int zc_topolar_host(const zc_params *p, const int32_t *xy_in, int32_t *mag, uint32_t *phase, size_t n,
		int device) {
	int rc = check_params(p, ZC_MODE_R2P);
	if (rc != ZC_OK) return rc;
	if (n && (!xy_in || !mag || !phase)) return set_error(ZC_EINVAL, "NULL buffer");
	const Lane in[2] = {{8, (const char *)xy_in, nullptr}, {0, nullptr, nullptr}};
	const Lane out[2] = {{4, nullptr, (char *)mag}, {4, nullptr, (char *)phase}};
	return host_pipeline(device, n, in, out,
		[&](size_t, size_t cnt, char *i0, char *, char *o0, char *o1, cudaStream_t st) {
			return zc_topolar_ex(p, (const int32_t *)i0, (int32_t *)o0, (uint32_t *)o1, cnt, device, st, ZC_F_DEFAULT);
		});
}

int zc_nco_rotate_host(const zc_params *p, int32_t x0, int32_t y0, uint32_t phase0, uint32_t step,
		uint64_t n0, int32_t *xy, size_t n, int device) {
	int rc = check_params(p, ZC_MODE_P2R);
	if (rc != ZC_OK) return rc;
	if (n && !xy) return set_error(ZC_EINVAL, "NULL buffer");
	const Lane in[2] = {{0, nullptr, nullptr}, {0, nullptr, nullptr}};
	const Lane out[2] = {{8, nullptr, (char *)xy}, {0, nullptr, nullptr}};
	return host_pipeline(device, n, in, out,
		[&](size_t off, size_t cnt, char *, char *, char *o0, char *, cudaStream_t st) {
			return zc_nco_rotate_ex(p, x0, y0, phase0, step, n0 + off, (int32_t *)o0, cnt, device, st, ZC_F_DEFAULT);
		});
}

// nco: {phase0, step} and n0 instead of a phase stream (zc_nco_lut_*_host)
static int lut_host(bool quarter, int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32,
		void *out, size_t n, int device, bool out16 = false, const uint32_t *nco = nullptr, uint64_t n0 = 0) {
	int rc = check_lut(quarter, pw, ow);
	if (rc != ZC_OK) return rc;
	if (out16 && ow > 16) return set_error(ZC_ERANGE, "packed int16 outputs need OW <= 16 (OW=%d)", ow);
	if (!tbl_host || (n && ((!nco && !phase32) || !out))) return set_error(ZC_EINVAL, "NULL buffer");
	DeviceInfo di;
	if ((rc = device_info(device, di)) != ZC_OK) return rc;
	if (n == 0) return ZC_OK;
	uint32_t *tbl_dev = nullptr;
	const size_t words = quarter ? ((size_t)1 << (pw - 2)) : ((size_t)1 << pw);
	{
		DeviceScope scope;
		if ((rc = scope.enter(device)) != ZC_OK) return rc;
		ZC_CUDA(cudaMalloc((void **)&tbl_dev, words * 4));
		cudaError_t e = cudaMemcpy(tbl_dev, tbl_host, words * 4, cudaMemcpyHostToDevice);
		if (e != cudaSuccess) {
			cudaFree(tbl_dev);
			return set_error(ZC_ECUDA, "table upload failed: %s", cudaGetErrorString(e));
		}
	}
	const Lane in[2] = {{(size_t)(nco ? 0 : 4), (const char *)phase32, nullptr}, {0, nullptr, nullptr}};
	const Lane outl[2] = {{(size_t)(out16 ? 2 : 4), nullptr, (char *)out}, {0, nullptr, nullptr}};
	rc = host_pipeline(device, n, in, outl,
		[&](size_t off, size_t cnt, char *i0, char *, char *o0, char *, cudaStream_t st) {
			if (nco)
				return quarter ? zc_nco_lut_qwav(pw, ow, tbl_dev, nco[0], nco[1], n0 + off, (int32_t *)o0, cnt, device, st)
					       : zc_nco_lut_sin(pw, ow, tbl_dev, nco[0], nco[1], n0 + off, (int32_t *)o0, cnt, device, st);
			if (out16)
				return quarter ? zc_lut_qwav_o16(pw, ow, tbl_dev, (const uint32_t *)i0, (int16_t *)o0, cnt, device, st)
					       : zc_lut_sin_o16(pw, ow, tbl_dev, (const uint32_t *)i0, (int16_t *)o0, cnt, device, st);
			return quarter ? zc_lut_qwav(pw, ow, tbl_dev, (const uint32_t *)i0, (int32_t *)o0, cnt, device, st)
				       : zc_lut_sin(pw, ow, tbl_dev, (const uint32_t *)i0, (int32_t *)o0, cnt, device, st);
		});
	{
		DeviceScope scope;
		if (scope.enter(device) == ZC_OK) cudaFree(tbl_dev);
	}
	return rc;
}

int zc_nco_mix_host(const zc_params *p, const int32_t *xy_in, uint32_t phase0, uint32_t step, uint64_t n0,
		int32_t *xy_out, size_t n, int device) {
	int rc = check_params(p, ZC_MODE_P2R);
	if (rc != ZC_OK) return rc;
	if (n && (!xy_in || !xy_out)) return set_error(ZC_EINVAL, "NULL buffer");
	const Lane in[2] = {{8, (const char *)xy_in, nullptr}, {0, nullptr, nullptr}};
	const Lane out[2] = {{8, nullptr, (char *)xy_out}, {0, nullptr, nullptr}};
	return host_pipeline(device, n, in, out,
		[&](size_t off, size_t cnt, char *i0, char *, char *o0, char *, cudaStream_t st) {
			return zc_nco_mix(p, (const int32_t *)i0, phase0, step, n0 + off, (int32_t *)o0, cnt, device, st);
		});
}

// $readmemh files in the layout of sw/hexfile.cpp:78-89 ("@%08x " every 8 words, "%0*lx " per word)
int zc_hex_write(const char *path, const uint32_t *words, size_t nwords, int bits) {
	if (!path || !words || bits < 1 || bits > 32) return set_error(ZC_EINVAL, "bad argument");
	FILE *fp = fopen(path, "w");
	if (!fp) return set_error(ZC_EINVAL, "cannot open %s for writing", path);
	const int nc = (bits + 3) / 4;
	const uint32_t mask = (bits >= 32) ? 0xffffffffu : ((1u << bits) - 1u);
	for (size_t k = 0; k < nwords; k++) {
		if (k % 8 == 0) fprintf(fp, "%s@%08x ", k ? "\n" : "", (unsigned)k);
		fprintf(fp, "%0*lx ", nc, (unsigned long)(words[k] & mask));
	}
	fprintf(fp, "\n");
	fclose(fp);
	return ZC_OK;
}

long zc_hex_read(const char *path, uint32_t *words, size_t max_words) {
	if (!path || !words) return set_error(ZC_EINVAL, "bad argument");
	FILE *fp = fopen(path, "r");
	if (!fp) return set_error(ZC_EINVAL, "cannot open %s", path);
	size_t addr = 0, count = 0;
	char tok[64];
	while (fscanf(fp, "%63s", tok) == 1) {
		if (tok[0] == '@') { addr = strtoul(tok + 1, nullptr, 16); continue; }
		if (addr >= max_words) { fclose(fp); return set_error(ZC_ERANGE, "%s holds more than %zu words", path, max_words); }
		words[addr++] = (uint32_t)strtoul(tok, nullptr, 16);
		if (addr > count) count = addr;
	}
	fclose(fp);
	return (long)count;
}

int zc_quadtbl_sin_host(const zc_quadtbl *q, const uint32_t *phase32, int32_t *out, size_t n, int device) {
	int rc = check_qtbl(q);
	if (rc != ZC_OK) return rc;
	if (n && (!phase32 || !out)) return set_error(ZC_EINVAL, "NULL buffer");
	const Lane in[2] = {{4, (const char *)phase32, nullptr}, {0, nullptr, nullptr}};
	const Lane outl[2] = {{4, nullptr, (char *)out}, {0, nullptr, nullptr}};
	return host_pipeline(device, n, in, outl,
		[&](size_t, size_t cnt, char *i0, char *, char *o0, char *, cudaStream_t st) {
			return zc_quadtbl_sin(q, (const uint32_t *)i0, (int32_t *)o0, cnt, device, st);
		});
}

int zc_lut_sin_host(int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32, int32_t *out,
		size_t n, int device) {
	return lut_host(false, pw, ow, tbl_host, phase32, out, n, device);
}
int zc_lut_qwav_host(int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32, int32_t *out,
		size_t n, int device) {
	return lut_host(true, pw, ow, tbl_host, phase32, out, n, device);
}
int zc_nco_lut_sin_host(int pw, int ow, const uint32_t *tbl_host, uint32_t phase0, uint32_t step, uint64_t n0, int32_t *out,
		size_t n, int device) {
	const uint32_t nco[2] = {phase0, step};
	return lut_host(false, pw, ow, tbl_host, nullptr, out, n, device, false, nco, n0);
}
int zc_nco_lut_qwav_host(int pw, int ow, const uint32_t *tbl_host, uint32_t phase0, uint32_t step, uint64_t n0, int32_t *out,
		size_t n, int device) {
	const uint32_t nco[2] = {phase0, step};
	return lut_host(true, pw, ow, tbl_host, nullptr, out, n, device, false, nco, n0);
}
int zc_lut_sin_o16_host(int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32, int16_t *out,
		size_t n, int device) {
	return lut_host(false, pw, ow, tbl_host, phase32, out, n, device, true);
}
int zc_lut_qwav_o16_host(int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32, int16_t *out,
		size_t n, int device) {
	return lut_host(true, pw, ow, tbl_host, phase32, out, n, device, true);
}

} // extern "C"
