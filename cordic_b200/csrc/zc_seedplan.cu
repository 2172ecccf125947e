// zc_seedplan.cu -- builds, caches and evicts the table plans of the table-seeded rotation kernels (see zc_seeded.cuh
// for what the tables mean).  Host code is phase arithmetic only; the x/y table is filled by a setup kernel that runs
// the real stage arithmetic.
#include "zc_seedplan.h"

#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

namespace zc {

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

// ---- host: plan construction and cache -----------------------------------------------------------

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
	const size_t b_t1 = (size_t)4 << LB, b_ts = (R * 4 + 15) & ~(size_t)15, b_t2 = R * (fl_dirs(flavour) ? 16 : 32),
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

// FL_PACKED_M: folds TS mod 2^16 (the byte offset of the interval's first direction row, modulo the table size bound)
// into the low bytes of the interval's four (x, y) records.
__global__ void k_seed_pack_records(int2 *__restrict__ t2, const uint32_t *__restrict__ ts, uint32_t R, int recsh) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 4u * R) return;
	const uint32_t c = ts[i >> 2] & 0xffffu;
	const int2 v = t2[i];
	t2[i] = make_int2((int)(((uint32_t)v.x << recsh) | (c & 0xffu)), (int)(((uint32_t)v.y << recsh) | (c >> 8)));
}

static std::mutex g_seed_mu;
static std::vector<SeedPlan> g_seed_cache;
static uint64_t g_seed_clock = 0;


// Builds (or finds) the plan for (p, constant vector, device).  Called with the device current.
int seed_plan_get(const zc_params *p, const CoreConsts &c, int device, int flavour, cudaStream_t st,
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
	if (ok && flavour == FL_DIRS_DP && pl.s.M + 1 < p->ww - 16) ok = false;	// same 16-bit condition for the suffix shifts
	if (ok && fl_merged(flavour)) {
		// eight free bits under x and y, and every TD byte offset below 2^16 (the kernel recovers it modulo 2^16)
		const size_t td_bytes = (size_t)(rmax - rmin + 1) << pl.s.lgrow;
		if (p->ww > 24 || td_bytes > 65536) ok = false;
		pl.s.recsh = 32 - p->ww;
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
		if (fl_dirs(flavour)) {		// per-interval prefix directions, one signed byte per stage, 16-byte rows
			unsigned char *tp = reinterpret_cast<unsigned char *>(host.data()) + s.off_t2;
			for (size_t k = 0; k < R; k++)
				for (int j = 0; j < DIRS_M; j++)
					tp[k * 16 + j] = (unsigned char)(((iv[k].neg >> j) & 1u) ? 0xff : 0x01);
		}
		for (size_t k = 0; k < R && ok; k++) {
			ts[k] = (uint32_t)(int32_t)((iv[k].S + half + rmin) * ((int64_t)1 << s.lgrow));
			rep[k] = (uint32_t)((uint64_t)iv[k].lo << pshift);
			// the per-interval direction rows have four spare bytes: the row offset rides there (one lookup fewer)
			if (fl_dirs(flavour)) host[s.off_t2 / 4 + k * 4 + 3] = ts[k];
		}
		const size_t nres = (size_t)(rmax - rmin + 1);
		for (int64_t res = rmin; res <= rmax && ok; res++) {
			int64_t ph = res;
			for (int j = 0; j < pl.NS; j++) {			// rtl/cordic.v:265-279, phase only
				const bool neg = ph < 0;
				if (packed)
					reinterpret_cast<unsigned char *>(td)[((size_t)(res - rmin) << s.lgrow) + j] = (unsigned char)(neg ? 0xff : 0x01);
				else if (fl_dp(flavour))	// bytes 3..0 = {0, d, 0, -d}
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
			if (e == cudaSuccess && fl_merged(flavour)) {
				const uint32_t nthreads = 4u * (uint32_t)R;
				k_seed_pack_records<<<(nthreads + 255) / 256, 256, 0, st>>>(
					reinterpret_cast<int2 *>(reinterpret_cast<char *>(pl.dev) + s.off_t2),
					reinterpret_cast<const uint32_t *>(reinterpret_cast<char *>(pl.dev) + s.off_ts), (uint32_t)R, s.recsh);
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
cudaError_t ensure_dynamic_smem(const void *kern, size_t smem) {
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

// Drops the cached plans of `device` (all devices when negative).
void seed_trim(int device) {
	std::lock_guard<std::mutex> lk(g_seed_mu);
	for (size_t k = 0; k < g_seed_cache.size();) {
		if (device < 0 || g_seed_cache[k].device == device) g_seed_cache.erase(g_seed_cache.begin() + k);
		else k++;
	}
}


} // namespace zc
