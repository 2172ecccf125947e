// zc_generic.cuh -- the kernels that take any configuration at run time: the generic rotation / vectoring kernels
// (every alignment, every width, the WW-bit register wrap modelled) and the LUT cores.  Included by zc_api.cu only.
#ifndef ZC_GENERIC_CUH
#define ZC_GENERIC_CUH

#include "zc_kernels.cuh"

namespace zc {

// ---- generic kernels: any configuration, any alignment, WW-bit wrap modelled ----------------
__device__ __forceinline__ int wrapw(int v, int wsh) { return (int)((uint32_t)v << wsh) >> wsh; }

template <int SRC>
__global__ void __launch_bounds__(256)
k_rotate_generic(const uint32_t *__restrict__ phase, const int32_t *__restrict__ xyin,
		int32_t *__restrict__ xyout, size_t n, const __grid_constant__ CoreConsts c, const int out16) {
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
		const int ox = wrapw((int)((uint32_t)x + (uint32_t)c.rc + (uint32_t)bx), c.wsh) >> c.D;
		const int oy = wrapw((int)((uint32_t)y + (uint32_t)c.rc + (uint32_t)by), c.wsh) >> c.D;
		if (out16) xyout[i] = pack16(ox, oy);
		else { xyout[2 * i] = ox; xyout[2 * i + 1] = oy; }
	}
}

__global__ void __launch_bounds__(256)
k_topolar_generic(const int32_t *__restrict__ xyin, int32_t *__restrict__ mag,
		uint32_t *__restrict__ phout, size_t n, const __grid_constant__ CoreConsts c, const int in16) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const int rx = in16 ? xyin[i] : xyin[2 * i], ry = in16 ? (xyin[i] >> 16) : xyin[2 * i + 1];
		const int ex = (rx << c.in_shl) >> c.in_shr;
		const int ey = (ry << c.in_shl) >> c.in_shr;
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

// which LUT kernel a probed batch goes to
enum { LUT_GATE_L2 = PROBE_LOCAL, LUT_GATE_SMEM = PROBE_SCATTERED };
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

// OUT16 (all three LUT kernels): o_val packed as int16, four samples per 8-byte store (tables with OW <= 16)
__device__ __forceinline__ void lut_store4(void *out, size_t g, const int4 o, bool out16) {
	if (out16) stg_stream64(reinterpret_cast<int2 *>(out) + g, make_int2(pack16(o.x, o.y), pack16(o.z, o.w)));
	else stg_stream(reinterpret_cast<int4 *>(out) + g, o);
}

template <bool QUARTER, bool OUT16 = false>
__global__ void __launch_bounds__(256)
k_lut(const int4 *__restrict__ phase4, void *__restrict__ out4, const uint32_t *__restrict__ tbl,
		size_t ngroups, const __grid_constant__ LutConsts c, const int probe_lim) {
	// the probe (same verdict in every CTA of both kernels) chose the shared-memory kernel for this batch
	if (probe_lim >= 0 && probe_local(reinterpret_cast<const uint32_t *>(phase4), ngroups << 2, 0, probe_lim) != LUT_GATE_L2) return;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
		const int4 pv = ldg_stream(phase4 + g);
		int4 o;
		o.x = lut_one<QUARTER>((uint32_t)pv.x, tbl, c);
		o.y = lut_one<QUARTER>((uint32_t)pv.y, tbl, c);
		o.z = lut_one<QUARTER>((uint32_t)pv.z, tbl, c);
		o.w = lut_one<QUARTER>((uint32_t)pv.w, tbl, c);
		lut_store4(out4, g, o, OUT16);
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
template <bool QUARTER, bool HI8, bool OUT16 = false>
__global__ void __launch_bounds__(1024, 1)
k_lut_smem(const int4 *__restrict__ phase4, void *__restrict__ out4, const uint32_t *__restrict__ tbl,
		size_t ngroups, const __grid_constant__ LutConsts c, const int probe_lim) {
	if (probe_lim >= 0 && probe_local(reinterpret_cast<const uint32_t *>(phase4), ngroups << 2, 0, probe_lim) != LUT_GATE_SMEM) return;
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
				lut_store4(out4, g + k * stride, make_int4(one((uint32_t)pv[k].x), one((uint32_t)pv[k].y),
					one((uint32_t)pv[k].z), one((uint32_t)pv[k].w)), OUT16);
		}
		for (; g < ngroups; g += stride) {
			const int4 pv = ldg_stream(phase4 + g);
			lut_store4(out4, g, make_int4(one((uint32_t)pv.x), one((uint32_t)pv.y), one((uint32_t)pv.z), one((uint32_t)pv.w)), OUT16);
		}
	} else {
		for (; g < ngroups; g += stride) {
			const int4 pv = ldg_stream(phase4 + g);
			int4 o;
			o.x = lut_one<QUARTER>((uint32_t)pv.x, tbl, c);
			o.y = lut_one<QUARTER>((uint32_t)pv.y, tbl, c);
			o.z = lut_one<QUARTER>((uint32_t)pv.z, tbl, c);
			o.w = lut_one<QUARTER>((uint32_t)pv.w, tbl, c);
			lut_store4(out4, g, o, OUT16);
		}
	}
}

template <bool QUARTER>
__global__ void __launch_bounds__(256)
k_lut_scalar(const uint32_t *__restrict__ phase, void *__restrict__ out,
		const uint32_t *__restrict__ tbl, size_t n, const __grid_constant__ LutConsts c, const int out16) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const int v = lut_one<QUARTER>(phase[i], tbl, c);
		if (out16) reinterpret_cast<short *>(out)[i] = (short)v;
		else reinterpret_cast<int32_t *>(out)[i] = v;
	}
}

} // namespace zc
#endif
