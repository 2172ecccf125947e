// zc_quadtbl.cuh -- the quadratically interpolated sine table of rtl/quadtbl.v as a batched kernel.
//
// Per sample (rtl/quadtbl.v:143-291, a 6-clock feed-forward pipeline in the RTL, a pure function here):
//   idx = i_phase[PW-1:DXBITS-1]; dx = {0, i_phase[DXBITS-2:0]}
//   lsum = (qtbl[idx]*dx)[..:DXBITS-1] + ltbl[idx]         (LBITS-bit register, wraps)
//   r    = (lsum*dx)[..:DXBITS-1] + ctbl[idx]              (CBITS-bit register, wraps)
//   o_sin = convergent-round r to OW bits -- except that r is passed through unrounded when its top OW bits are
//           0111..1 or 1100..0 (the two patterns :262-268 tests for) -- then drop XTRA bits.
// The three coefficient tables (64 entries each for the shipped core) sit in shared memory interleaved as one 16-byte
// row {c, l, q, 0} per index, so a sample costs one address computation and one LDS.128 instead of three of each
// (the ALU pipe is this kernel's busiest: 82 % before); a phase sweep reads the rows as broadcasts.  Streaming: 4 samples per thread, 128-bit loads/stores, 8 bytes per sample: HBM-bound.
#ifndef ZC_QUADTBL_CUH
#define ZC_QUADTBL_CUH

#include "zc_kernels.cuh"

namespace zc {

struct QtConsts {
	int32_t pshift;		// 32-PW
	int32_t dxs;		// DXBITS-1: index shift and product renormalisation
	uint32_t dxmask;	// 2^(DXBITS-1)-1
	int32_t qsh, lsh, csh;	// 32-QBITS, 32-LBITS, 32-CBITS: sign extension / register wrap
	int32_t xtra;		// XTRA = WW-OW
	int32_t rc;		// 2^(XTRA-1)-1
	int32_t keep_hi;	// 2^(OW-1)-1 : r>>XTRA patterns that must not be rounded (rtl/quadtbl.v:262-268)
	int32_t keep_lo;	// -2^(OW-2)
	int32_t osh;		// 32-OW
	int32_t ntbl;		// 2^LGTBL
};

__device__ __forceinline__ int wrap_bits(int v, int sh) { return (int)((uint32_t)v << sh) >> sh; }

// ct/lt/qt hold the coefficients already sign-extended (done once at upload).  NOWRAP: the host has checked,
// from the actual table contents, that neither the LBITS-bit lsum nor the CBITS-bit r_value register can
// overflow, so the two register wraps are skipped (they are kept otherwise: the RTL registers do wrap).
template <bool WIDE, bool NOWRAP>
__device__ __forceinline__ int quadtbl_one(uint32_t phase32, const int4 *__restrict__ rows, const QtConsts &c) {
	const uint32_t ip = phase32 >> c.pshift;
	const uint32_t idx = ip >> c.dxs;
	const int dx = (int)(ip & c.dxmask);
	const int4 e = rows[idx];
	const int cv = e.x, lv = e.y, qv = e.z;
	// (qv*dx)[QBITS+DXBITS-1 : DXBITS-1] sign-extended to LBITS bits == arithmetic shift (rtl/quadtbl.v:196-199)
	const int wq = WIDE ? (int)(((long long)qv * dx) >> c.dxs) : ((qv * dx) >> c.dxs);
	int lsum = wq + lv;
	if (!NOWRAP) lsum = wrap_bits(lsum, c.lsh);
	const int wl = WIDE ? (int)(((long long)lsum * dx) >> c.dxs) : ((lsum * dx) >> c.dxs);
	int r = wl + cv;
	if (!NOWRAP) r = wrap_bits(r, c.csh);
	const int t = r >> c.xtra;
	const int rounded = (r + c.rc + (t & 1)) >> c.xtra;	// cannot leave WW bits unless t is the guarded maximum
	return (t == c.keep_hi || t == c.keep_lo) ? t : (NOWRAP ? rounded : wrap_bits(rounded, c.osh));
}

template <bool WIDE, bool NOWRAP>
__global__ void __launch_bounds__(256)
k_quadtbl(const int4 *__restrict__ phase4, int4 *__restrict__ out4, const int *__restrict__ tables,
		size_t ngroups, const uint32_t *__restrict__ phase_tail, int32_t *__restrict__ out_tail, int ntail,
		const __grid_constant__ QtConsts c) {
	extern __shared__ __align__(16) int4 qsm[];
	for (int i = threadIdx.x; i < c.ntbl; i += blockDim.x) qsm[i] = reinterpret_cast<const int4 *>(tables)[i];
	__syncthreads();
	const int4 *rows = qsm;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	const size_t first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (size_t g = first; g < ngroups; g += stride) {
		const int4 pv = ldg_stream(phase4 + g);
		int4 o;
		o.x = quadtbl_one<WIDE, NOWRAP>((uint32_t)pv.x, rows, c);
		o.y = quadtbl_one<WIDE, NOWRAP>((uint32_t)pv.y, rows, c);
		o.z = quadtbl_one<WIDE, NOWRAP>((uint32_t)pv.z, rows, c);
		o.w = quadtbl_one<WIDE, NOWRAP>((uint32_t)pv.w, rows, c);
		stg_stream(out4 + g, o);
	}
	if (first < (size_t)ntail)		// ragged tail / misaligned buffers: scalar
		for (size_t i = first; i < (size_t)ntail; i += stride)
			out_tail[i] = quadtbl_one<WIDE, NOWRAP>(phase_tail[i], rows, c);
}

} // namespace zc
#endif
