// zc_rot_plain.cu -- instantiates the fully unrolled rotation kernels k_rotate<N, SRC, OUT16> (zc_kernels.cuh), one per
// live-stage count 1..32 and input source, and picks one at run time.  A translation unit of its own so that the
// library builds in parallel.
#include "zc_seedplan.h"

namespace zc {

template <int SRC, bool OUT16, int N>
struct RotTable {
	static void launch(int neff, int grid, cudaStream_t st, const int4 *ph, const int4 *xin, int4 *out,
			size_t groups, const CoreConsts &c) {
		if (neff == N) k_rotate<N, SRC, OUT16><<<grid, 256, 0, st>>>(ph, xin, out, groups, c);
		else RotTable<SRC, OUT16, N - 1>::launch(neff, grid, st, ph, xin, out, groups, c);
	}
};
template <int SRC, bool OUT16>
struct RotTable<SRC, OUT16, 0> {
	static void launch(int, int, cudaStream_t, const int4 *, const int4 *, int4 *, size_t, const CoreConsts &) {}
};


void launch_rotate_plain(int src, bool out16, int neff, int grid, cudaStream_t st, const int4 *ph, const int4 *xin,
		int4 *out, size_t groups, const CoreConsts &c) {
	if (out16) { RotTable<SRC_CONST, true, 32>::launch(neff, grid, st, ph, xin, out, groups, c); return; }	// packed outputs: zc_rotate_const_o16 only
	switch (src) {
	case SRC_CONST: RotTable<SRC_CONST, false, 32>::launch(neff, grid, st, ph, xin, out, groups, c); break;
	case SRC_XY:    RotTable<SRC_XY, false, 32>::launch(neff, grid, st, ph, xin, out, groups, c); break;
	case SRC_NCO:   RotTable<SRC_NCO, false, 32>::launch(neff, grid, st, ph, xin, out, groups, c); break;
	default:        RotTable<SRC_MIX, false, 32>::launch(neff, grid, st, ph, xin, out, groups, c); break;
	}
}

} // namespace zc
