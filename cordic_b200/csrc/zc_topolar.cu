// zc_topolar.cu -- instantiates the fully unrolled vectoring kernels k_topolar<N, NTAIL, IN16> (zc_kernels.cuh) and picks
// one at run time.  A translation unit of its own so that the library builds in parallel.
#include "zc_seedplan.h"

namespace zc {

// NTAIL is quantised to even counts (at most 16) to bound the number of instantiations
template <bool IN16, int N, int T>
struct VecTail {
	static void launch(int ntail, int grid, cudaStream_t st, const int4 *xin, int4 *mag, int4 *ph,
			size_t groups, const CoreConsts &c) {
		if constexpr (T > N) VecTail<IN16, N, T - 2>::launch(ntail, grid, st, xin, mag, ph, groups, c);
		else if (ntail >= T) k_topolar<N, T, IN16><<<grid, 256, 0, st>>>(xin, mag, ph, groups, c);
		else VecTail<IN16, N, T - 2>::launch(ntail, grid, st, xin, mag, ph, groups, c);
	}
};
template <bool IN16, int N>
struct VecTail<IN16, N, 0> {
	static void launch(int, int grid, cudaStream_t st, const int4 *xin, int4 *mag, int4 *ph,
			size_t groups, const CoreConsts &c) {
		k_topolar<N, 0, IN16><<<grid, 256, 0, st>>>(xin, mag, ph, groups, c);
	}
};

template <bool IN16, int N>
struct VecTable {
	static void launch(int neff, int ntail, int grid, cudaStream_t st, const int4 *xin, int4 *mag, int4 *ph,
			size_t groups, const CoreConsts &c) {
		if (neff == N) VecTail<IN16, N, 16>::launch(ntail, grid, st, xin, mag, ph, groups, c);
		else VecTable<IN16, N - 1>::launch(neff, ntail, grid, st, xin, mag, ph, groups, c);
	}
};
template <bool IN16>
struct VecTable<IN16, 0> {
	static void launch(int, int, int, cudaStream_t, const int4 *, int4 *, int4 *, size_t, const CoreConsts &) {}
};


void launch_topolar_plain(bool in16, int neff, int ntail, int grid, cudaStream_t st, const int4 *xin, int4 *mag, int4 *ph,
		size_t groups, const CoreConsts &c) {
	if (in16) VecTable<true, 32>::launch(neff, ntail, grid, st, xin, mag, ph, groups, c);
	else VecTable<false, 32>::launch(neff, ntail, grid, st, xin, mag, ph, groups, c);
}

} // namespace zc
