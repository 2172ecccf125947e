// zc_seeded.cuh -- table-seeded rotation kernel (constant input vector).  Placeholder until the
// seeded path lands: reports "not used" so the caller runs every stage in registers.
#ifndef ZC_SEEDED_CUH
#define ZC_SEEDED_CUH
#include "zc_kernels.cuh"
namespace zc {
template <int SRC>
static int seeded_rotate_try(const zc_params *, const CoreConsts &, const uint32_t *, int32_t *, size_t,
		int, int, cudaStream_t, uint32_t, bool &used) {
	used = false;
	return ZC_OK;
}
} // namespace zc
#endif
