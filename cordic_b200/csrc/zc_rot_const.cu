// zc_rot_const.cu -- instantiates the table-seeded rotation kernels for a phase stream with a constant input vector
// (zc_seeded.cuh: k_rotate_seeded<NS, SRC_CONST, ...>, 32-bit and packed 16-bit outputs).
#include "zc_seeded.cuh"

namespace zc {

int seeded_rotate_const(const zc_params *p, const CoreConsts &c, const uint32_t *phase, void *xy_out, bool out16, size_t n,
		int device, int sms, cudaStream_t st, uint32_t flags, size_t &done, int &launches) {
	if (out16) return seeded_rotate_try<SRC_CONST, true>(p, c, phase, xy_out, n, device, sms, st, flags, done, launches);
	return seeded_rotate_try<SRC_CONST, false>(p, c, phase, xy_out, n, device, sms, st, flags, done, launches);
}

} // namespace zc
