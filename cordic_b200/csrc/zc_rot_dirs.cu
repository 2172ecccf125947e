// zc_rot_dirs.cu -- instantiates the table-directed rotation kernels for per-sample input vectors
// (zc_seeded.cuh: k_rotate_dirs<NS, SRC_XY | SRC_MIX, RF>).
#include "zc_seeded.cuh"

namespace zc {

int dirs_rotate_xy(const zc_params *p, const CoreConsts &c, const uint32_t *phase, const int32_t *xy_in, int32_t *xy_out,
		size_t n, int device, int sms, cudaStream_t st, uint32_t flags, size_t &done, int &launches) {
	return dirs_rotate_try<SRC_XY>(p, c, phase, xy_in, xy_out, n, device, sms, st, flags, done, launches);
}
int dirs_rotate_mix(const zc_params *p, const CoreConsts &c, const int32_t *xy_in, int32_t *xy_out, size_t n, int device,
		int sms, cudaStream_t st, uint32_t flags, size_t &done, int &launches) {
	return dirs_rotate_try<SRC_MIX>(p, c, nullptr, xy_in, xy_out, n, device, sms, st, flags, done, launches);
}

} // namespace zc
