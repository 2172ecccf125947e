// zc_rot_nco.cu -- instantiates the table-seeded rotation kernels for the NCO (phase generated in registers, no input
// stream), block and comb sample mappings (zc_seeded.cuh: k_rotate_seeded<NS, SRC_NCO, ...>).
#include "zc_seeded.cuh"

namespace zc {

int seeded_rotate_nco(const zc_params *p, const CoreConsts &c, void *xy_out, size_t n, int device, int sms,
		cudaStream_t st, uint32_t flags, size_t &done, int &launches) {
	return seeded_rotate_try<SRC_NCO, false>(p, c, nullptr, xy_out, n, device, sms, st, flags, done, launches);
}

// Diagnostic behind zc_nco_comb_run: the run length the comb mapping would use for this step over n samples (0: none).
long long nco_comb_run(const zc_params *p, uint32_t step, size_t n) {
	const int32_t sstep = (int32_t)step;
	const uint32_t mag = (uint32_t)(sstep < 0 ? -(int64_t)sstep : (int64_t)sstep);
	if ((mag >> (32 - p->pw)) < 2u) return 0;		// neighbouring samples are neighbours already: block mapping
	return (long long)comb_search(step, 32 - p->pw, n);
}

} // namespace zc
