// zc_internal.h -- declarations shared by the host (.cpp) and device (.cu) halves of
// libzcordic.  Not part of the public ABI (that is include/zcordic.h).
#ifndef ZC_INTERNAL_H
#define ZC_INTERNAL_H

#include "zcordic.h"

namespace zc {

// Records a thread-local message and returns `code` (so `return set_error(...)` reads well).
int set_error(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

int derive_p2r(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *o);
int derive_r2p(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *o);
int derive_sp2r(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *o);
int derive_sr2p(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *o);
int derive_lut(bool quarter, int iw, int pw, int ow, int *pw_out, int *ow_out);
int check_lut(bool quarter, int pw, int ow);
int build_sintable(int pw, int ow, uint32_t *tbl);
int build_quarterwav(int pw, int ow, uint32_t *tbl);
int derive_qtbl(int iw, int ow, int xtra_user, int pw, zc_quadtbl *o);
int check_qtbl(const zc_quadtbl *q);

// Validates a zc_params handed back to us across the ABI (it is caller memory).
int check_params(const zc_params *p, int want_mode);

} // namespace zc

#endif
