/*
 * zcordic_nccl.h -- C ABI of libzcordic_nccl: one sample stream held by ONE device, worked on by all devices of the box.
 *
 * The reference has no call site for this (it has no communication of any kind); the shape comes from the north star:
 * "shard the sample stream across the 8 GPUs of one box with NCCL over NVLink only as a trivial scatter/gather of
 * independent chunks".  devices[0] owns n phase words and receives 2n output words (the layout of zc_rotate_const,
 * include/zcordic.h); the stream is cut into chunks, every chunk into ndev contiguous pieces, piece g is rotated on
 * devices[g] and lands at its own offset of the output, so the result equals zc_rotate_const on devices[0] alone, byte
 * for byte.  Two transports:
 *   ZC_XCHG_NCCL  ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd on two communicator sets (one for the scatter, one
 *                 for the gather), three streams per device, three staging buffers: scatter(k+1), kernel(k) and
 *                 gather(k-1) overlap;
 *   ZC_XCHG_PEER  no staging and no collective at all: with peer access enabled the rotation kernel on devices[g] loads
 *                 its phases from, and stores its outputs to, devices[0]'s memory directly over NVLink -- compute and
 *                 transfer fused in one kernel, overlapped word by word;
 *   ZC_XCHG_COPY  the staged pipeline of the NCCL transport with the transfers done by the copy engines
 *                 (cudaMemcpyPeerAsync on the peers' own streams): no communication kernel competes for the SMs.
 * Single process; the handle owns communicators, streams, events and staging buffers and is not thread-safe.
 */
#ifndef ZCORDIC_NCCL_H
#define ZCORDIC_NCCL_H

#include "zcordic.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { ZC_XCHG_NCCL = 0, ZC_XCHG_PEER = 1, ZC_XCHG_COPY = 2 };

typedef struct zc_exchange zc_exchange;

/* max_piece: the largest number of samples one device will be handed per chunk (sizes the NCCL staging buffers). */
int  zc_exchange_create(const int *devices, int ndev, int transport, size_t max_piece, zc_exchange **out);
void zc_exchange_destroy(zc_exchange *x);

/* rtl/cordic.v with constant (i_xval, i_yval) over the n phases at phase_dev0 (device memory of devices[0]); outputs to
 * xy_dev0 (2n words on devices[0]).  nchunks >= 1 pipeline chunks (NCCL transport; the peer transport has no stages to
 * overlap and always works on the whole stream at once).  Enqueues on the handle's own streams and returns
 * when everything is complete on devices[0]. */
int  zc_scatter_rotate_gather(zc_exchange *x, const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase_dev0,
		int32_t *xy_dev0, size_t n, int nchunks);

#ifdef __cplusplus
}
#endif
#endif /* ZCORDIC_NCCL_H */
