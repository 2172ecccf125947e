// zc_exchange.cpp -- libzcordic_nccl: scatter -> rotate -> gather of a stream owned by one device (include/zcordic_nccl.h).
#include "zcordic_nccl.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <vector>

namespace {

constexpr int NBUF = 3;

struct Dev {
	int id = -1;
	cudaStream_t s_sc = nullptr, s_k = nullptr, s_ga = nullptr;
	cudaEvent_t ev_sc[NBUF] = {}, ev_k[NBUF] = {}, ev_ga[NBUF] = {};
	uint32_t *in[NBUF] = {};
	int32_t *out[NBUF] = {};
};

thread_local char g_msg[256] = "";

} // namespace

struct zc_exchange {
	int ndev = 0, transport = ZC_XCHG_NCCL;
	size_t max_piece = 0;
	std::vector<Dev> d;
	std::vector<ncclComm_t> comm_sc, comm_ga;	// two sets: the scatter of chunk k+1 must not queue behind the gather of k-1
	bool have_comms = false;
};

#define XC(call)                                                                                   \
	do {                                                                                       \
		cudaError_t e_ = (call);                                                           \
		if (e_ != cudaSuccess) {                                                           \
			std::snprintf(g_msg, sizeof(g_msg), "%s: %s", #call, cudaGetErrorString(e_)); \
			std::fprintf(stderr, "zc_exchange: %s\n", g_msg);                         \
			return ZC_ECUDA;                                                           \
		}                                                                                  \
	} while (0)
#define XN(call)                                                                                   \
	do {                                                                                       \
		ncclResult_t r_ = (call);                                                          \
		if (r_ != ncclSuccess) {                                                           \
			std::snprintf(g_msg, sizeof(g_msg), "%s: %s", #call, ncclGetErrorString(r_)); \
			std::fprintf(stderr, "zc_exchange: %s\n", g_msg);                         \
			return ZC_ECUDA;                                                           \
		}                                                                                  \
	} while (0)

extern "C" {

void zc_exchange_destroy(zc_exchange *x) {
	if (!x) return;
	for (Dev &v : x->d) {
		if (v.id < 0) continue;
		cudaSetDevice(v.id);
		cudaDeviceSynchronize();
		for (int b = 0; b < NBUF; b++) {
			if (v.in[b]) cudaFree(v.in[b]);
			if (v.out[b]) cudaFree(v.out[b]);
			if (v.ev_sc[b]) cudaEventDestroy(v.ev_sc[b]);
			if (v.ev_k[b]) cudaEventDestroy(v.ev_k[b]);
			if (v.ev_ga[b]) cudaEventDestroy(v.ev_ga[b]);
		}
		if (v.s_sc) cudaStreamDestroy(v.s_sc);
		if (v.s_k) cudaStreamDestroy(v.s_k);
		if (v.s_ga) cudaStreamDestroy(v.s_ga);
	}
	if (x->have_comms) {
		for (ncclComm_t c : x->comm_sc) ncclCommDestroy(c);
		for (ncclComm_t c : x->comm_ga) ncclCommDestroy(c);
	}
	delete x;
}

int zc_exchange_create(const int *devices, int ndev, int transport, size_t max_piece, zc_exchange **out) {
	if (!devices || !out || ndev < 1 || ndev > 64 || transport < ZC_XCHG_NCCL || transport > ZC_XCHG_COPY) return ZC_EINVAL;
	zc_exchange *x = new zc_exchange();
	x->ndev = ndev; x->transport = transport; x->max_piece = max_piece;
	x->d.resize(ndev);
	int prev = 0;
	cudaGetDevice(&prev);
	auto fail = [&](int rc) { zc_exchange_destroy(x); cudaSetDevice(prev); return rc; };
	for (int g = 0; g < ndev; g++) {
		Dev &v = x->d[g];
		v.id = devices[g];
		if (cudaSetDevice(v.id) != cudaSuccess) return fail(ZC_ENODEV);
		cudaError_t e = cudaStreamCreateWithFlags(&v.s_sc, cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&v.s_k, cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&v.s_ga, cudaStreamNonBlocking);
		for (int b = 0; b < NBUF && e == cudaSuccess; b++) {
			e = cudaEventCreateWithFlags(&v.ev_sc[b], cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&v.ev_k[b], cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&v.ev_ga[b], cudaEventDisableTiming);
			if (e == cudaSuccess && transport != ZC_XCHG_PEER && g > 0) {
				e = cudaMalloc((void **)&v.in[b], max_piece * 4);
				if (e == cudaSuccess) e = cudaMalloc((void **)&v.out[b], max_piece * 8);
			}
		}
		if (e != cudaSuccess) { std::fprintf(stderr, "zc_exchange_create: %s\n", cudaGetErrorString(e)); return fail(ZC_ECUDA); }
		if (transport != ZC_XCHG_NCCL && g > 0) {		// devices[g] reads and writes devices[0]'s memory directly (kernels or copy engines)
			int can = 0;
			cudaDeviceCanAccessPeer(&can, v.id, devices[0]);
			if (!can) { std::fprintf(stderr, "zc_exchange_create: device %d cannot access device %d\n", v.id, devices[0]); return fail(ZC_ENODEV); }
			e = cudaDeviceEnablePeerAccess(devices[0], 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ZC_ECUDA);
			cudaGetLastError();
			// and the owner -> peer direction: the copy engines take the direct NVLink path only between mutually mapped devices
			// (measured at N=4: 128 Gsamples/s after NCCL had mapped both directions, 89 with the peer -> owner mapping alone)
			if (cudaSetDevice(devices[0]) == cudaSuccess) {
				e = cudaDeviceEnablePeerAccess(v.id, 0);
				cudaGetLastError();
			}
			cudaSetDevice(v.id);
		}
	}
	if (transport == ZC_XCHG_NCCL && ndev > 1) {
		x->comm_sc.resize(ndev); x->comm_ga.resize(ndev);
		if (ncclCommInitAll(x->comm_sc.data(), ndev, devices) != ncclSuccess) return fail(ZC_ECUDA);
		if (ncclCommInitAll(x->comm_ga.data(), ndev, devices) != ncclSuccess) {
			for (ncclComm_t c : x->comm_sc) ncclCommDestroy(c);
			return fail(ZC_ECUDA);
		}
		x->have_comms = true;
	}
	cudaSetDevice(prev);
	*out = x;
	return ZC_OK;
}

static int scatter_rotate_gather(zc_exchange *x, const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase,
		int32_t *xy, size_t n, int nchunks);

int zc_scatter_rotate_gather(zc_exchange *x, const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase,
		int32_t *xy, size_t n, int nchunks) {
	if (!x || !p || (n && (!phase || !xy)) || nchunks < 1) return ZC_EINVAL;
	int prev = 0;
	cudaGetDevice(&prev);
	const int rc = scatter_rotate_gather(x, p, x0, y0, phase, xy, n, nchunks);	// every exit restores the caller's device
	cudaSetDevice(prev);
	return rc;
}

static int scatter_rotate_gather(zc_exchange *x, const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase,
		int32_t *xy, size_t n, int nchunks) {
	const int G = x->ndev;
	// chunk c = [c0, c1); piece g of it = [c0 + g*len/G, c0 + (g+1)*len/G), boundaries at multiples of 128 samples so
	// that every piece keeps the alignment (and the table-seeded kernel's block size) of the whole
	// the peer transport has no stages to overlap: one chunk, the largest pieces
	if (x->transport == ZC_XCHG_PEER || G == 1) nchunks = 1;
	const size_t per_chunk = ((n + (size_t)nchunks - 1) / (size_t)nchunks + 127) & ~(size_t)127;
	auto piece_lo = [&](size_t c0, size_t len, int g) { return g >= G ? c0 + len : c0 + ((len * (size_t)g / (size_t)G) & ~(size_t)127); };
	int rc = ZC_OK;
	size_t ci = 0;
	for (size_t c0 = 0; c0 < n && rc == ZC_OK; c0 += per_chunk, ci++) {
		const size_t len = (n - c0 < per_chunk) ? (n - c0) : per_chunk;
		const int b = (int)(ci % NBUF);
		if (x->transport == ZC_XCHG_PEER || G == 1) {
			for (int g = 0; g < G && rc == ZC_OK; g++) {
				const size_t lo = piece_lo(c0, len, g), cnt = piece_lo(c0, len, g + 1) - lo;
				if (!cnt) continue;
				XC(cudaSetDevice(x->d[g].id));
				rc = zc_rotate_const(p, x0, y0, phase + lo, xy + 2 * lo, cnt, x->d[g].id, x->d[g].s_k);
			}
			continue;
		}
		// ---- staged transports: NCCL send/recv groups, or copy-engine peer copies ---------------------------------
		const bool copy = x->transport == ZC_XCHG_COPY;
		for (int g = 1; g < G; g++) {
			if (piece_lo(c0, len, g + 1) - piece_lo(c0, len, g) > x->max_piece) return ZC_ERANGE;
			if (ci >= NBUF) {	// staging buffer b of device g: its previous outputs must have left
				XC(cudaSetDevice(x->d[g].id));
				XC(cudaStreamWaitEvent(x->d[g].s_sc, x->d[g].ev_ga[b], 0));
			}
		}
		if (copy) {		// one copy engine transfer per peer, pulled on the peer's own scatter stream
			for (int g = 1; g < G; g++) {
				const size_t lo = piece_lo(c0, len, g), cnt = piece_lo(c0, len, g + 1) - lo;
				if (!cnt) continue;
				XC(cudaSetDevice(x->d[g].id));
				XC(cudaMemcpyPeerAsync(x->d[g].in[b], x->d[g].id, phase + lo, x->d[0].id, cnt * 4, x->d[g].s_sc));
			}
		} else {
			XN(ncclGroupStart());
			for (int g = 1; g < G; g++) {
				const size_t lo = piece_lo(c0, len, g), cnt = piece_lo(c0, len, g + 1) - lo;
				if (!cnt) continue;
				XN(ncclSend(phase + lo, cnt, ncclUint32, g, x->comm_sc[0], x->d[0].s_sc));
				XN(ncclRecv(x->d[g].in[b], cnt, ncclUint32, 0, x->comm_sc[g], x->d[g].s_sc));
			}
			XN(ncclGroupEnd());
		}
		for (int g = 0; g < G && rc == ZC_OK; g++) {
			const size_t lo = piece_lo(c0, len, g), cnt = piece_lo(c0, len, g + 1) - lo;
			Dev &v = x->d[g];
			XC(cudaSetDevice(v.id));
			if (g == 0) {		// the owner works in place
				if (cnt) rc = zc_rotate_const(p, x0, y0, phase + lo, xy + 2 * lo, cnt, v.id, v.s_k);
				continue;
			}
			XC(cudaEventRecord(v.ev_sc[b], v.s_sc));
			XC(cudaStreamWaitEvent(v.s_k, v.ev_sc[b], 0));
			if (ci >= NBUF) XC(cudaStreamWaitEvent(v.s_k, v.ev_ga[b], 0));
			if (cnt) rc = zc_rotate_const(p, x0, y0, v.in[b], v.out[b], cnt, v.id, v.s_k);
			XC(cudaEventRecord(v.ev_k[b], v.s_k));
			XC(cudaStreamWaitEvent(v.s_ga, v.ev_k[b], 0));
		}
		if (rc != ZC_OK) break;
		if (copy) {		// pushed by the peer's own gather stream into the owner's output
			for (int g = 1; g < G; g++) {
				const size_t lo = piece_lo(c0, len, g), cnt = piece_lo(c0, len, g + 1) - lo;
				if (!cnt) continue;
				XC(cudaSetDevice(x->d[g].id));
				XC(cudaMemcpyPeerAsync(xy + 2 * lo, x->d[0].id, x->d[g].out[b], x->d[g].id, cnt * 8, x->d[g].s_ga));
			}
		} else {
			XN(ncclGroupStart());
			for (int g = 1; g < G; g++) {
				const size_t lo = piece_lo(c0, len, g), cnt = piece_lo(c0, len, g + 1) - lo;
				if (!cnt) continue;
				XN(ncclSend(x->d[g].out[b], 2 * cnt, ncclInt32, 0, x->comm_ga[g], x->d[g].s_ga));
				XN(ncclRecv(xy + 2 * lo, 2 * cnt, ncclInt32, g, x->comm_ga[0], x->d[0].s_ga));
			}
			XN(ncclGroupEnd());
		}
		for (int g = 1; g < G; g++) {
			XC(cudaSetDevice(x->d[g].id));
			XC(cudaEventRecord(x->d[g].ev_ga[b], x->d[g].s_ga));
		}
	}
	for (int g = 0; g < G; g++) {		// complete on every device (the owner's gather stream last)
		cudaSetDevice(x->d[g].id);
		cudaError_t e = cudaStreamSynchronize(x->d[g].s_sc);
		if (e == cudaSuccess) e = cudaStreamSynchronize(x->d[g].s_k);
		if (e == cudaSuccess) e = cudaStreamSynchronize(x->d[g].s_ga);
		if (e != cudaSuccess && rc == ZC_OK) { std::fprintf(stderr, "zc_exchange: %s\n", cudaGetErrorString(e)); rc = ZC_ECUDA; }
	}
	return rc;
}

} // extern "C"
