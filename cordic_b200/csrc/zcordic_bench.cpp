// zcordic_bench.cpp -- a C++ client of the C ABI (include/zcordic.h), no Python and no torch in the process: what a
// maintainer of the reference would write first.  It derives a core from the generator's flags, fills a phase sweep
// the way bench/cpp/cordic_tb.cpp:127-138 does, runs it through zc_rotate_const on the device and through
// zc_rotate_const_host end to end, and prints Gsamples/s for both.  bench.py is the driver's contract; this is the
// same measurement from the reference's own language.
//   zcordic_bench [-i iw] [-o ow] [-p pw] [-n stages] [-x xtra] [-l lg2(samples)] [-s steps] [-d device] [-g gpus]
//                 [--scatter] [--transport nccl|peer|copy|both] [--chunks C] [--json] [--pcie-probe]
// -g G: the sample stream is sharded over devices 0..G-1 as independent chunks, one host thread per device (no collective:
// the path has no exchange step); the figure is all samples over the slowest device's time.
// -g G --scatter: device 0 owns the whole stream (G * 2^l samples) and every step scatters it over the G devices,
// rotates and gathers the outputs back (libzcordic_nccl: zc_scatter_rotate_gather), over NCCL send/recv with the three
// stages pipelined in C chunks, and with the kernels reading and writing device 0's memory directly over NVLink (peer);
// the gathered output is compared byte for byte with device 0 computing the whole stream alone.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "zcordic.h"
#include "zcordic_nccl.h"

#include <chrono>
#include <string>

#define CK(call)                                                                                  \
	do {                                                                                      \
		cudaError_t e_ = (call);                                                          \
		if (e_ != cudaSuccess) {                                                          \
			std::fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));         \
			return 2;                                                                 \
		}                                                                                 \
	} while (0)
#define ZC(call)                                                                                  \
	do {                                                                                      \
		int rc_ = (call);                                                                 \
		if (rc_ != ZC_OK) {                                                               \
			std::fprintf(stderr, "%s: %s (%s)\n", #call, zc_strerror(rc_), zc_last_error()); \
			return 2;                                                                 \
		}                                                                                 \
	} while (0)

// One shard on one device: n samples starting at phase index `first`; returns ms for `steps` calls (0 on failure).
static float shard(const zc_params *p, int device, size_t n, size_t first, int steps, std::atomic<int> *ready, int gpus) {
	const int32_t x0 = (1 << (p->iw - 1)) - 1;
	const uint32_t mask = p->pw >= 32 ? 0xFFFFFFFFu : ((1u << p->pw) - 1u);
	if (cudaSetDevice(device) != cudaSuccess) return 0;
	std::vector<uint32_t> h(n);
	for (size_t i = 0; i < n; i++) h[i] = (uint32_t)(first + i) & mask;
	uint32_t *d_phase = nullptr;
	int32_t *d_xy = nullptr;
	cudaStream_t st;
	cudaEvent_t e0, e1;
	if (cudaMalloc(&d_phase, n * 4) != cudaSuccess || cudaMalloc(&d_xy, n * 8) != cudaSuccess) return 0;
	cudaMemcpy(d_phase, h.data(), n * 4, cudaMemcpyHostToDevice);
	cudaStreamCreate(&st); cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int w = 0; w < 3; w++)
		if (zc_rotate_const(p, x0, 0, d_phase, d_xy, n, device, st) != ZC_OK) return 0;
	cudaStreamSynchronize(st);
	ready->fetch_add(1);
	while (ready->load() < gpus) std::this_thread::yield();		// all devices start their timed region together
	cudaEventRecord(e0, st);
	for (int s = 0; s < steps; s++)
		if (zc_rotate_const(p, x0, 0, d_phase, d_xy, n, device, st) != ZC_OK) return 0;
	cudaEventRecord(e1, st);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaFree(d_phase); cudaFree(d_xy);
	return ms;
}

// Copy-only probe: what the host links of this box deliver when G devices move 4 B/sample in and 8 B/sample out
// concurrently (the traffic of zc_rotate_const_host), each from its own pinned buffers -- plain cudaHostAlloc memory, then
// zc_host_alloc_sharded memory (each device's shard on its own NUMA node).  The ceiling of the end-to-end figure.
static int pcie_probe(int gpus, size_t n, bool json) {
	std::vector<int> devices(gpus);
	for (int g = 0; g < gpus; g++) devices[g] = g;
	std::string out = "{\"what\": \"copy-only probe: every device moves 4 B/sample H2D and 8 B/sample D2H concurrently on two streams, "
		"pinned host memory, all devices at once; Gsamples/s = the ceiling of zc_rotate_const_host_multi\", \"n_gpus\": " + std::to_string(gpus) +
		", \"samples_per_gpu\": " + std::to_string(n);
	for (int mode = 0; mode < 2; mode++) {
		char *hin = nullptr, *hout = nullptr;
		if (mode == 0) { hin = (char *)zc_host_alloc(n * 4 * gpus); hout = (char *)zc_host_alloc(n * 8 * gpus); }
		else { hin = (char *)zc_host_alloc_sharded(n * 4 * gpus, devices.data(), gpus); hout = (char *)zc_host_alloc_sharded(n * 8 * gpus, devices.data(), gpus); }
		if (!hin || !hout) { std::fprintf(stderr, "host alloc: %s\n", zc_last_error()); return 2; }
		std::memset(hin, 1, n * 4 * gpus);
		std::vector<double> secs(gpus, 0.0);
		std::vector<std::thread> th;
		std::atomic<int> ready{0};
		for (int g = 0; g < gpus; g++)
			th.emplace_back([&, g] {
				cudaSetDevice(g);
				char *din = nullptr, *dout = nullptr;
				cudaStream_t s1, s2;
				cudaMalloc(&din, n * 4); cudaMalloc(&dout, n * 8);
				cudaStreamCreate(&s1); cudaStreamCreate(&s2);
				auto pass = [&] {
					cudaMemcpyAsync(din, hin + (size_t)g * n * 4, n * 4, cudaMemcpyHostToDevice, s1);
					cudaMemcpyAsync(hout + (size_t)g * n * 8, dout, n * 8, cudaMemcpyDeviceToHost, s2);
					cudaStreamSynchronize(s1); cudaStreamSynchronize(s2);
				};
				pass();
				ready.fetch_add(1);
				while (ready.load() < gpus) std::this_thread::yield();
				const auto t0 = std::chrono::steady_clock::now();
				for (int r = 0; r < 3; r++) pass();
				secs[g] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 3;
				cudaFree(din); cudaFree(dout);
			});
		for (auto &t : th) t.join();
		double worst = 0;
		for (double v : secs) worst = v > worst ? v : worst;
		const double gs = (double)n * gpus / worst / 1e9;
		char buf[256];
		std::snprintf(buf, sizeof(buf), ", \"%s\": {\"gsamples_per_s\": %.2f, \"h2d_gbs\": %.1f, \"d2h_gbs\": %.1f}", mode ? "numa_sharded_pinned" : "plain_pinned",
			gs, gs * 4, gs * 8);
		out += buf;
		if (!json) std::printf("%d GPUs, %s: %.2f Gsamples/s (%.1f GB/s in + %.1f GB/s out)\n", gpus, mode ? "NUMA-sharded pinned" : "plain pinned", gs, gs * 4, gs * 8);
		zc_host_free(hin); zc_host_free(hout);
	}
	out += ", \"numa_nodes\": [";
	for (int g = 0; g < gpus; g++) out += (g ? ", " : "") + std::to_string(zc_device_numa_node(g));
	out += "]}";
	if (json) std::printf("%s\n", out.c_str());
	return 0;
}

// Device 0 owns G * n samples; per step: scatter -> rotate on G devices -> gather.  Returns 0 and prints one JSON line.
static int scatter_bench(const zc_params *p, int gpus, size_t n_per, int steps, int chunks, const std::string &transport, bool json) {
	const size_t n = n_per * (size_t)gpus;
	const int32_t x0 = (1 << (p->iw - 1)) - 1;
	const uint32_t mask = p->pw >= 32 ? 0xFFFFFFFFu : ((1u << p->pw) - 1u);
	std::vector<int> devices(gpus);
	for (int g = 0; g < gpus; g++) devices[g] = g;
	CK(cudaSetDevice(0));
	uint32_t *d_phase = nullptr;
	int32_t *d_xy = nullptr, *d_ref = nullptr;
	CK(cudaMalloc(&d_phase, n * 4));
	CK(cudaMalloc(&d_xy, n * 8));
	CK(cudaMalloc(&d_ref, n * 8));
	{
		const size_t slice = (size_t)1 << 26;
		std::vector<uint32_t> h(slice);
		for (size_t s0 = 0; s0 < n; s0 += slice) {
			const size_t cnt = n - s0 < slice ? n - s0 : slice;
			for (size_t i = 0; i < cnt; i++) h[i] = (uint32_t)(s0 + i) & mask;
			CK(cudaMemcpy(d_phase + s0, h.data(), cnt * 4, cudaMemcpyHostToDevice));
		}
	}
	ZC(zc_rotate_const(p, x0, 0, d_phase, d_ref, n, 0, nullptr));		// device 0 alone: the reference result
	CK(cudaDeviceSynchronize());
	const size_t max_piece = ((n + chunks - 1) / chunks / gpus + 256) & ~(size_t)127;
	double best = 0;
	std::string best_name, detail;
	bool parity = true;
	for (int tr = 0; tr < 3; tr++) {
		const char *name = tr == ZC_XCHG_NCCL ? "nccl" : tr == ZC_XCHG_PEER ? "peer" : "copy";
		if (transport != "both" && transport != name) continue;
		zc_exchange *x = nullptr;
		int rc = zc_exchange_create(devices.data(), gpus, tr, max_piece, &x);
		if (rc != ZC_OK) { std::fprintf(stderr, "zc_exchange_create(%s): %s\n", name, zc_strerror(rc)); return 2; }
		CK(cudaSetDevice(0));
		CK(cudaMemset(d_xy, 0xff, n * 8));
		ZC(zc_scatter_rotate_gather(x, p, x0, 0, d_phase, d_xy, n, chunks));	// warm-up: tables on every device, NCCL channels
		const auto t0 = std::chrono::steady_clock::now();
		for (int s = 0; s < steps; s++) ZC(zc_scatter_rotate_gather(x, p, x0, 0, d_phase, d_xy, n, chunks));	// returns complete
		const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		zc_exchange_destroy(x);
		CK(cudaSetDevice(0));
		// byte-for-byte against device 0 alone
		const size_t slice = (size_t)1 << 25;
		std::vector<int32_t> a(slice * 2), b(slice * 2);
		bool same = true;
		for (size_t s0 = 0; s0 < n && same; s0 += slice) {
			const size_t cnt = n - s0 < slice ? n - s0 : slice;
			CK(cudaMemcpy(a.data(), d_xy + 2 * s0, cnt * 8, cudaMemcpyDeviceToHost));
			CK(cudaMemcpy(b.data(), d_ref + 2 * s0, cnt * 8, cudaMemcpyDeviceToHost));
			same = std::memcmp(a.data(), b.data(), cnt * 8) == 0;
		}
		parity = parity && same;
		const double gs = (double)n * steps / dt / 1e9;
		// what crosses device 0's NVLink port per sample: 4 B out + 8 B back for the (G-1)/G of the stream that leaves
		const double link = gs * 8.0 * (gpus - 1) / gpus;
		char buf[256];
		std::snprintf(buf, sizeof(buf), "%s\"%s\": {\"value\": %.2f, \"ms_per_step\": %.3f, \"dev0_ingress_gbs\": %.1f, \"parity\": %s}",
			detail.empty() ? "" : ", ", name, gs, 1e3 * dt / steps, link, same ? "true" : "false");
		detail += buf;
		if (!json) std::printf("%d GPUs, scatter->rotate->gather over %s: %.1f Gsamples/s (%.1f GB/s into device 0), output %s device 0 alone\n",
			gpus, name, gs, link, same ? "==" : "!=");
		if (gs > best) { best = gs; best_name = name; }
	}
	if (json)
		std::printf("{\"value\": %.2f, \"unit\": \"Gsamples/s\", \"transport\": \"%s\", \"n_gpus\": %d, \"samples_per_gpu_per_step\": %zu, "
			"\"steps\": %d, \"chunks\": %d, \"parity\": %s, %s, \"what\": \"device 0 owns the whole phase stream; per step it is scattered over "
			"the devices, rotated and gathered back (libzcordic_nccl: ncclSend/ncclRecv groups, 3-stage pipeline | kernels on peer "
			"memory over NVLink | the same pipeline with copy-engine transfers); host wall clock around complete calls; parity = output byte-identical to device 0 alone\"}\n",
			best, best_name.c_str(), gpus, n_per, steps, chunks, parity ? "true" : "false", detail.c_str());
	cudaFree(d_phase); cudaFree(d_xy); cudaFree(d_ref);
	return parity ? 0 : 4;
}

int main(int argc, char **argv) {
	int iw = 18, ow = 18, pw = 24, ns = 20, xtra = 2, lg = 28, steps = 20, device = 0, gpus = 1, chunks = 8;
	bool scatter = false, json = false, probe = false;
	std::string transport = "both";
	for (int k = 1; k < argc; k++) {
		if (!std::strcmp(argv[k], "--scatter")) { scatter = true; continue; }
		if (!std::strcmp(argv[k], "--json")) { json = true; continue; }
		if (!std::strcmp(argv[k], "--pcie-probe")) { probe = true; continue; }
		if (k + 1 >= argc) { std::fprintf(stderr, "option %s needs a value\n", argv[k]); return 1; }
		if (!std::strcmp(argv[k], "--transport")) { transport = argv[++k]; continue; }
		const int v = std::atoi(argv[k + 1]);
		k++;
		if (!std::strcmp(argv[k - 1], "--chunks")) chunks = v;
		else if (!std::strcmp(argv[k - 1], "-i")) iw = v;
		else if (!std::strcmp(argv[k - 1], "-o")) ow = v;
		else if (!std::strcmp(argv[k - 1], "-p")) pw = v;
		else if (!std::strcmp(argv[k - 1], "-n")) ns = v;
		else if (!std::strcmp(argv[k - 1], "-x")) xtra = v;
		else if (!std::strcmp(argv[k - 1], "-l")) lg = v;
		else if (!std::strcmp(argv[k - 1], "-s")) steps = v;
		else if (!std::strcmp(argv[k - 1], "-d")) device = v;
		else if (!std::strcmp(argv[k - 1], "-g")) gpus = v;
		else { std::fprintf(stderr, "unknown option %s\n", argv[k - 1]); return 1; }
	}
	if (zc_device_count() <= 0) {
		std::fprintf(stderr, "no CUDA device: %s (libzcordic has no CPU path)\n", zc_last_error());
		return 3;
	}
	zc_params p;
	ZC(zc_derive_p2r(iw, ow, xtra, pw, ns, &p));		// sw/main.cpp:260-279
	if (!json) std::printf("core: IW=%d OW=%d WW=%d PW=%d NSTAGES=%d GAIN=%.12f\n", p.iw, p.ow, p.ww, p.pw, p.nstages, p.gain);
	const size_t n = (size_t)1 << lg;
	if (probe) {
		if (gpus > zc_device_count()) { std::fprintf(stderr, "-g %d but %d devices\n", gpus, zc_device_count()); return 1; }
		return pcie_probe(gpus, n, json);
	}
	if (scatter) {
		if (gpus > zc_device_count()) { std::fprintf(stderr, "-g %d but %d devices\n", gpus, zc_device_count()); return 1; }
		return scatter_bench(&p, gpus, n, steps, chunks < 1 ? 1 : chunks, transport, json);
	}
	const int32_t x0 = (1 << (p.iw - 1)) - 1, y0 = 0;	// cordic_tb.cpp:68-69
	if (gpus > 1) {		// independent shards of n samples each, one host thread per device
		if (gpus > zc_device_count()) { std::fprintf(stderr, "-g %d but %d devices\n", gpus, zc_device_count()); return 1; }
		std::vector<float> ms(gpus, 0.f);
		std::vector<std::thread> th;
		std::atomic<int> ready{0};
		for (int g = 0; g < gpus; g++)
			th.emplace_back([&, g] { ms[g] = shard(&p, g, n, (size_t)g * n, steps, &ready, gpus); });
		for (auto &t : th) t.join();
		float worst = 0;
		for (int g = 0; g < gpus; g++) {
			if (ms[g] <= 0) { std::fprintf(stderr, "device %d failed: %s\n", g, zc_last_error()); return 2; }
			if (ms[g] > worst) worst = ms[g];
		}
		std::printf("%d GPUs, device buffers: %.1f Gsamples/s  (%zu samples per GPU x %d steps, slowest device %.3f ms per step)\n",
			gpus, (double)n * gpus * steps / (worst * 1e-3) / 1e9, n, steps, worst / steps);
		return 0;
	}
	CK(cudaSetDevice(device));

	// host buffers (pinned) holding the sweep i -> i mod 2^PW, and the device copies
	uint32_t *h_phase = static_cast<uint32_t *>(zc_host_alloc(n * 4));
	int32_t *h_xy = static_cast<int32_t *>(zc_host_alloc(n * 8));
	if (!h_phase || !h_xy) { std::fprintf(stderr, "zc_host_alloc: %s\n", zc_last_error()); return 2; }
	const uint32_t mask = p.pw >= 32 ? 0xFFFFFFFFu : ((1u << p.pw) - 1u);
	for (size_t i = 0; i < n; i++) h_phase[i] = (uint32_t)i & mask;
	uint32_t *d_phase = nullptr;
	int32_t *d_xy = nullptr;
	CK(cudaMalloc(&d_phase, n * 4));
	CK(cudaMalloc(&d_xy, n * 8));
	CK(cudaMemcpy(d_phase, h_phase, n * 4, cudaMemcpyHostToDevice));
	cudaStream_t st;
	CK(cudaStreamCreate(&st));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));

	// device-resident
	for (int w = 0; w < 3; w++) ZC(zc_rotate_const(&p, x0, y0, d_phase, d_xy, n, device, st));
	CK(cudaStreamSynchronize(st));
	const uint64_t l0 = zc_launch_count();
	CK(cudaEventRecord(e0, st));
	for (int s = 0; s < steps; s++) ZC(zc_rotate_const(&p, x0, y0, d_phase, d_xy, n, device, st));
	CK(cudaEventRecord(e1, st));
	CK(cudaEventSynchronize(e1));
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, e0, e1));
	std::printf("device buffers: %.1f Gsamples/s  (%zu samples x %d steps, %.3f ms per step, %llu kernel launches)\n",
		(double)n * steps / (ms * 1e-3) / 1e9, n, steps, ms / steps, (unsigned long long)(zc_launch_count() - l0));

	// end to end from host memory
	ZC(zc_rotate_const_host(&p, x0, y0, h_phase, h_xy, n, device));
	CK(cudaEventRecord(e0, st));
	ZC(zc_rotate_const_host(&p, x0, y0, h_phase, h_xy, n, device));
	CK(cudaEventRecord(e1, st));
	CK(cudaEventSynchronize(e1));
	CK(cudaEventElapsedTime(&ms, e0, e1));
	std::printf("host buffers  : %.2f Gsamples/s  (H2D %zu MB + D2H %zu MB per call)\n", (double)n / (ms * 1e-3) / 1e9, n * 4 >> 20, n * 8 >> 20);

	// the device path and the host path must agree, and phase 0 must give the test bench's first sample
	std::vector<int32_t> head(16);
	CK(cudaMemcpy(head.data(), d_xy, head.size() * 4, cudaMemcpyDeviceToHost));
	if (std::memcmp(head.data(), h_xy, head.size() * 4) != 0) { std::fprintf(stderr, "device and host paths disagree\n"); return 4; }
	std::printf("o_xval,o_yval at phase 0..3: (%d,%d) (%d,%d) (%d,%d) (%d,%d)\n", h_xy[0], h_xy[1], h_xy[2], h_xy[3], h_xy[4], h_xy[5], h_xy[6], h_xy[7]);

	zc_host_free(h_phase);
	zc_host_free(h_xy);
	cudaFree(d_phase);
	cudaFree(d_xy);
	zc_trim(device);
	return 0;
}
