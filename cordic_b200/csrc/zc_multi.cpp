// zc_multi.cpp -- the host-buffer entry points over SEVERAL devices of one box, and the pinned host memory they want.
//
// One process, one host thread per device: the sample stream is cut into contiguous shards (independent units, no
// exchange -- SURVEY.md §8e), shard g runs through the single-device H2D -> kernel -> D2H pipeline of zc_*_host on
// devices[g], and the call returns when every shard's outputs are in host memory.  Concatenated multi-device output is
// byte-identical to the single-device output (tests/test_gpu_multi.py).
//
// On a two-socket box the copies only reach PCIe speed when a shard's host pages sit on the NUMA node its GPU hangs
// off: zc_host_alloc_sharded() places the g-th 1/ndev of a pinned buffer on the node of devices[g] (mbind(2) before
// first touch, then cudaHostRegister).  Plain zc_host_alloc() memory works too, at whatever the interconnect gives.
#include "zc_internal.h"

#include <cuda_runtime.h>

#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace zc {

// ---- pinned host memory ---------------------------------------------------------------------------------------
struct HostBlock { size_t bytes; bool mapped; };	// mapped: our own mmap + cudaHostRegister; else cudaHostAlloc
static std::mutex g_host_mu;
static std::map<void *, HostBlock> g_host_blocks;

static int device_numa_node(int device) {
	char id[32] = "";
	if (cudaDeviceGetPCIBusId(id, sizeof(id), device) != cudaSuccess) { cudaGetLastError(); return -1; }
	for (char *c = id; *c; c++) *c = (char)std::tolower((unsigned char)*c);
	const std::string path = std::string("/sys/bus/pci/devices/") + id + "/numa_node";
	FILE *fp = std::fopen(path.c_str(), "r");
	if (!fp) return -1;
	int node = -1;
	if (std::fscanf(fp, "%d", &node) != 1) node = -1;
	std::fclose(fp);
	return node;
}

static void *host_alloc_plain(size_t bytes) {
	void *p = nullptr;
	cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
	if (e != cudaSuccess) {
		cudaGetLastError();
		set_error(ZC_ENOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
		return nullptr;
	}
	std::lock_guard<std::mutex> lk(g_host_mu);
	g_host_blocks[p] = HostBlock{bytes, false};
	return p;
}

static void *host_alloc_sharded(size_t bytes, const int *devices, int ndev) {
	if (ndev < 1 || !devices) { set_error(ZC_EINVAL, "zc_host_alloc_sharded needs at least one device"); return nullptr; }
	const size_t page = (size_t)sysconf(_SC_PAGESIZE);
	const size_t len = ((bytes ? bytes : 1) + page - 1) / page * page;
	void *base = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	if (base == MAP_FAILED) { set_error(ZC_ENOMEM, "mmap(%zu) failed", len); return nullptr; }
#ifdef MADV_HUGEPAGE
	madvise(base, len, MADV_HUGEPAGE);		// fewer faults, fewer pinned-page descriptors; best effort
#endif
	// shard g = [g*bytes/ndev, (g+1)*bytes/ndev), widened to page boundaries; bind it, then touch it from its own thread
	std::vector<std::thread> th;
	for (int g = 0; g < ndev; g++) {
		size_t lo = bytes / (size_t)ndev * (size_t)g, hi = (g + 1 == ndev) ? len : bytes / (size_t)ndev * (size_t)(g + 1);
		lo = lo / page * page; hi = (hi + page - 1) / page * page;
		if (hi > len) hi = len;
		const int node = device_numa_node(devices[g]);
		char *a = static_cast<char *>(base) + lo;
		const size_t l = hi - lo;
		th.emplace_back([a, l, node, page] {
			if (node >= 0 && node < 1024) {
				unsigned long mask[16] = {};
				mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
				syscall(SYS_mbind, a, l, 2 /* MPOL_BIND */, mask, (unsigned long)(8 * sizeof(mask)), 0u);	// best effort
			}
			for (size_t off = 0; off < l; off += page) a[off] = 0;
		});
	}
	for (auto &t : th) t.join();
	cudaError_t e = cudaHostRegister(base, len, cudaHostRegisterPortable);
	if (e != cudaSuccess) {
		cudaGetLastError();
		munmap(base, len);
		set_error(ZC_ENOMEM, "cudaHostRegister(%zu): %s", len, cudaGetErrorString(e));
		return nullptr;
	}
	std::lock_guard<std::mutex> lk(g_host_mu);
	g_host_blocks[base] = HostBlock{len, true};
	return base;
}

static void host_free(void *ptr) {
	if (!ptr) return;
	HostBlock b{0, false};
	{
		std::lock_guard<std::mutex> lk(g_host_mu);
		auto it = g_host_blocks.find(ptr);
		if (it != g_host_blocks.end()) { b = it->second; g_host_blocks.erase(it); }
	}
	if (b.mapped) {
		cudaHostUnregister(ptr);
		munmap(ptr, b.bytes);
	} else {
		cudaFreeHost(ptr);
	}
}

// ---- sharding ---------------------------------------------------------------------------------------------------
// Shard g of G over n samples: [first(g), first(g+1)), boundaries multiples of 4 samples so that every shard keeps the
// 16-byte alignment of the caller's buffers (the vector kernels want it).
static inline size_t shard_first(size_t n, int g, int G) {
	if (g >= G) return n;
	return (size_t)((unsigned __int128)n * (unsigned)g / (unsigned)G) & ~(size_t)3;
}

template <class Call>
static int run_sharded(size_t n, const int *devices, int ndev, Call call) {
	if (ndev < 1 || !devices) return set_error(ZC_EINVAL, "need at least one device");
	if (ndev > 64) return set_error(ZC_EINVAL, "at most 64 devices");
	for (int g = 0; g < ndev; g++)
		for (int h = 0; h < g; h++)
			if (devices[g] == devices[h]) return set_error(ZC_EINVAL, "device %d listed twice", devices[g]);
	std::vector<int> rc(ndev, ZC_OK);
	std::vector<std::string> msg(ndev);
	std::vector<std::thread> th;
	for (int g = 0; g < ndev; g++) {
		const size_t first = shard_first(n, g, ndev), count = shard_first(n, g + 1, ndev) - first;
		th.emplace_back([&, g, first, count] {
			rc[g] = count ? call(devices[g], first, count) : ZC_OK;
			if (rc[g] != ZC_OK) msg[g] = zc_last_error();	// thread-local: carry it back to the caller's thread
		});
	}
	for (auto &t : th) t.join();
	for (int g = 0; g < ndev; g++)
		if (rc[g] != ZC_OK) return set_error(rc[g], "device %d: %s", devices[g], msg[g].c_str());
	return ZC_OK;
}

} // namespace zc

using namespace zc;

extern "C" {

void *zc_host_alloc(size_t bytes) { return host_alloc_plain(bytes); }
void *zc_host_alloc_sharded(size_t bytes, const int *devices, int ndev) { return host_alloc_sharded(bytes, devices, ndev); }
void zc_host_free(void *ptr) { host_free(ptr); }
int zc_device_numa_node(int device) { return device_numa_node(device); }

int zc_shard_range(size_t n, int ndev, int g, size_t *first, size_t *count) {
	if (ndev < 1 || g < 0 || g >= ndev || !first || !count) return set_error(ZC_EINVAL, "bad shard arguments");
	*first = shard_first(n, g, ndev);
	*count = shard_first(n, g + 1, ndev) - *first;
	return ZC_OK;
}

int zc_rotate_const_host_multi(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase, int32_t *xy, size_t n,
		const int *devices, int ndev) {
	if (n && (!phase || !xy)) return set_error(ZC_EINVAL, "NULL buffer");
	return run_sharded(n, devices, ndev, [&](int dev, size_t first, size_t count) {
		return zc_rotate_const_host(p, x0, y0, phase + first, xy + 2 * first, count, dev);
	});
}

int zc_rotate_host_multi(const zc_params *p, const int32_t *xy_in, const uint32_t *phase, int32_t *xy_out, size_t n,
		const int *devices, int ndev) {
	if (n && (!phase || !xy_in || !xy_out)) return set_error(ZC_EINVAL, "NULL buffer");
	return run_sharded(n, devices, ndev, [&](int dev, size_t first, size_t count) {
		return zc_rotate_host(p, xy_in + 2 * first, phase + first, xy_out + 2 * first, count, dev);
	});
}

int zc_topolar_host_multi(const zc_params *p, const int32_t *xy_in, int32_t *mag, uint32_t *phase, size_t n,
		const int *devices, int ndev) {
	if (n && (!xy_in || !mag || !phase)) return set_error(ZC_EINVAL, "NULL buffer");
	return run_sharded(n, devices, ndev, [&](int dev, size_t first, size_t count) {
		return zc_topolar_host(p, xy_in + 2 * first, mag + first, phase + first, count, dev);
	});
}

int zc_nco_rotate_host_multi(const zc_params *p, int32_t x0, int32_t y0, uint32_t phase0, uint32_t step, uint64_t n0,
		int32_t *xy, size_t n, const int *devices, int ndev) {
	if (n && !xy) return set_error(ZC_EINVAL, "NULL buffer");
	return run_sharded(n, devices, ndev, [&](int dev, size_t first, size_t count) {	// each shard starts in closed form
		return zc_nco_rotate_host(p, x0, y0, phase0, step, n0 + first, xy + 2 * first, count, dev);
	});
}

int zc_lut_sin_host_multi(int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32, int32_t *out, size_t n,
		const int *devices, int ndev) {
	if (n && (!phase32 || !out)) return set_error(ZC_EINVAL, "NULL buffer");
	return run_sharded(n, devices, ndev, [&](int dev, size_t first, size_t count) {
		return zc_lut_sin_host(pw, ow, tbl_host, phase32 + first, out + first, count, dev);
	});
}

int zc_lut_qwav_host_multi(int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32, int32_t *out, size_t n,
		const int *devices, int ndev) {
	if (n && (!phase32 || !out)) return set_error(ZC_EINVAL, "NULL buffer");
	return run_sharded(n, devices, ndev, [&](int dev, size_t first, size_t count) {
		return zc_lut_qwav_host(pw, ow, tbl_host, phase32 + first, out + first, count, dev);
	});
}

int zc_quadtbl_sin_host_multi(const zc_quadtbl *q, const uint32_t *phase32, int32_t *out, size_t n, const int *devices, int ndev) {
	if (n && (!phase32 || !out)) return set_error(ZC_EINVAL, "NULL buffer");
	return run_sharded(n, devices, ndev, [&](int dev, size_t first, size_t count) {
		return zc_quadtbl_sin_host(q, phase32 + first, out + first, count, dev);
	});
}

int zc_nco_mix_host_multi(const zc_params *p, const int32_t *xy_in, uint32_t phase0, uint32_t step, uint64_t n0,
		int32_t *xy_out, size_t n, const int *devices, int ndev) {
	if (n && (!xy_in || !xy_out)) return set_error(ZC_EINVAL, "NULL buffer");
	return run_sharded(n, devices, ndev, [&](int dev, size_t first, size_t count) {
		return zc_nco_mix_host(p, xy_in + 2 * first, phase0, step, n0 + first, xy_out + 2 * first, count, dev);
	});
}

int zc_topolar_i16_host_multi(const zc_params *p, const int16_t *xy16_in, int32_t *mag, uint32_t *phase, size_t n,
		const int *devices, int ndev) {
	if (n && (!xy16_in || !mag || !phase)) return set_error(ZC_EINVAL, "NULL buffer");
	return run_sharded(n, devices, ndev, [&](int dev, size_t first, size_t count) {
		return zc_topolar_i16_host(p, xy16_in + 2 * first, mag + first, phase + first, count, dev);
	});
}

int zc_rotate_const_o16_host_multi(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase, int16_t *xy16, size_t n,
		const int *devices, int ndev) {
	if (n && (!phase || !xy16)) return set_error(ZC_EINVAL, "NULL buffer");
	return run_sharded(n, devices, ndev, [&](int dev, size_t first, size_t count) {
		return zc_rotate_const_o16_host(p, x0, y0, phase + first, xy16 + 2 * first, count, dev);
	});
}

} // extern "C"
