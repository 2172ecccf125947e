// cordic_b200/vshim/zc_deferred.h -- a clocked, port-level facade over the batched engine.
//
// The generated cores are feed-forward pipelines: with i_ce high, the outputs after clock t belong to the
// inputs sampled at clock t-LAT, LAT = NSTAGES+1 register stages after the input register
// (rtl/cordic.v:118-190 stage 0, :232-283 stages 1..N, :303-314 output; aux chain :100-105,313).  This
// class keeps the inputs of the last LAT clocks in a FIFO and, when the test bench clocks out a sample whose
// result is not on the host yet, sends the whole backlog through ONE zc_* call on the GPU.  Reset empties the
// pipeline (every register is zeroed, rtl/cordic.v:119-124), which reads back as zeros for LAT clocks.
#ifndef ZC_VSHIM_DEFERRED_H
#define ZC_VSHIM_DEFERRED_H

#include "zcordic.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <deque>
#include <vector>

namespace zc_vshim {

inline void die(const char *what, int rc) {
	fprintf(stderr, "zcordic vshim: %s failed (%d): %s\n", what, rc, zc_last_error());
	exit(EXIT_FAILURE);
}
inline void cuda_ok(cudaError_t e, const char *what) {
	if (e != cudaSuccess) {
		fprintf(stderr, "zcordic vshim: %s: %s\n", what, cudaGetErrorString(e));
		exit(EXIT_FAILURE);
	}
}

// Totals over the process, reported at exit when ZC_VSHIM_STATS is set (the test benches leave through
// exit() without destroying the model, bench/cpp/cordic_tb.cpp:376-381).
struct Stats {
	unsigned long long samples = 0, batches = 0;
	static Stats &get() { static Stats s; return s; }
	static void report() {
		fprintf(stderr, "zcordic vshim: %llu samples in %llu GPU batches (%llu kernel launches)\n",
			get().samples, get().batches, (unsigned long long)zc_launch_count());
	}
};

struct Slot {
	uint32_t a, b, c;	// raw port words of one clock (meaning is the core's)
	uint32_t aux;
	bool bubble;		// a register stage zeroed by reset: reads back as 0
	bool done;
	uint32_t r0, r1;	// results
};

// CORE provides: static int lanes_in(), lanes_out();  void pack(const Slot&, uint32_t *hin, size_t k, size_t cap);
//   int run(const uint32_t *din, uint32_t *dout, size_t n, size_t cap, cudaStream_t);
//   void unpack(Slot&, const uint32_t *hout, size_t k, size_t cap);
// The staging buffers hold `cap` samples per lane, lane-major, and are copied whole (a few hundred bytes).
template <class CORE>
class Deferred {
protected:
	std::deque<Slot> m_fifo;
	int m_lat = 0;
	int m_device = 0;
	cudaStream_t m_stream = nullptr;
	uint32_t *m_hin = nullptr, *m_hout = nullptr, *m_din = nullptr, *m_dout = nullptr;
	size_t m_cap = 0;
	uint64_t m_batches = 0, m_samples = 0;

	void setup(int latency, size_t cap) {
		m_lat = latency; m_cap = cap;
		const char *dev = getenv("ZC_VSHIM_DEVICE");
		m_device = dev ? atoi(dev) : 0;
		cuda_ok(cudaSetDevice(m_device), "cudaSetDevice");
		cuda_ok(cudaStreamCreateWithFlags(&m_stream, cudaStreamNonBlocking), "cudaStreamCreate");
		const size_t wi = (size_t)CORE::lanes_in() * cap * 4, wo = (size_t)CORE::lanes_out() * cap * 4;
		m_hin = (uint32_t *)zc_host_alloc(wi); m_hout = (uint32_t *)zc_host_alloc(wo);
		if (!m_hin || !m_hout) die("zc_host_alloc", ZC_ENOMEM);
		cuda_ok(cudaMalloc((void **)&m_din, wi), "cudaMalloc");
		cuda_ok(cudaMalloc((void **)&m_dout, wo), "cudaMalloc");
		reset_pipe();
		static bool registered = false;
		if (!registered && getenv("ZC_VSHIM_STATS")) { atexit(Stats::report); registered = true; }
	}
	void teardown() {
		if (m_din) cudaFree(m_din);
		if (m_dout) cudaFree(m_dout);
		if (m_hin) zc_host_free(m_hin);
		if (m_hout) zc_host_free(m_hout);
		if (m_stream) cudaStreamDestroy(m_stream);
	}
	void reset_pipe() {
		m_fifo.clear();
		for (int k = 0; k < m_lat; k++) m_fifo.push_back(Slot{0, 0, 0, 0, true, true, 0, 0});
	}
	// One rising clock edge with i_ce: shift `in` into the pipe, return what falls out of the far end.
	Slot clock(const Slot &in) {
		m_fifo.push_back(in);
		Slot &head = m_fifo.front();
		if (!head.done) flush();
		Slot out = m_fifo.front();
		m_fifo.pop_front();
		return out;
	}
	void flush() {
		CORE *self = static_cast<CORE *>(this);
		size_t n = 0;
		const int li = CORE::lanes_in(), lo = CORE::lanes_out();
		for (Slot &s : m_fifo) {
			if (s.done) continue;
			if (n == m_cap) break;
			self->pack(s, m_hin, n, m_cap);
			n++;
		}
		if (!n) return;
		cuda_ok(cudaMemcpyAsync(m_din, m_hin, (size_t)li * m_cap * 4, cudaMemcpyHostToDevice, m_stream), "H2D");
		int rc = self->run(m_din, m_dout, n, m_cap, m_stream);
		if (rc != ZC_OK) die("engine call", rc);
		cuda_ok(cudaMemcpyAsync(m_hout, m_dout, (size_t)lo * m_cap * 4, cudaMemcpyDeviceToHost, m_stream), "D2H");
		cuda_ok(cudaStreamSynchronize(m_stream), "sync");
		size_t k = 0;
		for (Slot &s : m_fifo) {
			if (s.done) continue;
			if (k == n) break;
			self->unpack(s, m_hout, k, m_cap);
			s.done = true;
			k++;
		}
		m_batches++; m_samples += n;
		Stats::get().batches++; Stats::get().samples += n;
	}
};

} // namespace zc_vshim
#endif
