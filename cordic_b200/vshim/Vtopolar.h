// cordic_b200/vshim/Vtopolar.h -- the class Verilator would generate for rtl/topolar.v (ports :59-64), backed
// by the GPU engine (zc_topolar).  Drop-in for bench/cpp/topolar_tb.cpp via TESTB<Vtopolar>.
#ifndef ZC_VSHIM_VTOPOLAR_H
#define ZC_VSHIM_VTOPOLAR_H
#include "verilated.h"
#include "verilated_vcd_c.h"
#include "topolar.h"		// generated constants (rtl/topolar.h:46-58)
#include "zc_deferred.h"

class Vtopolar : public zc_vshim::Deferred<Vtopolar> {
	zc_params m_p;
	uint32_t m_lastclk = 0;
public:
	uint32_t i_clk = 0, i_reset = 0, i_ce = 0, i_xval = 0, i_yval = 0, i_aux = 0;
	uint32_t o_mag = 0, o_phase = 0, o_aux = 0;

	Vtopolar() {
		int rc = zc_derive_r2p(IW, OW, NEXTRA - 2, PW, NSTAGES, &m_p);
		if (rc != ZC_OK || m_p.ww != WW) zc_vshim::die("zc_derive_r2p", rc);
		setup(NSTAGES + 1, NSTAGES + 4);
	}
	~Vtopolar() { teardown(); }
	void trace(VerilatedVcdC *, int) {}
	void eval() {
		const bool rising = (i_clk & 1) && !(m_lastclk & 1);
		m_lastclk = i_clk;
		if (!rising) return;
		if (i_reset & 1) { reset_pipe(); o_mag = o_phase = o_aux = 0; return; }
		if (!(i_ce & 1)) return;
		const zc_vshim::Slot out = clock(zc_vshim::Slot{i_xval, i_yval, 0, i_aux & 1, false, false, 0, 0});
		const uint32_t omask = (OW >= 32) ? 0xffffffffu : ((1u << OW) - 1u);
		o_mag = out.r0 & omask; o_phase = out.r1; o_aux = out.aux;
	}
	// Staging layout: in = [xy pairs: 2*cap], out = [mag: cap][phase: cap]
	static int lanes_in() { return 2; }
	static int lanes_out() { return 2; }
	void pack(const zc_vshim::Slot &s, uint32_t *hin, size_t k, size_t) { hin[2 * k] = s.a; hin[2 * k + 1] = s.b; }
	int run(const uint32_t *din, uint32_t *dout, size_t n, size_t cap, cudaStream_t st) {
		return zc_topolar(&m_p, (const int32_t *)din, (int32_t *)dout, dout + cap, n, m_device, st);
	}
	void unpack(zc_vshim::Slot &s, const uint32_t *hout, size_t k, size_t cap) { s.r0 = hout[k]; s.r1 = hout[cap + k]; }
};
#endif
