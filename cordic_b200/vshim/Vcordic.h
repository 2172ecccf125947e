// cordic_b200/vshim/Vcordic.h -- the class Verilator would generate for rtl/cordic.v (ports :58-63), backed by
// the GPU engine (zc_rotate).  Drop-in for bench/cpp/cordic_tb.cpp via TESTB<Vcordic> (testb.h:49-136).
#ifndef ZC_VSHIM_VCORDIC_H
#define ZC_VSHIM_VCORDIC_H
#include "verilated.h"
#include "verilated_vcd_c.h"
#include "cordic.h"		// IW OW NEXTRA WW PW NSTAGES of the configuration under test (rtl/cordic.h:46-59)
#include "zc_deferred.h"

class Vcordic : public zc_vshim::Deferred<Vcordic> {
	zc_params m_p;
	uint32_t m_lastclk = 0;
public:
	uint32_t i_clk = 0, i_reset = 0, i_ce = 0, i_xval = 0, i_yval = 0, i_phase = 0, i_aux = 0;
	uint32_t o_xval = 0, o_yval = 0, o_aux = 0;

	Vcordic() {
		int rc = zc_derive_p2r(IW, OW, NEXTRA - 1, PW, NSTAGES, &m_p);
		if (rc != ZC_OK || m_p.ww != WW) zc_vshim::die("zc_derive_p2r", rc);
		setup(NSTAGES + 1, NSTAGES + 4);
	}
	~Vcordic() { teardown(); }
	void trace(VerilatedVcdC *, int) {}
	void eval() {
		const bool rising = (i_clk & 1) && !(m_lastclk & 1);
		m_lastclk = i_clk;
		if (!rising) return;
		if (i_reset & 1) { reset_pipe(); o_xval = o_yval = o_aux = 0; return; }
		if (!(i_ce & 1)) return;
		const zc_vshim::Slot out = clock(zc_vshim::Slot{i_xval, i_yval, i_phase, i_aux & 1, false, false, 0, 0});
		const uint32_t omask = (OW >= 32) ? 0xffffffffu : ((1u << OW) - 1u);
		o_xval = out.r0 & omask; o_yval = out.r1 & omask; o_aux = out.aux;
	}
	// --- Deferred<> hooks.  Staging layout: in = [xy pairs: 2*cap][phase: cap], out = [xy pairs: 2*cap]
	static int lanes_in() { return 3; }
	static int lanes_out() { return 2; }
	void pack(const zc_vshim::Slot &s, uint32_t *hin, size_t k, size_t cap) {
		hin[2 * k] = s.a; hin[2 * k + 1] = s.b; hin[2 * cap + k] = s.c;
	}
	int run(const uint32_t *din, uint32_t *dout, size_t n, size_t cap, cudaStream_t st) {
		return zc_rotate(&m_p, (const int32_t *)din, din + 2 * cap, (int32_t *)dout, n, m_device, st);
	}
	void unpack(zc_vshim::Slot &s, const uint32_t *hout, size_t k, size_t) { s.r0 = hout[2 * k]; s.r1 = hout[2 * k + 1]; }
};
#endif
