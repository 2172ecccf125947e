// cordic_b200/vshim/Vquadtbl.h -- the class Verilator would generate for rtl/quadtbl.v (ports :53-58), backed by
// the GPU engine (zc_quadtbl_sin).  Drop-in for bench/cpp/quadtbl_tb.cpp via TESTB<Vquadtbl>.  The RTL is a
// 6-register pipeline (o_aux = aux[NSTAGES-1], rtl/quadtbl.v:121-127): outputs lag inputs by 5 clocks.
#ifndef ZC_VSHIM_VQUADTBL_H
#define ZC_VSHIM_VQUADTBL_H
#include "verilated.h"
#include "verilated_vcd_c.h"
#include "quadtbl.h"		// OW NEXTRA PW TBL_LGSZ TBL_ERR (rtl/quadtbl.h)
#include "zc_deferred.h"

class Vquadtbl : public zc_vshim::Deferred<Vquadtbl> {
	zc_quadtbl *m_q;
	uint32_t m_lastclk = 0;
public:
	uint32_t i_clk = 0, i_reset = 0, i_ce = 0, i_aux = 0, i_phase = 0, o_sin = 0, o_aux = 0;

	Vquadtbl() {
		m_q = new zc_quadtbl;
		int rc = zc_derive_qtbl(0, OW, NEXTRA - 1, PW, m_q);
		if (rc != ZC_OK || m_q->lgtbl != TBL_LGSZ) zc_vshim::die("zc_derive_qtbl", rc);
		setup(5, 256);
	}
	~Vquadtbl() { teardown(); delete m_q; }
	void trace(VerilatedVcdC *, int) {}
	void eval() {
		const bool rising = (i_clk & 1) && !(m_lastclk & 1);
		m_lastclk = i_clk;
		if (!rising) return;
		if (i_reset & 1) { reset_pipe(); o_sin = o_aux = 0; return; }
		if (!(i_ce & 1)) return;
		const zc_vshim::Slot out = clock(zc_vshim::Slot{i_phase, 0, 0, i_aux & 1, false, false, 0, 0});
		o_sin = out.r0 & ((1u << OW) - 1u); o_aux = out.aux;
	}
	static int lanes_in() { return 1; }
	static int lanes_out() { return 1; }
	// the ABI takes 32-bit NCO words; the port word is its top PW bits
	void pack(const zc_vshim::Slot &s, uint32_t *hin, size_t k, size_t) { hin[k] = (PW >= 32) ? s.a : (s.a << (32 - PW)); }
	int run(const uint32_t *din, uint32_t *dout, size_t n, size_t, cudaStream_t st) {
		return zc_quadtbl_sin(m_q, din, (int32_t *)dout, n, m_device, st);
	}
	void unpack(zc_vshim::Slot &s, const uint32_t *hout, size_t k, size_t) { s.r0 = hout[k]; }
};
#endif
