// oracle/shim/Vseqcordic.h -- TEST INFRASTRUCTURE ONLY.
// Clock-by-clock stand-in for the Verilated model of rtl/seqcordic.v: the port fields and eval()/trace()
// surface bench/cpp/testb.h drives, with the state machine's registers restated one by one --
// prex/prey/preph (rtl/seqcordic.v:128-186), idle (:225-233), pre_valid (:236-243), cangle (:246-249),
// state (:252-263), xv/yv/ph (:266-299), o_done (:313-318), o_xval/o_yval/o_aux (:320-329), aux (:111-117).
// It does NOT call the oracle's batched seq function: the reference's unmodified cordic_tb.cpp
// (-DCLOCKS_PER_OUTPUT) over this model and zo_rotate1() with sequential=1 are two independent restatements,
// compared in tests/test_oracle_golden.py.
#ifndef ZC_SHIM_VSEQCORDIC_H
#define ZC_SHIM_VSEQCORDIC_H

#include "verilated.h"
#include "verilated_vcd_c.h"
#include "seqcordic.h"	// generated constants: IW OW NEXTRA WW PW NSTAGES CLOCKS_PER_OUTPUT
#include "zc_oracle.h"

class Vseqcordic {
	zo_params m_p;
	uint32_t m_prex, m_prey, m_preph, m_xv, m_yv, m_ph, m_cangle, m_state, m_idle, m_pre_valid, m_aux;
	uint32_t m_lastclk;
public:
	uint32_t i_clk, i_reset, i_stb, i_xval, i_yval, i_phase, i_aux;
	uint32_t o_busy, o_done, o_xval, o_yval, o_aux;

	Vseqcordic() {
		int rc = zo_derive_sp2r(IW, OW, NEXTRA - 1, PW, NSTAGES, &m_p);
		assert(rc == 0 && m_p.ww == WW && zo_clocks_per_output(&m_p) == CLOCKS_PER_OUTPUT);
		(void)rc;
		m_prex = m_prey = m_preph = m_xv = m_yv = m_ph = m_cangle = m_state = 0;
		m_idle = 1; m_pre_valid = 0; m_aux = 0; m_lastclk = 0;	// the `initial` values
		i_clk = i_reset = i_stb = i_xval = i_yval = i_phase = i_aux = 0;
		o_busy = o_done = o_xval = o_yval = o_aux = 0;
	}
	void trace(VerilatedVcdC *t, int) {
		t->declare("i_clk", 1, &i_clk);     t->declare("i_reset", 1, &i_reset);
		t->declare("i_stb", 1, &i_stb);     t->declare("i_xval", IW, &i_xval);
		t->declare("i_yval", IW, &i_yval);  t->declare("i_phase", PW, &i_phase);
		t->declare("i_aux", 1, &i_aux);     t->declare("o_busy", 1, &o_busy);
		t->declare("o_done", 1, &o_done);   t->declare("o_xval", OW, &o_xval);
		t->declare("o_yval", OW, &o_yval);  t->declare("o_aux", 1, &o_aux);
		t->declare("state", 6, &m_state);   t->declare("xv", WW, &m_xv);
		t->declare("yv", WW, &m_yv);        t->declare("ph", PW, &m_ph);
	}
	void eval() {
		bool rising = (i_clk & 1) && !(m_lastclk & 1);
		m_lastclk = i_clk;
		if (!rising) { o_busy = !m_idle; return; }
		const uint32_t wmask = (WW >= 32) ? 0xffffffffu : ((1u << WW) - 1u);
		const uint32_t pmask = (PW >= 32) ? 0xffffffffu : ((1u << PW) - 1u);
		const uint32_t omask = (OW >= 32) ? 0xffffffffu : ((1u << OW) - 1u);
		const bool rst = (i_reset & 1), stb = (i_stb & 1);
		const bool last = (m_state >= (uint32_t)(NSTAGES - 1));
		// every right-hand side below reads the registers as they were before the edge
		uint32_t n_aux = rst ? 0 : ((stb && m_idle) ? (i_aux & 1) : m_aux);
		int32_t px, py; uint32_t pph;
		zo_rotate_pre(&m_p, (int32_t)i_xval, (int32_t)i_yval, i_phase, &px, &py, &pph);
		uint32_t n_idle = rst ? 1 : (stb ? 0 : (m_state == (uint32_t)(NSTAGES - 1) ? 1 : m_idle));
		uint32_t n_pre_valid = rst ? 0 : (stb && m_idle);
		uint32_t n_cangle = (m_state < (uint32_t)NSTAGES && m_state < ZO_MAX_STAGES) ? m_p.angle[m_state] : 0;
		uint32_t n_state = (rst || m_idle || m_state == (uint32_t)(NSTAGES - 1)) ? 0 : m_state + 1;
		uint32_t n_xv, n_yv, n_ph;
		if (m_pre_valid) {
			n_xv = m_prex; n_yv = m_prey; n_ph = m_preph;
		} else {
			const int sh = (m_state > 31) ? 31 : (int)m_state;
			const int32_t x = sext(m_xv), y = sext(m_yv);
			if ((m_ph >> (PW - 1)) & 1) {
				n_xv = (uint32_t)(x + (y >> sh)); n_yv = (uint32_t)(y - (x >> sh)); n_ph = m_ph + m_cangle;
			} else {
				n_xv = (uint32_t)(x - (y >> sh)); n_yv = (uint32_t)(y + (x >> sh)); n_ph = m_ph - m_cangle;
			}
		}
		uint32_t n_done = rst ? 0 : (last ? 1 : 0);
		if (last) {
			o_xval = (uint32_t)zo_round_out(&m_p, sext(m_xv)) & omask;
			o_yval = (uint32_t)zo_round_out(&m_p, sext(m_yv)) & omask;
			o_aux = m_aux;
		}
		m_aux = n_aux; m_prex = (uint32_t)px & wmask; m_prey = (uint32_t)py & wmask; m_preph = pph & pmask;
		m_idle = n_idle; m_pre_valid = n_pre_valid; m_cangle = n_cangle & pmask; m_state = n_state;
		m_xv = n_xv & wmask; m_yv = n_yv & wmask; m_ph = n_ph & pmask;
		o_done = n_done;
		o_busy = !m_idle;
	}
private:
	static int32_t sext(uint32_t v) {
		return (WW >= 32) ? (int32_t)v : ((int32_t)(v << (32 - WW)) >> (32 - WW));
	}
};

#endif
