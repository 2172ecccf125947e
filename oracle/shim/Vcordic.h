// oracle/shim/Vcordic.h -- TEST INFRASTRUCTURE ONLY.
// Cycle-accurate stand-in for the Verilated model of rtl/cordic.v: the same port
// fields and eval()/trace() surface bench/cpp/testb.h drives, with one register per
// pipeline stage (rtl/cordic.v:118-190 stage 0, :232-283 stages 1..N, :303-314 output,
// :100-105 aux chain), updated on the rising edge of i_clk.  The per-stage arithmetic
// is the oracle's (zc_oracle.c), so running the reference's unmodified cordic_tb.cpp
// over this model checks the oracle against the reference's own acceptance thresholds.
#ifndef ZC_SHIM_VCORDIC_H
#define ZC_SHIM_VCORDIC_H

#include "verilated.h"
#include "verilated_vcd_c.h"
#include "cordic.h"	// generated constants: IW OW NEXTRA WW PW NSTAGES (rtl/cordic.h:46-59)
#include "zc_oracle.h"

class Vcordic {
	zo_params m_p;
	uint32_t m_xv[ZO_MAX_STAGES + 1], m_yv[ZO_MAX_STAGES + 1], m_ph[ZO_MAX_STAGES + 1];
	uint64_t m_ax;
	uint32_t m_lastclk;
public:
	uint32_t i_clk, i_reset, i_ce, i_xval, i_yval, i_phase, i_aux;
	uint32_t o_xval, o_yval, o_aux;

	Vcordic() {
		int rc = zo_derive_p2r(IW, OW, NEXTRA - 1, PW, NSTAGES, &m_p);
		assert(rc == 0 && m_p.ww == WW);
		(void)rc;
		memset(m_xv, 0, sizeof(m_xv)); memset(m_yv, 0, sizeof(m_yv));
		memset(m_ph, 0, sizeof(m_ph));
		m_ax = 0; m_lastclk = 0;
		i_clk = i_reset = i_ce = i_xval = i_yval = i_phase = i_aux = 0;
		o_xval = o_yval = o_aux = 0;
	}
	void trace(VerilatedVcdC *t, int) {
		t->declare("i_clk", 1, &i_clk);     t->declare("i_reset", 1, &i_reset);
		t->declare("i_ce", 1, &i_ce);       t->declare("i_xval", IW, &i_xval);
		t->declare("i_yval", IW, &i_yval);  t->declare("i_phase", PW, &i_phase);
		t->declare("i_aux", 1, &i_aux);     t->declare("o_xval", OW, &o_xval);
		t->declare("o_yval", OW, &o_yval);  t->declare("o_aux", 1, &o_aux);
		static char names[3 * (ZO_MAX_STAGES + 1)][16];
		for (int k = 0; k <= NSTAGES; k++) {
			snprintf(names[3 * k], 16, "xv(%d)", k);
			snprintf(names[3 * k + 1], 16, "yv(%d)", k);
			snprintf(names[3 * k + 2], 16, "ph(%d)", k);
			t->declare(names[3 * k], WW, &m_xv[k]);
			t->declare(names[3 * k + 1], WW, &m_yv[k]);
			t->declare(names[3 * k + 2], PW, &m_ph[k]);
		}
	}
	void eval() {
		bool rising = (i_clk & 1) && !(m_lastclk & 1);
		m_lastclk = i_clk;
		if (!rising) return;
		const uint32_t wmask = (WW >= 32) ? 0xffffffffu : ((1u << WW) - 1u);
		const uint32_t omask = (OW >= 32) ? 0xffffffffu : ((1u << OW) - 1u);
		if (i_reset & 1) {
			memset(m_xv, 0, sizeof(m_xv)); memset(m_yv, 0, sizeof(m_yv));
			memset(m_ph, 0, sizeof(m_ph));
			m_ax = 0; o_xval = o_yval = o_aux = 0;
			return;
		}
		if (!(i_ce & 1)) return;
		// Non-blocking assignments: evaluate from the tail of the pipe backwards.
		o_xval = (uint32_t)zo_round_out(&m_p, sext(m_xv[NSTAGES])) & omask;
		o_yval = (uint32_t)zo_round_out(&m_p, sext(m_yv[NSTAGES])) & omask;
		o_aux = (uint32_t)((m_ax >> NSTAGES) & 1);
		for (int i = NSTAGES - 1; i >= 0; i--) {
			int32_t x = sext(m_xv[i]), y = sext(m_yv[i]);
			uint32_t ph = m_ph[i];
			zo_rotate_stage(&m_p, i, &x, &y, &ph);
			m_xv[i + 1] = (uint32_t)x & wmask;
			m_yv[i + 1] = (uint32_t)y & wmask;
			m_ph[i + 1] = ph;
		}
		{
			int32_t x, y; uint32_t ph;
			zo_rotate_pre(&m_p, (int32_t)i_xval, (int32_t)i_yval, i_phase, &x, &y, &ph);
			m_xv[0] = (uint32_t)x & wmask; m_yv[0] = (uint32_t)y & wmask; m_ph[0] = ph;
		}
		m_ax = ((m_ax << 1) | (i_aux & 1)) & ((2ull << NSTAGES) - 1ull);
	}
private:
	static int32_t sext(uint32_t v) {
		return (WW >= 32) ? (int32_t)v : ((int32_t)(v << (32 - WW)) >> (32 - WW));
	}
};

#endif
