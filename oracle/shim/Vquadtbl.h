// oracle/shim/Vquadtbl.h -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the Verilated model of rtl/quadtbl.v (ports :53-58).  The core is a 6-register feed-forward
// pipeline (clock 1 table read :143-160, 2 qv*dx :173, 3 lsum :205, 4 lsum*dx :231, 5 r_value :254,
// 6 o_sin :277; o_aux = aux[NSTAGES-1], :121-127), modelled as the oracle's per-sample function
// (zo_quadtbl1) followed by a 6-deep delay line that reset clears.
#ifndef ZC_SHIM_VQUADTBL_H
#define ZC_SHIM_VQUADTBL_H
#include "verilated.h"
#include "verilated_vcd_c.h"
#include "quadtbl.h"		// OW NEXTRA PW TBL_LGSZ ... (rtl/quadtbl.h)
#include "zc_oracle.h"

class Vquadtbl {
	zo_quadtbl *m_q;
	uint32_t m_val[6], m_aux[6];
	uint32_t m_lastclk;
public:
	uint32_t i_clk, i_reset, i_ce, i_aux, i_phase, o_sin, o_aux;
	Vquadtbl() {
		m_q = new zo_quadtbl;
		int rc = zo_derive_qtbl(0, OW, NEXTRA - 1, PW, m_q);
		assert(rc == 0 && m_q->lgtbl == TBL_LGSZ);
		(void)rc;
		memset(m_val, 0, sizeof(m_val)); memset(m_aux, 0, sizeof(m_aux));
		m_lastclk = 0;
		i_clk = i_reset = i_ce = i_aux = i_phase = o_sin = o_aux = 0;
	}
	~Vquadtbl() { delete m_q; }
	void trace(VerilatedVcdC *t, int) {
		t->declare("i_clk", 1, &i_clk); t->declare("i_reset", 1, &i_reset); t->declare("i_ce", 1, &i_ce);
		t->declare("i_aux", 1, &i_aux); t->declare("i_phase", PW, &i_phase);
		t->declare("o_sin", OW, &o_sin); t->declare("o_aux", 1, &o_aux);
	}
	void eval() {
		bool rising = (i_clk & 1) && !(m_lastclk & 1);
		m_lastclk = i_clk;
		if (!rising) return;
		if (i_reset & 1) {
			memset(m_val, 0, sizeof(m_val)); memset(m_aux, 0, sizeof(m_aux));
			o_sin = o_aux = 0;
			return;
		}
		if (!(i_ce & 1)) return;
		for (int k = 5; k > 0; k--) { m_val[k] = m_val[k - 1]; m_aux[k] = m_aux[k - 1]; }
		const uint32_t omask = (1u << OW) - 1u;
		m_val[0] = (uint32_t)zo_quadtbl1(m_q, i_phase) & omask;
		m_aux[0] = i_aux & 1;
		o_sin = m_val[5]; o_aux = m_aux[5];
	}
};
#endif
