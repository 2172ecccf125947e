/*
 * zc_oracle.h -- CPU oracle for the zcordic hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the arithmetic that ZipCPU/cordic's generated
 * cores perform (rtl/cordic.v, rtl/topolar.v, rtl/sintable.v, rtl/quarterwav.v) and
 * of the parameter derivation in its generator (sw/main.cpp, sw/cordiclib.cpp).
 * Nothing under oracle/ is linked into, imported by, or executed from the product
 * (cordic_b200/, include/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker or as
 * the timed CPU baseline.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *  - parameters, angle tables, header constants and LUT / quadtbl words are pinned word-exactly against the
 *    reference's own generator built from /root/reference/sw (oracle/_ref/gencordic) and against the checked-in
 *    rtl/ artefacts (tests/golden/gen_*.json);
 *  - the datapaths are pinned per sample against the reference's RTL TEXT EXECUTED: oracle/vsim.py simulates
 *    rtl/cordic.v, topolar.v, seqcordic.v, seqpolar.v, quadtbl.v, sintable.v, quarterwav.v as checked in and the
 *    Verilog the real generator prints for 29 further command lines, and every one of its 21,504 vectors (tests/golden/rtl_vectors.json) is
 *    reproduced by this file (tests/test_rtl_vectors.py).  The reference itself holds no per-sample vectors (its
 *    tests are statistical, bench/cpp/cordic_tb.cpp:285-337, topolar_tb.cpp:303-315) and Verilator is not in this
 *    image, so "the reference run here" means: its generator run for real, its RTL run under our simulator;
 *  - the reference's own unmodified test benches compiled against oracle/shim/ and run over this oracle print
 *    SUCCESS with their own thresholds (oracle/_ref/cordic_tb_*, topolar_tb_*, quadtbl_tb_*, seqcordic_tb_*,
 *    seqpolar_tb_*).
 *
 * All citations are relative to /root/reference.
 */
#ifndef ZC_ORACLE_H
#define ZC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZO_MAX_STAGES 64

typedef struct zo_params {
	int	iw, ow;		/* input / output widths			*/
	int	nextra;		/* NEXTRA as printed in rtl/X.h (already bumped)	*/
	int	ww;		/* working width				*/
	int	pw;		/* phase width					*/
	int	nstages;
	int	vectoring;	/* 0: rtl/cordic.v (p2r), 1: rtl/topolar.v (r2p)	*/
	int	sequential;	/* 1: rtl/seqcordic.v / rtl/seqpolar.v (one sample at a time) */
	uint32_t angle[ZO_MAX_STAGES];
	double	gain;		/* GAIN constant of rtl/X.h			*/
	double	cordic_gain;	/* raw prod sqrt(1+2^-2(k+1))			*/
	double	qvar;		/* QUANTIZATION_VARIANCE			*/
	double	pvar_rad;	/* PHASE_VARIANCE_RAD				*/
	double	best_cnr;	/* BEST_POSSIBLE_CNR (p2r only)			*/
} zo_params;

/* sw/cordiclib.cpp */
double	zo_cordic_gain(int nstages);				/* :66-80   */
double	zo_phase_variance(int nstages, int pw);			/* :82-109  */
double	zo_quantization_variance(int nstages, int xtra, int dropped); /* :111-130 */
uint32_t zo_angle(int k, int pw);				/* :157-169 */
int	zo_calc_stages_ww(int ww, int pw);			/* :214-229 */
int	zo_calc_stages(int pw);					/* :231-244 */
int	zo_calc_phase_bits(int ow);				/* :246-268 */

/* sw/main.cpp:260-279 + sw/basiccordic.cpp:67-73,465-498.  Pass <=0 for "not given". */
int	zo_derive_p2r(int iw, int ow, int xtra_user, int pw, int nstages, zo_params *p);
/* sw/main.cpp:312-328 + sw/topolar.cpp:67-75,428-446 */
int	zo_derive_r2p(int iw, int ow, int xtra_user, int pw, int nstages, zo_params *p);
/* -t sp2r / -t sr2p (sw/main.cpp:183-198,301-304,347-350): the same constants, emitted by sw/seqcordic.cpp /
 * sw/seqpolar.cpp as a one-sample-at-a-time state machine.  What that machine computes differs from the
 * pipelined core (established by executing rtl/seqcordic.v, rtl/seqpolar.v and generated variants under
 * oracle/vsim.py): every iteration runs, zero angle or not (no pass-through test); seqcordic registers its
 * outputs when state == NSTAGES-1, before the last two iterations land (rtl/seqcordic.v:319-324), so
 * NSTAGES-2 iterations count; seqpolar's last_state is state >= NSTAGES+1 (rtl/seqpolar.v), NSTAGES
 * iterations.  sr2p returns -2 when NSTAGES+1 is a power of two: the reference's state register
 * (nextlg(NSTAGES+1) bits, sw/seqpolar.cpp:158-159) cannot reach NSTAGES+1 and o_done never rises. */
int	zo_derive_sp2r(int iw, int ow, int xtra_user, int pw, int nstages, zo_params *p);
int	zo_derive_sr2p(int iw, int ow, int xtra_user, int pw, int nstages, zo_params *p);
int	zo_iterations(const zo_params *p);	/* stage updates that reach the output */
int	zo_clocks_per_output(const zo_params *p);	/* CLOCKS_PER_OUTPUT of rtl/seq*.h; 1 for the pipelined cores */
/* sw/main.cpp:358-379 (tbl) and :401-422 (qtr): resolves (pw, ow) from -i/-p/-o */
int	zo_derive_tbl(int iw, int pw, int ow, int *pw_out, int *ow_out);
int	zo_derive_qtr(int iw, int pw, int ow, int *pw_out, int *ow_out);

/* rtl/cordic.v:85-86,131-188,253-280,290-295,311-312 -- one sample */
void	zo_rotate1(const zo_params *p, int32_t ix, int32_t iy, uint32_t phase,
		int32_t *ox, int32_t *oy);
/* rtl/topolar.v:83-84,122-152,217-243,253-255,268-269 -- one sample */
void	zo_topolar1(const zo_params *p, int32_t ix, int32_t iy,
		int32_t *omag, uint32_t *ophase);

/* Single pipeline-register steps, used by the cycle-accurate shim (oracle/shim) */
void	zo_rotate_pre(const zo_params *p, int32_t ix, int32_t iy, uint32_t phase,
		int32_t *x, int32_t *y, uint32_t *ph);
void	zo_rotate_stage(const zo_params *p, int i, int32_t *x, int32_t *y, uint32_t *ph);
void	zo_topolar_pre(const zo_params *p, int32_t ix, int32_t iy,
		int32_t *x, int32_t *y, uint32_t *ph);
void	zo_topolar_stage(const zo_params *p, int i, int32_t *x, int32_t *y, uint32_t *ph);
int32_t	zo_round_out(const zo_params *p, int32_t v);

/* Batched forms (xy interleaved).  nthreads<=1: scalar loop; else pthreads. */
void	zo_rotate_const(const zo_params *p, int32_t x0, int32_t y0,
		const uint32_t *phase, int32_t *xy, size_t n, int nthreads);
void	zo_rotate(const zo_params *p, const int32_t *xy_in, const uint32_t *phase,
		int32_t *xy_out, size_t n, int nthreads);
void	zo_topolar(const zo_params *p, const int32_t *xy_in, int32_t *mag,
		uint32_t *phase, size_t n, int nthreads);
/* NCO: phase32_n = phase0 + (n0+i)*step mod 2^32 ; i_phase = phase32 >> (32-pw) */
void	zo_nco_rotate(const zo_params *p, int32_t x0, int32_t y0, uint32_t phase0,
		uint32_t step, uint64_t n0, int32_t *xy, size_t n, int nthreads);

/* LUTs: sw/sintable.cpp:156-168 (full wave), :325-337 (quarter wave).
 * Tables hold the OW-bit words as $readmemh would load them (masked to OW bits). */
int	zo_sintable_build(int pw, int ow, uint32_t *tbl /* 2^pw */);
int	zo_quarterwav_build(int pw, int ow, uint32_t *tbl /* 2^(pw-2) */);
/* rtl/sintable.v:71-75 ; rtl/quarterwav.v:92-109.  Return o_val sign-extended from OW.
 * phase is a 32-bit NCO word; the core sees i_phase = phase >> (32-pw). */
int32_t	zo_sintable_lookup1(int pw, int ow, const uint32_t *tbl, uint32_t phase32);
int32_t	zo_quarterwav_lookup1(int pw, int ow, const uint32_t *tbl, uint32_t phase32);
void	zo_lut_sin(int pw, int ow, const uint32_t *tbl, const uint32_t *phase32,
		int32_t *out, size_t n, int nthreads);
void	zo_lut_qwav(int pw, int ow, const uint32_t *tbl, const uint32_t *phase32,
		int32_t *out, size_t n, int nthreads);

/* ---- quadratically interpolated sine table: sw/quadtbl.cpp, rtl/quadtbl.v ------------------ */
#define ZO_QT_MAXLG 12
typedef struct zo_quadtbl {
	int	ow, nextra;	/* OW, NEXTRA (XTRA) as printed in rtl/quadtbl.h		*/
	int	pw, ww;		/* PW ; WW = OW+XTRA						*/
	int	lgtbl, dxbits, cbits, lbits, qbits;	/* localparams of rtl/quadtbl.v:65-71		*/
	long	scale;		/* SCALE  = max_integer(ow)					*/
	double	itbl_err;	/* ITBL_ERR (the loop's tblerr)					*/
	double	tbl_err;	/* TBL_ERR							*/
	double	spurdb;		/* SPURDB							*/
	uint32_t ctbl[1 << ZO_QT_MAXLG], ltbl[1 << ZO_QT_MAXLG], qtbl[1 << ZO_QT_MAXLG];	/* $readmemh words */
} zo_quadtbl;
/* gencordic -t qtbl [-i iw] [-o ow] [-p pw] [-x xtra]: sw/main.cpp:444-484 + sw/quadtbl.cpp:270-304 */
int	zo_derive_qtbl(int iw, int ow, int xtra_user, int pw, zo_quadtbl *q);
/* rtl/quadtbl.v:143-291, one sample; phase is the PW-bit port word */
int32_t	zo_quadtbl1(const zo_quadtbl *q, uint32_t phase);
void	zo_quadtbl_batch(const zo_quadtbl *q, const uint32_t *phase, int32_t *out, size_t n, int nthreads);

/* sw/hexfile.cpp:78-89 -- parse a $readmemh file written by hextable(). Returns #words. */
long	zo_hex_load(const char *fname, uint32_t *words, long maxwords);

#ifdef __cplusplus
}
#endif
#endif
