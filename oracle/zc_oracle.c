/*
 * zc_oracle.c -- CPU oracle for the zcordic hot path.  TEST INFRASTRUCTURE ONLY.
 * See zc_oracle.h for the scope statement and how parity is pinned.
 * Citations are file:line relative to /root/reference.
 */
#define _GNU_SOURCE
#include "zc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- bit helpers ------------------------------------------------------- */

/* low w bits of v as two's complement (Verilog "signed [w-1:0]") */
static inline int64_t sx(int64_t v, int w) {
	uint64_t m = (w >= 64) ? ~0ull : ((1ull << w) - 1ull);
	uint64_t u = (uint64_t)v & m;
	if (w < 64 && (u >> (w - 1)) & 1ull)
		u |= ~m;
	return (int64_t)u;
}

static inline uint64_t ux(uint64_t v, int w) {
	return (w >= 64) ? v : (v & ((1ull << w) - 1ull));
}

/* ---- sw/cordiclib.cpp -------------------------------------------------- */

/* sw/cordiclib.cpp:66-80 */
double zo_cordic_gain(int nstages) {
	double gain = 1.0;
	for (int k = 0; k < nstages; k++) {
		double dgain = 1.0 + pow(2.0, -2. * (k + 1));
		dgain = sqrt(dgain);
		gain = gain * dgain;
	}
	return gain;
}

/* sw/cordiclib.cpp:82-109 */
double zo_phase_variance(int nstages, int pw) {
	double RAD_TO_PHASE = (double)(1ul << (pw - 1)) / M_PI;
	double variance = 1. / 12.;
	for (unsigned k = 0; k < (unsigned)nstages; k++) {
		double x, err;
		unsigned long phase_value;
		x = atan2(1., pow(2, k + 1)) * RAD_TO_PHASE;
		phase_value = (unsigned)x;
		err = phase_value - x;
		err *= err;
		variance += err;
	}
	variance /= pow(RAD_TO_PHASE, 2.);
	return variance;
}

/* sw/cordiclib.cpp:111-130 */
double zo_quantization_variance(int nstages, int xtrabits, int dropped_bits) {
	double v = pow(2, 2 * xtrabits) / 12.;
	for (int k = 0; k < nstages; k++)
		v = (1 + pow(4, -k - 1)) * v + 1. / 3.;
	if (dropped_bits > 0)
		v = pow(2, -2 * dropped_bits) * v + 1 / 12.;
	return v;
}

/* sw/cordiclib.cpp:157-169 -- truncation, not rounding */
uint32_t zo_angle(int k, int pw) {
	double x = atan2(1., pow(2, k + 1));
	x *= (4.0 * (double)(1ul << (pw - 2))) / (M_PI * 2.0);
	return (uint32_t)(unsigned)x;
}

/* sw/cordiclib.cpp:214-229 */
int zo_calc_stages_ww(int ww, int pw) {
	unsigned n;
	for (n = 0; n < 64; n++) {
		if (zo_angle((int)n, pw) == 0)
			break;
		if (ww <= (int)n)
			break;
	}
	return (int)n;
}

/* sw/cordiclib.cpp:231-244 */
int zo_calc_stages(int pw) {
	unsigned n;
	for (n = 0; n < 64; n++)
		if (zo_angle((int)n, pw) == 0)
			break;
	return (int)n;
}

/* sw/cordiclib.cpp:246-268 -- note (2^ow - 1), not (2^(ow-1) - 1) */
int zo_calc_phase_bits(int ow) {
	unsigned pb;
	for (pb = 3; pb < 64; pb++) {
		double a = (2.0 * M_PI / (double)(1ul << pb));
		double ds = sin(a);
		ds *= (double)((1ul << ow) - 1);
		if (ds < 0.5)
			break;
	}
	if (pb < 3)
		pb = 3;
	return (int)pb;
}

/* ---- parameter derivation ---------------------------------------------- */

static void fill_angles(zo_params *p) {
	memset(p->angle, 0, sizeof(p->angle));
	for (int k = 0; k < p->nstages && k < ZO_MAX_STAGES; k++)
		p->angle[k] = zo_angle(k, p->pw);
}

static int default_widths(int *iw, int *ow) {
	/* sw/main.cpp:262-270 (p2r), :314-322 (r2p) */
	if ((*iw <= 0) && (*ow > 0))
		*iw = *ow;
	if (*ow <= 0)
		*ow = *iw;
	if ((*iw <= 0) || (*ow <= 0)) {
		*iw = 24;
		*ow = 24;
	}
	return 0;
}

int zo_derive_p2r(int iw, int ow, int xtra_user, int pw, int nstages, zo_params *p) {
	memset(p, 0, sizeof(*p));
	default_widths(&iw, &ow);
	int mx = (ow > iw) ? ow : iw;
	int nxtra = xtra_user + 1;		/* sw/main.cpp:273 */
	int ww_main = mx + nxtra;		/* sw/main.cpp:272-274 */
	if (pw <= 0)
		pw = zo_calc_phase_bits(ww_main);	/* :276 */
	if (nstages <= 0)
		nstages = zo_calc_stages_ww(ww_main, pw);	/* :278 */
	if (nxtra < 1)				/* sw/basiccordic.cpp:67-68 */
		nxtra = 1;
	if (pw < 3 || pw > 32 || nstages > ZO_MAX_STAGES)
		return -1;
	p->iw = iw; p->ow = ow; p->nextra = nxtra;
	p->ww = mx + nxtra;			/* sw/basiccordic.cpp:71-73 */
	p->pw = pw; p->nstages = nstages; p->vectoring = 0;
	if (p->ww > 32)
		return -1;
	fill_angles(p);
	/* sw/basiccordic.cpp:471-496 */
	p->qvar = zo_quantization_variance(nstages, p->ww - iw, p->ww - ow);
	p->pvar_rad = zo_phase_variance(nstages, pw);
	p->cordic_gain = zo_cordic_gain(nstages);
	p->gain = p->cordic_gain;
	{
		double amplitude = (double)((1ul << (iw - 1))) - 1., sig, noise;
		amplitude *= (double)(1ul << (p->ww - iw));
		amplitude *= zo_cordic_gain(nstages);
		amplitude *= pow(2.0, -(p->ww - ow));
		sig = amplitude * amplitude;
		noise = zo_quantization_variance(nstages, p->ww - iw, p->ww - ow);
		noise += sig * zo_phase_variance(nstages, pw)
			* pow(2, zo_cordic_gain(nstages));	/* (sic) :490-491 */
		p->best_cnr = 10.0 * log(sig / noise) / log(10.0);
	}
	return 0;
}

int zo_derive_r2p(int iw, int ow, int xtra_user, int pw, int nstages, zo_params *p) {
	memset(p, 0, sizeof(*p));
	default_widths(&iw, &ow);
	int mx = (ow > iw) ? ow : iw;
	int nxtra = xtra_user + 2;		/* sw/main.cpp:323 */
	int ww_main = mx + nxtra;
	if (pw <= 0)
		pw = zo_calc_phase_bits(ww_main);	/* :325-326 */
	if (nstages <= 0)
		nstages = zo_calc_stages(pw);		/* :327-328 */
	if (nxtra < 2)				/* sw/topolar.cpp:67-68 */
		nxtra = 2;
	if (pw < 3 || pw > 32 || nstages > ZO_MAX_STAGES)
		return -1;
	p->iw = iw; p->ow = ow; p->nextra = nxtra;
	p->ww = mx + nxtra + nxtra;		/* sw/topolar.cpp:71-75: added twice */
	p->pw = pw; p->nstages = nstages; p->vectoring = 1;
	if (p->ww > 32)
		return -1;
	fill_angles(p);
	/* sw/topolar.cpp:434-440 */
	p->qvar = zo_quantization_variance(nstages, p->ww - iw, p->ww - ow);
	p->pvar_rad = zo_phase_variance(nstages, pw);
	p->cordic_gain = zo_cordic_gain(nstages);
	p->gain = p->cordic_gain * sqrt(2.0) / 2.;
	p->best_cnr = 0.0;
	return 0;
}

static int derive_lut(int iw, int pw, int ow, int qtr, int *pw_out, int *ow_out) {
	/* sw/main.cpp:358-379 (tbl): (iw>=0)&&(phase_bits<=0);  :401-422 (qtr): phase_bits<0.
	 * "not given" is -1 in main.cpp; callers pass <=0 which we map to -1 first. */
	if (iw <= 0) iw = -1;
	if (pw <= 0) pw = -1;
	if (ow <= 0) ow = -1;
	if ((iw >= 0) && (qtr ? (pw < 0) : (pw <= 0))) {
		pw = iw;
		iw = -1;
	}
	if ((pw > 3) && (ow <= 0)) {
		for (int k = pw - 2; k < pw + 3; k++) {
			if (zo_calc_phase_bits(k) == pw) {
				ow = k;
				break;
			}
		}
	}
	if (ow <= 0)
		ow = 24;
	if (pw <= 0)
		pw = zo_calc_phase_bits(ow);
	*pw_out = pw;
	*ow_out = ow;
	/* sw/sintable.cpp:62 (tbl >=24 refused), :190 (qtr >=26 refused), hexfile.cpp:52 */
	if (qtr ? (pw >= 26 || pw <= 2) : (pw >= 24))
		return -1;
	if (ow >= 31)
		return -1;
	return 0;
}

int zo_derive_tbl(int iw, int pw, int ow, int *pw_out, int *ow_out) {
	return derive_lut(iw, pw, ow, 0, pw_out, ow_out);
}
int zo_derive_qtr(int iw, int pw, int ow, int *pw_out, int *ow_out) {
	return derive_lut(iw, pw, ow, 1, pw_out, ow_out);
}

/* ---- rotation core: rtl/cordic.v --------------------------------------- */

/* rtl/cordic.v:85-86 (extend) and :131-188 (octant pre-rotation) */
/* ---- sequential cores: rtl/seqcordic.v, rtl/seqpolar.v ----------------------- */
int zo_derive_sp2r(int iw, int ow, int xtra_user, int pw, int nstages, zo_params *p) {
	int rc = zo_derive_p2r(iw, ow, xtra_user, pw, nstages, p);	/* same branch of sw/main.cpp:260-279 */
	if (rc != 0)
		return rc;
	if (p->nstages < 3)
		return -2;
	p->sequential = 1;
	return 0;
}

int zo_derive_sr2p(int iw, int ow, int xtra_user, int pw, int nstages, zo_params *p) {
	int rc = zo_derive_r2p(iw, ow, xtra_user, pw, nstages, p);	/* sw/main.cpp:312-328 */
	if (rc != 0)
		return rc;
	if (p->nstages < 1 || ((p->nstages + 1) & p->nstages) == 0)	/* state register too narrow: never done */
		return -2;
	p->sequential = 1;
	return 0;
}

int zo_iterations(const zo_params *p) {
	if (!p->sequential)
		return p->nstages;
	return p->vectoring ? p->nstages : p->nstages - 2;
}

int zo_clocks_per_output(const zo_params *p) {
	if (!p->sequential)
		return 1;
	return p->vectoring ? p->nstages + 3 : p->nstages + 1;	/* sw/seqpolar.cpp:396, sw/seqcordic.cpp:459 */
}

void zo_rotate_pre(const zo_params *p, int32_t ix, int32_t iy, uint32_t phase,
		int32_t *x, int32_t *y, uint32_t *ph) {
	const int WW = p->ww, IW = p->iw, PW = p->pw;
	/* { sign, i_xval, (WW-IW-1) zeros }  (sw/basiccordic.cpp:137-145) */
	int64_t ex = sx(sx(ix, IW) * ((int64_t)1 << (WW - IW - 1)), WW);
	int64_t ey = sx(sx(iy, IW) * ((int64_t)1 << (WW - IW - 1)), WW);
	uint64_t ip = ux(phase, PW);
	uint64_t Q = 1ull << (PW - 2);
	int64_t xv, yv;
	uint64_t pv;
	switch ((int)((ip >> (PW - 3)) & 7)) {
	case 0: case 7:
		xv = ex;  yv = ey;  pv = ip;		break;
	case 1: case 2:
		xv = -ey; yv = ex;  pv = ip - Q;	break;
	case 3: case 4:
		xv = -ex; yv = -ey; pv = ip - 2 * Q;	break;
	default: /* 5, 6 */
		xv = ey;  yv = -ex; pv = ip - 3 * Q;	break;
	}
	*x = (int32_t)sx(xv, WW);
	*y = (int32_t)sx(yv, WW);
	*ph = (uint32_t)ux(pv, PW);
}

/* rtl/cordic.v:253-280 */
void zo_rotate_stage(const zo_params *p, int i, int32_t *x, int32_t *y, uint32_t *ph) {
	const int WW = p->ww, PW = p->pw;
	/* the sequential machine has no such test: rtl/seqcordic.v:281-299 runs every state */
	if (!p->sequential && (p->angle[i] == 0 || i >= WW))
		return;
	int64_t xv = *x, yv = *y;
	uint64_t pv = *ph;
	const int sh = (i + 1 > 62) ? 62 : (i + 1);	/* >>> by WW or more leaves the sign */
	int64_t xs = xv >> sh, ys = yv >> sh;	/* >>> on signed regs */
	if ((pv >> (PW - 1)) & 1) {	/* negative phase */
		*x = (int32_t)sx(xv + ys, WW);
		*y = (int32_t)sx(yv - xs, WW);
		*ph = (uint32_t)ux(pv + p->angle[i], PW);
	} else {
		*x = (int32_t)sx(xv - ys, WW);
		*y = (int32_t)sx(yv + xs, WW);
		*ph = (uint32_t)ux(pv - p->angle[i], PW);
	}
}

/* rtl/cordic.v:290-295,311-312 ; no-rounding branch sw/basiccordic.cpp:407-444 */
int32_t zo_round_out(const zo_params *p, int32_t v) {
	const int WW = p->ww, OW = p->ow, D = WW - OW;
	int64_t x = v;
	if (WW > OW + 1) {
		int b = (int)((x >> D) & 1);
		int64_t add = b ? ((int64_t)1 << (D - 1)) : (((int64_t)1 << (D - 1)) - 1);
		x = sx(x + add, WW);
	}
	/* bits [WW-1:D] of the WW-bit word, as an OW-bit signed output */
	return (int32_t)sx((int64_t)(ux((uint64_t)x, WW) >> D), OW);
}

void zo_rotate1(const zo_params *p, int32_t ix, int32_t iy, uint32_t phase,
		int32_t *ox, int32_t *oy) {
	int32_t x, y;
	uint32_t ph;
	zo_rotate_pre(p, ix, iy, phase, &x, &y, &ph);
	const int n = zo_iterations(p);
	for (int i = 0; i < n; i++)
		zo_rotate_stage(p, i, &x, &y, &ph);
	*ox = zo_round_out(p, x);
	*oy = zo_round_out(p, y);
}

/* ---- vectoring core: rtl/topolar.v ------------------------------------- */

/* rtl/topolar.v:83-84 (extend; variants sw/topolar.cpp:139-151), :122-152 (quadrant) */
void zo_topolar_pre(const zo_params *p, int32_t ix, int32_t iy,
		int32_t *x, int32_t *y, uint32_t *ph) {
	const int WW = p->ww, IW = p->iw, PW = p->pw;
	int64_t sxv = sx(ix, IW), syv = sx(iy, IW);
	int64_t ex, ey;
	if (WW - IW > 2) {
		ex = sxv * ((int64_t)1 << (WW - IW - 2));
		ey = syv * ((int64_t)1 << (WW - IW - 2));
	} else if (WW - IW > 1) {
		ex = sxv; ey = syv;
	} else {
		ex = sxv >> 1; ey = syv >> 1;
	}
	ex = sx(ex, WW); ey = sx(ey, WW);
	uint64_t E = 1ull << (PW - 3);
	int xneg = (sxv < 0), yneg = (syv < 0);
	int64_t xv, yv;
	uint64_t pv;
	if (!xneg && yneg) {		/* 2'b01 */
		xv = ex - ey;  yv = ex + ey;  pv = 7 * E;
	} else if (xneg && !yneg) {	/* 2'b10 */
		xv = -ex + ey; yv = -ex - ey; pv = 3 * E;
	} else if (xneg && yneg) {	/* 2'b11 */
		xv = -ex - ey; yv = ex - ey;  pv = 5 * E;
	} else {			/* default */
		xv = ex + ey;  yv = -ex + ey; pv = E;
	}
	*x = (int32_t)sx(xv, WW);
	*y = (int32_t)sx(yv, WW);
	*ph = (uint32_t)ux(pv, PW);
}

/* rtl/topolar.v:217-243 */
void zo_topolar_stage(const zo_params *p, int i, int32_t *x, int32_t *y, uint32_t *ph) {
	const int WW = p->ww, PW = p->pw;
	if (!p->sequential && (p->angle[i] == 0 || i >= WW))	/* rtl/seqpolar.v runs every state */
		return;
	int64_t xv = *x, yv = *y;
	uint64_t pv = *ph;
	const int sh = (i + 1 > 62) ? 62 : (i + 1);
	int64_t xs = xv >> sh, ys = yv >> sh;
	if (yv < 0) {		/* yv[WW-1]: below the axis */
		*x = (int32_t)sx(xv - ys, WW);
		*y = (int32_t)sx(yv + xs, WW);
		*ph = (uint32_t)ux(pv - p->angle[i], PW);
	} else {
		*x = (int32_t)sx(xv + ys, WW);
		*y = (int32_t)sx(yv - xs, WW);
		*ph = (uint32_t)ux(pv + p->angle[i], PW);
	}
}

/* rtl/topolar.v:253-255,268-269 */
void zo_topolar1(const zo_params *p, int32_t ix, int32_t iy,
		int32_t *omag, uint32_t *ophase) {
	int32_t x, y;
	uint32_t ph;
	zo_topolar_pre(p, ix, iy, &x, &y, &ph);
	const int n = zo_iterations(p);
	for (int i = 0; i < n; i++)
		zo_topolar_stage(p, i, &x, &y, &ph);
	*omag = zo_round_out(p, x);
	*ophase = ph;
}

/* ---- LUT cores --------------------------------------------------------- */

/* sw/sintable.cpp:156-168 + sw/hexfile.cpp:78-89 (mask to OW bits) */
int zo_sintable_build(int pw, int ow, uint32_t *tbl) {
	if (pw >= 24 || pw < 1 || ow >= 31 || ow < 2)
		return -1;
	int tbl_entries = (1 << pw);
	long maxv = (1l << (ow - 1)) - 1l;
	long msk = (1l << ow) - 1l;
	for (int k = 0; k < tbl_entries; k++) {
		double ph = 2.0 * M_PI * (double)k / (double)tbl_entries;
		long v = (long)((double)maxv * sin(ph));
		tbl[k] = (uint32_t)(v & msk);
	}
	return 0;
}

/* sw/sintable.cpp:325-337 */
int zo_quarterwav_build(int pw, int ow, uint32_t *tbl) {
	if (pw >= 26 || pw <= 2 || ow >= 31 || ow < 2)
		return -1;
	int tbl_entries = (1 << pw);
	long maxv = (1l << (ow - 1)) - 1l;
	long msk = (1l << ow) - 1l;
	for (int k = 0; k < tbl_entries / 4; k++) {
		double ph = 2.0 * M_PI * (double)k / (double)tbl_entries;
		ph += M_PI / (double)tbl_entries;
		long v = (long)((double)maxv * sin(ph));
		tbl[k] = (uint32_t)(v & msk);
	}
	return 0;
}

/* rtl/sintable.v:71-75 */
int32_t zo_sintable_lookup1(int pw, int ow, const uint32_t *tbl, uint32_t phase32) {
	uint32_t ip = phase32 >> (32 - pw);
	return (int32_t)sx(tbl[ip], ow);
}

/* rtl/quarterwav.v:92-109 */
int32_t zo_quarterwav_lookup1(int pw, int ow, const uint32_t *tbl, uint32_t phase32) {
	uint32_t ip = phase32 >> (32 - pw);
	uint32_t lowmask = (1u << (pw - 2)) - 1u;
	int negate = (ip >> (pw - 1)) & 1;
	uint32_t index = ((ip >> (pw - 2)) & 1) ? (~ip & lowmask) : (ip & lowmask);
	uint64_t v = tbl[index];
	if (negate)
		v = ux((uint64_t)(-(int64_t)v), ow);
	return (int32_t)sx((int64_t)v, ow);
}

/* ---- quadratic-interpolation core: sw/quadtbl.cpp + rtl/quadtbl.v -------- */

/* sw/quadtbl.cpp:54-57 */
static double qt_sinc(double v) {
	double x = v * M_PI;
	return sin(x) / x;
}

/* sw/quadtbl.cpp:72-114 */
static double qt_est_max_err(double c, double l, double q, double idx, int N) {
	double lft, rht, mid, ph;
	ph = 2.0 * M_PI * idx / (double)N;
	lft = c - sin(ph);
	ph = 2.0 * M_PI * (idx + 1) / (double)N;
	rht = c + l + q - sin(ph);
	mid = 0;
	for (int k = 0; k < 64; k++) {
		double mer, mph, mdx;
		mdx = k / 64.0;
		mph = 2.0 * M_PI * (idx + mdx) / N;
		mer = c + (l + q * mdx) * mdx - sin(mph);
		if (fabs(mer) > fabs(mid))
			mid = mer;
	}
	double er = lft;
	if (fabs(er) < fabs(rht))
		er = rht;
	if (fabs(er) < fabs(mid))
		er = mid;
	return er;
}

/* sw/quadtbl.cpp:136-266: coefficient tables, their widths and the table error */
static void qt_build(int lgsz, int wid, int *cbits, int *lbits, int *qbits, double *tblerr,
		long *ct, long *lt, long *qt) {
	int ln = (1 << lgsz);
	long maxv = (1l << (wid - 1)) - 2l;		/* max_integer(): :59-61 */
	double dl = M_PI / (double)ln, dph = dl * 2.;
	double *table = (double *)malloc(sizeof(double) * ln);
	double *slope = (double *)malloc(sizeof(double) * ln);
	double *dslope = (double *)malloc(sizeof(double) * ln);
	int i;
	for (i = 0; i < ln; i++)
		table[i] = sin(dph * i + dl);
	for (i = 1; i < ln - 1; i++)
		slope[i] = (table[i + 1] - table[i - 1]) / 2.0;
	slope[0] = (table[1] - table[ln - 1]) / 2.0;
	slope[ln - 1] = (table[0] - table[ln - 2]) / 2.0;
	for (i = 1; i < ln - 1; i++)
		dslope[i] = -(table[i] - 0.5 * (table[i + 1] + table[i - 1]));
	dslope[0] = -(table[0] - 0.5 * (table[1] + table[ln - 1]));
	dslope[ln - 1] = -(table[ln - 1] - 0.5 * (table[0] + table[ln - 2]));
	for (i = 0; i < ln; i++)
		table[i] = 0.75 * sin(dph * i + dl)
			+ (sin(dph * (i - 1) + dl) + sin(dph * (i + 1) + dl)) / 8.0;
	const double del = 1.0, hlfdel = del / 2.0;
	for (i = 0; i < ln; i++)
		table[i] = dslope[i] * hlfdel * hlfdel - slope[i] * hlfdel + table[i];
	for (i = 0; i < ln; i++)
		slope[i] = slope[i] - del * dslope[i];
	double fctr = pow(1. / qt_sinc(dl), 3);
	for (i = 0; i < ln; i++) table[i] *= fctr;
	for (i = 0; i < ln; i++) slope[i] *= fctr;
	for (i = 0; i < ln; i++) dslope[i] *= fctr;
	double mxtbl = 0.0, mxslope = 0.0, mxdslope = 0.0;
	for (i = 0; i < ln; i++)
		mxtbl = (mxtbl > fabs(table[i])) ? mxtbl : fabs(table[i]);
	for (i = 0; i < ln; i++) table[i] *= 1. / mxtbl;
	for (i = 0; i < ln; i++) slope[i] *= 1. / mxtbl;
	for (i = 0; i < ln; i++) dslope[i] *= 1. / mxtbl;
	double mxerr = 0.0, err;
	for (i = 0; i < ln; i++) {
		err = qt_est_max_err(table[i], slope[i], dslope[i], i, ln);
		if (fabs(err) > fabs(mxerr))
			mxerr = err;
	}
	mxerr *= maxv;
	*tblerr = mxerr;
	mxtbl = 0.0;
	for (i = 0; i < ln; i++)
		mxtbl = (mxtbl > fabs(table[i])) ? mxtbl : fabs(table[i]);
	for (i = 0; i < ln; i++) {
		mxslope = (mxslope > fabs(slope[i])) ? mxslope : fabs(slope[i]);
		mxdslope = (mxdslope > fabs(dslope[i])) ? mxdslope : fabs(dslope[i]);
	}
	*cbits = wid + (int)ceil(log(mxtbl) / log(2.0));
	*lbits = wid + (int)ceil(-log(1. / mxslope) / log(2.0));
	*qbits = wid + (int)ceil(-log(1. / mxdslope) / log(2.0));
	for (i = 0; i < ln; i++) {
		ct[i] = (long)(maxv * table[i]);
		lt[i] = (long)(maxv * slope[i]);
		qt[i] = (long)(maxv * dslope[i]);
	}
	free(table); free(slope); free(dslope);
}

int zo_derive_qtbl(int iw, int ow, int xtra_user, int pw, zo_quadtbl *q) {
	memset(q, 0, sizeof(*q));
	default_widths(&iw, &ow);			/* sw/main.cpp:446-454 */
	int mx = (ow > iw) ? ow : iw;
	int nxtra = xtra_user + 1;			/* :456 */
	int ww_main = mx + nxtra;
	if (pw <= 0)
		pw = zo_calc_phase_bits(ww_main);	/* :458-459 */
	if (nxtra < 0 || pw <= 4 || pw > 32)
		return -1;
	/* sw/quadtbl.cpp:295-301: grow the table until the table error is below one unit */
	int lgtbl = 3, cbits = 0, lbits = 0, qbits = 0;
	double tblerr = 0;
	static long ct[1 << ZO_QT_MAXLG], lt[1 << ZO_QT_MAXLG], qt[1 << ZO_QT_MAXLG];
	if (ow + nxtra <= 6 || ow + nxtra > 30)
		return -1;
	do {
		lgtbl++;
		if (lgtbl > ZO_QT_MAXLG)
			return -1;
		qt_build(lgtbl, ow + nxtra, &cbits, &lbits, &qbits, &tblerr, ct, lt, qt);
	} while ((fabs(tblerr) > 1.0) && (lgtbl < 20));
	if (pw <= lgtbl)
		return -1;
	int wid = ow + nxtra;
	if (nxtra < 2)					/* :315-316 */
		nxtra = 2;
	q->ow = ow; q->nextra = nxtra; q->pw = pw; q->ww = ow + nxtra;
	q->lgtbl = lgtbl; q->dxbits = pw - lgtbl + 1;
	q->cbits = cbits; q->lbits = lbits; q->qbits = qbits;
	q->scale = (1l << (ow - 1)) - 2l;		/* :789-790 */
	q->itbl_err = tblerr;
	q->tbl_err = tblerr * pow(0.5, wid);		/* :793-795 (ow + nxtra before the clamp) */
	{
		double spur = pow(qt_sinc(1.0 - (1. / (1 << lgtbl))), 3.);
		q->spurdb = 20. * log(spur) / log(10.0);	/* :797-799 */
	}
	long cm = (1l << cbits) - 1l, lm = (1l << lbits) - 1l, qm = (1l << qbits) - 1l;
	for (int k = 0; k < (1 << lgtbl); k++) {
		q->ctbl[k] = (uint32_t)(ct[k] & cm);
		q->ltbl[k] = (uint32_t)(lt[k] & lm);
		q->qtbl[k] = (uint32_t)(qt[k] & qm);
	}
	/* widths this restatement (and the engine) accept: what every sane command line produces */
	if (q->cbits != q->ww || q->lbits < q->qbits + 1 || q->cbits < q->lbits + 1 || q->dxbits < 2
			|| q->qbits < 2 || q->cbits > 30 || q->lbits + q->dxbits > 62)
		return -2;
	return 0;
}

/* rtl/quadtbl.v:143-291 */
int32_t zo_quadtbl1(const zo_quadtbl *q, uint32_t phase) {
	const int PW = q->pw, DX = q->dxbits, QB = q->qbits, LB = q->lbits, CB = q->cbits;
	const int WW = q->ww, OW = q->ow, XTRA = q->nextra;
	uint64_t ip = ux(phase, PW);
	uint32_t idx = (uint32_t)(ip >> (DX - 1));			/* i_phase[(PW-1):(DXBITS-1)] */
	int64_t qv = sx(q->qtbl[idx], QB), lv = sx(q->ltbl[idx], LB), cv = sx(q->ctbl[idx], CB);
	int64_t dx = (int64_t)(ip & ((1ull << (DX - 1)) - 1ull));	/* { 1'b0, i_phase[(DXBITS-2):0] } */
	int64_t qprod = sx(qv * dx, QB + DX);				/* :173 */
	/* :196-199: w_qprod = { sign..., qprod[(QBITS+DXBITS-1):(DXBITS-1)] } -- an arithmetic shift */
	int64_t w_qprod = sx(qprod >> (DX - 1), LB);
	int64_t lsum = sx(w_qprod + lv, LB);				/* :205 */
	int64_t lprod = sx(lsum * dx, LB + DX);				/* :231 */
	int64_t w_lprod = sx(lprod >> (DX - 1), CB);			/* :247-248 */
	int64_t r = sx(w_lprod + cv, CB);				/* :254 */
	uint64_t rb = ux((uint64_t)r, CB);
	uint64_t w;
	/* :262-271: do not round when that would overflow the WW-bit word */
	uint64_t mid_hi = (rb >> XTRA) & ((1ull << (WW - 1 - XTRA)) - 1ull);	/* r[(WW-2):XTRA] */
	uint64_t mid_lo = (WW - 2 - XTRA > 0) ? ((rb >> XTRA) & ((1ull << (WW - 2 - XTRA)) - 1ull)) : 0;	/* r[(WW-3):XTRA] */
	int top = (int)((rb >> (WW - 1)) & 1), top2 = (int)((rb >> (WW - 2)) & 3);
	if (!top && mid_hi == ((1ull << (WW - 1 - XTRA)) - 1ull))
		w = rb;
	else if (top2 == 3 && mid_lo == 0)
		w = rb;
	else {
		int D = WW - OW;
		int b = (int)((rb >> D) & 1);
		uint64_t add = b ? (1ull << (D - 1)) : ((1ull << (D - 1)) - 1ull);
		w = rb + add;
	}
	w = ux(w, WW);
	return (int32_t)sx((int64_t)(w >> XTRA), OW);			/* o_sin <= w_value[(WW-1):XTRA] */
}

/* ---- batched forms ----------------------------------------------------- */

typedef struct job {
	int kind;
	const zo_params *p;
	int32_t x0, y0;
	const int32_t *xy_in;
	const uint32_t *phase_in;
	int32_t *out0;
	uint32_t *out1;
	uint32_t phase0, step;
	uint64_t n0;
	int pw, ow;
	const uint32_t *tbl;
	const zo_quadtbl *qt;
	size_t lo, hi;
} job;

enum { K_ROTC, K_ROT, K_TOPOLAR, K_NCO, K_SIN, K_QWAV, K_QTBL };

static void run_range(const job *j) {
	size_t i;
	switch (j->kind) {
	case K_ROTC:
		for (i = j->lo; i < j->hi; i++)
			zo_rotate1(j->p, j->x0, j->y0, j->phase_in[i],
				&j->out0[2 * i], &j->out0[2 * i + 1]);
		break;
	case K_ROT:
		for (i = j->lo; i < j->hi; i++)
			zo_rotate1(j->p, j->xy_in[2 * i], j->xy_in[2 * i + 1], j->phase_in[i],
				&j->out0[2 * i], &j->out0[2 * i + 1]);
		break;
	case K_TOPOLAR:
		for (i = j->lo; i < j->hi; i++)
			zo_topolar1(j->p, j->xy_in[2 * i], j->xy_in[2 * i + 1],
				&j->out0[i], &j->out1[i]);
		break;
	case K_NCO:
		for (i = j->lo; i < j->hi; i++) {
			uint32_t ph32 = j->phase0 + (uint32_t)(j->n0 + i) * j->step;
			zo_rotate1(j->p, j->x0, j->y0, ph32 >> (32 - j->p->pw),
				&j->out0[2 * i], &j->out0[2 * i + 1]);
		}
		break;
	case K_SIN:
		for (i = j->lo; i < j->hi; i++)
			j->out0[i] = zo_sintable_lookup1(j->pw, j->ow, j->tbl, j->phase_in[i]);
		break;
	case K_QWAV:
		for (i = j->lo; i < j->hi; i++)
			j->out0[i] = zo_quarterwav_lookup1(j->pw, j->ow, j->tbl, j->phase_in[i]);
		break;
	case K_QTBL:
		for (i = j->lo; i < j->hi; i++)
			j->out0[i] = zo_quadtbl1(j->qt, j->phase_in[i]);
		break;
	}
}

static void *thread_main(void *arg) {
	run_range((const job *)arg);
	return NULL;
}

static void run_parallel(job *proto, size_t n, int nthreads) {
	if (nthreads <= 1 || n < 4096) {
		proto->lo = 0; proto->hi = n;
		run_range(proto);
		return;
	}
	if (nthreads > 256)
		nthreads = 256;
	pthread_t tid[256];
	job jobs[256];
	size_t per = (n + (size_t)nthreads - 1) / (size_t)nthreads;
	int started = 0;
	for (int t = 0; t < nthreads; t++) {
		size_t lo = per * (size_t)t, hi = lo + per;
		if (lo >= n) break;
		if (hi > n) hi = n;
		jobs[t] = *proto;
		jobs[t].lo = lo; jobs[t].hi = hi;
		if (pthread_create(&tid[t], NULL, thread_main, &jobs[t]) != 0) {
			run_range(&jobs[t]);	/* degrade to inline */
			tid[t] = 0;
		}
		started = t + 1;
	}
	for (int t = 0; t < started; t++)
		if (tid[t])
			pthread_join(tid[t], NULL);
}

void zo_rotate_const(const zo_params *p, int32_t x0, int32_t y0,
		const uint32_t *phase, int32_t *xy, size_t n, int nthreads) {
	job j; memset(&j, 0, sizeof(j));
	j.kind = K_ROTC; j.p = p; j.x0 = x0; j.y0 = y0; j.phase_in = phase; j.out0 = xy;
	run_parallel(&j, n, nthreads);
}

void zo_rotate(const zo_params *p, const int32_t *xy_in, const uint32_t *phase,
		int32_t *xy_out, size_t n, int nthreads) {
	job j; memset(&j, 0, sizeof(j));
	j.kind = K_ROT; j.p = p; j.xy_in = xy_in; j.phase_in = phase; j.out0 = xy_out;
	run_parallel(&j, n, nthreads);
}

void zo_topolar(const zo_params *p, const int32_t *xy_in, int32_t *mag,
		uint32_t *phase, size_t n, int nthreads) {
	job j; memset(&j, 0, sizeof(j));
	j.kind = K_TOPOLAR; j.p = p; j.xy_in = xy_in; j.out0 = mag; j.out1 = phase;
	run_parallel(&j, n, nthreads);
}

void zo_nco_rotate(const zo_params *p, int32_t x0, int32_t y0, uint32_t phase0,
		uint32_t step, uint64_t n0, int32_t *xy, size_t n, int nthreads) {
	job j; memset(&j, 0, sizeof(j));
	j.kind = K_NCO; j.p = p; j.x0 = x0; j.y0 = y0; j.phase0 = phase0; j.step = step;
	j.n0 = n0; j.out0 = xy;
	run_parallel(&j, n, nthreads);
}

void zo_lut_sin(int pw, int ow, const uint32_t *tbl, const uint32_t *phase32,
		int32_t *out, size_t n, int nthreads) {
	job j; memset(&j, 0, sizeof(j));
	j.kind = K_SIN; j.pw = pw; j.ow = ow; j.tbl = tbl; j.phase_in = phase32; j.out0 = out;
	run_parallel(&j, n, nthreads);
}

void zo_lut_qwav(int pw, int ow, const uint32_t *tbl, const uint32_t *phase32,
		int32_t *out, size_t n, int nthreads) {
	job j; memset(&j, 0, sizeof(j));
	j.kind = K_QWAV; j.pw = pw; j.ow = ow; j.tbl = tbl; j.phase_in = phase32; j.out0 = out;
	run_parallel(&j, n, nthreads);
}

void zo_quadtbl_batch(const zo_quadtbl *q, const uint32_t *phase, int32_t *out, size_t n, int nthreads) {
	job j; memset(&j, 0, sizeof(j));
	j.kind = K_QTBL; j.qt = q; j.phase_in = phase; j.out0 = out;
	run_parallel(&j, n, nthreads);
}

/* ---- $readmemh fixture loader (format of sw/hexfile.cpp:78-89) ---------- */

long zo_hex_load(const char *fname, uint32_t *words, long maxwords) {
	FILE *fp = fopen(fname, "r");
	if (!fp)
		return -1;
	long addr = 0, count = 0;
	char tok[64];
	while (fscanf(fp, "%63s", tok) == 1) {
		if (tok[0] == '@') {
			addr = strtol(tok + 1, NULL, 16);
			continue;
		}
		if (addr >= maxwords) {
			fclose(fp);
			return -2;
		}
		words[addr++] = (uint32_t)strtoul(tok, NULL, 16);
		if (addr > count)
			count = addr;
	}
	fclose(fp);
	return count;
}
