// zc_params.cpp -- host-side configuration of the zcordic engine: the same parameter
// surface as the ZipCPU/cordic core generator (phase width, output width, stage count,
// extra bits, CORDIC gain), derived with the same double-precision libm calls so every
// integer constant matches what the generator prints.  Citations: /root/reference.
#include "zc_internal.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace zc {

// sw/cordiclib.cpp:157-169.  The double is truncated (not rounded) to an unsigned.
static uint32_t angle_word(int k, int pw) {
	const double turns_to_units = (4.0 * (double)(1ul << (pw - 2))) / (M_PI * 2.0);
	double a = std::atan2(1., std::pow(2, k + 1)) * turns_to_units;
	return (uint32_t)(unsigned)a;
}

// sw/cordiclib.cpp:66-80
static double cordic_gain(int nstages) {
	double g = 1.0;
	for (int k = 0; k < nstages; k++)
		g = g * std::sqrt(1.0 + std::pow(2.0, -2. * (k + 1)));
	return g;
}

// sw/cordiclib.cpp:82-109
static double phase_variance(int nstages, int pw) {
	const double rad_to_phase = (double)(1ul << (pw - 1)) / M_PI;
	double var = 1. / 12.;
	for (unsigned k = 0; k < (unsigned)nstages; k++) {
		double x = std::atan2(1., std::pow(2, k + 1)) * rad_to_phase;
		unsigned long q = (unsigned)x;
		double e = q - x;
		var += e * e;
	}
	return var / std::pow(rad_to_phase, 2.);
}

// sw/cordiclib.cpp:111-130
static double quantization_variance(int nstages, int xtrabits, int dropped) {
	double v = std::pow(2, 2 * xtrabits) / 12.;
	for (int k = 0; k < nstages; k++)
		v = (1 + std::pow(4, -k - 1)) * v + 1. / 3.;
	if (dropped > 0)
		v = std::pow(2, -2 * dropped) * v + 1 / 12.;
	return v;
}

// sw/cordiclib.cpp:246-268: smallest pb>=3 with sin(2pi/2^pb)*(2^ow - 1) < 1/2
static int calc_phase_bits(int ow) {
	unsigned pb = 3;
	for (; pb < 64; pb++) {
		double step = 2.0 * M_PI / (double)(1ul << pb);
		if (std::sin(step) * (double)((1ul << ow) - 1) < 0.5)
			break;
	}
	return (int)pb;
}

// sw/cordiclib.cpp:214-229 (bounded by the working width) and :231-244 (unbounded)
static int calc_stages(int pw, int ww_bound /* <0: none */) {
	int n = 0;
	for (; n < 64; n++) {
		if (angle_word(n, pw) == 0)
			break;
		if (ww_bound >= 0 && ww_bound <= n)
			break;
	}
	return n;
}

static void resolve_widths(int &iw, int &ow) {
	// sw/main.cpp:262-270 / :314-322
	if (iw <= 0 && ow > 0) iw = ow;
	if (ow <= 0) ow = iw;
	if (iw <= 0 || ow <= 0) iw = ow = 24;
}

static int finish(zc_params *o, int mode, int iw, int ow, int nxtra, int ww, int pw, int nstages) {
	if (pw < 3 || pw > 32 || ww > 32 || ww < 2 || iw < 1 || ow < 1 ||
	    nstages < 0 || nstages > ZC_MAX_STAGES || ow > ww || iw > ww)
		return set_error(ZC_ERANGE, "configuration needs PW in [3,32], WW<=32, NSTAGES<=64 "
			"(got IW=%d OW=%d WW=%d PW=%d NSTAGES=%d)", iw, ow, ww, pw, nstages);
	std::memset(o, 0, sizeof(*o));
	o->mode = mode;
	o->iw = iw; o->ow = ow; o->nextra = nxtra; o->ww = ww; o->pw = pw; o->nstages = nstages;
	for (int k = 0; k < nstages; k++)
		o->angle[k] = angle_word(k, pw);
	o->cordic_gain = cordic_gain(nstages);
	o->qvar = quantization_variance(nstages, ww - iw, ww - ow);
	o->pvar_rad = phase_variance(nstages, pw);
	return ZC_OK;
}

int derive_p2r(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *o) {
	if (!o) return set_error(ZC_EINVAL, "NULL zc_params");
	resolve_widths(iw, ow);
	const int wide = (ow > iw) ? ow : iw;
	int nxtra = xtra_user + 1;				// sw/main.cpp:273
	const int ww_cli = wide + nxtra;			// sw/main.cpp:272-274
	if (ww_cli < 1 || ww_cli > 62)
		return set_error(ZC_ERANGE, "working width %d out of range", ww_cli);
	if (pw <= 0) pw = calc_phase_bits(ww_cli);		// :276
	if (pw < 3 || pw > 32)
		return set_error(ZC_ERANGE, "phase width %d outside [3,32]", pw);
	if (nstages <= 0) nstages = calc_stages(pw, ww_cli);	// :278
	if (nxtra < 1) nxtra = 1;				// sw/basiccordic.cpp:67-68
	const int ww = wide + nxtra;				// sw/basiccordic.cpp:71-73
	int rc = finish(o, ZC_MODE_P2R, iw, ow, nxtra, ww, pw, nstages);
	if (rc != ZC_OK) return rc;
	o->gain = o->cordic_gain;				// sw/basiccordic.cpp:477-478
	// sw/basiccordic.cpp:479-496 (including the pow(2, gain) factor as written there)
	double amp = (double)(1ul << (iw - 1)) - 1.;
	amp *= (double)(1ul << (ww - iw));
	amp *= o->cordic_gain;
	amp *= std::pow(2.0, -(ww - ow));
	const double sig = amp * amp;
	const double noise = o->qvar + sig * o->pvar_rad * std::pow(2, o->cordic_gain);
	o->best_cnr = 10.0 * std::log(sig / noise) / std::log(10.0);
	return ZC_OK;
}

int derive_r2p(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *o) {
	if (!o) return set_error(ZC_EINVAL, "NULL zc_params");
	resolve_widths(iw, ow);
	const int wide = (ow > iw) ? ow : iw;
	int nxtra = xtra_user + 2;				// sw/main.cpp:323
	const int ww_cli = wide + nxtra;
	if (ww_cli < 1 || ww_cli > 62)
		return set_error(ZC_ERANGE, "working width %d out of range", ww_cli);
	if (pw <= 0) pw = calc_phase_bits(ww_cli);		// :325-326
	if (pw < 3 || pw > 32)
		return set_error(ZC_ERANGE, "phase width %d outside [3,32]", pw);
	if (nstages <= 0) nstages = calc_stages(pw, -1);	// :327-328
	if (nxtra < 2) nxtra = 2;				// sw/topolar.cpp:67-68
	const int ww = wide + 2 * nxtra;			// sw/topolar.cpp:71-75 (added twice)
	int rc = finish(o, ZC_MODE_R2P, iw, ow, nxtra, ww, pw, nstages);
	if (rc != ZC_OK) return rc;
	o->gain = o->cordic_gain * std::sqrt(2.0) / 2.;		// sw/topolar.cpp:439-440
	o->best_cnr = 0.0;
	return ZC_OK;
}

// sw/main.cpp:358-379 (tbl) / :401-422 (qtr).  In main() "not given" is -1; here <=0.
// -t sp2r / -t sr2p (sw/main.cpp:183-198): the constants are those of the pipelined core (same branch of main(),
// sw/seqcordic.cpp:455-498 and sw/seqpolar.cpp:393-415 print the same set plus CLOCKS_PER_OUTPUT); what changes is
// the schedule the state machine runs (include/zcordic.h, zc_params.seq).
int derive_sp2r(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *o) {
	int rc = derive_p2r(iw, ow, xtra_user, pw, nstages, o);
	if (rc != ZC_OK) return rc;
	if (o->nstages < 3)
		return set_error(ZC_ERANGE, "sequential p2r needs NSTAGES >= 3 (got %d): its output is taken two iterations early",
			o->nstages);
	o->seq = 1;
	return ZC_OK;
}

int derive_sr2p(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *o) {
	int rc = derive_r2p(iw, ow, xtra_user, pw, nstages, o);
	if (rc != ZC_OK) return rc;
	if (o->nstages < 1 || ((o->nstages + 1) & o->nstages) == 0)
		return set_error(ZC_ERANGE, "sequential r2p with NSTAGES=%d: the reference's state register "
			"(nextlg(NSTAGES+1) bits, sw/seqpolar.cpp:158-159) cannot reach NSTAGES+1, o_done never rises", o->nstages);
	o->seq = 1;
	return ZC_OK;
}

int derive_lut(bool quarter, int iw, int pw, int ow, int *pw_out, int *ow_out) {
	if (!pw_out || !ow_out) return set_error(ZC_EINVAL, "NULL output");
	if (iw <= 0) iw = -1;
	if (pw <= 0) pw = -1;
	if (ow <= 0) ow = -1;
	if (iw >= 0 && pw < 0) { pw = iw; iw = -1; }
	if (pw > 3 && ow <= 0) {
		for (int k = pw - 2; k < pw + 3; k++)
			if (k > 0 && k < 63 && calc_phase_bits(k) == pw) { ow = k; break; }
	}
	if (ow <= 0) ow = 24;
	if (pw <= 0) pw = calc_phase_bits(ow);
	*pw_out = pw; *ow_out = ow;
	return check_lut(quarter, pw, ow);
}

int check_lut(bool quarter, int pw, int ow) {
	// sw/sintable.cpp:62-69 (tbl refuses >=24), :188-197 (qtr asserts >2, refuses >=26),
	// sw/hexfile.cpp:52-55 (ow < 31)
	if (ow < 2 || ow >= 31)
		return set_error(ZC_ERANGE, "LUT output width %d outside [2,30]", ow);
	if (quarter ? (pw <= 2 || pw >= 26) : (pw < 1 || pw >= 24))
		return set_error(ZC_ERANGE, "LUT phase width %d outside the generator's limits", pw);
	return ZC_OK;
}

// sw/sintable.cpp:156-168: tbl[k] = (long)(maxv*sin(2 pi k / 2^pw)), C truncation, & (2^ow - 1)
int build_sintable(int pw, int ow, uint32_t *tbl) {
	if (!tbl) return set_error(ZC_EINVAL, "NULL table");
	int rc = check_lut(false, pw, ow);
	if (rc != ZC_OK) return rc;
	const long entries = 1l << pw;
	const long maxv = (1l << (ow - 1)) - 1l, mask = (1l << ow) - 1l;
	for (long k = 0; k < entries; k++) {
		double ph = 2.0 * M_PI * (double)k / (double)entries;
		long w = (long)(maxv * std::sin(ph));
		tbl[k] = (uint32_t)(w & mask);
	}
	return ZC_OK;
}

// sw/sintable.cpp:325-337: half-sample offset, first quadrant only
int build_quarterwav(int pw, int ow, uint32_t *tbl) {
	if (!tbl) return set_error(ZC_EINVAL, "NULL table");
	int rc = check_lut(true, pw, ow);
	if (rc != ZC_OK) return rc;
	const long entries = 1l << pw;
	const long maxv = (1l << (ow - 1)) - 1l, mask = (1l << ow) - 1l;
	for (long k = 0; k < entries / 4; k++) {
		double ph = 2.0 * M_PI * (double)k / (double)entries;
		ph += M_PI / (double)entries;
		long w = (long)(maxv * std::sin(ph));
		tbl[k] = (uint32_t)(w & mask);
	}
	return ZC_OK;
}

// ---- quadratically interpolated table: gencordic -t qtbl ------------------------------------------
namespace {

struct QuadFit {			// one pass of sw/quadtbl.cpp:136-266 for a table of 2^lg entries
	std::vector<double> c, l, q;	// constant, linear, quadratic coefficient per entry, scaled to |c| <= 1
	double worst;			// largest signed fit error over the table (sw/quadtbl.cpp:72-114)
};

double sinc_pi(double v) { const double x = v * M_PI; return std::sin(x) / x; }	// sw/quadtbl.cpp:54-57

// Largest-magnitude error of c + (l + q t) t against the sine over one table step, sampled as the
// generator samples it: both ends and 64 interior points (sw/quadtbl.cpp:72-114).
double fit_error(double c, double l, double q, double idx, int n) {
	double ang = 2.0 * M_PI * idx / (double)n;
	const double at_left = c - std::sin(ang);
	ang = 2.0 * M_PI * (idx + 1) / (double)n;
	const double at_right = c + l + q - std::sin(ang);
	double inside = 0;
	for (int k = 0; k < 64; k++) {
		const double t = k / 64.0;
		const double e = c + (l + q * t) * t - std::sin(2.0 * M_PI * (idx + t) / n);
		if (std::fabs(e) > std::fabs(inside)) inside = e;
	}
	double worst = at_left;
	if (std::fabs(worst) < std::fabs(at_right)) worst = at_right;
	if (std::fabs(worst) < std::fabs(inside)) worst = inside;
	return worst;
}

QuadFit quad_fit(int lg) {
	const int n = 1 << lg;
	const double half_step = M_PI / (double)n, step = half_step * 2.;
	QuadFit f;
	f.c.resize(n); f.l.resize(n); f.q.resize(n);
	std::vector<double> &c = f.c, &l = f.l, &q = f.q;
	for (int i = 0; i < n; i++) c[i] = std::sin(step * i + half_step);		// mid-interval samples
	for (int i = 1; i < n - 1; i++) l[i] = (c[i + 1] - c[i - 1]) / 2.0;		// central difference
	l[0] = (c[1] - c[n - 1]) / 2.0;
	l[n - 1] = (c[0] - c[n - 2]) / 2.0;
	for (int i = 1; i < n - 1; i++) q[i] = -(c[i] - 0.5 * (c[i + 1] + c[i - 1]));	// second difference
	q[0] = -(c[0] - 0.5 * (c[1] + c[n - 1]));
	q[n - 1] = -(c[n - 1] - 0.5 * (c[0] + c[n - 2]));
	for (int i = 0; i < n; i++)							// the quadratic's own smoothing
		c[i] = 0.75 * std::sin(step * i + half_step)
			+ (std::sin(step * (i - 1) + half_step) + std::sin(step * (i + 1) + half_step)) / 8.0;
	const double del = 1.0, hdel = del / 2.0;					// re-centre on the interval's left edge
	for (int i = 0; i < n; i++) c[i] = q[i] * hdel * hdel - l[i] * hdel + c[i];
	for (int i = 0; i < n; i++) l[i] = l[i] - del * q[i];
	const double gain = std::pow(1. / sinc_pi(half_step), 3);			// undo the interpolator's droop
	for (int i = 0; i < n; i++) c[i] *= gain;
	for (int i = 0; i < n; i++) l[i] *= gain;
	for (int i = 0; i < n; i++) q[i] *= gain;
	double peak = 0.0;
	for (int i = 0; i < n; i++) peak = (peak > std::fabs(c[i])) ? peak : std::fabs(c[i]);
	for (int i = 0; i < n; i++) c[i] *= 1. / peak;
	for (int i = 0; i < n; i++) l[i] *= 1. / peak;
	for (int i = 0; i < n; i++) q[i] *= 1. / peak;
	f.worst = 0.0;
	for (int i = 0; i < n; i++) {
		const double e = fit_error(c[i], l[i], q[i], i, n);
		if (std::fabs(e) > std::fabs(f.worst)) f.worst = e;
	}
	return f;
}

double peak_of(const std::vector<double> &v) {
	double m = 0.0;
	for (double x : v) m = (m > std::fabs(x)) ? m : std::fabs(x);
	return m;
}

} // namespace

int derive_qtbl(int iw, int ow, int xtra_user, int pw, zc_quadtbl *o) {
	if (!o) return set_error(ZC_EINVAL, "NULL zc_quadtbl");
	resolve_widths(iw, ow);						// sw/main.cpp:446-454
	const int wide = (ow > iw) ? ow : iw;
	int nxtra = xtra_user + 1;					// :456
	if (pw <= 0) {
		const int ww_cli = wide + nxtra;
		if (ww_cli < 1 || ww_cli > 62) return set_error(ZC_ERANGE, "working width %d out of range", ww_cli);
		pw = calc_phase_bits(ww_cli);				// :458-459
	}
	const int wid = ow + nxtra;					// the width the tables are built for
	if (nxtra < 0 || pw <= 4 || pw > 32 || wid <= 6 || wid > 30)
		return set_error(ZC_ERANGE, "quadtbl needs XTRA>=0, 4<PW<=32, 6<OW+XTRA<=30 (got OW=%d XTRA=%d PW=%d)", ow, nxtra, pw);
	const long fullscale = (1l << (wid - 1)) - 2l;			// max_integer(), sw/quadtbl.cpp:59-61
	// sw/quadtbl.cpp:295-301: double the table until the fit error is under one output unit
	int lg = 3;
	QuadFit fit;
	double tblerr;
	do {
		lg++;
		if (lg > ZC_QT_MAXLG) return set_error(ZC_ERANGE, "quadtbl would need more than 2^%d entries", ZC_QT_MAXLG);
		fit = quad_fit(lg);
		tblerr = fit.worst * fullscale;
	} while (std::fabs(tblerr) > 1.0 && lg < 20);
	if (pw <= lg) return set_error(ZC_ERANGE, "PW=%d too small for a 2^%d-entry table", pw, lg);
	std::memset(o, 0, sizeof(*o));
	// coefficient widths: sw/quadtbl.cpp:236-238
	o->cbits = wid + (int)std::ceil(std::log(peak_of(fit.c)) / std::log(2.0));
	o->lbits = wid + (int)std::ceil(-std::log(1. / peak_of(fit.l)) / std::log(2.0));
	o->qbits = wid + (int)std::ceil(-std::log(1. / peak_of(fit.q)) / std::log(2.0));
	if (nxtra < 2) nxtra = 2;					// sw/quadtbl.cpp:315-316 (after the tables!)
	o->ow = ow; o->nextra = nxtra; o->pw = pw; o->ww = ow + nxtra;
	o->lgtbl = lg; o->dxbits = pw - lg + 1;
	o->scale = (1l << (ow - 1)) - 2l;				// :789-790
	o->itbl_err = tblerr;
	o->tbl_err = tblerr * std::pow(0.5, wid);			// :793-795
	o->spurdb = 20. * std::log(std::pow(sinc_pi(1.0 - (1. / (1 << lg))), 3.)) / std::log(10.0);	// :797-799
	if (o->cbits < 2 || o->cbits > 30 || o->lbits < 2 || o->qbits < 2)
		return set_error(ZC_ERANGE, "quadtbl coefficient widths out of range");
	const long cm = (1l << o->cbits) - 1l, lm = (1l << o->lbits) - 1l, qm = (1l << o->qbits) - 1l;
	for (int k = 0; k < (1 << lg); k++) {				// :250-266 + sw/hexfile.cpp:78-89
		o->ctbl[k] = (uint32_t)((long)(fullscale * fit.c[k]) & cm);
		o->ltbl[k] = (uint32_t)((long)(fullscale * fit.l[k]) & lm);
		o->qtbl[k] = (uint32_t)((long)(fullscale * fit.q[k]) & qm);
	}
	return check_qtbl(o);
}

// What the kernel implements: CBITS == WW (true whenever -x >= 1), nested coefficient widths, and
// products that fit 62 bits.
int check_qtbl(const zc_quadtbl *q) {
	if (!q) return set_error(ZC_EINVAL, "NULL zc_quadtbl");
	if (q->lgtbl < 1 || q->lgtbl > ZC_QT_MAXLG || q->pw <= q->lgtbl || q->pw > 32 || q->dxbits != q->pw - q->lgtbl + 1 ||
	    q->dxbits < 2 || q->ow < 2 || q->nextra < 2 || q->ww != q->ow + q->nextra || q->cbits != q->ww || q->cbits > 30 ||
	    q->lbits < q->qbits + 1 || q->cbits < q->lbits + 1 || q->qbits < 2 || q->lbits + q->dxbits > 62)
		return set_error(ZC_ERANGE, "unsupported quadtbl geometry OW=%d XTRA=%d PW=%d LGTBL=%d CBITS=%d LBITS=%d QBITS=%d",
			q->ow, q->nextra, q->pw, q->lgtbl, q->cbits, q->lbits, q->qbits);
	return ZC_OK;
}

} // namespace zc
