/*
 * zcordic.h -- C ABI of libzcordic: a Blackwell (sm_100a) batched CORDIC engine that
 * evaluates, bit-exactly, the functions computed by the cores ZipCPU/cordic generates.
 *
 * The reference has no FFI of its own.  Its de-facto boundaries are (1) the port list
 * of the generated cores as consumed through the Verilated model by TESTB<VA>
 * (bench/cpp/testb.h:49-136) and (2) the generator's command line / generated header
 * constants (sw/main.cpp:139-232, rtl/cordic.h:46-59).  Each entry point below cites
 * the reference interface it replaces; citations are relative to the reference tree.
 *
 * Conventions
 *   - plain C, caller-owned buffers, no torch / C++ types in any signature;
 *   - every function returns ZC_OK (0) or a negative zc_status; nothing throws;
 *   - a zc_params is a POD filled by zc_derive_*; it is immutable afterwards and may be
 *     used from any thread and any device;
 *   - "device" entry points take DEVICE pointers valid on CUDA device `device` and a
 *     cudaStream_t passed as void* (NULL = the legacy default stream); they enqueue work
 *     and return without synchronising;
 *   - "_host" entry points take HOST pointers (pinned memory from zc_host_alloc gives full
 *     PCIe overlap; pageable memory works, slower), run a chunked
 *     H2D -> kernel -> D2H pipeline on `device`, and return when the outputs are complete;
 *   - sample words: ix/iy/phase inputs are the raw port bit-vectors in the low IW / PW
 *     bits of a 32-bit word (higher bits are ignored, exactly as the port would truncate
 *     them); o_xval / o_yval / o_mag / o_val are returned SIGN-EXTENDED from OW bits to
 *     int32; o_phase is returned zero-extended in PW bits;
 *   - there is NO CPU fallback: without a usable CUDA device the compute entry points
 *     fail with ZC_ENODEV / ZC_ECUDA.
 */
#ifndef ZCORDIC_H
#define ZCORDIC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZC_VERSION_MAJOR 0
#define ZC_VERSION_MINOR 1
#define ZC_MAX_STAGES    64

typedef enum zc_status {
	ZC_OK       =  0,
	ZC_EINVAL   = -1,	/* NULL pointer, bad size, bad argument                         */
	ZC_ERANGE   = -2,	/* configuration outside what the engine (or the reference) supports */
	ZC_ECUDA    = -3,	/* a CUDA runtime call failed; see zc_last_error()              */
	ZC_ENODEV   = -4,	/* no CUDA device / bad device ordinal                          */
	ZC_ENOMEM   = -5
} zc_status;

enum { ZC_MODE_P2R = 0, ZC_MODE_R2P = 1 };

/* The generated-header constant set (rtl/cordic.h:46-59, rtl/topolar.h:46-58) plus the
 * cordic_angle table printed into the Verilog (sw/cordiclib.cpp:157-169). */
typedef struct zc_params {
	int32_t	 mode;			/* ZC_MODE_P2R (rtl/cordic.v) or ZC_MODE_R2P (rtl/topolar.v) */
	int32_t	 iw, ow;		/* IW, OW                                          */
	int32_t	 nextra;		/* NEXTRA as printed (already incremented)          */
	int32_t	 ww, pw, nstages;	/* WW, PW, NSTAGES                                  */
	int32_t	 seq;			/* 0: pipelined core; 1: sequential core (zc_derive_sp2r/_sr2p), see below */
	uint32_t angle[ZC_MAX_STAGES];	/* cordic_angle[k], PW-bit, truncated               */
	double	 gain;			/* GAIN                                             */
	double	 cordic_gain;		/* prod sqrt(1+2^-2(k+1)) (== GAIN for p2r)         */
	double	 qvar;			/* QUANTIZATION_VARIANCE                            */
	double	 pvar_rad;		/* PHASE_VARIANCE_RAD                               */
	double	 best_cnr;		/* BEST_POSSIBLE_CNR (p2r only, else 0)             */
} zc_params;

/* The quadratically interpolated sine table of `gencordic -t qtbl` (rtl/quadtbl.v, rtl/quadtbl.h):
 * its localparams (rtl/quadtbl.v:49-51,65-71), the header constants (rtl/quadtbl.h) and the three
 * $readmemh coefficient tables (rtl/quadtbl_{c,l,q}tbl.hex), words masked to CBITS/LBITS/QBITS. */
#define ZC_QT_MAXLG 12
typedef struct zc_quadtbl {
	int32_t	 ow, nextra;		/* OW, NEXTRA (= XTRA)                               */
	int32_t	 pw, ww;		/* PW, WW = OW+XTRA                                  */
	int32_t	 lgtbl, dxbits;		/* LGTBL, DXBITS = PW-LGTBL+1                        */
	int32_t	 cbits, lbits, qbits;	/* CBITS, LBITS, QBITS                               */
	int32_t	 reserved;
	int64_t	 scale;			/* SCALE                                             */
	double	 itbl_err, tbl_err;	/* ITBL_ERR, TBL_ERR                                 */
	double	 spurdb;		/* SPURDB                                            */
	uint32_t ctbl[1 << ZC_QT_MAXLG], ltbl[1 << ZC_QT_MAXLG], qtbl[1 << ZC_QT_MAXLG];
} zc_quadtbl;

/* flags for the *_ex entry points */
enum {
	ZC_F_DEFAULT       = 0,
	ZC_F_FORCE_GENERIC = 1,	/* runtime-parameter kernel that models the WW-bit wrap       */
	ZC_F_NO_SEED       = 2,	/* rotate_const/nco: run every stage in registers (no table-seeded prefix) */
	ZC_F_FORCE_SEED    = 4,	/* rotate_const/nco: use the table-seeded prefix even for small n */
	ZC_F_SEED_PACKED   = 8,	/* seeded kernel: suffix directions as one byte per stage (less shared-memory traffic: the better
				   choice when neighbouring samples have scattered phases) */
	ZC_F_SEED_REGS     = 16,	/* seeded kernel: run the suffix phase recursion in registers (no direction table) */
	ZC_F_NO_DP2A       = 128,	/* seeded kernel, word table: multiply-adds as IMAD with an explicit negation instead of IDP.2A */
	ZC_F_NO_TAIL       = 64,	/* topolar: every stage in its full form (by default the late stages, where y has provably
				   converged below the shift, run a shorter instruction sequence with identical results) */
	ZC_F_NO_MERGE      = 512,	/* seeded kernel, byte table: keep the interval's row offset in its own table (by default it rides in
				   the low bytes of the (x, y) records when WW <= 24: three dependent shared-memory lookups per
				   sample instead of four; same results) */
	ZC_F_NO_COMB       = 256,	/* NCO with a scattering step: keep the block sample mapping + byte table (by default the
				   engine looks for a comb mapping -- runs K samples apart with K*step ~ 0 mod 2^32 share a
				   quarter-warp -- under which the lanes of a quarter-warp share word-table rows; same results) */
	ZC_F_SEED_WORDS    = 32	/* seeded kernel: suffix directions as one word per stage (fastest for phase sweeps, slow NCOs).
				   With none of the three: NCO picks by step size on the host; phase streams of >= 4 Mi
				   samples are probed on the device: both flavours are enqueued, every CTA of both evaluates
				   the same pure function of the phase stream and exactly one flavour proceeds (one skipped
				   launch per call; no device-side state is shared between calls, streams or graph replays). */
};

int         zc_version(void);				/* major*1000 + minor */
const char *zc_strerror(int status);
const char *zc_last_error(void);			/* thread-local detail of the last failure */
int         zc_device_count(void);			/* >=0, or ZC_ECUDA */

/* ---- configuration: replaces the generator command line ------------------------------ */

/* gencordic -t p2r -i iw -o ow -x xtra_user [-p pw] [-n nstages]
 *   sw/main.cpp:260-279 (width defaulting, nxtra+1, phase bits, stages),
 *   sw/basiccordic.cpp:67-73 (working width), :465-498 (header constants).
 * Pass <=0 for iw / ow / pw / nstages that were not given on the command line. */
int zc_derive_p2r(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *out);
/* gencordic -t r2p ...   sw/main.cpp:312-328, sw/topolar.cpp:67-75 (nxtra added twice), :428-446 */
int zc_derive_r2p(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *out);
/* gencordic -t sp2r / -t sr2p ...  the sequential cores rtl/seqcordic.v, rtl/seqpolar.v (sw/main.cpp:183-198,
 * sw/seqcordic.cpp, sw/seqpolar.cpp): one sample every CLOCKS_PER_OUTPUT clocks behind an i_stb/o_busy/o_done
 * handshake (rtl/seqcordic.v:63-70).  The constants are those of the pipelined core; zc_params.seq = 1 selects the
 * schedule the state machine really runs, which is NOT that of the pipelined core:
 *   - every iteration executes, zero cordic_angle or shift >= WW included (no pass-through test;
 *     rtl/seqcordic.v:281-299, rtl/seqpolar.v),
 *   - seqcordic registers o_xval/o_yval when state == NSTAGES-1 (rtl/seqcordic.v:319-324), i.e. after NSTAGES-2
 *     iterations; seqpolar's last_state is state >= NSTAGES+1, i.e. NSTAGES iterations.
 * A batch call with such a zc_params returns, per sample, exactly what the sequential core would present with
 * o_done.  zc_derive_sr2p returns ZC_ERANGE when NSTAGES+1 is a power of two: the reference's state register
 * (sw/seqpolar.cpp:158-159) cannot count that far and the core never finishes. */
int zc_derive_sp2r(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *out);
int zc_derive_sr2p(int iw, int ow, int xtra_user, int pw, int nstages, zc_params *out);
/* Stage updates that reach the output (NSTAGES; NSTAGES-2 for sp2r) and CLOCKS_PER_OUTPUT as printed in
 * rtl/seqcordic.h:49 / rtl/seqpolar.h:49 (1 for the pipelined cores). */
int zc_iterations(const zc_params *p);
int zc_clocks_per_output(const zc_params *p);
/* Diagnostic: how many of the last vectoring stages zc_topolar runs in the short form (0 for rotation cores).  From
 * stage i on the engine has proved -2^(i+1) <= y < 2^(i+1) for every input, so rtl/topolar.v:227-243's
 * y>>>(i+1) is the sign word and x' = x - sign: identical results, six instructions instead of eight. */
int zc_topolar_tail_stages(const zc_params *p);
/* Diagnostic: the run length K of the comb sample mapping zc_nco_rotate would use for a phase accumulator advancing by
 * `step` over n samples (0: the block mapping is used -- a slow NCO, or no suitable K; negative: zc_status).  K*step is
 * within a third of a phase LSB of a multiple of 2^32, so samples K apart mostly read the same table rows; lane (a, b)
 * of a warp computes samples a*K + 16m + 4b + {0,1,2,3} of every 8K-sample tile.  Same results as the block mapping. */
long long zc_nco_comb_run(const zc_params *p, uint32_t step, size_t n);
/* gencordic -t tbl [-i n] [-p pw] [-o ow]   sw/main.cpp:358-379 ; limit sw/sintable.cpp:62 */
int zc_derive_tbl(int iw, int pw, int ow, int *pw_out, int *ow_out);
/* gencordic -t qtr ...                       sw/main.cpp:401-422 ; limit sw/sintable.cpp:190 */
int zc_derive_qtr(int iw, int pw, int ow, int *pw_out, int *ow_out);

/* gencordic -t qtbl [-i iw] [-o ow] [-p pw] [-x xtra]   sw/main.cpp:444-484, sw/quadtbl.cpp:136-304
 * (table growth until the table error is below one unit, coefficient widths, the three tables). */
int zc_derive_qtbl(int iw, int ow, int xtra_user, int pw, zc_quadtbl *out);

/* The $readmemh table contents (host side; same libm as the generator so the words equal
 * rtl/sintable.hex / rtl/quarterwav.hex).  Words are masked to OW bits as hextable() writes
 * them (sw/hexfile.cpp:78-89).  sintable: 2^pw words (sw/sintable.cpp:156-168);
 * quarterwav: 2^(pw-2) words (sw/sintable.cpp:325-337). */
int zc_lut_build_sintable(int pw, int ow, uint32_t *tbl_host);
int zc_lut_build_quarterwav(int pw, int ow, uint32_t *tbl_host);

/* ---- the data path, device buffers --------------------------------------------------- */

/* rtl/cordic.v with constant (i_xval,i_yval) = (x0,y0), one i_phase per sample
 * (ports rtl/cordic.v:58-63; this is the sweep of bench/cpp/cordic_tb.cpp:127-178).
 * xy[2*i] = o_xval, xy[2*i+1] = o_yval. */
int zc_rotate_const(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase,
		int32_t *xy, size_t n, int device, void *stream);
int zc_rotate_const_ex(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase,
		int32_t *xy, size_t n, int device, void *stream, uint32_t flags);
/* rtl/cordic.v with per-sample (i_xval,i_yval) = (xy_in[2i], xy_in[2i+1]) */
int zc_rotate(const zc_params *p, const int32_t *xy_in, const uint32_t *phase,
		int32_t *xy_out, size_t n, int device, void *stream);
int zc_rotate_ex(const zc_params *p, const int32_t *xy_in, const uint32_t *phase,
		int32_t *xy_out, size_t n, int device, void *stream, uint32_t flags);
/* rtl/topolar.v (ports :59-64; sweep of bench/cpp/topolar_tb.cpp:127-189):
 * mag[i] = o_mag, phase[i] = o_phase */
int zc_topolar(const zc_params *p, const int32_t *xy_in, int32_t *mag, uint32_t *phase,
		size_t n, int device, void *stream);
int zc_topolar_ex(const zc_params *p, const int32_t *xy_in, int32_t *mag, uint32_t *phase,
		size_t n, int device, void *stream, uint32_t flags);
/* NCO: a 32-bit phase accumulator feeding rtl/cordic.v.  Sample i (0<=i<n) uses
 * phase32 = phase0 + (n0+i)*step (mod 2^32) and i_phase = phase32 >> (32-PW), the
 * truncation bench/cpp/cordic_tb.cpp:128-138 applies for shift>=0.  No input stream. */
int zc_nco_rotate(const zc_params *p, int32_t x0, int32_t y0, uint32_t phase0, uint32_t step,
		uint64_t n0, int32_t *xy, size_t n, int device, void *stream);
int zc_nco_rotate_ex(const zc_params *p, int32_t x0, int32_t y0, uint32_t phase0, uint32_t step,
		uint64_t n0, int32_t *xy, size_t n, int device, void *stream, uint32_t flags);
/* NCO -> complex mixer: rotate the per-sample vector (xy_in[2i], xy_in[2i+1]) by the accumulating phase of
 * zc_nco_rotate -- rtl/cordic.v fed by a phase accumulator, the down-converter the README's blog list builds. */
int zc_nco_mix(const zc_params *p, const int32_t *xy_in, uint32_t phase0, uint32_t step, uint64_t n0,
		int32_t *xy_out, size_t n, int device, void *stream);
/* rtl/sintable.v:71-75 / rtl/quarterwav.v:92-109.  phase32 is a 32-bit NCO word; the core
 * sees i_phase = phase32 >> (32-pw).  tbl_dev: the table of zc_lut_build_* in device memory -- or any other words:
 * the lookup is rtl/sintable.v's / rtl/quarterwav.v's whatever the contents.  (Batches of 4 Mi samples and more with
 * scattered phases are served from a compressed copy of the table staged in shared memory when the table allows it --
 * half-wave odd symmetry and 16-bit range for sintable, 16/24-bit magnitudes for quarterwav, checked on the device per
 * call; same results either way.) */
int zc_lut_sin(int pw, int ow, const uint32_t *tbl_dev, const uint32_t *phase32, int32_t *out,
		size_t n, int device, void *stream);
int zc_lut_qwav(int pw, int ow, const uint32_t *tbl_dev, const uint32_t *phase32, int32_t *out,
		size_t n, int device, void *stream);

/* NCO -> LUT core: the 32-bit phase accumulator of zc_nco_rotate feeding rtl/sintable.v / rtl/quarterwav.v instead of
 * rtl/cordic.v (the table-based oscillator the generator's -t tbl / -t qtr cores are built for; BASELINE configs[3] asks
 * for the LUT modes "head to head with the CORDIC kernel", configs[4] for the streaming NCO).  Sample i uses
 * phase32 = phase0 + (n0+i)*step (mod 2^32), i_phase = phase32 >> (32-pw); out[i] = o_val.  No input stream: 4
 * algorithmic bytes per sample, all written. */
int zc_nco_lut_sin(int pw, int ow, const uint32_t *tbl_dev, uint32_t phase0, uint32_t step, uint64_t n0, int32_t *out,
		size_t n, int device, void *stream);
int zc_nco_lut_qwav(int pw, int ow, const uint32_t *tbl_dev, uint32_t phase0, uint32_t step, uint64_t n0, int32_t *out,
		size_t n, int device, void *stream);

/* ---- packed port words ("report separately") ----------------------------------------------
 * The ports of the generated cores are IW / OW bits wide (rtl/cordic.v:58-63, rtl/topolar.v:59-64); the entry points
 * above carry each port in a 32-bit word.  For cores with IW <= 16 / OW <= 16 these variants carry a pair of ports
 * in ONE 32-bit word -- half the bytes on the PCIe-bound host path -- with bit-identical values:
 *   zc_topolar_i16       xy16_in[2i] = i_xval, xy16_in[2i+1] = i_yval as int16 (the low IW bits are the port);
 *                        mag / phase as in zc_topolar.  12 algorithmic bytes per sample instead of 16.
 *   zc_rotate_const_o16  xy16[2i] = o_xval, xy16[2i+1] = o_yval as int16 (OW <= 16: the sign-extended port value
 *                        fits).  8 algorithmic bytes per sample instead of 12.
 *   zc_lut_sin_o16 / zc_lut_qwav_o16   out[i] = o_val as int16 (tables with OW <= 16: rtl/sintable.v:53-58,
 *                        rtl/quarterwav.v:54-59).  6 algorithmic bytes per sample instead of 8.
 * Buffers must be 4-byte aligned (2-byte for the LUT outputs). */
int zc_lut_sin_o16(int pw, int ow, const uint32_t *tbl_dev, const uint32_t *phase32, int16_t *out, size_t n,
		int device, void *stream);
int zc_lut_qwav_o16(int pw, int ow, const uint32_t *tbl_dev, const uint32_t *phase32, int16_t *out, size_t n,
		int device, void *stream);
int zc_topolar_i16(const zc_params *p, const int16_t *xy16_in, int32_t *mag, uint32_t *phase, size_t n,
		int device, void *stream);
int zc_rotate_const_o16(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase, int16_t *xy16,
		size_t n, int device, void *stream);

/* rtl/quadtbl.v:143-291: table lookup + quadratic interpolation + overflow-safe convergent rounding.
 * phase32 is a 32-bit NCO word; the core sees i_phase = phase32 >> (32-PW).  out[i] = o_sin. */
int zc_quadtbl_sin(const zc_quadtbl *q, const uint32_t *phase32, int32_t *out, size_t n, int device, void *stream);

/* ---- the data path, host buffers (end-to-end) ---------------------------------------- */

void *zc_host_alloc(size_t bytes);		/* pinned; NULL on failure */
void  zc_host_free(void *ptr);			/* memory of zc_host_alloc or zc_host_alloc_sharded */
/* Pinned memory for the *_host_multi entry points: the g-th 1/ndev of the buffer is placed on the NUMA node of
 * devices[g] (read from sysfs; mbind(2) before first touch, best effort), so that each device's shard is copied over
 * its own socket's memory controllers.  NULL on failure. */
void *zc_host_alloc_sharded(size_t bytes, const int *devices, int ndev);
int   zc_device_numa_node(int device);		/* NUMA node the device's PCIe slot hangs off; -1 when unknown */

int zc_rotate_const_host(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase,
		int32_t *xy, size_t n, int device);
int zc_rotate_host(const zc_params *p, const int32_t *xy_in, const uint32_t *phase,
		int32_t *xy_out, size_t n, int device);
int zc_topolar_host(const zc_params *p, const int32_t *xy_in, int32_t *mag, uint32_t *phase,
		size_t n, int device);
int zc_nco_rotate_host(const zc_params *p, int32_t x0, int32_t y0, uint32_t phase0,
		uint32_t step, uint64_t n0, int32_t *xy, size_t n, int device);
int zc_lut_sin_host(int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32,
		int32_t *out, size_t n, int device);
int zc_lut_qwav_host(int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32,
		int32_t *out, size_t n, int device);
int zc_quadtbl_sin_host(const zc_quadtbl *q, const uint32_t *phase32, int32_t *out, size_t n, int device);
int zc_nco_mix_host(const zc_params *p, const int32_t *xy_in, uint32_t phase0, uint32_t step, uint64_t n0,
		int32_t *xy_out, size_t n, int device);

int zc_nco_lut_sin_host(int pw, int ow, const uint32_t *tbl_host, uint32_t phase0, uint32_t step, uint64_t n0,
		int32_t *out, size_t n, int device);
int zc_nco_lut_qwav_host(int pw, int ow, const uint32_t *tbl_host, uint32_t phase0, uint32_t step, uint64_t n0,
		int32_t *out, size_t n, int device);
int zc_lut_sin_o16_host(int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32, int16_t *out, size_t n,
		int device);
int zc_lut_qwav_o16_host(int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32, int16_t *out, size_t n,
		int device);
int zc_topolar_i16_host(const zc_params *p, const int16_t *xy16_in, int32_t *mag, uint32_t *phase, size_t n,
		int device);
int zc_rotate_const_o16_host(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase, int16_t *xy16,
		size_t n, int device);

/* ---- the data path, host buffers, several devices of one box -----------------------------------------------------
 * The sample stream is cut into ndev contiguous shards (boundaries at multiples of 4 samples; samples are independent
 * units -- SURVEY.md section 8e -- so there is no exchange step); shard g runs through the H2D -> kernel -> D2H
 * pipeline of the matching zc_*_host call on devices[g], one host thread per device, concurrently.  Returns when every
 * output word is in host memory.  Concatenated output == single-device output, byte for byte.  The NCO derives every
 * shard's start phase in closed form (phase0 + (n0 + first)*step), the truncation of bench/cpp/cordic_tb.cpp:128-138. */
/* The sharding rule itself: shard g of ndev over n samples is [first, first + count), boundaries at multiples of 4
 * samples (every shard keeps the 16-byte alignment of the whole), the last shard takes the ragged tail.  A caller that
 * keeps device-resident shards (one process per GPU under torchrun/MPI) uses the same rule, and for the NCO passes
 * n0 + first as the shard's start index. */
int zc_shard_range(size_t n, int ndev, int g, size_t *first, size_t *count);
int zc_rotate_const_host_multi(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase, int32_t *xy, size_t n,
		const int *devices, int ndev);
int zc_rotate_host_multi(const zc_params *p, const int32_t *xy_in, const uint32_t *phase, int32_t *xy_out, size_t n,
		const int *devices, int ndev);
int zc_topolar_host_multi(const zc_params *p, const int32_t *xy_in, int32_t *mag, uint32_t *phase, size_t n,
		const int *devices, int ndev);
int zc_nco_rotate_host_multi(const zc_params *p, int32_t x0, int32_t y0, uint32_t phase0, uint32_t step, uint64_t n0,
		int32_t *xy, size_t n, const int *devices, int ndev);
int zc_lut_sin_host_multi(int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32, int32_t *out, size_t n,
		const int *devices, int ndev);
int zc_lut_qwav_host_multi(int pw, int ow, const uint32_t *tbl_host, const uint32_t *phase32, int32_t *out, size_t n,
		const int *devices, int ndev);
int zc_quadtbl_sin_host_multi(const zc_quadtbl *q, const uint32_t *phase32, int32_t *out, size_t n, const int *devices,
		int ndev);
int zc_nco_mix_host_multi(const zc_params *p, const int32_t *xy_in, uint32_t phase0, uint32_t step, uint64_t n0,
		int32_t *xy_out, size_t n, const int *devices, int ndev);
int zc_topolar_i16_host_multi(const zc_params *p, const int16_t *xy16_in, int32_t *mag, uint32_t *phase, size_t n,
		const int *devices, int ndev);
int zc_rotate_const_o16_host_multi(const zc_params *p, int32_t x0, int32_t y0, const uint32_t *phase, int16_t *xy16, size_t n,
		const int *devices, int ndev);

/* $readmemh files in the layout hextable() writes (sw/hexfile.cpp:78-89): exchange LUTs with FPGA flows.
 * zc_hex_read returns the number of words read (>= 0) or a negative zc_status. */
int  zc_hex_write(const char *path, const uint32_t *words, size_t nwords, int bits);
long zc_hex_read(const char *path, uint32_t *words, size_t max_words);

/* Number of kernel launches this library has enqueued from the calling process so far
 * (all devices); lets a benchmark report how many of OUR kernels ran in a timed region. */
uint64_t zc_launch_count(void);

/* The library keeps a few device allocations between calls: the seed/direction tables of recently used
 * configurations, quadtbl coefficient tables, and the staging buffers + streams of the *_host pipelines
 * (the model the reference's test bench `new`s once and clocks many times, bench/cpp/testb.h:56-63).
 * zc_trim releases those of `device` (every device when negative); work already enqueued finishes first.
 * The first call for a new configuration uploads its tables and synchronises the stream -- warm up before
 * capturing calls into a CUDA graph.
 * A captured graph holds raw pointers to the tables of the configurations it uses and thereby PINS them: between
 * capture and the last replay do not call zc_trim for that device and do not cycle more than 15 other
 * (configuration, input vector) pairs through the engine (the table caches keep the 16 most recently used and free an
 * evicted entry once its last in-flight user returns -- a replay is not a user the library can see).  Nothing else is
 * shared between calls: a graph may replay concurrently with any other call on any other stream. */
int zc_trim(int device);

#ifdef __cplusplus
}
#endif
#endif /* ZCORDIC_H */
