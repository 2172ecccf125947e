"""cordic_b200 -- host-side Python mirror of libzcordic's C ABI (include/zcordic.h).

The product is the shared library (hand-written sm_100a kernels behind a C ABI); this
module only binds it with ctypes and moves pointers around.  PyTorch is used for device
memory and streams, nothing else.  There is no CPU implementation here: every compute
call goes to the CUDA kernels and raises ``ZcError`` if the library or a GPU is missing.

Vocabulary follows the reference (ZipCPU/cordic): a *core* is configured with the
generator's flags (``-i -o -x -p -n``, sw/main.cpp:139-232) and exposes the generated
header constants (``IW OW NEXTRA WW PW NSTAGES GAIN ...``, rtl/cordic.h:46-59).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzcordic.so")
ZC_MAX_STAGES = 64

F_DEFAULT = 0
F_FORCE_GENERIC = 1
F_NO_SEED = 2
F_FORCE_SEED = 4
F_SEED_PACKED = 8
F_SEED_REGS = 16
F_SEED_WORDS = 32
F_NO_TAIL = 64
F_NO_DP2A = 128
F_NO_COMB = 256
F_NO_MERGE = 512

MODE_P2R, MODE_R2P = 0, 1


class ZcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("zcordic error %d: %s" % (code, msg))
        self.code = code


class Params(ctypes.Structure):
    """zc_params (include/zcordic.h)."""
    _fields_ = [
        ("mode", ctypes.c_int32), ("iw", ctypes.c_int32), ("ow", ctypes.c_int32),
        ("nextra", ctypes.c_int32), ("ww", ctypes.c_int32), ("pw", ctypes.c_int32),
        ("nstages", ctypes.c_int32), ("seq", ctypes.c_int32),
        ("angle", ctypes.c_uint32 * ZC_MAX_STAGES),
        ("gain", ctypes.c_double), ("cordic_gain", ctypes.c_double), ("qvar", ctypes.c_double),
        ("pvar_rad", ctypes.c_double), ("best_cnr", ctypes.c_double),
    ]

    def angles(self):
        return [int(self.angle[k]) for k in range(self.nstages)]

    def header(self):
        """The constant set of the generated rtl/X.h, by the reference's names."""
        h = dict(IW=self.iw, OW=self.ow, NEXTRA=self.nextra, WW=self.ww, PW=self.pw,
                 NSTAGES=self.nstages, QUANTIZATION_VARIANCE=self.qvar,
                 PHASE_VARIANCE_RAD=self.pvar_rad, GAIN=self.gain)
        if self.mode == MODE_P2R:
            h["BEST_POSSIBLE_CNR"] = self.best_cnr
        if self.seq:
            h["CLOCKS_PER_OUTPUT"] = int(lib().zc_clocks_per_output(ctypes.byref(self)))
        return h


class QuadTblParams(ctypes.Structure):
    """zc_quadtbl (include/zcordic.h)."""
    _fields_ = [
        ("ow", ctypes.c_int32), ("nextra", ctypes.c_int32), ("pw", ctypes.c_int32), ("ww", ctypes.c_int32),
        ("lgtbl", ctypes.c_int32), ("dxbits", ctypes.c_int32), ("cbits", ctypes.c_int32), ("lbits", ctypes.c_int32),
        ("qbits", ctypes.c_int32), ("reserved", ctypes.c_int32), ("scale", ctypes.c_int64),
        ("itbl_err", ctypes.c_double), ("tbl_err", ctypes.c_double), ("spurdb", ctypes.c_double),
        ("ctbl", ctypes.c_uint32 * 4096), ("ltbl", ctypes.c_uint32 * 4096), ("qtbl", ctypes.c_uint32 * 4096),
    ]

    def header(self):
        """The constant set of the generated rtl/quadtbl.h, by the reference's names."""
        return dict(OW=self.ow, NEXTRA=self.nextra, PW=self.pw, TBL_LGSZ=self.lgtbl, TBL_SZ=1 << self.lgtbl,
                    SCALE=self.scale, ITBL_ERR=self.itbl_err, TBL_ERR=self.tbl_err, SPURDB=self.spurdb)


_lib = None

_SIGNATURES = {
    # name: (restype, argtypes)
    "zc_version": (ctypes.c_int, []),
    "zc_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "zc_last_error": (ctypes.c_char_p, []),
    "zc_device_count": (ctypes.c_int, []),
    "zc_launch_count": (ctypes.c_uint64, []),
    "zc_trim": (ctypes.c_int, [ctypes.c_int]),
    "zc_derive_p2r": (ctypes.c_int, [ctypes.c_int] * 5 + [ctypes.POINTER(Params)]),
    "zc_derive_r2p": (ctypes.c_int, [ctypes.c_int] * 5 + [ctypes.POINTER(Params)]),
    "zc_derive_sp2r": (ctypes.c_int, [ctypes.c_int] * 5 + [ctypes.POINTER(Params)]),
    "zc_derive_sr2p": (ctypes.c_int, [ctypes.c_int] * 5 + [ctypes.POINTER(Params)]),
    "zc_iterations": (ctypes.c_int, [ctypes.POINTER(Params)]),
    "zc_clocks_per_output": (ctypes.c_int, [ctypes.POINTER(Params)]),
    "zc_topolar_tail_stages": (ctypes.c_int, [ctypes.POINTER(Params)]),
    "zc_nco_comb_run": (ctypes.c_longlong, [ctypes.POINTER(Params), ctypes.c_uint32, ctypes.c_size_t]),
    "zc_derive_tbl": (ctypes.c_int, [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_int)] * 2),
    "zc_derive_qtr": (ctypes.c_int, [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_int)] * 2),
    "zc_lut_build_sintable": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "zc_lut_build_quarterwav": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "zc_rotate_const": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                       ctypes.c_void_p]),
    "zc_rotate_const_ex": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_uint32]),
    "zc_rotate": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "zc_rotate_ex": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32]),
    "zc_topolar": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "zc_topolar_ex": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32]),
    "zc_nco_rotate": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32,
                                     ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t,
                                     ctypes.c_int, ctypes.c_void_p]),
    "zc_nco_rotate_ex": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32,
                                        ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32]),
    "zc_lut_sin": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "zc_lut_qwav": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "zc_nco_lut_sin": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                      ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "zc_nco_lut_qwav": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "zc_nco_lut_sin_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                           ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "zc_nco_lut_qwav_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                            ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "zc_lut_sin_o16": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "zc_lut_qwav_o16": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "zc_lut_sin_o16_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "zc_lut_qwav_o16_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "zc_derive_qtbl": (ctypes.c_int, [ctypes.c_int] * 4 + [ctypes.POINTER(QuadTblParams)]),
    "zc_quadtbl_sin": (ctypes.c_int, [ctypes.POINTER(QuadTblParams), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                      ctypes.c_int, ctypes.c_void_p]),
    "zc_quadtbl_sin_host": (ctypes.c_int, [ctypes.POINTER(QuadTblParams), ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_size_t, ctypes.c_int]),
    "zc_nco_mix": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64,
                                  ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "zc_nco_mix_host": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "zc_hex_write": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "zc_hex_read": (ctypes.c_long, [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]),
    "zc_host_alloc": (ctypes.c_void_p, [ctypes.c_size_t]),
    "zc_host_alloc_sharded": (ctypes.c_void_p, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "zc_device_numa_node": (ctypes.c_int, [ctypes.c_int]),
    "zc_shard_range": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t),
                                      ctypes.POINTER(ctypes.c_size_t)]),
    "zc_topolar_i16": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "zc_rotate_const_o16": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                           ctypes.c_void_p]),
    "zc_topolar_i16_host": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_size_t, ctypes.c_int]),
    "zc_rotate_const_o16_host": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_int32, ctypes.c_int32,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "zc_rotate_const_host_multi": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_int32, ctypes.c_int32,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                  ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "zc_rotate_host_multi": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "zc_topolar_host_multi": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "zc_nco_rotate_host_multi": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32,
                                                ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t,
                                                ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "zc_lut_sin_host_multi": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "zc_lut_qwav_host_multi": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "zc_quadtbl_sin_host_multi": (ctypes.c_int, [ctypes.POINTER(QuadTblParams), ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "zc_nco_mix_host_multi": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t,
                                             ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "zc_topolar_i16_host_multi": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "zc_rotate_const_o16_host_multi": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_int32, ctypes.c_int32,
                                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                      ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "zc_host_free": (None, [ctypes.c_void_p]),
    "zc_rotate_const_host": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_int32, ctypes.c_int32,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "zc_rotate_host": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_size_t, ctypes.c_int]),
    "zc_topolar_host": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_size_t, ctypes.c_int]),
    "zc_nco_rotate_host": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32,
                                          ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t,
                                          ctypes.c_int]),
    "zc_lut_sin_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "zc_lut_qwav_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
}

EXPORTED_SYMBOLS = sorted(_SIGNATURES)


def lib():
    """Loads libzcordic.so (building it with nvcc first if it is not there). Never falls back."""
    global _lib
    if _lib is None:
        path = os.environ.get("ZCORDIC_LIB") or LIB_PATH     # ZCORDIC_LIB: an experiment build (cordic_b200/build.py --out)
        if path == LIB_PATH and not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        L = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise ZcError(rc, (lib().zc_last_error() or b"").decode() or lib().zc_strerror(rc).decode())


def launch_count():
    return int(lib().zc_launch_count())


def trim(device=-1):
    """Releases the library's cached device allocations (tables, staging buffers) -- zc_trim."""
    _check(lib().zc_trim(int(device)))


def shard_range(n, ndev, g):
    """(first, count) of shard g of ndev over n samples -- zc_shard_range, the rule the *_host_multi calls apply."""
    a, b = ctypes.c_size_t(), ctypes.c_size_t()
    _check(lib().zc_shard_range(int(n), int(ndev), int(g), ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


# ---- configuration ------------------------------------------------------------------------

def derive_p2r(iw=0, ow=0, xtra=2, pw=0, nstages=0):
    """``gencordic -t p2r -i iw -o ow -x xtra [-p pw] [-n nstages]`` (sw/main.cpp:260-279)."""
    p = Params()
    _check(lib().zc_derive_p2r(iw or 0, ow or 0, xtra, pw or 0, nstages or 0, ctypes.byref(p)))
    return p


def derive_r2p(iw=0, ow=0, xtra=2, pw=0, nstages=0):
    """``gencordic -t r2p ...`` (sw/main.cpp:312-328, sw/topolar.cpp:67-75)."""
    p = Params()
    _check(lib().zc_derive_r2p(iw or 0, ow or 0, xtra, pw or 0, nstages or 0, ctypes.byref(p)))
    return p


def derive_sp2r(iw=0, ow=0, xtra=2, pw=0, nstages=0):
    """``gencordic -t sp2r ...``: the sequential core rtl/seqcordic.v (zc_params.seq = 1)."""
    p = Params()
    _check(lib().zc_derive_sp2r(iw or 0, ow or 0, xtra, pw or 0, nstages or 0, ctypes.byref(p)))
    return p


def derive_sr2p(iw=0, ow=0, xtra=2, pw=0, nstages=0):
    """``gencordic -t sr2p ...``: the sequential core rtl/seqpolar.v (zc_params.seq = 1)."""
    p = Params()
    _check(lib().zc_derive_sr2p(iw or 0, ow or 0, xtra, pw or 0, nstages or 0, ctypes.byref(p)))
    return p


def derive_tbl(iw=0, pw=0, ow=0):
    a, b = ctypes.c_int(), ctypes.c_int()
    _check(lib().zc_derive_tbl(iw or 0, pw or 0, ow or 0, ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def derive_qtr(iw=0, pw=0, ow=0):
    a, b = ctypes.c_int(), ctypes.c_int()
    _check(lib().zc_derive_qtr(iw or 0, pw or 0, ow or 0, ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def derive_qtbl(iw=0, ow=0, xtra=2, pw=0):
    """``gencordic -t qtbl [-i iw] [-o ow] [-x xtra] [-p pw]`` (sw/main.cpp:444-484, sw/quadtbl.cpp)."""
    q = QuadTblParams()
    _check(lib().zc_derive_qtbl(iw or 0, ow or 0, xtra, pw or 0, ctypes.byref(q)))
    return q


def hex_write(path, words, bits):
    """Write a $readmemh file in the reference's layout (sw/hexfile.cpp:78-89)."""
    w = np.ascontiguousarray(words, dtype=np.uint32)
    _check(lib().zc_hex_write(path.encode(), w.ctypes.data, w.size, bits))


def hex_read(path, max_words=1 << 26):
    w = np.zeros(max_words, dtype=np.uint32)
    n = lib().zc_hex_read(path.encode(), w.ctypes.data, max_words)
    if n < 0:
        _check(int(n))
    return w[:n].copy()


def build_sintable(pw, ow):
    """The words of sintable.hex (sw/sintable.cpp:156-168) as a uint32 numpy array."""
    tbl = np.empty(1 << pw, dtype=np.uint32) if 0 < pw < 31 else np.empty(1, dtype=np.uint32)
    _check(lib().zc_lut_build_sintable(pw, ow, tbl.ctypes.data))
    return tbl


def build_quarterwav(pw, ow):
    """The words of quarterwav.hex (sw/sintable.cpp:325-337) as a uint32 numpy array."""
    tbl = np.empty(1 << (pw - 2), dtype=np.uint32) if 2 < pw < 31 else np.empty(1, dtype=np.uint32)
    _check(lib().zc_lut_build_quarterwav(pw, ow, tbl.ctypes.data))
    return tbl


# ---- buffers --------------------------------------------------------------------------------

def _torch():
    import torch
    return torch


def _is_tensor(x):
    return type(x).__module__.startswith("torch")


def _dev_ptr(t, nwords=None, itemsize=4):
    torch = _torch()
    if not (_is_tensor(t) and t.is_cuda):
        raise ZcError(-1, "expected a CUDA tensor")
    if t.element_size() != itemsize or not t.is_contiguous():
        raise ZcError(-1, "expected a contiguous %d-bit tensor, got %s" % (8 * itemsize, t.dtype))
    if nwords is not None and t.numel() != nwords:
        raise ZcError(-1, "tensor has %d words, expected %d" % (t.numel(), nwords))
    return ctypes.c_void_p(t.data_ptr())


def _devices(devices):
    arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
    return arr, len(devices)


def _stream_ptr(device_index, stream):
    torch = _torch()
    s = stream if stream is not None else torch.cuda.current_stream(device_index)
    return ctypes.c_void_p(s.cuda_stream)


def _host_ptr(a, nwords=None, itemsize=4):
    """numpy array or CPU (possibly pinned) torch tensor -> pointer."""
    if _is_tensor(a):
        if a.is_cuda or a.element_size() != itemsize or not a.is_contiguous():
            raise ZcError(-1, "expected a contiguous %d-bit CPU tensor" % (8 * itemsize))
        if nwords is not None and a.numel() != nwords:
            raise ZcError(-1, "tensor has %d words, expected %d" % (a.numel(), nwords))
        return ctypes.c_void_p(a.data_ptr())
    if not isinstance(a, np.ndarray) or a.dtype.itemsize != itemsize or not a.flags["C_CONTIGUOUS"]:
        raise ZcError(-1, "expected a C-contiguous %d-bit numpy array" % (8 * itemsize))
    if nwords is not None and a.size != nwords:
        raise ZcError(-1, "array has %d words, expected %d" % (a.size, nwords))
    return ctypes.c_void_p(a.ctypes.data)


class PinnedBuffer:
    """Page-locked host memory from zc_host_alloc, exposed as a numpy array."""

    def __init__(self, nwords, dtype=np.int32):
        self.nbytes = int(nwords) * 4
        self.ptr = lib().zc_host_alloc(self.nbytes)
        if not self.ptr:
            raise ZcError(-5, (lib().zc_last_error() or b"").decode())
        buf = (ctypes.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(nwords))

    def free(self):
        if self.ptr:
            self.array = None
            lib().zc_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ShardedPinnedBuffer(PinnedBuffer):
    """Pinned host memory from zc_host_alloc_sharded: the g-th 1/len(devices) of the buffer lives on the NUMA node of
    devices[g] -- the buffers the *_host_multi calls want.  ``placement`` lists (device, node) pairs."""

    def __init__(self, nwords, devices, dtype=np.int32):
        self.nbytes = int(nwords) * 4
        arr, nd = _devices(devices)
        self.ptr = lib().zc_host_alloc_sharded(self.nbytes, arr, nd)
        if not self.ptr:
            raise ZcError(-5, (lib().zc_last_error() or b"").decode())
        buf = (ctypes.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(nwords))
        self.placement = [(int(d), int(lib().zc_device_numa_node(int(d)))) for d in devices]


# ---- cores ----------------------------------------------------------------------------------

class Cordic:
    """Rotation-mode core: the function of rtl/cordic.v (phase -> rotated (x, y))."""

    def __init__(self, iw=0, ow=0, xtra=2, phase_bits=0, nstages=0, sequential=False):
        """``sequential=True``: the function of rtl/seqcordic.v (``-t sp2r``) instead."""
        self.params = (derive_sp2r if sequential else derive_p2r)(iw, ow, xtra, phase_bits, nstages)
        for k, v in self.params.header().items():
            setattr(self, k, v)

    # device buffers -------------------------------------------------------------------
    def rotate_const(self, x0, y0, phase, out=None, stream=None, flags=F_DEFAULT):
        """(i_xval, i_yval) = (x0, y0) for every sample; ``phase``: CUDA tensor of n words.
        Returns an int32 CUDA tensor [n, 2] of (o_xval, o_yval)."""
        torch = _torch()
        n = phase.numel()
        dev = phase.device.index or 0
        if out is None:
            out = torch.empty((n, 2), dtype=torch.int32, device=phase.device)
        _check(lib().zc_rotate_const_ex(ctypes.byref(self.params), int(x0), int(y0), _dev_ptr(phase),
                                        _dev_ptr(out, 2 * n), n, dev, _stream_ptr(dev, stream), flags))
        return out

    def rotate_const_o16(self, x0, y0, phase, out=None, stream=None):
        """As rotate_const for a core with OW <= 16, outputs packed: int16 CUDA tensor [n, 2] (zc_rotate_const_o16)."""
        torch = _torch()
        n = phase.numel()
        dev = phase.device.index or 0
        if out is None:
            out = torch.empty((n, 2), dtype=torch.int16, device=phase.device)
        _check(lib().zc_rotate_const_o16(ctypes.byref(self.params), int(x0), int(y0), _dev_ptr(phase),
                                         _dev_ptr(out, 2 * n, 2), n, dev, _stream_ptr(dev, stream)))
        return out

    def rotate(self, xy, phase, out=None, stream=None, flags=F_DEFAULT):
        """Per-sample (i_xval, i_yval) = xy[i]; xy: int32 CUDA tensor [n, 2]."""
        torch = _torch()
        n = phase.numel()
        dev = phase.device.index or 0
        if out is None:
            out = torch.empty((n, 2), dtype=torch.int32, device=phase.device)
        _check(lib().zc_rotate_ex(ctypes.byref(self.params), _dev_ptr(xy, 2 * n), _dev_ptr(phase),
                                  _dev_ptr(out, 2 * n), n, dev, _stream_ptr(dev, stream), flags))
        return out

    def nco(self, x0, y0, phase0, step, n, n0=0, out=None, device=None, stream=None, flags=F_DEFAULT):
        """Streaming NCO: phase32 = phase0 + (n0+i)*step, i_phase = phase32 >> (32-PW)."""
        torch = _torch()
        if out is None:
            out = torch.empty((n, 2), dtype=torch.int32, device=device if device is not None else "cuda")
        dev = out.device.index or 0
        _check(lib().zc_nco_rotate_ex(ctypes.byref(self.params), int(x0), int(y0), int(phase0) & 0xFFFFFFFF,
                                      int(step) & 0xFFFFFFFF, int(n0), _dev_ptr(out, 2 * n), n, dev,
                                      _stream_ptr(dev, stream), flags))
        return out

    def mix(self, xy, phase0, step, n0=0, out=None, stream=None):
        """NCO mixer: rotate each (x, y) of ``xy`` by phase0 + (n0+i)*step (32-bit accumulator)."""
        torch = _torch()
        n = xy.numel() // 2
        dev = xy.device.index or 0
        if out is None:
            out = torch.empty((n, 2), dtype=torch.int32, device=xy.device)
        _check(lib().zc_nco_mix(ctypes.byref(self.params), _dev_ptr(xy, 2 * n), int(phase0) & 0xFFFFFFFF,
                                int(step) & 0xFFFFFFFF, int(n0), _dev_ptr(out, 2 * n), n, dev, _stream_ptr(dev, stream)))
        return out

    # host buffers (end to end) -----------------------------------------------------------
    def rotate_const_host(self, x0, y0, phase, out, device=0):
        n = phase.size if isinstance(phase, np.ndarray) else phase.numel()
        _check(lib().zc_rotate_const_host(ctypes.byref(self.params), int(x0), int(y0), _host_ptr(phase),
                                          _host_ptr(out, 2 * n), n, device))
        return out

    def rotate_const_o16_host(self, x0, y0, phase, out, device=0):
        n = phase.size if isinstance(phase, np.ndarray) else phase.numel()
        _check(lib().zc_rotate_const_o16_host(ctypes.byref(self.params), int(x0), int(y0), _host_ptr(phase),
                                              _host_ptr(out, 2 * n, 2), n, device))
        return out

    def rotate_const_host_multi(self, x0, y0, phase, out, devices):
        """zc_rotate_const_host_multi: the stream sharded over ``devices``, one host thread + pipeline per device."""
        n = phase.size if isinstance(phase, np.ndarray) else phase.numel()
        arr, nd = _devices(devices)
        _check(lib().zc_rotate_const_host_multi(ctypes.byref(self.params), int(x0), int(y0), _host_ptr(phase),
                                                _host_ptr(out, 2 * n), n, arr, nd))
        return out

    def rotate_const_o16_host_multi(self, x0, y0, phase, out, devices):
        n = phase.size if isinstance(phase, np.ndarray) else phase.numel()
        arr, nd = _devices(devices)
        _check(lib().zc_rotate_const_o16_host_multi(ctypes.byref(self.params), int(x0), int(y0), _host_ptr(phase),
                                                    _host_ptr(out, 2 * n, 2), n, arr, nd))
        return out

    def mix_host_multi(self, xy, phase0, step, out, devices, n0=0):
        n = (xy.size if isinstance(xy, np.ndarray) else xy.numel()) // 2
        arr, nd = _devices(devices)
        _check(lib().zc_nco_mix_host_multi(ctypes.byref(self.params), _host_ptr(xy, 2 * n), int(phase0) & 0xFFFFFFFF,
                                           int(step) & 0xFFFFFFFF, int(n0), _host_ptr(out, 2 * n), n, arr, nd))
        return out

    def rotate_host_multi(self, xy, phase, out, devices):
        n = phase.size if isinstance(phase, np.ndarray) else phase.numel()
        arr, nd = _devices(devices)
        _check(lib().zc_rotate_host_multi(ctypes.byref(self.params), _host_ptr(xy, 2 * n), _host_ptr(phase),
                                          _host_ptr(out, 2 * n), n, arr, nd))
        return out

    def nco_host_multi(self, x0, y0, phase0, step, out, devices, n0=0):
        n = (out.size if isinstance(out, np.ndarray) else out.numel()) // 2
        arr, nd = _devices(devices)
        _check(lib().zc_nco_rotate_host_multi(ctypes.byref(self.params), int(x0), int(y0), int(phase0) & 0xFFFFFFFF,
                                              int(step) & 0xFFFFFFFF, int(n0), _host_ptr(out), n, arr, nd))
        return out

    def rotate_host(self, xy, phase, out, device=0):
        n = phase.size if isinstance(phase, np.ndarray) else phase.numel()
        _check(lib().zc_rotate_host(ctypes.byref(self.params), _host_ptr(xy, 2 * n), _host_ptr(phase),
                                    _host_ptr(out, 2 * n), n, device))
        return out

    def nco_host(self, x0, y0, phase0, step, out, n0=0, device=0):
        n = (out.size if isinstance(out, np.ndarray) else out.numel()) // 2
        _check(lib().zc_nco_rotate_host(ctypes.byref(self.params), int(x0), int(y0), int(phase0) & 0xFFFFFFFF,
                                        int(step) & 0xFFFFFFFF, int(n0), _host_ptr(out), n, device))
        return out


class Topolar:
    """Vectoring-mode core: the function of rtl/topolar.v ((x, y) -> magnitude, phase)."""

    def __init__(self, iw=0, ow=0, xtra=2, phase_bits=0, nstages=0, sequential=False):
        """``sequential=True``: the function of rtl/seqpolar.v (``-t sr2p``) instead."""
        self.params = (derive_sr2p if sequential else derive_r2p)(iw, ow, xtra, phase_bits, nstages)
        for k, v in self.params.header().items():
            setattr(self, k, v)

    def topolar(self, xy, mag=None, phase=None, stream=None, flags=F_DEFAULT):
        torch = _torch()
        n = xy.numel() // 2
        dev = xy.device.index or 0
        if mag is None:
            mag = torch.empty(n, dtype=torch.int32, device=xy.device)
        if phase is None:
            phase = torch.empty(n, dtype=torch.int32, device=xy.device)
        _check(lib().zc_topolar_ex(ctypes.byref(self.params), _dev_ptr(xy, 2 * n), _dev_ptr(mag, n),
                                   _dev_ptr(phase, n), n, dev, _stream_ptr(dev, stream), flags))
        return mag, phase

    def topolar_i16(self, xy16, mag=None, phase=None, stream=None):
        """As topolar for a core with IW <= 16, inputs packed: xy16 is an int16 CUDA tensor [n, 2] (zc_topolar_i16)."""
        torch = _torch()
        n = xy16.numel() // 2
        dev = xy16.device.index or 0
        if mag is None:
            mag = torch.empty(n, dtype=torch.int32, device=xy16.device)
        if phase is None:
            phase = torch.empty(n, dtype=torch.int32, device=xy16.device)
        _check(lib().zc_topolar_i16(ctypes.byref(self.params), _dev_ptr(xy16, 2 * n, 2), _dev_ptr(mag, n),
                                    _dev_ptr(phase, n), n, dev, _stream_ptr(dev, stream)))
        return mag, phase

    def topolar_i16_host(self, xy16, mag, phase, device=0):
        n = (xy16.size if isinstance(xy16, np.ndarray) else xy16.numel()) // 2
        _check(lib().zc_topolar_i16_host(ctypes.byref(self.params), _host_ptr(xy16, 2 * n, 2), _host_ptr(mag, n),
                                         _host_ptr(phase, n), n, device))
        return mag, phase

    def topolar_i16_host_multi(self, xy16, mag, phase, devices):
        n = (xy16.size if isinstance(xy16, np.ndarray) else xy16.numel()) // 2
        arr, nd = _devices(devices)
        _check(lib().zc_topolar_i16_host_multi(ctypes.byref(self.params), _host_ptr(xy16, 2 * n, 2), _host_ptr(mag, n),
                                               _host_ptr(phase, n), n, arr, nd))
        return mag, phase

    def topolar_host_multi(self, xy, mag, phase, devices):
        n = (xy.size if isinstance(xy, np.ndarray) else xy.numel()) // 2
        arr, nd = _devices(devices)
        _check(lib().zc_topolar_host_multi(ctypes.byref(self.params), _host_ptr(xy), _host_ptr(mag, n),
                                           _host_ptr(phase, n), n, arr, nd))
        return mag, phase

    def topolar_host(self, xy, mag, phase, device=0):
        n = (xy.size if isinstance(xy, np.ndarray) else xy.numel()) // 2
        _check(lib().zc_topolar_host(ctypes.byref(self.params), _host_ptr(xy), _host_ptr(mag, n),
                                     _host_ptr(phase, n), n, device))
        return mag, phase


class _Lut:
    QUARTER = False

    def __init__(self, iw=0, phase_bits=0, ow=0):
        derive = derive_qtr if self.QUARTER else derive_tbl
        self.PW, self.OW = derive(iw, phase_bits, ow)
        build = build_quarterwav if self.QUARTER else build_sintable
        self.table = build(self.PW, self.OW)          # host copy == the $readmemh words
        self._dev = {}

    def _table_on(self, device):
        torch = _torch()
        key = str(device)
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(self.table.view(np.int32)).to(device)
        return self._dev[key]

    def lookup(self, phase32, out=None, stream=None):
        """phase32: CUDA tensor of 32-bit NCO phase words; i_phase = phase32 >> (32-PW)."""
        torch = _torch()
        n = phase32.numel()
        dev = phase32.device.index or 0
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=phase32.device)
        fn = lib().zc_lut_qwav if self.QUARTER else lib().zc_lut_sin
        _check(fn(self.PW, self.OW, _dev_ptr(self._table_on(phase32.device)), _dev_ptr(phase32),
                  _dev_ptr(out, n), n, dev, _stream_ptr(dev, stream)))
        return out

    def nco(self, phase0, step, n, n0=0, out=None, device=None, stream=None):
        """Streaming NCO through the table: phase32 = phase0 + (n0+i)*step, no input stream (zc_nco_lut_sin / _qwav)."""
        torch = _torch()
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=device if device is not None else "cuda")
        dev = out.device.index or 0
        fn = lib().zc_nco_lut_qwav if self.QUARTER else lib().zc_nco_lut_sin
        _check(fn(self.PW, self.OW, _dev_ptr(self._table_on(out.device)), int(phase0) & 0xFFFFFFFF, int(step) & 0xFFFFFFFF,
                  int(n0), _dev_ptr(out, n), n, dev, _stream_ptr(dev, stream)))
        return out

    def nco_host(self, phase0, step, out, n0=0, device=0):
        n = out.size if isinstance(out, np.ndarray) else out.numel()
        fn = lib().zc_nco_lut_qwav_host if self.QUARTER else lib().zc_nco_lut_sin_host
        _check(fn(self.PW, self.OW, self.table.ctypes.data, int(phase0) & 0xFFFFFFFF, int(step) & 0xFFFFFFFF, int(n0),
                  _host_ptr(out, n), n, device))
        return out

    def lookup_o16(self, phase32, out=None, stream=None):
        """As lookup for a table with OW <= 16, outputs as an int16 CUDA tensor (zc_lut_sin_o16 / zc_lut_qwav_o16)."""
        torch = _torch()
        n = phase32.numel()
        dev = phase32.device.index or 0
        if out is None:
            out = torch.empty(n, dtype=torch.int16, device=phase32.device)
        fn = lib().zc_lut_qwav_o16 if self.QUARTER else lib().zc_lut_sin_o16
        _check(fn(self.PW, self.OW, _dev_ptr(self._table_on(phase32.device)), _dev_ptr(phase32),
                  _dev_ptr(out, n, 2), n, dev, _stream_ptr(dev, stream)))
        return out

    def lookup_host(self, phase32, out, device=0):
        n = phase32.size if isinstance(phase32, np.ndarray) else phase32.numel()
        fn = lib().zc_lut_qwav_host if self.QUARTER else lib().zc_lut_sin_host
        _check(fn(self.PW, self.OW, self.table.ctypes.data, _host_ptr(phase32), _host_ptr(out, n), n, device))
        return out

    def lookup_o16_host(self, phase32, out, device=0):
        n = phase32.size if isinstance(phase32, np.ndarray) else phase32.numel()
        fn = lib().zc_lut_qwav_o16_host if self.QUARTER else lib().zc_lut_sin_o16_host
        _check(fn(self.PW, self.OW, self.table.ctypes.data, _host_ptr(phase32), _host_ptr(out, n, 2), n, device))
        return out


    def lookup_host_multi(self, phase32, out, devices):
        n = phase32.size if isinstance(phase32, np.ndarray) else phase32.numel()
        fn = lib().zc_lut_qwav_host_multi if self.QUARTER else lib().zc_lut_sin_host_multi
        arr, nd = _devices(devices)
        _check(fn(self.PW, self.OW, self.table.ctypes.data, _host_ptr(phase32), _host_ptr(out, n), n, arr, nd))
        return out


class SinTable(_Lut):
    """rtl/sintable.v: o_val = tbl[i_phase] (``gencordic -t tbl``)."""
    QUARTER = False


class QuarterWav(_Lut):
    """rtl/quarterwav.v: quarter-wave table with index fold and negate (``gencordic -t qtr``)."""
    QUARTER = True


class QuadTbl:
    """rtl/quadtbl.v: table lookup with quadratic interpolation (``gencordic -t qtbl``)."""

    def __init__(self, iw=0, ow=0, xtra=2, phase_bits=0):
        self.params = derive_qtbl(iw, ow, xtra, phase_bits)
        for k, v in self.params.header().items():
            setattr(self, k, v)

    def tables(self):
        n = 1 << self.params.lgtbl
        return (np.array(self.params.ctbl[:n], dtype=np.uint32), np.array(self.params.ltbl[:n], dtype=np.uint32),
                np.array(self.params.qtbl[:n], dtype=np.uint32))

    def lookup(self, phase32, out=None, stream=None):
        """phase32: CUDA tensor of 32-bit NCO phase words; i_phase = phase32 >> (32-PW).  Returns o_sin."""
        torch = _torch()
        n = phase32.numel()
        dev = phase32.device.index or 0
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=phase32.device)
        _check(lib().zc_quadtbl_sin(ctypes.byref(self.params), _dev_ptr(phase32), _dev_ptr(out, n), n, dev,
                                    _stream_ptr(dev, stream)))
        return out

    def lookup_host(self, phase32, out, device=0):
        n = phase32.size if isinstance(phase32, np.ndarray) else phase32.numel()
        _check(lib().zc_quadtbl_sin_host(ctypes.byref(self.params), _host_ptr(phase32), _host_ptr(out, n), n, device))
        return out

    def lookup_host_multi(self, phase32, out, devices):
        n = phase32.size if isinstance(phase32, np.ndarray) else phase32.numel()
        arr, nd = _devices(devices)
        _check(lib().zc_quadtbl_sin_host_multi(ctypes.byref(self.params), _host_ptr(phase32), _host_ptr(out, n), n, arr, nd))
        return out
