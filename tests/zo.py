"""ctypes binding of the CPU oracle (oracle/libzc_oracle.so).  Test-side only."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


class ZoParams(ctypes.Structure):
    _fields_ = [("iw", ctypes.c_int), ("ow", ctypes.c_int), ("nextra", ctypes.c_int), ("ww", ctypes.c_int),
                ("pw", ctypes.c_int), ("nstages", ctypes.c_int), ("vectoring", ctypes.c_int), ("sequential", ctypes.c_int),
                ("angle", ctypes.c_uint32 * 64), ("gain", ctypes.c_double), ("cordic_gain", ctypes.c_double),
                ("qvar", ctypes.c_double), ("pvar_rad", ctypes.c_double), ("best_cnr", ctypes.c_double)]


class ZoQuadTbl(ctypes.Structure):
    _fields_ = [("ow", ctypes.c_int), ("nextra", ctypes.c_int), ("pw", ctypes.c_int), ("ww", ctypes.c_int),
                ("lgtbl", ctypes.c_int), ("dxbits", ctypes.c_int), ("cbits", ctypes.c_int), ("lbits", ctypes.c_int),
                ("qbits", ctypes.c_int), ("scale", ctypes.c_long), ("itbl_err", ctypes.c_double),
                ("tbl_err", ctypes.c_double), ("spurdb", ctypes.c_double),
                ("ctbl", ctypes.c_uint32 * 4096), ("ltbl", ctypes.c_uint32 * 4096), ("qtbl", ctypes.c_uint32 * 4096)]


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(os.path.join(ROOT, "oracle", "libzc_oracle.so"))
        P, vp, sz, i32, u32, ci = ctypes.POINTER(ZoParams), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32, ctypes.c_uint32, ctypes.c_int
        L.zo_derive_p2r.argtypes = [ci] * 5 + [P]
        L.zo_derive_r2p.argtypes = [ci] * 5 + [P]
        L.zo_derive_sp2r.argtypes = [ci] * 5 + [P]
        L.zo_derive_sr2p.argtypes = [ci] * 5 + [P]
        L.zo_iterations.argtypes = [P]
        L.zo_clocks_per_output.argtypes = [P]
        L.zo_derive_tbl.argtypes = [ci] * 3 + [ctypes.POINTER(ci)] * 2
        L.zo_derive_qtr.argtypes = [ci] * 3 + [ctypes.POINTER(ci)] * 2
        L.zo_rotate1.argtypes = [P, i32, i32, u32, ctypes.POINTER(i32), ctypes.POINTER(i32)]
        L.zo_topolar1.argtypes = [P, i32, i32, ctypes.POINTER(i32), ctypes.POINTER(u32)]
        L.zo_rotate_const.argtypes = [P, i32, i32, vp, vp, sz, ci]
        L.zo_rotate.argtypes = [P, vp, vp, vp, sz, ci]
        L.zo_topolar.argtypes = [P, vp, vp, vp, sz, ci]
        L.zo_nco_rotate.argtypes = [P, i32, i32, u32, u32, ctypes.c_uint64, vp, sz, ci]
        L.zo_sintable_build.argtypes = [ci, ci, vp]
        L.zo_quarterwav_build.argtypes = [ci, ci, vp]
        L.zo_lut_sin.argtypes = [ci, ci, vp, vp, vp, sz, ci]
        L.zo_lut_qwav.argtypes = [ci, ci, vp, vp, vp, sz, ci]
        L.zo_derive_qtbl.argtypes = [ci] * 4 + [ctypes.POINTER(ZoQuadTbl)]
        L.zo_quadtbl1.argtypes = [ctypes.POINTER(ZoQuadTbl), u32]
        L.zo_quadtbl1.restype = i32
        L.zo_quadtbl_batch.argtypes = [ctypes.POINTER(ZoQuadTbl), vp, vp, sz, ci]
        L.zo_hex_load.argtypes = [ctypes.c_char_p, vp, ctypes.c_long]
        L.zo_hex_load.restype = ctypes.c_long
        _lib = L
    return _lib


NTHREADS = min(16, os.cpu_count() or 1)


def derive_p2r(iw=0, ow=0, xtra=2, pw=0, nstages=0):
    p = ZoParams()
    rc = lib().zo_derive_p2r(iw or 0, ow or 0, xtra, pw or 0, nstages or 0, ctypes.byref(p))
    return rc, p


def derive_r2p(iw=0, ow=0, xtra=2, pw=0, nstages=0):
    p = ZoParams()
    rc = lib().zo_derive_r2p(iw or 0, ow or 0, xtra, pw or 0, nstages or 0, ctypes.byref(p))
    return rc, p


def derive_sp2r(iw=0, ow=0, xtra=2, pw=0, nstages=0):
    p = ZoParams()
    rc = lib().zo_derive_sp2r(iw or 0, ow or 0, xtra, pw or 0, nstages or 0, ctypes.byref(p))
    return rc, p


def derive_sr2p(iw=0, ow=0, xtra=2, pw=0, nstages=0):
    p = ZoParams()
    rc = lib().zo_derive_sr2p(iw or 0, ow or 0, xtra, pw or 0, nstages or 0, ctypes.byref(p))
    return rc, p


def clocks_per_output(p):
    return lib().zo_clocks_per_output(ctypes.byref(p))


def derive_lut(mode, iw=0, pw=0, ow=0):
    a, b = ctypes.c_int(), ctypes.c_int()
    f = lib().zo_derive_qtr if mode == "qtr" else lib().zo_derive_tbl
    rc = f(iw or 0, pw or 0, ow or 0, ctypes.byref(a), ctypes.byref(b))
    return rc, a.value, b.value


def rotate1(p, ix, iy, phase):
    ox, oy = ctypes.c_int32(), ctypes.c_int32()
    lib().zo_rotate1(ctypes.byref(p), ix, iy, phase, ctypes.byref(ox), ctypes.byref(oy))
    return ox.value, oy.value


def topolar1(p, ix, iy):
    m, ph = ctypes.c_int32(), ctypes.c_uint32()
    lib().zo_topolar1(ctypes.byref(p), ix, iy, ctypes.byref(m), ctypes.byref(ph))
    return m.value, ph.value


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def rotate_const(p, x0, y0, phase):
    phase = _c(phase, np.uint32)
    out = np.empty((phase.size, 2), dtype=np.int32)
    lib().zo_rotate_const(ctypes.byref(p), int(x0), int(y0), phase.ctypes.data, out.ctypes.data, phase.size, NTHREADS)
    return out


def rotate(p, xy, phase):
    phase = _c(phase, np.uint32)
    xy = _c(xy, np.int32)
    out = np.empty((phase.size, 2), dtype=np.int32)
    lib().zo_rotate(ctypes.byref(p), xy.ctypes.data, phase.ctypes.data, out.ctypes.data, phase.size, NTHREADS)
    return out


def topolar(p, xy):
    xy = _c(xy, np.int32)
    n = xy.size // 2
    mag = np.empty(n, dtype=np.int32)
    ph = np.empty(n, dtype=np.uint32)
    lib().zo_topolar(ctypes.byref(p), xy.ctypes.data, mag.ctypes.data, ph.ctypes.data, n, NTHREADS)
    return mag, ph


def nco(p, x0, y0, phase0, step, n, n0=0):
    out = np.empty((n, 2), dtype=np.int32)
    lib().zo_nco_rotate(ctypes.byref(p), int(x0), int(y0), phase0 & 0xFFFFFFFF, step & 0xFFFFFFFF, n0, out.ctypes.data, n, NTHREADS)
    return out


def sintable(pw, ow):
    t = np.empty(1 << pw, dtype=np.uint32)
    assert lib().zo_sintable_build(pw, ow, t.ctypes.data) == 0
    return t


def quarterwav(pw, ow):
    t = np.empty(1 << (pw - 2), dtype=np.uint32)
    assert lib().zo_quarterwav_build(pw, ow, t.ctypes.data) == 0
    return t


def lut_sin(pw, ow, tbl, phase32):
    phase32 = _c(phase32, np.uint32)
    out = np.empty(phase32.size, dtype=np.int32)
    lib().zo_lut_sin(pw, ow, tbl.ctypes.data, phase32.ctypes.data, out.ctypes.data, phase32.size, NTHREADS)
    return out


def lut_qwav(pw, ow, tbl, phase32):
    phase32 = _c(phase32, np.uint32)
    out = np.empty(phase32.size, dtype=np.int32)
    lib().zo_lut_qwav(pw, ow, tbl.ctypes.data, phase32.ctypes.data, out.ctypes.data, phase32.size, NTHREADS)
    return out


def hex_load(path, maxwords):
    w = np.zeros(maxwords, dtype=np.uint32)
    n = lib().zo_hex_load(path.encode(), w.ctypes.data, maxwords)
    assert n >= 0, n
    return w[:n]


def derive_qtbl(iw=0, ow=0, xtra=2, pw=0):
    q = ZoQuadTbl()
    rc = lib().zo_derive_qtbl(iw or 0, ow or 0, xtra, pw or 0, ctypes.byref(q))
    return rc, q


def quadtbl(q, phase_port):
    """phase_port: the PW-bit i_phase words (not 32-bit NCO words)."""
    phase_port = _c(phase_port, np.uint32)
    out = np.empty(phase_port.size, dtype=np.int32)
    lib().zo_quadtbl_batch(ctypes.byref(q), phase_port.ctypes.data, out.ctypes.data, phase_port.size, NTHREADS)
    return out
