"""CPU tier: the product's host side -- parameter derivation and LUT construction against the
real generator's output, the C ABI's symbol table against include/zcordic.h, and error
behaviour.  No compute call is made (there is no GPU here and no CPU fallback to call)."""
import ctypes
import hashlib
import json
import os
import re

import numpy as np
import pytest

import cordic_b200 as zc
from . import zo
from .conftest import ROOT
from .test_oracle_golden import (LUTS, PARAMS, QTBLS, SEQS, check_against_generator, check_quadtbl_against_generator,
                                 check_seq_against_generator)


@pytest.mark.parametrize("name", sorted(PARAMS))
def test_derive_matches_generator(name):
    g = PARAMS[name]
    a = g["args"]
    derive = zc.derive_p2r if g["mode"] == "p2r" else zc.derive_r2p
    p = derive(a["iw"], a["ow"], 2 if a["xtra"] is None else a["xtra"], a["pw"], a["nstages"])
    assert p.mode == (zc.MODE_P2R if g["mode"] == "p2r" else zc.MODE_R2P)
    check_against_generator(name, p, g["mode"])


@pytest.mark.parametrize("name", sorted(PARAMS))
def test_derive_matches_oracle_bitwise(name):
    g = PARAMS[name]
    a = g["args"]
    x = 2 if a["xtra"] is None else a["xtra"]
    if g["mode"] == "p2r":
        p, (rc, o) = zc.derive_p2r(a["iw"], a["ow"], x, a["pw"], a["nstages"]), zo.derive_p2r(a["iw"], a["ow"], x, a["pw"], a["nstages"])
    else:
        p, (rc, o) = zc.derive_r2p(a["iw"], a["ow"], x, a["pw"], a["nstages"]), zo.derive_r2p(a["iw"], a["ow"], x, a["pw"], a["nstages"])
    assert rc == 0
    for f in ("iw", "ow", "nextra", "ww", "pw", "nstages"):
        assert getattr(p, f) == getattr(o, f)
    for f in ("gain", "cordic_gain", "qvar", "pvar_rad", "best_cnr"):
        assert getattr(p, f) == getattr(o, f), f       # same libm, same order of operations
    assert list(p.angle) == list(o.angle)


@pytest.mark.parametrize("name", sorted(SEQS))
def test_sequential_derive_matches_generator(name):
    """zc_derive_sp2r / zc_derive_sr2p against `gencordic -t sp2r|sr2p` (header, angle table, CLOCKS_PER_OUTPUT)."""
    g = SEQS[name]
    a = g["args"]
    p = (zc.derive_sp2r if g["mode"] == "sp2r" else zc.derive_sr2p)(a["iw"], a["ow"], a["xtra"], a["pw"], a["nstages"])
    assert p.seq == 1 and p.mode == (zc.MODE_P2R if g["mode"] == "sp2r" else zc.MODE_R2P)
    check_seq_against_generator(name, p, zc.lib().zc_clocks_per_output(ctypes.byref(p)))
    assert p.header()["CLOCKS_PER_OUTPUT"] == g["clocks_per_output"]
    want_iters = p.nstages - 2 if g["mode"] == "sp2r" else p.nstages
    assert zc.lib().zc_iterations(ctypes.byref(p)) == want_iters


def test_sequential_configurations_the_reference_cannot_run():
    """sr2p with NSTAGES+1 a power of two never raises o_done in the reference's RTL (vector file:
    sr2p_n15_never_done); sp2r takes its output two iterations early and needs three stages."""
    for n in (15, 31):
        with pytest.raises(zc.ZcError) as e:
            zc.derive_sr2p(10, 10, 2, 0, n)
        assert e.value.code == -2 and "o_done never rises" in str(e.value)      # ZC_ERANGE
    assert zc.derive_sr2p(10, 10, 2, 0, 14).seq == 1
    with pytest.raises(zc.ZcError):
        zc.derive_sp2r(10, 10, 2, 0, 2)
    pipe = zc.derive_p2r(13, 13, 2)
    assert pipe.seq == 0 and zc.lib().zc_clocks_per_output(ctypes.byref(pipe)) == 1
    assert zc.lib().zc_iterations(ctypes.byref(pipe)) == pipe.nstages


@pytest.mark.parametrize("name", sorted(SEQS))
def test_zcordic_gen_sequential_header_matches_generator(name, tmp_path):
    g = SEQS[name]
    a = g["args"]
    fname = "seqcordic.v" if g["mode"] == "sp2r" else "seqpolar.v"
    args = ["-ca", "-t", g["mode"], "-f", fname]
    for flag, key in (("-i", "iw"), ("-o", "ow"), ("-x", "xtra"), ("-p", "pw"), ("-n", "nstages")):
        if a[key] is not None:
            args += [flag, str(a[key])]
    r = _gen(args, str(tmp_path))
    assert r.returncode == 0, r.stderr
    text = open(os.path.join(str(tmp_path), fname[:-2] + ".h")).read()
    got = {m.group(2): m.group(3).strip() for m in re.finditer(r"const\s+(int|double|bool)\s+(\w+)\s*=\s*([^;]+);", text)}
    assert got == g["header"]
    assert int(re.search(r"#define\s+CLOCKS_PER_OUTPUT\s+(\d+)", text).group(1)) == g["clocks_per_output"]


@pytest.mark.parametrize("name", sorted(LUTS))
def test_lut_build_matches_generator(name):
    g = LUTS[name]
    a = g["args"]
    derive = zc.derive_qtr if g["mode"] == "qtr" else zc.derive_tbl
    assert derive(a["iw"], a["pw"], a["ow"]) == (g["pw"], g["ow"])
    tbl = (zc.build_quarterwav if g["mode"] == "qtr" else zc.build_sintable)(g["pw"], g["ow"])
    assert tbl.size == g["nwords"]
    assert hashlib.sha256(tbl.astype("<u4").tobytes()).hexdigest() == g["sha256_le_u32"]
    assert [int(v) for v in tbl[::g["stride"]]] == g["samples"]


@pytest.mark.parametrize("name", sorted(QTBLS))
def test_quadtbl_derive_matches_generator_and_oracle(name):
    a = QTBLS[name]["args"]
    x = 2 if a["xtra"] is None else a["xtra"]
    q = zc.derive_qtbl(a["iw"], a["ow"], x, a["pw"])
    check_quadtbl_against_generator(name, q)
    rc, o = zo.derive_qtbl(a["iw"], a["ow"], x, a["pw"])
    assert rc == 0 and (q.itbl_err, q.tbl_err, q.spurdb) == (o.itbl_err, o.tbl_err, o.spurdb)


def test_generator_limits_are_enforced():
    # sw/sintable.cpp:62 refuses tables of 2^24 and up; :190 refuses quarter tables of 2^26 and up
    with pytest.raises(zc.ZcError) as e:
        zc.derive_tbl(pw=24, ow=12)
    assert e.value.code == -2
    with pytest.raises(zc.ZcError):
        zc.derive_qtr(pw=26, ow=12)
    with pytest.raises(zc.ZcError):
        zc.derive_qtr(pw=2, ow=12)
    # beyond what the engine's 32-bit lanes hold
    with pytest.raises(zc.ZcError) as e:
        zc.derive_p2r(iw=30, ow=30, xtra=4)
    assert e.value.code == -2
    with pytest.raises(zc.ZcError):
        zc.derive_r2p(iw=28, ow=28, xtra=2)


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "zcordic.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_functions()
    assert len(names) >= 25
    L = ctypes.CDLL(zc.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libzcordic.so does not export %s" % n
    # and the Python binding covers the same set
    assert names == zc.EXPORTED_SYMBOLS


def test_exchange_library_exports_every_declared_symbol():
    """include/zcordic_nccl.h -> cordic_b200/libzcordic_nccl.so (the only part that links NCCL): loads without a GPU,
    exports exactly what the header declares, refuses bad arguments, and has no CPU path either."""
    import torch          # first: torch must bind its own bundled libnccl.so.2 (2.28) before the system one (2.27) gets loaded
    src = open(os.path.join(ROOT, "include", "zcordic_nccl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(zc_[a-z0-9_]+)\s*\(", src)))
    assert names == ["zc_exchange_create", "zc_exchange_destroy", "zc_scatter_rotate_gather"]
    path = os.path.join(ROOT, "cordic_b200", "libzcordic_nccl.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cordic_b200", "vshim"), path])
    ctypes.CDLL(zc.LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    L = ctypes.CDLL(path)
    for n in names:
        assert hasattr(L, n), "libzcordic_nccl.so does not export %s" % n
    L.zc_exchange_create.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int, ctypes.c_size_t,
                                     ctypes.POINTER(ctypes.c_void_p)]
    h = ctypes.c_void_p()
    devs = (ctypes.c_int * 1)(0)
    assert L.zc_exchange_create(None, 1, 0, 1024, ctypes.byref(h)) == -1            # ZC_EINVAL
    assert L.zc_exchange_create(devs, 1, 7, 1024, ctypes.byref(h)) == -1            # unknown transport
    if not torch.cuda.is_available():
        assert L.zc_exchange_create(devs, 1, 0, 1024, ctypes.byref(h)) < 0          # no device: no fallback
    L.zc_scatter_rotate_gather.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    assert L.zc_scatter_rotate_gather(None, None, 0, 0, None, None, 0, 1) == -1


def test_header_is_plain_c():
    """The ABI header must compile as C (no torch / C++ types)."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write('#include "zcordic.h"\n#include "zcordic_nccl.h"\nint main(void){zc_params p; (void)p; return sizeof(zc_params)==%d?0:1;}\n'
                           % ctypes.sizeof(zc.Params))
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c,
                               "-o", os.path.join(td, "t")])
        assert subprocess.call([os.path.join(td, "t")]) == 0


def test_argument_errors():
    L = zc.lib()
    p = zc.derive_p2r(18, 18, 2, 24, 20)
    assert L.zc_derive_p2r(18, 18, 2, 24, 20, None) == -1
    # a vectoring configuration handed to the rotation entry point
    r = zc.derive_r2p(16, 16, 2)
    assert L.zc_rotate_const(ctypes.byref(r), 1, 0, None, None, 0, 0, None) == -1
    assert b"ZC_MODE_P2R" in L.zc_last_error()
    # NULL buffers with n>0
    assert L.zc_rotate_const(ctypes.byref(p), 1, 0, None, None, 16, 0, None) == -1
    assert L.zc_strerror(-2) == b"configuration out of range"
    assert L.zc_version() >= 1


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the compute entry points must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = zc.lib()
    p = zc.derive_p2r(18, 18, 2, 24, 20)
    ph = np.zeros(16, dtype=np.uint32)
    out = np.full(32, 12345, dtype=np.int32)
    rc = L.zc_rotate_const_host(ctypes.byref(p), 131071, 0, ph.ctypes.data, out.ctypes.data, 16, 0)
    assert rc in (-4, -3)
    assert (out == 12345).all()
    assert L.zc_device_count() <= 0 or rc != 0


def _gen(args, cwd):
    import subprocess
    exe = os.path.join(ROOT, "cordic_b200", "zcordic_gen")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cordic_b200", "vshim"), exe])
    return subprocess.run([exe] + args, cwd=cwd, capture_output=True, text=True)


@pytest.mark.parametrize("name", sorted(PARAMS))
def test_zcordic_gen_header_matches_generator(name, tmp_path):
    """zcordic_gen takes the reference generator's flags and must write the same constant set."""
    g = PARAMS[name]
    a = g["args"]
    fname = "cordic.v" if g["mode"] == "p2r" else "topolar.v"
    args = ["-ca", "-t", g["mode"], "-f", fname]
    for flag, key in (("-i", "iw"), ("-o", "ow"), ("-x", "xtra"), ("-p", "pw"), ("-n", "nstages")):
        if a[key] is not None:
            args += [flag, str(a[key])]
    r = _gen(args, str(tmp_path))
    assert r.returncode == 0, r.stderr
    got = {}
    for line in open(os.path.join(str(tmp_path), fname[:-2] + ".h")):
        m = re.match(r"const\s+(int|double|bool)\s+(\w+)\s*=\s*([^;]+);", line)
        if m:
            got[m.group(2)] = m.group(3).strip()
    assert got == g["header"]


@pytest.mark.parametrize("name", sorted(LUTS))
def test_zcordic_gen_hex_matches_generator(name, tmp_path):
    g = LUTS[name]
    a = g["args"]
    fname = "sintable.v" if g["mode"] == "tbl" else "quarterwav.v"
    args = ["-t", g["mode"], "-f", fname]
    for flag, key in (("-i", "iw"), ("-p", "pw"), ("-o", "ow")):
        if a[key] is not None:
            args += [flag, str(a[key])]
    r = _gen(args, str(tmp_path))
    assert r.returncode == 0, r.stderr
    hexfile = os.path.join(str(tmp_path), fname[:-2] + ".hex")
    words = zo.hex_load(hexfile, g["nwords"] + 8)
    assert words.size == g["nwords"]
    assert hashlib.sha256(words.astype("<u4").tobytes()).hexdigest() == g["sha256_le_u32"]
    if g["mode"] == "tbl" and (g["pw"], g["ow"]) == (17, 13) and os.path.exists("/root/reference/rtl/sintable.hex"):
        assert open(hexfile, "rb").read() == open("/root/reference/rtl/sintable.hex", "rb").read()


@pytest.mark.parametrize("name", sorted(QTBLS))
def test_zcordic_gen_quadtbl_matches_generator(name, tmp_path):
    g = QTBLS[name]
    a = g["args"]
    args = ["-ca", "-t", "qtbl", "-f", "quadtbl.v"]
    for flag, key in (("-i", "iw"), ("-o", "ow"), ("-x", "xtra"), ("-p", "pw")):
        if a[key] is not None:
            args += [flag, str(a[key])]
    r = _gen(args, str(tmp_path))
    assert r.returncode == 0, r.stderr
    got = {}
    for line in open(os.path.join(str(tmp_path), "quadtbl.h")):
        m = re.match(r"const\s+(int|long|double|bool)\s+(\w+)\s*=\s*([^;]+);", line)
        if m:
            got[m.group(2)] = m.group(3).strip()
    assert got == g["header"]
    for nm in "clq":
        words = zo.hex_load(os.path.join(str(tmp_path), "quadtbl_%stbl.hex" % nm), 4096)
        assert words.tolist() == g[nm + "tbl"]


def test_hex_roundtrip_and_reference_layout(tmp_path):
    """zc_hex_write produces the byte layout of sw/hexfile.cpp:78-89 and zc_hex_read parses it back."""
    tbl = zc.build_sintable(17, 13)
    path = os.path.join(str(tmp_path), "sintable.hex")
    zc.hex_write(path, tbl, 13)
    assert np.array_equal(zc.hex_read(path), tbl)
    if os.path.exists("/root/reference/rtl/sintable.hex"):
        assert open(path, "rb").read() == open("/root/reference/rtl/sintable.hex", "rb").read()
        assert np.array_equal(zc.hex_read("/root/reference/rtl/quarterwav.hex"), zc.build_quarterwav(18, 24))


def test_cpp_client_builds_and_refuses_to_run_without_a_gpu():
    """cordic_b200/zcordic_bench: a C++ program over the C ABI (no Python in the process).  Without a CUDA device it must
    say so and exit 3 -- there is no CPU path to fall back to."""
    import shutil
    import subprocess
    exe = os.path.join(ROOT, "cordic_b200", "zcordic_bench")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cordic_b200", "vshim"), exe])
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: covered by tests/test_gpu_runtime.py::test_cpp_client")
    except ImportError:
        pass
    r = subprocess.run([exe, "-l", "10"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU path" in r.stderr, (r.returncode, r.stderr)
