"""GPU tier: the CUDA path, called through the C ABI (cordic_b200 binds it with ctypes), against
the CPU oracle on the same seeded inputs -- bit-exact, every word.  Marked ``gpu``."""
import ctypes
import json
import os

import numpy as np
import pytest

import cordic_b200 as zc
from . import zo
from .conftest import ROOT

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

SEED = int(os.environ.get("ZC_TEST_SEED", "20261017"))       # another seed: a soak run of the randomised tests
KATS = json.load(open(os.path.join(ROOT, "tests", "golden", "survey_kats.json")))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def both_p2r(iw=0, ow=0, xtra=2, pw=0, n=0):
    core = zc.Cordic(iw, ow, xtra, pw, n)
    rc, op = zo.derive_p2r(iw, ow, xtra, pw, n)
    assert rc == 0
    return core, op


def both_r2p(iw=0, ow=0, xtra=2, pw=0, n=0):
    core = zc.Topolar(iw, ow, xtra, pw, n)
    rc, op = zo.derive_r2p(iw, ow, xtra, pw, n)
    assert rc == 0
    return core, op


P2R_CONFIGS = {
    "shipped": dict(iw=13, ow=13, xtra=2),                      # rtl/cordic.h: WW16 PW20 N16
    "cfg0": dict(iw=16, ow=16, xtra=2, pw=16),                  # BASELINE configs[0]
    "cfg1": dict(iw=18, ow=18, xtra=2, pw=24, n=20),            # BASELINE configs[1]
    "8_8_x0": dict(iw=8, ow=8, xtra=0),
    "12_16_x1": dict(iw=12, ow=16, xtra=1),
    "20_10_p18": dict(iw=20, ow=10, xtra=2, pw=18),
    "manystages": dict(iw=6, ow=6, xtra=2, pw=10, n=30),         # zero-angle / i>=WW pass-through stages
    "negx": dict(iw=10, ow=10, xtra=-5),                        # WW=IW+1: no zero padding, no rounding... D=1
}


@pytest.mark.parametrize("flags", [zc.F_DEFAULT, zc.F_NO_SEED, zc.F_FORCE_GENERIC, zc.F_FORCE_SEED, zc.F_FORCE_SEED | zc.F_NO_DP2A,
                                   zc.F_FORCE_SEED | zc.F_SEED_PACKED, zc.F_FORCE_SEED | zc.F_SEED_PACKED | zc.F_NO_DP2A,
                                   zc.F_FORCE_SEED | zc.F_SEED_REGS])
@pytest.mark.parametrize("name", sorted(P2R_CONFIGS))
def test_rotate_const_full_phase_sweep(name, flags):
    """The sweep of bench/cpp/cordic_tb.cpp:127-178: every one of the 2^PW phases, full-scale
    input (x,y) = (2^(IW-1)-1, 0) (:68-69).  Bit-exact against the oracle."""
    core, op = both_p2r(**P2R_CONFIGS[name])
    n = 1 << core.PW
    phase = np.arange(n, dtype=np.uint32)
    x0 = (1 << (core.IW - 1)) - 1
    got = host(core.rotate_const(x0, 0, dev(phase), flags=flags))
    want = zo.rotate_const(op, x0, 0, phase)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("flags", [zc.F_DEFAULT, zc.F_FORCE_SEED, zc.F_SEED_PACKED, zc.F_SEED_REGS])
def test_rotate_const_other_vectors_and_random_phase(flags):
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    rng = np.random.default_rng(SEED)
    phase = rng.integers(0, 1 << 32, size=(1 << 20) + 3, dtype=np.uint64).astype(np.uint32)  # high bits are ignored
    for x0, y0 in [(131071, 0), (-131072, -131072), (131071, 131071), (0, 0), (-1, 1), (12345, -54321),
                   (0x7FFFFFFF, 0x40000)]:   # last: bits above IW must be dropped like the port would
        got = host(core.rotate_const(x0, y0, dev(phase), flags=flags))
        want = zo.rotate_const(op, x0, y0, phase)
        assert np.array_equal(got, want), (x0, y0)


def test_rotate_const_auto_selected_table_flavour():
    """Streams of >= 4 Mi phases are probed on the device: a sweep must take the word-table kernel, scattered
    phases the byte-table kernel; either way the result is the oracle's, and two table launches are enqueued
    (the chosen kernel + the one that returns at its probe)."""
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    rng = np.random.default_rng(SEED + 31)
    n = (1 << 22) + 5
    for phase in (np.arange(n, dtype=np.uint32) & 0xFFFFFF,
                  rng.integers(0, 1 << 24, size=n, dtype=np.uint64).astype(np.uint32),
                  ((np.arange(n, dtype=np.uint64) * 3) & 0xFFFFFF).astype(np.uint32)):
        before = zc.launch_count()
        got = host(core.rotate_const(131071, 0, dev(phase)))
        # 2 seeded launches (one returns at its probe) + the 5-sample rest: one group of 4 on the plain
        # fast kernel and one sample on the generic kernel
        assert zc.launch_count() - before == 4
        assert np.array_equal(got, zo.rotate_const(op, 131071, 0, phase))


@pytest.mark.parametrize("name", sorted(P2R_CONFIGS))
@pytest.mark.parametrize("flags", [zc.F_DEFAULT, zc.F_FORCE_GENERIC, zc.F_FORCE_SEED, zc.F_FORCE_SEED | zc.F_NO_DP2A, zc.F_NO_SEED])
def test_rotate_per_sample_inputs(name, flags):
    core, op = both_p2r(**P2R_CONFIGS[name])
    rng = np.random.default_rng(SEED + 1)
    n = (1 << 19) + 1
    lo, hi = -(1 << (core.IW - 1)), (1 << (core.IW - 1)) - 1
    xy = rng.integers(lo, hi + 1, size=(n, 2), dtype=np.int64).astype(np.int32)
    # edge vectors first: all four corners, axes, zero, +-1
    edges = [(lo, lo), (lo, hi), (hi, lo), (hi, hi), (0, 0), (hi, 0), (0, hi), (lo, 0), (0, lo), (1, 0), (-1, 0), (0, -1)]
    xy[:len(edges)] = edges
    phase = rng.integers(0, 1 << core.PW, size=n, dtype=np.uint64).astype(np.uint32)
    phase[:8] = [0, 1, (1 << core.PW) - 1, 1 << (core.PW - 1), 1 << (core.PW - 2), 1 << (core.PW - 3),
                 (1 << (core.PW - 3)) - 1, 7 << (core.PW - 3)]
    got = host(core.rotate(dev(xy), dev(phase), flags=flags))
    want = zo.rotate(op, xy, phase)
    assert np.array_equal(got, want)


def test_rotate_extreme_inputs_every_octant_boundary():
    """Corner inputs x every octant boundary +-2: where a missing guard bit or a wrong
    pre-rotation constant (rtl/cordic.v:143-178) would show."""
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    lo, hi = -(1 << 17), (1 << 17) - 1
    corners = [(lo, lo), (lo, hi), (hi, lo), (hi, hi), (hi, 0), (lo, 0), (0, hi), (0, lo)]
    ph = []
    for o in range(8):
        for d in (-2, -1, 0, 1, 2):
            ph.append(((o << 21) + d) & 0xFFFFFF)
    xy = np.array([c for c in corners for _ in ph], dtype=np.int32)
    phase = np.array(ph * len(corners), dtype=np.uint32)
    got = host(core.rotate(dev(xy), dev(phase)))
    assert np.array_equal(got, zo.rotate(op, xy, phase))


@pytest.mark.parametrize("flags", [zc.F_DEFAULT, zc.F_FORCE_GENERIC])
def test_wide_cores(flags):
    """The widest cores the 32-bit lanes hold: p2r 24/24 (WW27 PW31 N27), r2p 20/20 (WW28 PW28 N25), r2p 24/24
    (WW32 PW32 N29: the working registers fill the whole lane, wrap = native overflow) and r2p 22/24 (WW30 PW31)."""
    rng = np.random.default_rng(SEED + 21)
    n = (1 << 20) + 3
    core, op = both_p2r(iw=24, ow=24, xtra=2)
    assert (core.WW, core.PW, core.NSTAGES) == (27, 31, 27)
    xy = rng.integers(-(1 << 23), 1 << 23, size=(n, 2), dtype=np.int64).astype(np.int32)
    xy[:4] = [(-(1 << 23), -(1 << 23)), ((1 << 23) - 1, (1 << 23) - 1), ((1 << 23) - 1, -(1 << 23)), (0, 0)]
    phase = rng.integers(0, 1 << 31, size=n, dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(host(core.rotate(dev(xy), dev(phase), flags=flags)), zo.rotate(op, xy, phase))
    assert np.array_equal(host(core.rotate(dev(xy), dev(phase), flags=flags | zc.F_FORCE_SEED)), zo.rotate(op, xy, phase))
    assert np.array_equal(host(core.rotate_const((1 << 23) - 1, 0, dev(phase), flags=flags)),
                          zo.rotate_const(op, (1 << 23) - 1, 0, phase))
    for kw, lim in ((dict(iw=20, ow=20, xtra=2), 19), (dict(iw=24, ow=24, xtra=2), 23), (dict(iw=22, ow=24, xtra=1), 21)):
        vcore, vop = both_r2p(**kw)
        v = rng.integers(-(1 << lim), 1 << lim, size=(n, 2), dtype=np.int64).astype(np.int32)
        lo, hi = -(1 << lim), (1 << lim) - 1
        v[:6] = [(lo, lo), (hi, hi), (lo, hi), (hi, lo), (0, 0), (hi, 0)]
        mag, ph = vcore.topolar(dev(v), flags=flags)
        wm, wp = zo.topolar(vop, v)
        assert np.array_equal(host(mag), wm), kw
        assert np.array_equal(host(ph).view(np.uint32), wp), kw


def test_fast_path_exactness_proof_exhaustively_on_a_small_core():
    """The host only takes the non-wrapping fast kernels after bounding the register growth
    (zc_api.cu: fast_path_is_exact).  For a 6-bit core (WW=9, PW=10, 30 nominal stages) that bound holds with
    little margin, and every one of the 64*64*1024 (x, y, phase) inputs can be checked against the oracle,
    which models the WW-bit wrap."""
    core, op = both_p2r(iw=6, ow=6, xtra=2, pw=10, n=30)
    assert core.WW == 9
    v = np.arange(-32, 32, dtype=np.int32)
    ph = np.arange(1 << 10, dtype=np.uint32)
    X, Y, P = np.meshgrid(v, v, ph, indexing="ij")
    xy = np.stack([X.ravel(), Y.ravel()], axis=1).astype(np.int32)
    phase = P.ravel().astype(np.uint32)
    want = zo.rotate(op, xy, phase)
    assert np.array_equal(host(core.rotate(dev(xy), dev(phase))), want)                       # fast path
    assert np.array_equal(host(core.rotate(dev(xy), dev(phase), flags=zc.F_FORCE_GENERIC)), want)
    vcore, vop = both_r2p(iw=8, ow=8, xtra=0)                                                  # WW=12
    v8 = np.arange(-128, 128, dtype=np.int32)
    xy8 = np.stack(np.meshgrid(v8, v8, indexing="ij"), axis=-1).reshape(-1, 2)
    wm, wp = zo.topolar(vop, xy8)
    for fl in (zc.F_DEFAULT, zc.F_FORCE_GENERIC):
        mag, ph8 = vcore.topolar(dev(xy8), flags=fl)
        assert np.array_equal(host(mag), wm) and np.array_equal(host(ph8).view(np.uint32), wp)


def test_random_configurations_differential():
    """Property test over the generator's parameter surface: random (iw, ow, xtra, pw, nstages), random inputs,
    every kernel-selection flag (fast / seeded word / seeded byte / seeded registers / generic) against the
    oracle.  Exercises the seed-plan geometry (bucket width, prefix depth, residual range) far from cfg1."""
    rng = np.random.default_rng(SEED + 41)
    n = (1 << 16) + 3
    tried = seeded_ok = 0
    while tried < 48:
        iw, ow = int(rng.integers(4, 25)), int(rng.integers(4, 25))
        xtra = int(rng.integers(-1, 5))
        pw = int(rng.choice([0, 0, int(rng.integers(8, 31))]))
        ns = int(rng.choice([0, 0, int(rng.integers(3, 33))]))
        rc, op = zo.derive_p2r(iw, ow, xtra, pw, ns)
        if rc != 0:
            continue
        try:
            core = zc.Cordic(iw, ow, xtra, pw, ns)
        except zc.ZcError:
            continue
        tried += 1
        phase = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
        lo, hi = -(1 << (iw - 1)), (1 << (iw - 1)) - 1
        x0, y0 = int(rng.integers(lo, hi + 1)), int(rng.integers(lo, hi + 1))
        want = zo.rotate_const(op, x0, y0, phase)
        for fl in (zc.F_DEFAULT, zc.F_FORCE_SEED | zc.F_SEED_WORDS, zc.F_FORCE_SEED | zc.F_SEED_WORDS | zc.F_NO_DP2A,
                   zc.F_FORCE_SEED | zc.F_SEED_PACKED, zc.F_FORCE_SEED | zc.F_SEED_PACKED | zc.F_NO_DP2A,
                   zc.F_FORCE_SEED | zc.F_SEED_REGS, zc.F_FORCE_GENERIC):
            got = host(core.rotate_const(x0, y0, dev(phase), flags=fl))
            assert np.array_equal(got, want), (iw, ow, xtra, pw, ns, fl)
        xy = rng.integers(lo, hi + 1, size=(n, 2), dtype=np.int64).astype(np.int32)
        wxy = zo.rotate(op, xy, phase)
        for fl in (zc.F_DEFAULT, zc.F_FORCE_SEED, zc.F_FORCE_SEED | zc.F_NO_DP2A):        # plain fast kernel / table-directed kernels
            assert np.array_equal(host(core.rotate(dev(xy), dev(phase), flags=fl)), wxy), (iw, ow, xtra, pw, ns, fl)
        got = host(core.nco(x0, y0, 12345, 0x9E3779B1, n, n0=7, flags=zc.F_FORCE_SEED))
        assert np.array_equal(got, zo.nco(op, x0, y0, 12345, 0x9E3779B1, n, n0=7)), (iw, ow, xtra, pw, ns)
    tried = 0
    while tried < 24:
        iw, ow = int(rng.integers(4, 23)), int(rng.integers(4, 23))
        xtra = int(rng.integers(-2, 4))
        pw = int(rng.choice([0, 0, int(rng.integers(8, 31))]))
        ns = int(rng.choice([0, 0, int(rng.integers(3, 33))]))
        rc, op = zo.derive_r2p(iw, ow, xtra, pw, ns)
        if rc != 0:
            continue
        try:
            vcore = zc.Topolar(iw, ow, xtra, pw, ns)
        except zc.ZcError:
            continue
        tried += 1
        lo, hi = -(1 << (iw - 1)), (1 << (iw - 1)) - 1
        xy = rng.integers(lo, hi + 1, size=(n, 2), dtype=np.int64).astype(np.int32)
        wm, wp = zo.topolar(op, xy)
        for fl in (zc.F_DEFAULT, zc.F_FORCE_GENERIC):
            mag, ph = vcore.topolar(dev(xy), flags=fl)
            assert np.array_equal(host(mag), wm) and np.array_equal(host(ph).view(np.uint32), wp), (iw, ow, xtra, pw, ns, fl)


SEQ_P2R = {
    "shipped": dict(iw=13, ow=13, xtra=2),                       # rtl/seqcordic.v: NSTAGES 16, 14 iterations count
    "cfg1": dict(iw=18, ow=18, xtra=2, pw=24, n=20),
    "manystages": dict(iw=6, ow=6, xtra=2, pw=10, n=30),         # zero angles and shifts >= WW still rotate (x, y)
    "n64": dict(iw=10, ow=10, xtra=2, pw=16, n=64),              # more iterations than the fast kernels unroll
}


@pytest.mark.parametrize("flags", [zc.F_DEFAULT, zc.F_NO_SEED, zc.F_FORCE_GENERIC, zc.F_FORCE_SEED,
                                   zc.F_FORCE_SEED | zc.F_SEED_PACKED])
@pytest.mark.parametrize("name", sorted(SEQ_P2R))
def test_sequential_rotation_full_phase_sweep(name, flags):
    """zc_params.seq = 1 (rtl/seqcordic.v): every phase, constant and per-sample inputs, every kernel choice."""
    kw = SEQ_P2R[name]
    core = zc.Cordic(kw["iw"], kw["ow"], kw["xtra"], kw.get("pw", 0), kw.get("n", 0), sequential=True)
    rc, op = zo.derive_sp2r(kw["iw"], kw["ow"], kw["xtra"], kw.get("pw", 0), kw.get("n", 0))
    assert rc == 0 and core.params.seq == 1
    n = 1 << core.PW
    phase = np.arange(n, dtype=np.uint32)
    x0 = (1 << (core.IW - 1)) - 1
    assert np.array_equal(host(core.rotate_const(x0, 0, dev(phase), flags=flags)), zo.rotate_const(op, x0, 0, phase))
    rng = np.random.default_rng(SEED + 77)
    m = min(n, 1 << 20)
    xy = rng.integers(-(1 << (core.IW - 1)), 1 << (core.IW - 1), size=(m, 2), dtype=np.int64).astype(np.int32)
    assert np.array_equal(host(core.rotate(dev(xy), dev(phase[:m]), flags=flags)), zo.rotate(op, xy, phase[:m]))
    # and it is not the pipelined core's function
    rcp, pp = zo.derive_p2r(kw["iw"], kw["ow"], kw["xtra"], kw.get("pw", 0), kw.get("n", 0))
    assert not np.array_equal(zo.rotate_const(pp, x0, 0, phase), zo.rotate_const(op, x0, 0, phase))


@pytest.mark.parametrize("flags", [zc.F_DEFAULT, zc.F_FORCE_GENERIC])
@pytest.mark.parametrize("kw", [dict(iw=13, ow=13, xtra=2), dict(iw=16, ow=16, xtra=2), dict(iw=10, ow=10, xtra=2, pw=14, n=20),
                                dict(iw=8, ow=8, xtra=0), dict(iw=10, ow=10, xtra=2, pw=12, n=40)])
def test_sequential_vectoring(kw, flags):
    """zc_params.seq = 1 (rtl/seqpolar.v): all NSTAGES iterations run, zero angles included."""
    core = zc.Topolar(kw["iw"], kw["ow"], kw["xtra"], kw.get("pw", 0), kw.get("n", 0), sequential=True)
    rc, op = zo.derive_sr2p(kw["iw"], kw["ow"], kw["xtra"], kw.get("pw", 0), kw.get("n", 0))
    assert rc == 0
    rng = np.random.default_rng(SEED + 78)
    n = (1 << 20) + 3
    lo, hi = -(1 << (core.IW - 1)), (1 << (core.IW - 1)) - 1
    xy = rng.integers(lo, hi + 1, size=(n, 2), dtype=np.int64).astype(np.int32)
    xy[:4] = [[lo, lo], [hi, hi], [0, 0], [lo, hi]]
    mag, ph = core.topolar(dev(xy), flags=flags)
    wm, wp = zo.topolar(op, xy)
    assert np.array_equal(host(mag), wm) and np.array_equal(host(ph).view(np.uint32), wp)


def test_rotate_narrow_core_wraps_like_the_rtl():
    """WW=5 is too narrow for the CORDIC gain: the RTL registers wrap.  The engine must detect
    that its non-wrapping fast path is not provably exact and reproduce the wrap."""
    core, op = both_p2r(iw=4, ow=4, xtra=0, pw=8, n=6)
    assert core.WW == 5
    xs = np.arange(-8, 8, dtype=np.int32)
    xy = np.array([(x, y) for x in xs for y in xs for _ in range(256)], dtype=np.int32)
    phase = np.tile(np.arange(256, dtype=np.uint32), 256)
    got = host(core.rotate(dev(xy), dev(phase)))
    assert np.array_equal(got, zo.rotate(op, xy, phase))
    got = host(core.rotate_const(-8, -8, dev(phase)))
    assert np.array_equal(got, zo.rotate_const(op, -8, -8, phase))


R2P_CONFIGS = {
    "shipped": dict(iw=13, ow=13, xtra=2),              # rtl/topolar.h: WW21 PW21 N18
    "cfg2": dict(iw=16, ow=16, xtra=2),                 # BASELINE configs[2]: WW24 PW24 N21
    "8_8_x0": dict(iw=8, ow=8, xtra=0),
    "12_16_x1": dict(iw=12, ow=16, xtra=1),
    "10_10_p14_n20": dict(iw=10, ow=10, xtra=2, pw=14, n=20),
}


def tb_circle(iw, pw, n):
    """Stimulus of bench/cpp/topolar_tb.cpp:133-147: a full-scale circle, two revolutions."""
    i = np.arange(n, dtype=np.int64)
    ip = ((i << 1) & 0xFFFFFFFF).astype(np.uint32).view(np.int32).astype(np.float64) if pw == 32 else \
        ((i << 1)).astype(np.int32).astype(np.float64)
    ph = ip * np.pi / float(1 << (pw - 1))
    mg = float((1 << (iw - 1)) - 1)
    return np.stack([(mg * np.cos(ph)).astype(np.int32), (mg * np.sin(ph)).astype(np.int32)], axis=1)


@pytest.mark.parametrize("flags", [zc.F_DEFAULT, zc.F_FORCE_GENERIC])
@pytest.mark.parametrize("name", sorted(R2P_CONFIGS))
def test_topolar_testbench_circle(name, flags):
    core, op = both_r2p(**R2P_CONFIGS[name])
    xy = tb_circle(core.IW, core.PW, 1 << core.PW)
    mag, ph = core.topolar(dev(xy), flags=flags)
    wm, wp = zo.topolar(op, xy)
    assert np.array_equal(host(mag), wm)
    assert np.array_equal(host(ph).view(np.uint32), wp)


@pytest.mark.parametrize("name", sorted(R2P_CONFIGS))
def test_topolar_random_and_edges(name):
    core, op = both_r2p(**R2P_CONFIGS[name])
    rng = np.random.default_rng(SEED + 2)
    n = (1 << 20) + 2
    lo, hi = -(1 << (core.IW - 1)), (1 << (core.IW - 1)) - 1
    xy = rng.integers(lo, hi + 1, size=(n, 2), dtype=np.int64).astype(np.int32)
    edges = [(0, 0), (1, 0), (-1, 0), (0, 1), (0, -1), (hi, 0), (lo, 0), (0, hi), (0, lo), (hi, hi), (lo, lo),
             (hi, lo), (lo, hi), (1, 1), (-1, -1), (1, -1), (-1, 1)]
    xy[:len(edges)] = edges
    mag, ph = core.topolar(dev(xy))
    wm, wp = zo.topolar(op, xy)
    assert np.array_equal(host(mag), wm)
    assert np.array_equal(host(ph).view(np.uint32), wp)


def test_topolar_exhaustive_8bit():
    core, op = both_r2p(iw=8, ow=8, xtra=0)
    v = np.arange(-128, 128, dtype=np.int32)
    xy = np.stack(np.meshgrid(v, v, indexing="ij"), axis=-1).reshape(-1, 2)
    mag, ph = core.topolar(dev(xy))
    wm, wp = zo.topolar(op, xy)
    assert np.array_equal(host(mag), wm)
    assert np.array_equal(host(ph).view(np.uint32), wp)


TAIL_CASES = [dict(iw=8, ow=8, xtra=2), dict(iw=10, ow=10, xtra=2), dict(iw=13, ow=13, xtra=2), dict(iw=16, ow=16, xtra=2),
              dict(iw=12, ow=20, xtra=1), dict(iw=18, ow=18, xtra=2), dict(iw=22, ow=20, xtra=1),
              dict(iw=16, ow=16, xtra=2, pw=30, n=28), dict(iw=16, ow=16, xtra=2, seq=True), dict(iw=8, ow=8, xtra=2, seq=True)]


@pytest.mark.parametrize("kw", TAIL_CASES, ids=lambda k: "_".join("%s%s" % kv for kv in k.items()))
def test_topolar_short_late_stages(kw):
    """The kernel runs its last zc_topolar_tail_stages() stages in a six-instruction form that is only valid because
    |y| has provably converged below the shift (tests/test_tail_stages.py checks the proof on the CPU).  Exhaustive
    inputs for IW <= 10, otherwise corners, random, near-axis and near-diagonal vectors; default (short) and
    ZC_F_NO_TAIL (every stage in full) must both equal the oracle."""
    from .test_tail_stages import _inputs
    seq = kw.get("seq", False)
    core = zc.Topolar(kw["iw"], kw["ow"], kw["xtra"], kw.get("pw", 0), kw.get("n", 0), sequential=seq)
    rc, op = (zo.derive_sr2p if seq else zo.derive_r2p)(kw["iw"], kw["ow"], kw["xtra"], kw.get("pw", 0), kw.get("n", 0))
    assert rc == 0
    tail = zc.lib().zc_topolar_tail_stages(ctypes.byref(core.params))
    assert tail >= 2, "this configuration should exercise the short form"
    xy = _inputs(core.IW, np.random.default_rng(SEED + 91))
    xy = xy[:(len(xy) // 4) * 4]                        # whole groups of four: all of it through the fast kernel
    wm, wp = zo.topolar(op, xy)
    for fl in (zc.F_DEFAULT, zc.F_NO_TAIL):
        mag, ph = core.topolar(dev(xy), flags=fl)
        assert np.array_equal(host(mag), wm), fl
        assert np.array_equal(host(ph).view(np.uint32), wp), fl


@pytest.mark.parametrize("kind,pw,ow", [("tbl", 17, 13), ("tbl", 10, 8), ("tbl", 23, 16), ("tbl", 20, 30),
                                        ("qtr", 18, 24), ("qtr", 12, 12), ("qtr", 25, 16), ("qtr", 3, 6)])
def test_lut_modes(kind, pw, ow):
    """rtl/sintable.v:71-75 and rtl/quarterwav.v:92-109 over every table index and random NCO words."""
    cls = zc.QuarterWav if kind == "qtr" else zc.SinTable
    core = cls(phase_bits=pw, ow=ow)
    assert (core.PW, core.OW) == (pw, ow)
    otbl = zo.quarterwav(pw, ow) if kind == "qtr" else zo.sintable(pw, ow)
    assert np.array_equal(core.table, otbl)
    rng = np.random.default_rng(SEED + 3)
    sweep = (np.arange(1 << min(pw, 22), dtype=np.uint64) << (32 - min(pw, 22))).astype(np.uint32)
    rnd = rng.integers(0, 1 << 32, size=(1 << 20) + 1, dtype=np.uint64).astype(np.uint32)
    for phase in (sweep, rnd, rnd[:5], rnd[:3]):
        got = host(core.lookup(dev(phase)))
        want = (zo.lut_qwav if kind == "qtr" else zo.lut_sin)(pw, ow, otbl, phase)
        assert np.array_equal(got, want)


QTBL_CONFIGS = {
    "shipped": dict(ow=13, pw=18),                     # rtl/quadtbl.h
    "o13_pauto": dict(ow=13),
    "o16_p20": dict(ow=16, pw=20),
    "o10_p14_x1": dict(ow=10, xtra=1, pw=14),
    "o20_p24": dict(ow=20, pw=24),
    "o8_p12": dict(ow=8, pw=12),
    "o16_p32": dict(ow=16, pw=32),                     # 25-bit dx: the 64-bit product path
}


@pytest.mark.parametrize("name", sorted(QTBL_CONFIGS))
def test_quadtbl_every_phase(name):
    """rtl/quadtbl.v over every phase (bench/cpp/quadtbl_tb.cpp:96-121 sweeps the same), ragged and random too."""
    cfg = QTBL_CONFIGS[name]
    core = zc.QuadTbl(ow=cfg["ow"], xtra=cfg.get("xtra", 2), phase_bits=cfg.get("pw", 0))
    rc, q = zo.derive_qtbl(0, cfg["ow"], cfg.get("xtra", 2), cfg.get("pw", 0))
    assert rc == 0 and core.PW == q.pw
    pw = q.pw
    rng = np.random.default_rng(SEED + 9)
    if pw <= 24:
        port = np.arange(1 << pw, dtype=np.uint32)
    else:
        port = rng.integers(0, 1 << pw, size=1 << 22, dtype=np.uint64).astype(np.uint32)
        port[:4] = [0, 1, (1 << pw) - 1, 1 << (pw - 1)]
    words = (port.astype(np.uint64) << (32 - pw)).astype(np.uint32) | rng.integers(0, 1 << (32 - pw), size=port.size, dtype=np.uint64).astype(np.uint32)
    got = host(core.lookup(dev(words)))
    assert np.array_equal(got, zo.quadtbl(q, port))
    for m in (1, 3, 5, 1023):            # ragged sizes, misaligned start
        got = host(core.lookup(dev(words)[1:1 + m]))
        assert np.array_equal(got, zo.quadtbl(q, port[1:1 + m]))
    out = np.empty(port.size, dtype=np.int32)
    core.lookup_host(words, out)
    assert np.array_equal(out, zo.quadtbl(q, port))


@pytest.mark.parametrize("flags", [zc.F_DEFAULT, zc.F_NO_SEED, zc.F_FORCE_GENERIC, zc.F_SEED_PACKED, zc.F_SEED_REGS, zc.F_SEED_WORDS,
                                   zc.F_SEED_WORDS | zc.F_NO_DP2A])
def test_nco_stream(flags):
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    n = (1 << 20) + 5
    for phase0, step, n0 in [(0, 0x01234567, 0), (0xDEADBEEF, 0xFFFFFFFF, 12345), (7, 1, (1 << 33) + 3)]:
        got = host(core.nco(131071, 0, phase0, step, n, n0=n0, flags=flags))
        want = zo.nco(op, 131071, 0, phase0, step, n, n0=n0)
        assert np.array_equal(got, want), (phase0, step, n0)


def test_nco_mixer():
    """zc_nco_mix == zc_rotate fed with the NCO's phases (and the oracle), including a ragged, offset stream."""
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    rng = np.random.default_rng(SEED + 11)
    n = (1 << 20) + 7
    xy = rng.integers(-(1 << 17), 1 << 17, size=(n, 2), dtype=np.int64).astype(np.int32)
    for phase0, step, n0 in [(0, 0x01234567, 0), (0xCAFEF00D, 0xFFFF0001, 999)]:
        phase = ((phase0 + (n0 + np.arange(n, dtype=np.uint64)) * step) & 0xFFFFFFFF).astype(np.uint32) >> 8
        want = zo.rotate(op, xy, phase)
        assert np.array_equal(host(core.mix(dev(xy), phase0, step, n0=n0)), want)
        assert np.array_equal(host(core.mix(dev(xy)[1:], phase0, step, n0=n0 + 1)), want[1:])


def test_nco_chunks_concatenate():
    """Rank r of G computes n in [r*N/G, (r+1)*N/G) from phase0 + n*step alone."""
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    n, parts = 1 << 18, 8
    whole = host(core.nco(131071, 0, 99, 0x01234567, n))
    pieces = [host(core.nco(131071, 0, 99, 0x01234567, n // parts, n0=r * (n // parts))) for r in range(parts)]
    assert np.array_equal(np.concatenate(pieces), whole)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 255, 1023, 4097])
def test_ragged_sizes_and_misaligned_buffers(n):
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    vcore, vop = both_r2p(**R2P_CONFIGS["cfg2"])
    rng = np.random.default_rng(SEED + n)
    phase = rng.integers(0, 1 << 24, size=n + 1, dtype=np.uint64).astype(np.uint32)
    xy = rng.integers(-32768, 32768, size=(n + 1, 2), dtype=np.int64).astype(np.int32)
    dphase, dxy = dev(phase), dev(xy)
    for off in (0, 1):          # off=1: pointers 4 / 8 bytes past a 16-byte boundary
        m = n + 1 - off if off else n
        ph_h, xy_h = phase[off:off + m], xy[off:off + m]
        ph_d, xy_d = dphase[off:off + m], dxy[off:off + m]
        if m == 0:
            assert core.rotate_const(131071, 0, ph_d).numel() == 0
            continue
        assert np.array_equal(host(core.rotate_const(131071, 0, ph_d)), zo.rotate_const(op, 131071, 0, ph_h))
        assert np.array_equal(host(core.rotate(xy_d, ph_d)), zo.rotate(op, xy_h, ph_h))
        mag, ph = vcore.topolar(xy_d)
        wm, wp = zo.topolar(vop, xy_h)
        assert np.array_equal(host(mag), wm) and np.array_equal(host(ph).view(np.uint32), wp)


def test_large_buffers_off_the_16_byte_grid():
    """Big streams whose pointers are only naturally aligned (4-byte phases, 8-byte pairs): the table kernels take
    them; the ragged rest goes through the generic kernel.  Same words as the oracle."""
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    rng = np.random.default_rng(SEED + 51)
    n = (1 << 20) + 131
    phase = rng.integers(0, 1 << 24, size=n + 1, dtype=np.uint64).astype(np.uint32)
    xy = rng.integers(-(1 << 17), 1 << 17, size=(n + 1, 2), dtype=np.int64).astype(np.int32)
    dphase, dxy = dev(phase), dev(xy)
    out = torch.empty((n + 1, 2), dtype=torch.int32, device="cuda")
    before = zc.launch_count()
    core.rotate_const(131071, 0, dphase[1:], out=out[1:])
    assert zc.launch_count() - before == 2                     # seeded kernel + generic rest
    assert np.array_equal(host(out[1:]), zo.rotate_const(op, 131071, 0, phase[1:]))
    core.rotate(dxy[1:], dphase[1:], out=out[1:])
    assert np.array_equal(host(out[1:]), zo.rotate(op, xy[1:], phase[1:]))
    core.mix(dxy[1:], 77, 0x01234567, n0=5, out=out[1:])
    ph = (((77 + (5 + np.arange(n, dtype=np.uint64)) * 0x01234567) & 0xFFFFFFFF).astype(np.uint32)) >> 8
    assert np.array_equal(host(out[1:]), zo.rotate(op, xy[1:], ph))


@pytest.mark.parametrize("name", [k for k in sorted(KATS) if not k.startswith("_")])
def test_survey_known_answers_on_gpu(name):
    """tests/golden/survey_kats.json straight through the CUDA path (no oracle involved)."""
    c = KATS[name]
    d = c["derive"]
    if name.startswith("r2p"):
        core = zc.Topolar(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"])
        xy = np.array([[v[0], v[1]] for v in c["vectors"]], dtype=np.int32)
        mag, ph = core.topolar(dev(xy))
        assert host(mag).tolist() == [v[2] for v in c["vectors"]]
        assert host(ph).view(np.uint32).tolist() == [int(v[3], 16) for v in c["vectors"]]
        return
    core = zc.Cordic(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"])
    vec = c["vectors"]
    if vec and len(vec[0]) == 3:
        phase = np.array([int(v[0], 16) for v in vec], dtype=np.uint32)
        got = host(core.rotate_const(c["x0"], c["y0"], dev(phase)))
        assert got.tolist() == [[v[1], v[2]] for v in vec]
    elif vec:
        phase = np.array([int(v[2], 16) for v in vec], dtype=np.uint32)
        xy = np.array([[v[0], v[1]] for v in vec], dtype=np.int32)
        got = host(core.rotate(dev(xy), dev(phase)))
        assert got.tolist() == [[v[3], v[4]] for v in vec]
    if "sweep" in c:
        n = 1 << core.PW
        out = core.rotate_const(c["x0"], c["y0"], torch.arange(n, dtype=torch.int32, device="cuda"))
        assert int(out[:, 0].sum(dtype=torch.int64)) == c["sweep"]["sum_x"]
        assert int(out[:, 1].sum(dtype=torch.int64)) == c["sweep"]["sum_y"]


def test_host_buffer_entry_points():
    """The *_host ABI (H2D -> kernel -> D2H pipeline, several chunks) with pageable and pinned memory."""
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    vcore, vop = both_r2p(**R2P_CONFIGS["cfg2"])
    rng = np.random.default_rng(SEED + 4)
    n = (9 << 20) + 3                                   # > 2 pipeline chunks, ragged tail
    phase = rng.integers(0, 1 << 24, size=n, dtype=np.uint64).astype(np.uint32)
    out = np.empty((n, 2), dtype=np.int32)
    core.rotate_const_host(131071, 0, phase, out)
    assert np.array_equal(out, zo.rotate_const(op, 131071, 0, phase))
    # pinned
    pin_in, pin_out = zc.PinnedBuffer(n, np.uint32), zc.PinnedBuffer(2 * n, np.int32)
    pin_in.array[:] = phase
    core.rotate_const_host(131071, 0, pin_in.array, pin_out.array)
    assert np.array_equal(pin_out.array.reshape(n, 2), out)
    pin_in.free(); pin_out.free()
    m = (5 << 20) + 1
    xy = rng.integers(-32768, 32768, size=(m, 2), dtype=np.int64).astype(np.int32)
    out2 = np.empty((m, 2), dtype=np.int32)
    core.rotate_host(xy, phase[:m], out2)
    assert np.array_equal(out2, zo.rotate(op, xy, phase[:m]))
    mag, ph = np.empty(m, dtype=np.int32), np.empty(m, dtype=np.uint32)
    vcore.topolar_host(xy, mag, ph)
    wm, wp = zo.topolar(vop, xy)
    assert np.array_equal(mag, wm) and np.array_equal(ph, wp)
    out3 = np.empty((m, 2), dtype=np.int32)
    core.nco_host(131071, 0, 5, 0x01234567, out3, n0=77)
    assert np.array_equal(out3, zo.nco(op, 131071, 0, 5, 0x01234567, m, n0=77))
    lut = zc.SinTable(phase_bits=17, ow=13)
    words = rng.integers(0, 1 << 32, size=m, dtype=np.uint64).astype(np.uint32)
    o4 = np.empty(m, dtype=np.int32)
    lut.lookup_host(words, o4)
    assert np.array_equal(o4, zo.lut_sin(17, 13, zo.sintable(17, 13), words))
    q = zc.QuarterWav(phase_bits=18, ow=24)
    q.lookup_host(words, o4)
    assert np.array_equal(o4, zo.lut_qwav(18, 24, zo.quarterwav(18, 24), words))


def test_full_size_cfg1_properties():
    """BASELINE configs[1] at its full size (2^30 samples = 64 sweeps of the 2^24 phases): every
    sweep must reproduce the first one word for word, the first one must equal the oracle, and the
    checksum of checksums is 64 x the single-sweep sums (-39316 each)."""
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    n, period = 1 << 30, 1 << 24
    phase = torch.arange(n, dtype=torch.int32, device="cuda").bitwise_and_(period - 1)
    out = core.rotate_const(131071, 0, phase)
    del phase
    first = out[:period]
    assert np.array_equal(host(first), zo.rotate_const(op, 131071, 0, np.arange(period, dtype=np.uint32)))
    v = out.view(n // period, period, 2)
    for k in range(1, n // period):
        assert torch.equal(v[k], first), k
    assert int(out[:, 0].sum(dtype=torch.int64)) == 64 * -39316
    assert int(out[:, 1].sum(dtype=torch.int64)) == 64 * -39316
    # x(phase + half turn) = -x(phase) up to the rounding asymmetry of the core (a few LSB)
    half = period // 2
    d = (first[:half].to(torch.int64) + first[half:].to(torch.int64)).abs().max()
    assert int(d) <= 4
    del out, v, first
    torch.cuda.empty_cache()


def test_full_size_cfg2_round_trip():
    """BASELINE configs[2] at full size (2^28 complex samples): rotate a full-scale vector by a phase
    sweep, feed the result to the vectoring core, and recover the phase to within the cores' noise;
    a 2^24-sample slice is checked word for word against the oracle."""
    core, op = both_p2r(iw=16, ow=16, xtra=2, pw=24)
    vcore, vop = both_r2p(**R2P_CONFIGS["cfg2"])
    n = 1 << 28
    phase = torch.arange(n, dtype=torch.int32, device="cuda").mul_(16 + 1).bitwise_and_((1 << 24) - 1)
    xy = core.rotate_const(16000, 0, phase)        # amplitude 16000*1.1644/2 fits 16 bits
    mag, ph = vcore.topolar(xy)
    sl = slice(123 << 16, (123 << 16) + (1 << 24))
    wm, wp = zo.topolar(vop, host(xy[sl]))
    assert np.array_equal(host(mag[sl]), wm) and np.array_equal(host(ph[sl]).view(np.uint32), wp)
    err = (ph - phase).bitwise_and_((1 << 24) - 1)
    err = torch.where(err >= (1 << 23), err - (1 << 24), err)
    # 16-bit inputs at radius ~9300: angular resolution ~ 2^24/(2*pi*9300) = 287 phase units
    assert int(err.abs().max()) < 1000
    m = mag.to(torch.float64)
    # topolar_tb.cpp:242-246: o_mag ~ |in| * 2^(IW-1-OW) * GAIN, with |in| = 16000 * 1.16443.. / 2 from the rotator
    assert abs(float(m.mean()) - 16000 * 1.16443534550574 / 2 * vcore.GAIN * 0.5) < 2.0
    del phase, xy, mag, ph, err, m
    torch.cuda.empty_cache()


def test_more_than_2_32_samples_in_one_call():
    """One call over 2^32 + 2^24 + 133 samples (16 GiB of phases, 32 GiB of outputs): sample and byte offsets leave 32
    bits, the table kernels' 32-bit block counters and the NCO's modulo-2^32 sample index must not.  The phase stream is
    257 sweeps of the 2^24 phases plus a ragged tail; every sweep must equal the first, the first must equal the
    reference RTL's outputs (tests/golden/rtl_sweeps.json), and the NCO that generates the same phases must reproduce
    the stream's last 2^24 + 133 samples from a starting index beyond 2^32."""
    from . import rtl_sweeps as rs
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < (60 << 30):
        pytest.skip("needs 60 GB of free device memory")
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    period, tail = 1 << 24, 133
    n = (1 << 32) + period + tail
    phase = torch.empty(n, dtype=torch.int32, device="cuda")
    sweep = torch.arange(period, dtype=torch.int32, device="cuda")
    for k in range(0, n, period):
        m = min(period, n - k)
        phase[k:k + m] = sweep[:m]
    out = core.rotate_const(131071, 0, phase)
    torch.cuda.synchronize()
    del phase
    first = out[:period]
    w = host(first)
    rs.check("p2r_cfg1_sweep", rs.port_words([w[:, 0], w[:, 1]], [core.OW, core.OW]))
    for k in range(period, n, period):
        m = min(period, n - k)
        assert torch.equal(out[k:k + m], first[:m]), k
    # phase32 = i << 8 for sample i: step 0x100, and n0 carries the 64-bit sample index
    n0 = (1 << 32) - 77
    cnt = n - n0
    nco = core.nco(131071, 0, 0, 0x100, cnt, n0=n0)
    assert torch.equal(nco, out[n0:])
    del out, nco, first
    torch.cuda.empty_cache()


@pytest.mark.parametrize("kind,pw,ow", [("tbl", 17, 13), ("tbl", 10, 8), ("tbl", 16, 16), ("qtr", 18, 24), ("qtr", 18, 13),
                                        ("qtr", 12, 12), ("tbl", 18, 13), ("qtr", 20, 26)])
def test_lut_large_batches_any_phase_pattern(kind, pw, ow):
    """Batches of 4 Mi samples and more take the kernel that keeps a compressed table in shared memory (when it fits:
    the last two cases do not, or are too wide, and stay on the L2 path).  Random, swept and ragged."""
    lut = (zc.SinTable if kind == "tbl" else zc.QuarterWav)(phase_bits=pw, ow=ow)
    tbl = zo.sintable(pw, ow) if kind == "tbl" else zo.quarterwav(pw, ow)
    ref = zo.lut_sin if kind == "tbl" else zo.lut_qwav
    rng = np.random.default_rng(SEED + 17)
    n = (1 << 22) + 4 * 1025 + 3
    for words in (rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32),
                  (np.arange(n, dtype=np.uint64) * 1021 & 0xFFFFFFFF).astype(np.uint32)):
        got = host(lut.lookup(dev(words)))
        assert np.array_equal(got, ref(pw, ow, tbl, words))


@pytest.mark.parametrize("kind,pw,ow", [("tbl", 12, 13), ("tbl", 14, 16), ("qtr", 14, 24), ("qtr", 14, 16), ("qtr", 13, 30)])
def test_lut_arbitrary_table_contents(kind, pw, ow):
    """The table is caller memory (zc_lut_sin / zc_lut_qwav take a device pointer): words that are not a sine wave at all
    -- no half-wave symmetry, magnitudes beyond what the compressed copy can hold -- must still be looked up exactly as
    rtl/sintable.v:71-75 / rtl/quarterwav.v:92-109 would (the shared-memory kernel detects it and reads global memory)."""
    rng = np.random.default_rng(SEED + 18)
    nwords = 1 << (pw if kind == "tbl" else pw - 2)
    tbl = rng.integers(0, 1 << ow, size=nwords, dtype=np.uint64).astype(np.uint32)
    n = (1 << 22) + 8
    words = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    d_tbl, d_words = dev(tbl), dev(words)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    fn = zc.lib().zc_lut_sin if kind == "tbl" else zc.lib().zc_lut_qwav
    rc = fn(pw, ow, ctypes.c_void_p(d_tbl.data_ptr()), ctypes.c_void_p(d_words.data_ptr()), ctypes.c_void_p(out.data_ptr()), n, 0, None)
    assert rc == 0, zc.lib().zc_last_error()
    torch.cuda.synchronize()
    want = (zo.lut_sin if kind == "tbl" else zo.lut_qwav)(pw, ow, tbl, words)
    assert np.array_equal(host(out), want)


@pytest.mark.parametrize("ow,pw", [(13, 18), (16, 20), (10, 14)])
def test_quadtbl_tables_that_do_wrap(ow, pw):
    """The engine drops the LBITS / CBITS register wraps of rtl/quadtbl.v:196-260 only after walking the caller's tables
    and proving that no phase can make them wrap.  Coefficients that are not a sine wave (random words of the right
    widths) do wrap: the kernel that models the registers must take over, and the words must still be the RTL's."""
    core = zc.QuadTbl(ow=ow, phase_bits=pw)
    rc, q = zo.derive_qtbl(0, ow, 2, pw)
    assert rc == 0 and q.pw == core.PW
    rng = np.random.default_rng(SEED + 77)
    n = 1 << q.lgtbl
    for name, bits in (("ctbl", q.cbits), ("ltbl", q.lbits), ("qtbl", q.qbits)):
        words = rng.integers(0, 1 << bits, size=n, dtype=np.uint64)
        for k in range(n):
            getattr(q, name)[k] = int(words[k])
            getattr(core.params, name)[k] = int(words[k])
    port = np.arange(1 << core.PW, dtype=np.uint32) if core.PW <= 20 else rng.integers(0, 1 << core.PW, size=1 << 20, dtype=np.uint64).astype(np.uint32)
    words32 = (port.astype(np.uint64) << (32 - core.PW)).astype(np.uint32)
    got = host(core.lookup(dev(words32)))
    assert np.array_equal(got, zo.quadtbl(q, port))
