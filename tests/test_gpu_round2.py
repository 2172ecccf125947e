"""GPU tier, round 2: the NCO comb mapping, the packed-port entry points, the multi-device host entry points and the
multi-device parity check of SURVEY.md §4 T3 -- all through the C ABI, bit-exact against the oracle.  Marked ``gpu``."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import cordic_b200 as zc
from . import zo
from .conftest import ROOT
from .test_gpu_parity import P2R_CONFIGS, R2P_CONFIGS, SEED, both_p2r, both_r2p, dev, host

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def comb_run(core, step, n):
    return int(zc.lib().zc_nco_comb_run(ctypes.byref(core.params), step & 0xFFFFFFFF, n))


# ---- NCO comb mapping ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("step", [0x01234567, 0x80000001, 0x7FFFFFFF, 0x00300000, 0x00010001, 0xFEDCBA99, 0x00000200])
def test_nco_comb_mapping_is_bit_exact(step):
    """A scattering NCO step takes the comb mapping (lane (a, b) works in run a of every 8K-sample tile); the words must
    equal the oracle's and the block mapping's, whatever K, n0, phase0 and the size of the remainder."""
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    for n, phase0, n0 in [((1 << 20) + 5, 0, 0), ((3 << 20) + 1234, 0xDEADBEEF, (1 << 33) + 12345), (1 << 21, 7, 999)]:
        want = zo.nco(op, 131071, 0, phase0, step, n, n0=n0)
        got = host(core.nco(131071, 0, phase0, step, n, n0=n0))
        assert np.array_equal(got, want), (hex(step), n, phase0, n0, comb_run(core, step, n))
        blk = host(core.nco(131071, 0, phase0, step, n, n0=n0, flags=zc.F_NO_COMB))
        assert np.array_equal(blk, want), (hex(step), n, "block mapping")


def test_nco_comb_is_taken_when_a_near_period_exists():
    """Step 0x80000001: 64 steps are 32 turns plus 64/2^32, a quarter of a phase LSB -- the lanes of a quarter-warp share
    table rows under K = 64.  The diagnostic says so, and the launch count shows one comb pass covering all of n.  The
    BASELINE configs[4] step has no such K below n/8 (its near-period 225 is odd, 900 is 1.9 LSB off) and keeps the block
    mapping with the byte table; slow NCOs keep the block mapping with the word table."""
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    n = 1 << 22
    assert comb_run(core, 0x80000001, n) == 64
    assert comb_run(core, 0x01234567, n) == 0
    assert comb_run(core, 0x100, n) == 0 and comb_run(core, 0xFFFFFF00, n) == 0
    l0 = zc.launch_count()
    out = core.nco(131071, 0, 0, 0x80000001, n)
    torch.cuda.synchronize()
    assert zc.launch_count() - l0 == 1          # n is a multiple of the 512-sample tile: the comb pass is all there is
    assert np.array_equal(host(out), zo.nco(op, 131071, 0, 0, 0x80000001, n))


def test_nco_comb_misaligned_output_and_other_cores():
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    n = (1 << 20) + 64
    buf = torch.empty((n + 1, 2), dtype=torch.int32, device="cuda")
    got = core.nco(131071, 0, 3, 0x01234567, n, out=buf[1:])          # 8-byte aligned only: no comb, still exact
    assert np.array_equal(host(got), zo.nco(op, 131071, 0, 3, 0x01234567, n))
    for name in ("shipped", "cfg0", "12_16_x1"):
        c2, o2 = both_p2r(**P2R_CONFIGS[name])
        x0 = (1 << (c2.IW - 1)) - 1
        for step in (0x01234567, 0x9E3779B9, 0x40000001):
            got = host(c2.nco(x0, 0, 11, step, n, flags=zc.F_FORCE_SEED))
            assert np.array_equal(got, zo.nco(o2, x0, 0, 11, step, n)), (name, hex(step))


def test_nco_comb_chunks_concatenate():
    """Sharding the NCO by n0 (the 8-GPU configuration): pieces computed with the comb mapping equal the
    whole, including pieces that start in the middle of a tile."""
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    n, parts = 1 << 23, 4
    whole = host(core.nco(131071, 0, 99, 0x01234567, n))
    pieces = [host(core.nco(131071, 0, 99, 0x01234567, n // parts, n0=r * (n // parts))) for r in range(parts)]
    assert np.array_equal(np.concatenate(pieces), whole)
    assert np.array_equal(whole[:1 << 20], zo.nco(op, 131071, 0, 99, 0x01234567, 1 << 20))


# ---- byte table with merged records ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(P2R_CONFIGS))
def test_rotate_const_merged_records(name):
    """The byte-table kernel with the interval's row offset folded into its (x, y) records (three shared-memory lookups per
    sample; cores with WW <= 24) against the four-lookup form (ZC_F_NO_MERGE) and the oracle: scattered phases, a sweep,
    other input vectors, the NCO with a scattering step, packed outputs."""
    core, op = both_p2r(**P2R_CONFIGS[name])
    rng = np.random.default_rng(SEED + 26)
    n = (1 << 20) + 77
    lim = 1 << (core.IW - 1)
    for x0, y0 in [(lim - 1, 0), (-lim, lim - 1), (3, -2)]:
        for pat in ("random", "sweep"):
            ph = (np.arange(n, dtype=np.uint32) & ((1 << core.PW) - 1)) if pat == "sweep" else \
                rng.integers(0, 1 << core.PW, size=n, dtype=np.uint64).astype(np.uint32)
            want = zo.rotate_const(op, x0, y0, ph)
            for flags in (zc.F_SEED_PACKED | zc.F_FORCE_SEED, zc.F_SEED_PACKED | zc.F_FORCE_SEED | zc.F_NO_MERGE):
                assert np.array_equal(host(core.rotate_const(x0, y0, dev(ph), flags=flags)), want), (name, x0, y0, pat, flags)
    want = zo.nco(op, lim - 1, 0, 9, 0x01234567, n, n0=5)
    for flags in (zc.F_FORCE_SEED, zc.F_FORCE_SEED | zc.F_NO_MERGE):
        assert np.array_equal(host(core.nco(lim - 1, 0, 9, 0x01234567, n, n0=5, flags=flags)), want), (name, "nco", flags)


# ---- per-sample vectors: word-table suffix ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cfg1", "shipped", "cfg0"])
def test_rotate_per_sample_word_suffix(name):
    """zc_rotate with the suffix directions as IDP.2A word planes (forced, and auto-selected by the in-kernel probe for
    streams of >= 4 Mi samples): sweeps, scattered phases and a sweep with a stride, corner vectors included."""
    core, op = both_p2r(**P2R_CONFIGS[name])
    rng = np.random.default_rng(SEED + 25)
    lim = 1 << (core.IW - 1)
    n = (1 << 22) + 133
    xy = rng.integers(-lim, lim, size=(n, 2), dtype=np.int64).astype(np.int32)
    xy[:4] = [[lim - 1, lim - 1], [-lim, -lim], [0, 0], [-lim, lim - 1]]
    mask = (1 << core.PW) - 1
    for pat in ("sweep", "random", "stride3"):
        ph = (np.arange(n, dtype=np.uint32) & mask) if pat == "sweep" else \
            rng.integers(0, 1 << core.PW, size=n, dtype=np.uint64).astype(np.uint32) if pat == "random" else \
            ((np.arange(n, dtype=np.uint64) * 3) & mask).astype(np.uint32)
        want = zo.rotate(op, xy, ph)
        for flags in (zc.F_DEFAULT, zc.F_SEED_WORDS, zc.F_SEED_PACKED, zc.F_NO_DP2A):
            l0 = zc.launch_count()
            got = host(core.rotate(dev(xy), dev(ph), flags=flags))
            assert np.array_equal(got, want), (name, pat, flags)
            if flags == zc.F_DEFAULT and name == "cfg1":
                # two table launches (one returns at its probe) + the 133-sample rest: 132 on the plain kernel, 1 generic
                # (cores whose table geometry does not fit -- cfg0's 16-bit phase -- run on the plain kernels: 2 launches)
                assert zc.launch_count() - l0 == 4, (name, pat)
    m = (1 << 20) + 7                                  # the NCO mixer: a slow step takes the words, a fast one the bytes
    for step in (0x40, 0x01234567):
        phase = (((5 + np.arange(m, dtype=np.uint64) * step) & 0xFFFFFFFF) >> (32 - core.PW)).astype(np.uint32)
        assert np.array_equal(host(core.mix(dev(xy[:m]), 5, step)), zo.rotate(op, xy[:m], phase)), (name, hex(step))


# ---- packed port words -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cfg2", "shipped"])
def test_topolar_i16_equals_topolar(name):
    core, op = both_r2p(**R2P_CONFIGS[name])
    rng = np.random.default_rng(SEED + 21)
    for n in [1, 3, 4, 1023, (1 << 20) + 6]:
        iq = rng.integers(-32768, 32768, size=(n, 2), dtype=np.int64).astype(np.int16)
        iq[:4] = [[32767, 0], [-32768, -32768], [0, 0], [-1, 1]][:min(4, n)]
        wm, wp = zo.topolar(op, iq.astype(np.int32))           # the oracle masks to IW bits like the port does
        d = torch.from_numpy(iq).cuda()
        mag, ph = core.topolar_i16(d)
        assert np.array_equal(host(mag), wm) and np.array_equal(host(ph).view(np.uint32), wp), (name, n)
        if n > 8:                                              # off the 16-byte grid: generic kernel
            mag, ph = core.topolar_i16(d[1:])
            assert np.array_equal(host(mag), wm[1:]) and np.array_equal(host(ph).view(np.uint32), wp[1:])
    m = (5 << 20) + 2
    iq = rng.integers(-32768, 32768, size=(m, 2), dtype=np.int64).astype(np.int16)
    mag, ph = np.empty(m, dtype=np.int32), np.empty(m, dtype=np.uint32)
    core.topolar_i16_host(iq, mag, ph)
    wm, wp = zo.topolar(op, iq.astype(np.int32))
    assert np.array_equal(mag, wm) and np.array_equal(ph, wp)


def test_topolar_i16_refuses_wide_inputs():
    core = zc.Topolar(18, 18, 2)
    with pytest.raises(zc.ZcError) as e:
        core.topolar_i16(torch.zeros((8, 2), dtype=torch.int16, device="cuda"))
    assert e.value.code == -2


@pytest.mark.parametrize("name", ["cfg0", "shipped", "8_8_x0"])
def test_rotate_const_o16_equals_rotate_const(name):
    core, op = both_p2r(**P2R_CONFIGS[name])
    rng = np.random.default_rng(SEED + 22)
    x0 = (1 << (core.IW - 1)) - 1
    for n, pat in [(5, "rand"), (4099, "rand"), ((1 << 20) + 128, "sweep"), ((1 << 22) + 3, "sweep"), ((1 << 22) + 3, "rand")]:
        ph = (np.arange(n, dtype=np.uint32) & ((1 << core.PW) - 1)) if pat == "sweep" else \
            rng.integers(0, 1 << core.PW, size=n, dtype=np.uint64).astype(np.uint32)
        want = zo.rotate_const(op, x0, -3, ph).astype(np.int16)
        got = core.rotate_const_o16(x0, -3, dev(ph))
        assert got.dtype == torch.int16 and np.array_equal(host(got), want), (name, n, pat)
    n = (5 << 20) + 1
    ph = rng.integers(0, 1 << core.PW, size=n, dtype=np.uint64).astype(np.uint32)
    out = np.empty((n, 2), dtype=np.int16)
    core.rotate_const_o16_host(x0, 0, ph, out)
    assert np.array_equal(out, zo.rotate_const(op, x0, 0, ph).astype(np.int16))


@pytest.mark.parametrize("kind,pw,ow", [("tbl", 17, 13), ("tbl", 23, 16), ("tbl", 10, 8), ("qtr", 25, 16), ("qtr", 18, 13), ("qtr", 12, 16)])
def test_lut_o16_equals_lut(kind, pw, ow):
    """zc_lut_sin_o16 / zc_lut_qwav_o16: the same o_val words as int16 -- sweeps (L2 kernel), scattered phases (shared-memory
    kernel where the table fits), ragged sizes and odd alignments (scalar kernel), the host entry point."""
    lut = (zc.SinTable if kind == "tbl" else zc.QuarterWav)(phase_bits=pw, ow=ow)
    tbl = (zo.sintable if kind == "tbl" else zo.quarterwav)(pw, ow)
    ref = zo.lut_sin if kind == "tbl" else zo.lut_qwav
    rng = np.random.default_rng(SEED + 27)
    for n, pat in [(7, "rand"), (4099, "rand"), ((1 << 22) + 6, "sweep"), ((1 << 22) + 6, "rand")]:
        w = ((np.arange(n, dtype=np.uint64) * 1024) & 0xFFFFFFFF).astype(np.uint32) if pat == "sweep" else \
            rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
        want = ref(pw, ow, tbl, w).astype(np.int16)
        got = lut.lookup_o16(dev(w))
        assert got.dtype == torch.int16 and np.array_equal(host(got), want), (kind, pw, ow, n, pat)
        if n > 100:
            got = lut.lookup_o16(dev(w)[4:], out=torch.empty(n - 3, dtype=torch.int16, device="cuda")[1:])   # 2-byte aligned only
            assert np.array_equal(host(got), want[4:])
    m = (5 << 20) + 3
    w = rng.integers(0, 1 << 32, size=m, dtype=np.uint64).astype(np.uint32)
    out = np.empty(m, dtype=np.int16)
    lut.lookup_o16_host(w, out)
    assert np.array_equal(out, ref(pw, ow, tbl, w).astype(np.int16))


@pytest.mark.parametrize("kind,pw,ow", [("tbl", 17, 13), ("tbl", 23, 16), ("qtr", 18, 24), ("qtr", 25, 16), ("qtr", 10, 8)])
def test_nco_through_the_lut_cores(kind, pw, ow):
    """zc_nco_lut_sin / _qwav: the accumulator's phases generated in registers must give what the lookup gives on the same
    phases written out -- slow steps (L2 kernel), scattering steps (shared-memory kernel), 64-bit sample offsets, shards that
    concatenate, ragged sizes, the host entry point."""
    lut = (zc.SinTable if kind == "tbl" else zc.QuarterWav)(phase_bits=pw, ow=ow)
    tbl = (zo.sintable if kind == "tbl" else zo.quarterwav)(pw, ow)
    ref = zo.lut_sin if kind == "tbl" else zo.lut_qwav
    for n, phase0, step, n0 in [(5, 1, 3, 0), (4099, 0xDEADBEEF, 0x01234567, 7), ((1 << 22) + 3, 0, 0x100, (1 << 33) + 5),
                                ((1 << 22) + 3, 99, 0x01234567, 123456789), ((1 << 22) + 1, 5, 0xFFFFFF00, 0)]:
        w = ((phase0 + (n0 + np.arange(n, dtype=np.uint64)) * step) & 0xFFFFFFFF).astype(np.uint32)
        want = ref(pw, ow, tbl, w)
        assert np.array_equal(host(lut.nco(phase0, step, n, n0=n0)), want), (kind, pw, ow, n, hex(step))
    n = 1 << 20
    whole = host(lut.nco(3, 0x01234567, n))
    parts = [host(lut.nco(3, 0x01234567, n // 4, n0=r * (n // 4))) for r in range(4)]
    assert np.array_equal(np.concatenate(parts), whole)
    m = (5 << 20) + 2
    out = np.empty(m, dtype=np.int32)
    lut.nco_host(7, 0x01234567, out, n0=11)
    w = ((7 + (11 + np.arange(m, dtype=np.uint64)) * 0x01234567) & 0xFFFFFFFF).astype(np.uint32)
    assert np.array_equal(out, ref(pw, ow, tbl, w))


def test_lut_o16_refuses_wide_tables():
    lut = zc.QuarterWav(phase_bits=18, ow=24)
    with pytest.raises(zc.ZcError) as e:
        lut.lookup_o16(torch.zeros(8, dtype=torch.int32, device="cuda"))
    assert e.value.code == -2


def test_rotate_const_o16_refuses_wide_outputs():
    core = zc.Cordic(18, 18, 2, 24, 20)
    with pytest.raises(zc.ZcError) as e:
        core.rotate_const_o16(1, 0, torch.zeros(8, dtype=torch.int32, device="cuda"))
    assert e.value.code == -2


# ---- full size ---------------------------------------------------------------------------------------------------------
def test_full_size_scattered_phases_and_nco_properties():
    """BASELINE sizes (2^30 samples) through the round-2 paths, checked by size-independent properties: scattered phases
    through the byte table with merged records -- a random permutation-free check: the stream is 64 copies of one 2^24-phase
    scramble, every copy must equal the first, and the first equals the oracle -- and the NCO (the cfg4 step through the
    byte table, a near-period step through the comb mapping): both halves of the stream computed as separate shards with
    n0 offsets equal the whole, the first 2^20 samples equal the oracle, and every sample equals the plain kernel (all 20
    stages in registers) fed with the accumulator's phases written out explicitly."""
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    n, period = 1 << 30, 1 << 24
    base = (torch.arange(period, dtype=torch.int64, device="cuda") * 0x9E3779B1).bitwise_and_(period - 1).to(torch.int32)
    phase = base.repeat(n // period)
    out = core.rotate_const(131071, 0, phase)
    del phase
    first = out[:period]
    assert np.array_equal(host(first[:1 << 20]), zo.rotate_const(op, 131071, 0, host(base[:1 << 20]).view(np.uint32)))
    v = out.view(n // period, period, 2)
    for k in range(1, n // period):
        assert torch.equal(v[k], first), k
    # 0x9E3779B1 is odd: the scramble is a permutation of the 2^24 phases
    assert int(first[:, 0].sum(dtype=torch.int64)) == -39316 and int(first[:, 1].sum(dtype=torch.int64)) == -39316
    del out, v, first, base
    torch.cuda.empty_cache()
    for step in (0x01234567, 0x80000001):
        whole = core.nco(131071, 0, 0, step, n)
        assert np.array_equal(host(whole[:1 << 20]), zo.nco(op, 131071, 0, 0, step, 1 << 20))
        half = core.nco(131071, 0, 0, step, n // 2, n0=n // 2)
        assert torch.equal(half, whole[n // 2:]), hex(step)
        del half
        piece = 1 << 26
        for s0 in range(0, n, piece):
            idx = torch.arange(s0, s0 + piece, dtype=torch.int64, device="cuda")
            ph = ((idx * step) & 0xFFFFFFFF) >> 8                 # the truncation of bench/cpp/cordic_tb.cpp:128-138
            ref = core.rotate_const(131071, 0, ph.to(torch.int32), flags=zc.F_NO_SEED)
            assert torch.equal(ref, whole[s0:s0 + piece]), (hex(step), s0)
            del idx, ph, ref
        del whole
        torch.cuda.empty_cache()


# ---- several devices ------------------------------------------------------------------------------------------------
def all_devices():
    return list(range(zc.lib().zc_device_count()))


def test_host_multi_entry_points():
    """zc_*_host_multi over every device of the box (one device: the degenerate shard list) with NUMA-placed pinned
    buffers and with pageable memory: byte-identical to the oracle, i.e. to the single-device result."""
    devices = all_devices()
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    vcore, vop = both_r2p(**R2P_CONFIGS["cfg2"])
    rng = np.random.default_rng(SEED + 23)
    n = (9 << 20) + 3
    hin = zc.ShardedPinnedBuffer(n, devices, np.uint32)
    hout = zc.ShardedPinnedBuffer(2 * n, devices, np.int32)
    assert [d for d, _ in hin.placement] == devices
    hin.array[:] = rng.integers(0, 1 << 24, size=n, dtype=np.uint64).astype(np.uint32)
    core.rotate_const_host_multi(131071, 0, hin.array, hout.array, devices)
    want = zo.rotate_const(op, 131071, 0, hin.array)
    assert np.array_equal(hout.array.reshape(n, 2), want)
    single = np.empty((n, 2), dtype=np.int32)
    core.rotate_const_host(131071, 0, hin.array, single, device=0)
    assert np.array_equal(single, want)
    m = (3 << 20) + 5
    xy = rng.integers(-32768, 32768, size=(m, 2), dtype=np.int64).astype(np.int32)
    out = np.empty((m, 2), dtype=np.int32)
    core.rotate_host_multi(xy, hin.array[:m], out, devices)
    assert np.array_equal(out, zo.rotate(op, xy, hin.array[:m]))
    mag, ph = np.empty(m, dtype=np.int32), np.empty(m, dtype=np.uint32)
    vcore.topolar_host_multi(xy, mag, ph, devices)
    wm, wp = zo.topolar(vop, xy)
    assert np.array_equal(mag, wm) and np.array_equal(ph, wp)
    core.nco_host_multi(131071, 0, 5, 0x01234567, out, devices, n0=(1 << 32) + 77)
    assert np.array_equal(out, zo.nco(op, 131071, 0, 5, 0x01234567, m, n0=(1 << 32) + 77))
    lut = zc.QuarterWav(phase_bits=18, ow=24)
    words = rng.integers(0, 1 << 32, size=m, dtype=np.uint64).astype(np.uint32)
    o4 = np.empty(m, dtype=np.int32)
    lut.lookup_host_multi(words, o4, devices)
    assert np.array_equal(o4, zo.lut_qwav(18, 24, zo.quarterwav(18, 24), words))
    q = zc.QuadTbl(ow=13, phase_bits=18)
    rc, oq = zo.derive_qtbl(0, 13, 2, 18)
    q.lookup_host_multi(words, o4, devices)
    assert np.array_equal(o4, zo.quadtbl(oq, words >> 14))
    core.mix_host_multi(xy, 7, 0x01234567, out, devices, n0=3)
    mph = (((7 + (3 + np.arange(m, dtype=np.uint64)) * 0x01234567) & 0xFFFFFFFF) >> 8).astype(np.uint32)
    assert np.array_equal(out, zo.rotate(op, xy, mph))
    iq = xy.astype(np.int16)
    vcore.topolar_i16_host_multi(iq, mag, ph, devices)
    assert np.array_equal(mag, wm) and np.array_equal(ph, wp)
    c0, o0 = both_p2r(**P2R_CONFIGS["cfg0"])
    ph0 = (hin.array[:m] & 0xFFFF).astype(np.uint32)
    o16 = np.empty((m, 2), dtype=np.int16)
    c0.rotate_const_o16_host_multi(32767, 0, ph0, o16, devices)
    assert np.array_equal(o16, zo.rotate_const(o0, 32767, 0, ph0).astype(np.int16))
    hin.free(); hout.free()
    with pytest.raises(zc.ZcError):
        core.rotate_const_host_multi(131071, 0, words, out, [0, 0])          # a device listed twice


def test_multi_device_parity_through_the_c_abi():
    """the sample stream sharded over ALL devices of the box through the C ABI -- device buffers on
    each GPU, rank r owning [r*N/G, (r+1)*N/G) -- concatenates to exactly what device 0 computes alone.  Rotation with a
    phase stream, the NCO in closed form, vectoring.  Needs two devices; the 1-GPU tier skips it."""
    devices = all_devices()
    if len(devices) < 2:
        pytest.skip("one device: multi-device parity runs on the multi-GPU tier (gpurun --gpus N)")
    core, op = both_p2r(**P2R_CONFIGS["cfg1"])
    vcore, vop = both_r2p(**R2P_CONFIGS["cfg2"])
    rng = np.random.default_rng(SEED + 24)
    G = len(devices)
    n = G * ((1 << 21) + 128)
    per = n // G
    phase = rng.integers(0, 1 << 24, size=n, dtype=np.uint64).astype(np.uint32)
    xy = rng.integers(-32768, 32768, size=(n, 2), dtype=np.int64).astype(np.int32)
    alone = host(core.rotate_const(131071, 0, dev(phase)))
    alone_nco = host(core.nco(131071, 0, 9, 0x01234567, n, device="cuda:0"))
    am, ap = vcore.topolar(dev(xy))
    am, ap = host(am), host(ap)
    parts, parts_nco, pm, pp = [], [], [], []
    for r, d in enumerate(devices):
        sl = slice(r * per, (r + 1) * per)
        with torch.cuda.device(d):
            ph_d = torch.from_numpy(phase[sl].view(np.int32)).to("cuda:%d" % d)
            xy_d = torch.from_numpy(xy[sl]).to("cuda:%d" % d)
            parts.append(core.rotate_const(131071, 0, ph_d))
            parts_nco.append(core.nco(131071, 0, 9, 0x01234567, per, n0=r * per, device="cuda:%d" % d))
            m_, p_ = vcore.topolar(xy_d)
            pm.append(m_); pp.append(p_)
    for d in devices:
        torch.cuda.synchronize(d)
    assert np.array_equal(np.concatenate([t.cpu().numpy() for t in parts]), alone)
    assert np.array_equal(np.concatenate([t.cpu().numpy() for t in parts_nco]), alone_nco)
    assert np.array_equal(np.concatenate([t.cpu().numpy() for t in pm]), am)
    assert np.array_equal(np.concatenate([t.cpu().numpy() for t in pp]), ap)
    assert np.array_equal(alone[:1 << 20], zo.rotate_const(op, 131071, 0, phase[:1 << 20]))


def test_scatter_rotate_gather_cpp():
    """The C++ client's NCCL path (zcordic_bench --scatter): device 0 owns the stream, chunks are scattered with
    ncclSend/ncclRecv, rotated on every device and gathered back, pipelined; the binary itself compares the gathered
    output byte for byte with device 0 computing the whole stream alone."""
    devices = all_devices()
    exe = os.path.join(ROOT, "cordic_b200", "zcordic_bench")
    if len(devices) < 2 or not os.path.exists(exe):
        pytest.skip("needs two devices and the built C++ client")
    r = subprocess.run([exe, "-g", str(len(devices)), "--scatter", "-l", "24", "-s", "2", "--json"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    import json
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["parity"] is True and d["value"] > 0
