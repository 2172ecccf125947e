"""CPU tier: the host side of the NCO comb mapping (cordic_b200/csrc/zc_seeded.cuh: comb_search, MAP_COMB) -- which
run length K a step gets, and that the lane arithmetic of the kernel visits every sample of the covered prefix exactly
once.  No GPU call is made."""
import ctypes

import numpy as np
import pytest

import cordic_b200 as zc


def comb_run(p, step, n):
    return int(zc.lib().zc_nco_comb_run(ctypes.byref(p), step & 0xFFFFFFFF, n))


def signed_delta(K, step):
    d = (K * step) & 0xFFFFFFFF
    return d - (1 << 32) if d >= (1 << 31) else d


def test_comb_search_known_steps():
    p = zc.derive_p2r(18, 18, 2, 24, 20)
    # BASELINE configs[4]: 225 steps are one turn minus 0.47 phase LSB, but K must be a multiple of 4 and 900 steps are
    # 1.9 LSB off -- eight distinct table rows per quarter-warp again: no comb (measured: slower than the byte table)
    assert comb_run(p, 0x01234567, 1 << 30) == 0
    assert comb_run(p, 0x00300000, 1 << 24) == 4096              # an exact period: every lane of a quarter-warp on one row
    assert comb_run(p, 0x80000001, 1 << 24) == 64                # 64 steps: a quarter of a phase LSB
    assert comb_run(p, 1, 1 << 30) == 0 and comb_run(p, 0x1FF, 1 << 30) == 0   # below 2 phase LSBs: block mapping
    assert comb_run(p, 0x01234567, 2000) == 0                    # a tile (8K samples) does not fit


@pytest.mark.parametrize("pw", [16, 20, 24])
def test_comb_search_properties(pw):
    """Whatever it returns is a multiple of 4, at least 64, fits n, and keeps neighbouring lanes of a quarter-warp within
    0.35 phase LSB."""
    p = zc.derive_p2r(16, 16, 2, pw, 0)
    rng = np.random.default_rng(20261017 + pw)
    found = 0
    for step in [int(s) for s in rng.integers(1, 1 << 32, size=300, dtype=np.uint64)]:
        for n in (1 << 22, 1 << 30):
            K = comb_run(p, step, n)
            if K == 0:
                continue
            found += 1
            assert K % 4 == 0 and K >= 64 and 8 * K <= n
            assert abs(signed_delta(K, step)) <= 0.35 * (1 << (32 - pw)) + 1, (hex(step), K)
    assert found > 100               # 2^30 samples: Dirichlet guarantees a run with |delta| <= 32 below n/8


def test_comb_lane_arithmetic_covers_every_sample_once():
    """Mirror of the index arithmetic in k_rotate_seeded<.., MAP_COMB, ..>: unit (t, m) -> store lane (a', b') -> samples
    t*8K + a'*K + 16m + 4b' + {0, 1, 2, 3}, pairs stored when 16m + 4b' (+2) < K; warps stride over the units with
    (dt, dm).  The compute lane (a, b) = (l & 7, l >> 3) hands its four results to store lane 4a + b."""
    for K, n, nwarps in [(452, 50000, 7), (64, 4096, 5), (68, 9000, 300), (452, 452 * 8 * 3 + 17, 148 * 32)]:
        cpr, tile = (K + 15) // 16, 8 * K
        tiles = n // tile
        nunits = tiles * cpr
        dt, dm = nwarps // cpr, nwarps % cpr
        seen = np.zeros(tiles * tile, dtype=np.int32)
        for w in range(min(nwarps, nunits)):
            blk = w
            t = blk // cpr
            m = blk - t * cpr
            while blk < nunits:
                for lane in range(32):
                    a, b = lane >> 2, lane & 3                      # store layout
                    src = (lane >> 2) + 8 * (lane & 3)              # the compute lane that produced these four samples
                    assert (src & 7, src >> 3) == (a, b)
                    base = t * tile + a * K + 4 * b + 16 * m
                    j0 = 16 * m + 4 * b
                    if j0 < K:                                      # K % 4 == 0: all four or none
                        seen[base:base + 4] += 1
                m += dm; t += dt
                if m >= cpr:
                    m -= cpr; t += 1
                blk += nwarps
        assert seen.min() == 1 and seen.max() == 1, (K, n, nwarps)
