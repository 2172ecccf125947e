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
    assert comb_run(p, 0x01234567, 1 << 30) == 450               # BASELINE configs[4]: 225 steps ~ one turn
    assert comb_run(p, 0x00300000, 1 << 24) == 4096              # an exact period: every lane of a quarter-warp on one row
    assert comb_run(p, 1, 1 << 30) == 0 and comb_run(p, 0x1FF, 1 << 30) == 0   # below 2 phase LSBs: block mapping
    assert comb_run(p, 0x01234567, 2000) == 0                    # a tile (8K samples) does not fit


@pytest.mark.parametrize("pw", [16, 20, 24])
def test_comb_search_properties(pw):
    """Whatever it returns is even, at least 64, fits n, and keeps the lanes of a quarter-warp within a dozen phase LSBs."""
    p = zc.derive_p2r(16, 16, 2, pw, 0)
    rng = np.random.default_rng(20261017 + pw)
    found = 0
    for step in [int(s) for s in rng.integers(1, 1 << 32, size=300, dtype=np.uint64)]:
        for n in (1 << 22, 1 << 30):
            K = comb_run(p, step, n)
            if K == 0:
                continue
            found += 1
            assert K % 2 == 0 and K >= 64 and 8 * K <= n
            assert abs(signed_delta(K, step)) <= 12 * (1 << (32 - pw)), (hex(step), K)
    assert found > 100               # 2^30 samples: Dirichlet guarantees a run with |delta| <= 32 below n/8


def test_comb_lane_arithmetic_covers_every_sample_once():
    """Mirror of the index arithmetic in k_rotate_seeded<.., MAP_COMB, ..>: unit (t, m) -> lane (a, b) -> samples
    t*8K + a*K + 16m + 2b + {0, 1, 8, 9}, stored when 16m + 2b (+8) < K; warps stride over the units with (dt, dm)."""
    for K, n, nwarps in [(450, 50000, 7), (64, 4096, 5), (66, 9000, 300), (450, 450 * 8 * 3 + 17, 148 * 32)]:
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
                    a, b = lane & 7, lane >> 3
                    base = t * tile + a * K + 2 * b + 16 * m
                    j0 = 16 * m + 2 * b
                    if j0 < K:
                        seen[base] += 1; seen[base + 1] += 1
                    if j0 + 8 < K:
                        seen[base + 8] += 1; seen[base + 9] += 1
                m += dm; t += dt
                if m >= cpr:
                    m -= cpr; t += 1
                blk += nwarps
        assert seen.min() == 1 and seen.max() == 1, (K, n, nwarps)
