"""The short late stages of the vectoring kernel (zc_kernels.cuh: vec_step_tail) rest on a claim the host proves per
configuration (zc_api.cu: vec_tail_start): from stage i = NSTAGES_live - zc_topolar_tail_stages() on,
-2^(i+1) <= y < 2^(i+1) for every input, so rtl/topolar.v:227-243's y>>>(i+1) equals the sign word.  This file checks
the claim, and the short form itself, on the CPU: a numpy model of the stage recursion (validated against the oracle,
which follows the RTL text) over exhaustive input sets of small cores and corner/random inputs of large ones."""
import ctypes

import numpy as np
import pytest

import cordic_b200 as zc
from tests import zo


def _live(p, seq):
    if seq:
        return p.nstages
    n = 0
    while n < p.nstages and n < p.ww and p.angle[n] != 0:
        n += 1
    return n


def _model(p, xy, seq, tail):
    """rtl/topolar.v:83-84,122-152,227-243,253-255 in int64 numpy.  Returns (mag, phase, worst), worst[i] = the
    largest value of max(y, -y-1) >> (i+1) seen entering stage i (0 means the claim holds there).  Stages
    >= neff-tail are computed in the kernel's short form."""
    iw, ww, ow, pw = p.iw, p.ww, p.ow, p.pw
    sh = 64 - iw
    ex = ((xy[:, 0].astype(np.int64) << sh) >> sh) << (ww - iw - 2)
    ey = ((xy[:, 1].astype(np.int64) << sh) >> sh) << (ww - iw - 2)
    xn, yn = ex < 0, ey < 0
    s, d = ex + ey, ex - ey
    x = np.where(xn, np.where(yn, -s, -d), np.where(yn, d, s))
    y = np.where(xn, np.where(yn, d, -s), np.where(yn, s, -d))
    E = 1 << (pw - 3)
    ph = np.where(xn, np.where(yn, 5 * E, 3 * E), np.where(yn, 7 * E, 1 * E)).astype(np.int64)
    neff = _live(p, seq)
    worst = []
    for i in range(neff):
        k = min(i + 1, 31)
        worst.append(int(np.max(np.maximum(y, -y - 1) >> k)))
        a = int(p.angle[i]) if i < 64 else 0
        m = y >> 63
        if i >= neff - tail:
            ns = -2 * m - 1
            y, x = y + ns * (x >> k), x - m
            ph = ph + ns * (-a)
        else:
            sg = 2 * m + 1
            x, y = x + sg * (y >> k), y - sg * (x >> k)
            ph = ph + sg * a
    assert int(np.max(np.abs(x))) < (1 << (ww - 1)) and int(np.max(np.abs(y))) < (1 << (ww - 1))
    D = ww - ow
    if ww > ow + 1:
        x = (x + (1 << (D - 1)) - 1 + ((x >> D) & 1))
    mag = ((x << (64 - ww)) >> (64 - ww)) >> D
    return mag.astype(np.int32), (ph & ((1 << pw) - 1)).astype(np.uint32), worst


def _inputs(iw, rng):
    if iw <= 10:
        v = np.arange(-(1 << (iw - 1)), 1 << (iw - 1), dtype=np.int32)
        return np.stack(np.meshgrid(v, v, indexing="ij"), axis=-1).reshape(-1, 2)
    lo, hi = -(1 << (iw - 1)), (1 << (iw - 1)) - 1
    corners = np.array([lo, lo + 1, -2, -1, 0, 1, 2, hi - 1, hi], dtype=np.int64)
    grid = np.stack(np.meshgrid(corners, corners, indexing="ij"), axis=-1).reshape(-1, 2)
    rnd = rng.integers(lo, hi + 1, size=(1 << 18, 2), dtype=np.int64)
    # near-axis and near-diagonal vectors: where y converges slowest / the octant fold is at its edge
    t = rng.integers(lo, hi + 1, size=(1 << 16), dtype=np.int64)
    e = rng.integers(-3, 4, size=(1 << 16), dtype=np.int64)
    near = np.concatenate([np.stack([t, e], 1), np.stack([e, t], 1), np.stack([t, np.clip(t + e, lo, hi)], 1),
                           np.stack([t, np.clip(-t + e, lo, hi)], 1)])
    return np.concatenate([grid, rnd, near]).astype(np.int32)


CASES = [  # (iw, ow, xtra, pw, nstages, seq)
    (8, 8, 2, 0, 0, 0), (9, 12, 0, 0, 0, 0), (10, 8, 3, 0, 0, 0), (10, 10, 2, 14, 9, 0), (8, 8, 2, 0, 0, 1),
    (13, 13, 2, 0, 0, 0), (16, 16, 2, 0, 0, 0), (16, 16, 2, 0, 0, 1), (12, 20, 1, 0, 0, 0), (20, 16, 4, 0, 0, 0),
    (24, 24, 2, 0, 0, 0), (16, 16, 2, 30, 28, 0), (18, 18, 2, 0, 0, 0), (22, 20, 1, 0, 0, 0),
]


@pytest.mark.parametrize("iw,ow,xtra,pw,nstages,seq", CASES)
def test_tail_claim_and_short_form(iw, ow, xtra, pw, nstages, seq):
    derive_o = zo.derive_sr2p if seq else zo.derive_r2p
    derive_p = zc.derive_sr2p if seq else zc.derive_r2p
    rc, op = derive_o(iw, ow, xtra, pw, nstages)
    assert rc == 0
    p = derive_p(iw, ow, xtra, pw, nstages)
    tail = zc.lib().zc_topolar_tail_stages(ctypes.byref(p))
    neff = _live(p, seq)
    assert 0 <= tail <= min(16, neff) and tail % 2 == 0
    xy = _inputs(p.iw, np.random.default_rng(20261017 + iw))
    wm, wp = zo.topolar(op, xy)
    # the model with every stage in full is the oracle (so `worst` is measured on the RTL's recursion) ...
    m0, p0, worst = _model(p, xy, seq, 0)
    assert np.array_equal(m0, wm) and np.array_equal(p0, wp)
    # ... the claim holds wherever the engine relies on it ...
    assert all(w == 0 for w in worst[neff - tail:]), (tail, worst)
    # ... and the short form is the same function
    m1, p1, _ = _model(p, xy, seq, tail)
    assert np.array_equal(m1, wm) and np.array_equal(p1, wp)


def test_tail_is_used_for_the_benchmark_core_and_not_for_rotation():
    p = zc.derive_r2p(16, 16, 2)
    assert zc.lib().zc_topolar_tail_stages(ctypes.byref(p)) == 10      # stages 11..20 of 21
    assert zc.lib().zc_topolar_tail_stages(ctypes.byref(zc.derive_p2r(18, 18, 2, 24, 20))) == 0


def test_tail_start_is_not_far_from_tight():
    """The bound is sufficient, not necessary; on an exhaustively enumerated core it should not give away more than
    a few stages (a regression here means lost speed, not lost exactness)."""
    p = zc.derive_r2p(10, 10, 2)
    rc, op = zo.derive_r2p(10, 10, 2)
    xy = _inputs(10, None)
    _, _, worst = _model(p, xy, 0, 0)
    neff = _live(p, 0)
    first_ok = next(i for i in range(neff) if all(w == 0 for w in worst[i:]))
    tail = zc.lib().zc_topolar_tail_stages(ctypes.byref(p))
    assert neff - tail >= first_ok and (neff - tail) - first_ok <= 3, (neff, tail, first_ok)
