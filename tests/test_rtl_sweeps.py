"""Large-volume parity against the reference's RTL, executed: tests/golden/rtl_sweeps.json holds digests of 30.2 million
outputs that oracle/vsim.py obtained by clocking the reference's Verilog text (every phase of rtl/cordic.v as shipped, of
BASELINE configs[0] and of configs[1] -- the headline core; every phase of the shipped table cores; millions of seeded
inputs for per-sample rotation and for both vectoring cores).  CPU tier: the oracle reproduces every digest.  GPU tier:
the CUDA path reproduces them through every kernel flavour WITHOUT going through the oracle."""
import numpy as np
import pytest

import cordic_b200 as zc
from . import zo
from . import rtl_sweeps as rs

GOLD = rs.golden()
NAMES = [n for n in rs.CASES if n in GOLD]


def _missing(name):
    return name not in GOLD


def _params_match(op, want):
    """the localparams the simulated Verilog declares (the sequential cores have no NSTAGES) against the oracle's derivation"""
    have = {"IW": op.iw, "OW": op.ow, "WW": op.ww, "PW": op.pw, "NSTAGES": op.nstages}
    return all(have[k] == v for k, v in want.items())


def test_every_case_has_digests():
    assert [n for n in rs.CASES if _missing(n)] == []
    assert sum(GOLD[n]["n"] for n in rs.CASES) == sum(len(rs.case_inputs(n)[0]) for n in rs.CASES) and sum(GOLD[n]["n"] for n in rs.CASES) > 30_000_000


# ---------------------------------------------------------------------------------------------- CPU tier: the oracle
@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_the_rtl(name):
    c = rs.CASES[name]
    cols = rs.case_inputs(name)
    k = c["kind"]
    if k in ("p2r_const", "p2r_xy"):
        d = c["derive"]
        rc, op = (zo.derive_sp2r if c.get("seq") else zo.derive_p2r)(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"])
        assert rc == 0 and _params_match(op, GOLD[name]["params"])
        if k == "p2r_const":
            out = zo.rotate_const(op, c["x0"], c["y0"], cols[2])
        else:
            out = zo.rotate(op, np.stack([cols[0], cols[1]], 1).astype(np.int32), cols[2])
        rs.check(name, rs.port_words([out[:, 0], out[:, 1]], [op.ow, op.ow]))
    elif k in ("r2p", "r2p_all"):
        d = c["derive"]
        rc, op = (zo.derive_sr2p if c.get("seq") else zo.derive_r2p)(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"])
        assert rc == 0 and _params_match(op, GOLD[name]["params"])
        mag, ph = zo.topolar(op, np.stack([cols[0], cols[1]], 1).astype(np.int32))
        rs.check(name, rs.port_words([mag, ph], [op.ow, op.pw]))
    elif k == "tbl":
        out = zo.lut_sin(c["pw"], c["ow"], zo.sintable(c["pw"], c["ow"]), cols[0] << np.uint32(32 - c["pw"]))
        rs.check(name, rs.port_words([out], [c["ow"]]))
    elif k == "qtr":
        out = zo.lut_qwav(c["pw"], c["ow"], zo.quarterwav(c["pw"], c["ow"]), cols[0] << np.uint32(32 - c["pw"]))
        rs.check(name, rs.port_words([out], [c["ow"]]))
    else:
        rc, q = zo.derive_qtbl(0, c["ow"], 2, c["pw"])
        assert rc == 0
        rs.check(name, rs.port_words([zo.quadtbl(q, cols[0])], [c["ow"]]))


def test_a_wrong_word_is_noticed():
    name = "p2r_cfg0_sweep"
    c = rs.CASES[name]
    rc, op = zo.derive_p2r(16, 16, 2, 16, 0)
    out = zo.rotate_const(op, c["x0"], c["y0"], rs.case_inputs(name)[2])
    out[12345, 1] ^= 1
    with pytest.raises(AssertionError, match="block 0"):
        rs.check(name, rs.port_words([out[:, 0], out[:, 1]], [op.ow, op.ow]))


# ---------------------------------------------------------------------------------------------- GPU tier: the engine
def _dev(a):
    import torch
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).cuda()


def _host(t):
    return t.cpu().numpy()


P2R_FLAGS = [zc.F_DEFAULT, zc.F_NO_SEED, zc.F_FORCE_GENERIC, zc.F_FORCE_SEED | zc.F_SEED_WORDS, zc.F_FORCE_SEED | zc.F_SEED_WORDS | zc.F_NO_DP2A,
             zc.F_FORCE_SEED | zc.F_SEED_PACKED, zc.F_FORCE_SEED | zc.F_SEED_PACKED | zc.F_NO_DP2A, zc.F_FORCE_SEED | zc.F_SEED_REGS]


@pytest.mark.gpu
@pytest.mark.parametrize("flags", P2R_FLAGS)
@pytest.mark.parametrize("name", [n for n in NAMES if rs.CASES[n]["kind"] == "p2r_const"])
def test_gpu_rotation_sweeps_equal_the_rtl(name, flags):
    c, d = rs.CASES[name], rs.CASES[name]["derive"]
    core = zc.Cordic(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"], sequential=bool(c.get("seq")))
    out = _host(core.rotate_const(c["x0"], c["y0"], _dev(rs.case_inputs(name)[2]), flags=flags))
    rs.check(name, rs.port_words([out[:, 0], out[:, 1]], [core.OW, core.OW]))


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [zc.F_DEFAULT, zc.F_NO_SEED, zc.F_FORCE_GENERIC, zc.F_FORCE_SEED, zc.F_FORCE_SEED | zc.F_NO_DP2A])
def test_gpu_rotation_per_sample_vectors_equal_the_rtl(flags):
    name = "p2r_cfg1_xy"
    d = rs.CASES[name]["derive"]
    core = zc.Cordic(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"])
    x, y, p = rs.case_inputs(name)
    out = _host(core.rotate(_dev(np.stack([x, y], 1).astype(np.int32)), _dev(p), flags=flags))
    rs.check(name, rs.port_words([out[:, 0], out[:, 1]], [core.OW, core.OW]))


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [zc.F_DEFAULT, zc.F_NO_TAIL, zc.F_FORCE_GENERIC])
@pytest.mark.parametrize("name", [n for n in NAMES if rs.CASES[n]["kind"] in ("r2p", "r2p_all")])
def test_gpu_vectoring_equals_the_rtl(name, flags):
    d = rs.CASES[name]["derive"]
    core = zc.Topolar(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"], sequential=bool(rs.CASES[name].get("seq")))
    x, y = rs.case_inputs(name)
    mag, ph = core.topolar(_dev(np.stack([x, y], 1).astype(np.int32)), flags=flags)
    rs.check(name, rs.port_words([_host(mag), _host(ph)], [core.OW, core.PW]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in NAMES if rs.CASES[n]["kind"] in ("tbl", "qtr", "qtbl")])
def test_gpu_table_cores_equal_the_rtl(name):
    c = rs.CASES[name]
    lut = {"tbl": zc.SinTable, "qtr": zc.QuarterWav}.get(c["kind"])
    lut = lut(phase_bits=c["pw"], ow=c["ow"]) if lut else zc.QuadTbl(ow=c["ow"], phase_bits=c["pw"])
    out = _host(lut.lookup(_dev(rs.case_inputs(name)[0] << np.uint32(32 - c["pw"]))))
    rs.check(name, rs.port_words([out], [c["ow"]]))
