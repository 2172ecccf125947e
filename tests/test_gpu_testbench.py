"""GPU tier: the reference's own, unmodified test benches (bench/cpp/cordic_tb.cpp, topolar_tb.cpp) running over
the GPU engine through the Verilator-shaped adaptors of cordic_b200/vshim (built by `make -C cordic_b200/vshim`
in the build container, where the reference tree is mounted; the binaries travel with the snapshot).  Their
printed statistics must equal, digit for digit, what the same benches print over the CPU oracle."""
import os
import subprocess

import pytest

from .conftest import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "cordic_b200", "vshim", "_ref")


def run(name, timeout=300):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip("%s not built (make -C cordic_b200/vshim needs the reference tree)" % name)
    env = dict(os.environ, ZC_VSHIM_STATS="1")
    return subprocess.run([exe], cwd="/tmp", env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("name,lines", [
    ("cordic_tb_gpu_shipped", ["AVG Err: 0.558302", "MAX Err: 1.924713", "CNR    : 78.63 dB", "SFDR =   93.88 dBc"]),
    ("cordic_tb_gpu_cfg0", ["AVG Err: 3.290393", "MAX Err: 11.192690", "CNR    : 81.29 dB", "SFDR =   88.35 dBc"]),
])
def test_reference_cordic_tb_over_gpu(name, lines):
    r = run(name)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SUCCESS!!" in r.stdout
    for l in lines:
        assert l in r.stdout, (l, r.stdout)
    assert "GPU batches" in r.stderr            # the samples really went through libzcordic


def test_reference_topolar_tb_over_gpu():
    r = run("topolar_tb_gpu_shipped")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SUCCESS" in r.stdout
    assert "Max phase     error: 6.40" in r.stdout and "Max magnitude error:  0.870814" in r.stdout


def test_reference_quadtbl_tb_over_gpu():
    r = run("quadtbl_tb_gpu_shipped")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SUCCESS!!" in r.stdout and "MXERR: 1.565887" in r.stdout and "SFDR =   89.37 dBc" in r.stdout


@pytest.mark.skipif(not os.environ.get("ZC_SLOW"), reason="2^24 clocked ticks through 22-sample batches: ~1 min; set ZC_SLOW=1")
def test_reference_cordic_tb_cfg1_over_gpu():
    r = run("cordic_tb_gpu_cfg1", timeout=900)
    assert r.returncode == 0 and "SUCCESS!!" in r.stdout
    assert "AVG Err: 0.597045" in r.stdout and "MAX Err: 2.345266" in r.stdout and "SFDR =  124.03 dBc" in r.stdout


def test_gpu_scoring_equals_the_test_bench_printout():
    """cordic_b200.score restates the TB maths on the GPU; its figures must equal what the reference's own test
    benches print for the same cores (the strings asserted above), to the printed precision."""
    import cordic_b200 as zc
    from cordic_b200 import score
    r = score.score_rotation(zc.Cordic(iw=13, ow=13, xtra=2))
    assert r["passed"] and abs(r["avg_err"] - 0.558302) < 2e-6 and abs(r["max_err"] - 1.924713) < 2e-6
    assert abs(r["cnr_db"] - 78.63) < 0.006 and abs(r["sfdr_dbc"] - 93.88) < 0.006 and abs(r["alpha"] - 1.000001) < 2e-6
    r = score.score_rotation(zc.Cordic(iw=18, ow=18, xtra=2, phase_bits=24, nstages=20))
    assert r["passed"] and abs(r["avg_err"] - 0.597045) < 2e-6 and abs(r["max_err"] - 2.345266) < 2e-6
    assert abs(r["cnr_db"] - 108.15) < 0.006 and abs(r["sfdr_dbc"] - 124.03) < 0.02
    t = score.score_topolar(zc.Topolar(iw=13, ow=13, xtra=2))
    assert t["passed"] and abs(t["max_phase_err"] - 6.40) < 0.006 and abs(t["max_mag_err"] - 0.870814) < 2e-6
    t = score.score_topolar(zc.Topolar(iw=16, ow=16, xtra=2))
    assert not t["passed"] and abs(t["max_phase_err"] - 9.23) < 0.006      # see test_oracle_golden: TB tuned for 13 bits
    q = zc.QuadTbl(ow=13, phase_bits=18)
    s = score.score_sine(q.lookup, 18, 13)
    assert abs(s["max_err"] - 1.565887) < 2e-6 and (s["max"], s["min"]) == (4095, -4096) and abs(s["sfdr_dbc"] - 89.37) < 0.006
