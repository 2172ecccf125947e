"""Golden vectors obtained by executing the reference's RTL text (tests/golden/rtl_vectors.json, made by
tests/golden/make_rtl_vectors.py with oracle/vsim.py over the reference generator's output and the checked-in
rtl/*.v).  CPU tier: the oracle must reproduce every vector; the simulator itself is checked against the language
rules it implements and shown to be sensitive to the RTL.  GPU tier: the CUDA path must reproduce every vector."""
import json
import os
import shutil

import numpy as np
import pytest

from . import zo
from .conftest import ROOT, has_reference

VEC = json.load(open(os.path.join(ROOT, "tests", "golden", "rtl_vectors.json")))
CORES = sorted(k for k, v in VEC.items() if not k.startswith("_") and "in" in v)


def mask(w):
    return (1 << w) - 1


def test_vector_file_covers_every_core_family():
    kinds = {VEC[k]["kind"] for k in CORES}
    assert kinds == {"p2r", "r2p", "qtbl", "tbl", "qtr", "sp2r", "sr2p"}
    assert sum(len(VEC[k]["in"]) for k in CORES) >= 21000
    # the two no-rounding command lines yield Verilog no tool can load (generator bug, sw/basiccordic.cpp:419-420)
    assert {k for k, v in VEC.items() if "unparseable_rtl" in v} == {"p2r_8_8_x0", "p2r_negx"}
    # sequential flavours of the same trouble: the no-rounding branch of sw/seqcordic.cpp tests an i_ce the core
    # does not have, and a seqpolar whose NSTAGES+1 is a power of two never raises o_done
    assert {k for k, v in VEC.items() if "unusable_rtl" in v} == {"sp2r_8_8_x0"}
    assert {k for k, v in VEC.items() if isinstance(v, dict) and v.get("never_done")} == {"sr2p_n15_never_done"}


def test_sequential_cores_the_reference_cannot_run_are_refused():
    d = VEC["sr2p_n15_never_done"]["derive"]
    rc, _ = zo.derive_sr2p(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"])
    assert rc != 0
    assert zo.derive_sr2p(10, 10, 2, 0, 31)[0] != 0 and zo.derive_sr2p(10, 10, 2, 0, 14)[0] == 0
    assert zo.derive_sp2r(10, 10, 2, 0, 2)[0] != 0          # output taken two iterations early: needs >= 3


@pytest.mark.parametrize("name", CORES)
def test_oracle_reproduces_the_simulated_rtl(name):
    v = VEC[name]
    d, prm = v["derive"], v["params"]
    if v["kind"] in ("sp2r", "sr2p"):
        # rtl/seqcordic.v, rtl/seqpolar.v and generated variants, driven through i_stb/o_done as the TB does
        rc, p = (zo.derive_sp2r if v["kind"] == "sp2r" else zo.derive_sr2p)(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"])
        assert rc == 0 and p.sequential == 1 and (p.iw, p.ow, p.ww, p.pw) == (prm["IW"], prm["OW"], prm["WW"], prm["PW"])
        assert zo.clocks_per_output(p) == v["clocks_per_output"]
        inp = np.array(v["in"], dtype=np.int64)
        want = np.array(v["out"], dtype=np.int64)
        if v["kind"] == "sp2r":
            got = zo.rotate(p, inp[:, :2].astype(np.int32), inp[:, 2].astype(np.uint32))
            assert ((got.astype(np.int64) & mask(p.ow)) == want).all()
        else:
            mag, ph = zo.topolar(p, inp.astype(np.int32))
            assert ((mag.astype(np.int64) & mask(p.ow)) == want[:, 0]).all() and (ph.astype(np.int64) == want[:, 1]).all()
    elif v["kind"] == "p2r":
        rc, p = zo.derive_p2r(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"])
        assert rc == 0 and (p.iw, p.ow, p.ww, p.pw, p.nstages) == (prm["IW"], prm["OW"], prm["WW"], prm["PW"], prm["NSTAGES"])
        inp = np.array(v["in"], dtype=np.int64)
        got = zo.rotate(p, inp[:, :2].astype(np.int32), inp[:, 2].astype(np.uint32))
        assert ((got.astype(np.int64) & mask(p.ow)) == np.array(v["out"], dtype=np.int64)).all()
    elif v["kind"] == "r2p":
        rc, p = zo.derive_r2p(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"])
        assert rc == 0 and (p.iw, p.ow, p.ww, p.pw, p.nstages) == (prm["IW"], prm["OW"], prm["WW"], prm["PW"], prm["NSTAGES"])
        mag, ph = zo.topolar(p, np.array(v["in"], dtype=np.int64).astype(np.int32))
        want = np.array(v["out"], dtype=np.int64)
        assert ((mag.astype(np.int64) & mask(p.ow)) == want[:, 0]).all() and (ph.astype(np.int64) == want[:, 1]).all()
    elif v["kind"] == "qtbl":
        rc, q = zo.derive_qtbl(d["iw"], d["ow"], d["xtra"], d["pw"])
        assert rc == 0 and (q.pw, q.ow, q.nextra, q.lgtbl, q.cbits, q.lbits, q.qbits) == tuple(
            prm[k] for k in ("PW", "OW", "XTRA", "LGTBL", "CBITS", "LBITS", "QBITS"))
        got = zo.quadtbl(q, np.array(v["in"], dtype=np.uint32))
        assert ((got.astype(np.int64) & mask(q.ow)) == np.array(v["out"], dtype=np.int64)).all()
    else:
        pw, ow = d["pw"], d["ow"]
        words = (np.array(v["in"], dtype=np.uint64) << (32 - pw)).astype(np.uint32)
        got = zo.lut_sin(pw, ow, zo.sintable(pw, ow), words) if v["kind"] == "tbl" else zo.lut_qwav(pw, ow, zo.quarterwav(pw, ow), words)
        assert ((got.astype(np.int64) & mask(ow)) == np.array(v["out"], dtype=np.int64)).all()


# ---- the simulator against the language rules it implements (IEEE 1364-2005 5.4/5.5) --------------------------
def _sim(tmp_path, body, **inputs):
    from oracle import vsim
    src = "module t(input wire clk, input wire signed [7:0] a, input wire [7:0] b, input wire signed [7:0] c);\n%s\nendmodule\n" % body
    path = os.path.join(str(tmp_path), "t.v")
    open(path, "w").write(src)
    m = vsim.Module(path)
    m.set(**inputs)
    return m


def test_simulator_expression_rules(tmp_path):
    m = _sim(tmp_path, """
        wire [15:0] y_mixed = a + b;            // one unsigned operand: both zero-extended
        wire [15:0] y_signed = a + c;           // all signed: sign-extended to the 16-bit context
        wire [15:0] y_cast = a + $signed(b);
        wire [7:0]  sra = a >>> 2;              // signed context: arithmetic
        wire [7:0]  srl = (a + b) >>> 2;        // unsigned context: logical
        wire [7:0]  shr = a >> 2;
        wire [15:0] part = a[7:4] + 16'd0;      // part-selects are unsigned
        wire [15:0] neg = -a;                   // unary minus at the context width
        wire [3:0]  rep = {2{a[7], b[1]}};
        wire        red = &a[7:1];
        wire        cmp_s = (a < c);            // both signed: signed compare
        wire        cmp_u = (a < b);            // mixed: unsigned compare
        wire [15:0] mul = a * c;                // signed multiply at 16 bits
        wire [8:0]  carry = a + c;              // 9-bit context keeps the carry of a signed add
    """, a=0xFF, b=0x01, c=0x02)
    g = m.get
    assert g("y_mixed") == 0x0100 and g("y_signed") == 0x0001 and g("y_cast") == 0x0000
    assert g("sra") == 0xFF and g("srl") == 0x00 and g("shr") == 0x3F
    assert g("part") == 0x000F and g("neg") == 0x0001 and g("rep") == 0b1010 and g("red") == 1
    assert g("cmp_s") == 1 and g("cmp_u") == 0
    assert g("mul") == 0xFFFE and g("carry") == 0x001


def test_simulator_nonblocking_and_generate(tmp_path):
    m = _sim(tmp_path, """
        localparam N = 4;
        reg [7:0] pipe [0:N];
        reg [7:0] swap_a, swap_b;
        genvar i;
        generate for (i = 0; i < N; i = i + 1) begin : stage
            always @(posedge clk) pipe[i+1] <= pipe[i] + 8'd1;
        end endgenerate
        always @(posedge clk) pipe[0] <= b;
        always @(posedge clk) begin swap_a <= swap_b; swap_b <= b; end
    """, b=10)
    for _ in range(5):
        m.tick()
    assert m.get("pipe") == [10, 11, 12, 13, 14]       # one register per clock, old values read
    m.set(b=99)
    m.tick()
    assert m.get("swap_a") == 10 and m.get("swap_b") == 99


@pytest.mark.skipif(not has_reference(), reason="reference tree not mounted")
def test_simulator_is_sensitive_to_the_rtl(tmp_path):
    """Change one arctan constant / one pre-rotation constant in a copy of rtl/cordic.v: the vectors must move."""
    from oracle import vsim
    v = VEC["p2r_shipped"]
    vecs = [dict(i_xval=x, i_yval=y, i_phase=p) for x, y, p in v["in"][:400]]
    ref = [tuple(o) for o in v["out"][:400]]
    src = open("/root/reference/rtl/cordic.v").read()
    assert vsim.run_pipeline(vsim.Module("/root/reference/rtl/cordic.v"), vecs, ["o_xval", "o_yval"]) == ref
    for old, new in (("20'h0_0051", "20'h0_0052"), ("i_phase - 20'h80000", "i_phase - 20'h80001"),
                     ("xv[i] + (yv[i]>>>(i+1))", "xv[i] + (yv[i]>>(i+1))")):
        assert old in src
        path = os.path.join(str(tmp_path), "cordic.v")
        open(path, "w").write(src.replace(old, new, 1))
        assert vsim.run_pipeline(vsim.Module(path), vecs, ["o_xval", "o_yval"]) != ref, old


# ---- GPU tier ------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", CORES)
def test_cuda_path_reproduces_the_simulated_rtl(name):
    torch = pytest.importorskip("torch")
    import cordic_b200 as zc
    v = VEC[name]
    d = v["derive"]

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()

    def host(t):
        torch.cuda.synchronize()
        return t.cpu().numpy().astype(np.int64)

    if v["kind"] in ("p2r", "sp2r"):
        core = zc.Cordic(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"], sequential=(v["kind"] == "sp2r"))
        if v["kind"] == "sp2r":
            assert core.CLOCKS_PER_OUTPUT == v["clocks_per_output"]
        inp = np.array(v["in"], dtype=np.int64)
        want = np.array(v["out"], dtype=np.int64)
        for fl in (zc.F_DEFAULT, zc.F_FORCE_SEED, zc.F_FORCE_GENERIC):
            got = host(core.rotate(dev(inp[:, :2].astype(np.int32)), dev(inp[:, 2].astype(np.uint32)), flags=fl))
            assert ((got & mask(core.OW)) == want).all(), fl
        # constant-vector entry point (seeded kernels): group the vectors by their (x, y)
        x0, y0 = int(inp[0, 0]), int(inp[0, 1])
        sel = (inp[:, 0] == x0) & (inp[:, 1] == y0)
        for fl in (zc.F_DEFAULT, zc.F_FORCE_SEED | zc.F_SEED_WORDS, zc.F_FORCE_SEED | zc.F_SEED_WORDS | zc.F_NO_DP2A,
                   zc.F_FORCE_SEED | zc.F_SEED_PACKED):
            got = host(core.rotate_const(x0, y0, dev(inp[sel, 2].astype(np.uint32)), flags=fl))
            assert ((got & mask(core.OW)) == want[sel]).all(), fl
    elif v["kind"] in ("r2p", "sr2p"):
        core = zc.Topolar(d["iw"], d["ow"], d["xtra"], d["pw"], d["nstages"], sequential=(v["kind"] == "sr2p"))
        want = np.array(v["out"], dtype=np.int64)
        for fl in (zc.F_DEFAULT, zc.F_FORCE_GENERIC):
            mag, ph = core.topolar(dev(np.array(v["in"], dtype=np.int64).astype(np.int32)), flags=fl)
            assert ((host(mag) & mask(core.OW)) == want[:, 0]).all() and ((host(ph) & 0xFFFFFFFF) == want[:, 1]).all()
    elif v["kind"] == "qtbl":
        core = zc.QuadTbl(d["iw"], d["ow"], d["xtra"], d["pw"])
        words = (np.array(v["in"], dtype=np.uint64) << (32 - core.PW)).astype(np.uint32)
        assert ((host(core.lookup(dev(words))) & mask(core.OW)) == np.array(v["out"], dtype=np.int64)).all()
    else:
        core = (zc.SinTable if v["kind"] == "tbl" else zc.QuarterWav)(phase_bits=d["pw"], ow=d["ow"])
        words = (np.array(v["in"], dtype=np.uint64) << (32 - core.PW)).astype(np.uint32)
        assert ((host(core.lookup(dev(words))) & mask(core.OW)) == np.array(v["out"], dtype=np.int64)).all()


@pytest.mark.skipif(not has_reference(), reason="reference tree not mounted")
def test_pipeline_latencies_match_what_the_adaptors_assume():
    """o_aux comes out NSTAGES+2 clocks after i_aux goes in for the CORDIC cores, 6 for quadtbl,
    1 for sintable, 3 for quarterwav -- the latencies the Verilator-shaped adaptors (oracle/shim, cordic_b200/vshim)
    are built around."""
    from oracle import vsim
    want = {"cordic.v": lambda m: m.consts["NSTAGES"] + 2, "topolar.v": lambda m: m.consts["NSTAGES"] + 2,
            "quadtbl.v": lambda m: 6, "sintable.v": lambda m: 1, "quarterwav.v": lambda m: 3}
    for fname, lat in want.items():
        m = vsim.Module(os.path.join("/root/reference/rtl", fname))
        m.set(i_reset=1, i_ce=1)
        m.tick()
        m.set(i_reset=0, i_aux=1)
        n = 0
        while not m.get("o_aux"):
            m.tick()
            n += 1
            assert n < 100
        assert n == lat(m), fname
