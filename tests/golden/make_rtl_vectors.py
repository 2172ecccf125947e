#!/usr/bin/env python3
"""Regenerate tests/golden/rtl_vectors.json: per-sample golden vectors obtained by EXECUTING the reference's RTL.

For each configuration the reference's real generator (oracle/_ref/gencordic, compiled from /root/reference/sw by
oracle/Makefile) prints the Verilog core, and oracle/vsim.py -- a simulator for the Verilog subset the generator
emits, IEEE 1364 expression rules -- clocks it exactly as bench/cpp/cordic_tb.cpp does (i_ce = i_aux = 1, collect
when o_aux).  For the shipped configurations the checked-in rtl/*.v are simulated as they lie.  Inputs are seeded
random port words plus the corner cases; outputs are the raw port words.  Runs in the build container only (the GPU
box has no reference tree); the JSON is committed.

    python tests/golden/make_rtl_vectors.py
"""
import json
import os
import random
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import vsim  # noqa: E402

GEN = os.path.join(ROOT, "oracle", "_ref", "gencordic")
REF_RTL = "/root/reference/rtl"
SEED = 20261017
NVEC = 768

# name -> (generator args or None for "simulate the checked-in file", checked-in file)
P2R = {
    "p2r_shipped": (None, "cordic.v", dict(iw=13, ow=13, xtra=2, pw=0, nstages=0)),
    "p2r_cfg0": ("-i 16 -o 16 -p 16 -x 2", None, dict(iw=16, ow=16, xtra=2, pw=16, nstages=0)),
    "p2r_cfg1": ("-i 18 -o 18 -p 24 -n 20 -x 2", None, dict(iw=18, ow=18, xtra=2, pw=24, nstages=20)),
    "p2r_8_8_x0": ("-i 8 -o 8 -x 0", None, dict(iw=8, ow=8, xtra=0, pw=0, nstages=0)),
    "p2r_12_16_x1": ("-i 12 -o 16 -x 1", None, dict(iw=12, ow=16, xtra=1, pw=0, nstages=0)),
    "p2r_16_12_x3": ("-i 16 -o 12 -x 3", None, dict(iw=16, ow=12, xtra=3, pw=0, nstages=0)),
    "p2r_manystages": ("-i 6 -o 6 -x 2 -p 10 -n 30", None, dict(iw=6, ow=6, xtra=2, pw=10, nstages=30)),
    "p2r_negx": ("-i 10 -o 10 -x -5", None, dict(iw=10, ow=10, xtra=-5, pw=0, nstages=0)),
    "p2r_24_24": ("-i 24 -o 24 -x 2", None, dict(iw=24, ow=24, xtra=2, pw=0, nstages=0)),
    "p2r_narrow_wraps": ("-i 4 -o 3 -x 0 -p 8 -n 6", None, dict(iw=4, ow=3, xtra=0, pw=8, nstages=6)),
}
R2P = {
    "r2p_shipped": (None, "topolar.v", dict(iw=13, ow=13, xtra=2, pw=0, nstages=0)),
    "r2p_cfg2": ("-i 16 -o 16 -x 2", None, dict(iw=16, ow=16, xtra=2, pw=0, nstages=0)),
    "r2p_8_8_x0": ("-i 8 -o 8 -x 0", None, dict(iw=8, ow=8, xtra=0, pw=0, nstages=0)),
    "r2p_12_16_x1": ("-i 12 -o 16 -x 1", None, dict(iw=12, ow=16, xtra=1, pw=0, nstages=0)),
    "r2p_10_10_p14_n20": ("-i 10 -o 10 -x 2 -p 14 -n 20", None, dict(iw=10, ow=10, xtra=2, pw=14, nstages=20)),
    "r2p_20_20": ("-i 20 -o 20 -x 2", None, dict(iw=20, ow=20, xtra=2, pw=0, nstages=0)),
}
QTBL = {
    "qtbl_shipped": (None, "quadtbl.v", dict(iw=0, ow=13, xtra=2, pw=18)),
    "qtbl_o16_p20": ("-o 16 -p 20", None, dict(iw=0, ow=16, xtra=2, pw=20)),
    "qtbl_o10_p14_x1": ("-o 10 -p 14 -x 1", None, dict(iw=0, ow=10, xtra=1, pw=14)),
    "qtbl_o20_p24": ("-o 20 -p 24", None, dict(iw=0, ow=20, xtra=2, pw=24)),
}
# the sequential cores (-t sp2r / -t sr2p; rtl/seqcordic.v, rtl/seqpolar.v), driven through their i_stb/o_done handshake
SP2R = {
    "sp2r_shipped": (None, "seqcordic.v", dict(iw=13, ow=13, xtra=2, pw=0, nstages=0)),
    "sp2r_cfg0": ("-i 16 -o 16 -p 16 -x 2", None, dict(iw=16, ow=16, xtra=2, pw=16, nstages=0)),
    "sp2r_cfg1": ("-i 18 -o 18 -p 24 -n 20 -x 2", None, dict(iw=18, ow=18, xtra=2, pw=24, nstages=20)),
    "sp2r_12_16_x1": ("-i 12 -o 16 -x 1", None, dict(iw=12, ow=16, xtra=1, pw=0, nstages=0)),
    "sp2r_manystages": ("-i 6 -o 6 -x 2 -p 10 -n 30", None, dict(iw=6, ow=6, xtra=2, pw=10, nstages=30)),
    "sp2r_n17": ("-i 10 -o 10 -x 2 -n 17", None, dict(iw=10, ow=10, xtra=2, pw=0, nstages=17)),
    "sp2r_8_8_x0": ("-i 8 -o 8 -x 0", None, dict(iw=8, ow=8, xtra=0, pw=0, nstages=0)),
}
SR2P = {
    "sr2p_shipped": (None, "seqpolar.v", dict(iw=13, ow=13, xtra=2, pw=0, nstages=0)),
    "sr2p_cfg2": ("-i 16 -o 16 -x 2", None, dict(iw=16, ow=16, xtra=2, pw=0, nstages=0)),
    "sr2p_8_8_x0": ("-i 8 -o 8 -x 0", None, dict(iw=8, ow=8, xtra=0, pw=0, nstages=0)),
    "sr2p_12_16_x1": ("-i 12 -o 16 -x 1", None, dict(iw=12, ow=16, xtra=1, pw=0, nstages=0)),
    "sr2p_10_10_p14_n20": ("-i 10 -o 10 -x 2 -p 14 -n 20", None, dict(iw=10, ow=10, xtra=2, pw=14, nstages=20)),
    "sr2p_n14": ("-i 10 -o 10 -x 2 -n 14", None, dict(iw=10, ow=10, xtra=2, pw=0, nstages=14)),
    "sr2p_n15_never_done": ("-i 10 -o 10 -x 2 -n 15", None, dict(iw=10, ow=10, xtra=2, pw=0, nstages=15)),
}
NSEQ = 384

LUT = {
    "tbl_shipped": (None, "sintable.v", dict(pw=17, ow=13)),
    "qtr_shipped": (None, "quarterwav.v", dict(pw=18, ow=24)),
    "tbl_p10_o8": ("-t tbl -p 10 -o 8", None, dict(pw=10, ow=8)),
    "qtr_p12_o12": ("-t qtr -p 12 -o 12", None, dict(pw=12, ow=12)),
}


def rtl_file(td, name, mode, gen_args, checked_in, fname):
    if gen_args is None:
        return os.path.join(REF_RTL, checked_in)
    d = os.path.join(td, name)
    os.makedirs(d)
    args = [GEN, "-a", "-c"] + gen_args.split() + (["-t", mode] if "-t" not in gen_args else []) + ["-f", fname]
    r = subprocess.run(args, cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, (name, r.stderr)
    return os.path.join(d, fname)


def corners(w):
    hi, lo = (1 << (w - 1)) - 1, 1 << (w - 1)
    return [0, 1, (1 << w) - 1, hi, lo, hi - 1, lo + 1]


def main():
    if not (os.path.exists(GEN) and os.path.exists(REF_RTL)):
        sys.exit("needs the reference tree and oracle/_ref/gencordic (make -C oracle)")
    rng = random.Random(SEED)
    out = {"_comment": "outputs of the reference RTL text executed by oracle/vsim.py; see make_rtl_vectors.py"}
    with tempfile.TemporaryDirectory() as td:
        for name, (gen_args, checked_in, derive) in P2R.items():
            try:
                m = vsim.Module(rtl_file(td, name, "p2r", gen_args, checked_in, "cordic.v"))
            except SyntaxError as e:
                # WW == OW+1 ("No rounding required", sw/basiccordic.cpp:407-444): the generator prints the output
                # register's `always @(posedge i_clk)` on the same line as a `// }}}` comment (:419-420), so the
                # emitted Verilog has a bare `if` at module level and no simulator can load it.  Recorded, not fixed.
                out[name] = {"kind": "p2r", "derive": derive, "unparseable_rtl": str(e)}
                continue
            IW, PW = m.consts["IW"], m.consts["PW"]
            vecs = [(x, y, p) for x in corners(IW)[:5] for y in corners(IW)[:5] for p in
                    [0, 1, (1 << PW) - 1] + [(o << (PW - 3)) + d & ((1 << PW) - 1) for o in range(8) for d in (-1, 0)]][:400]
            while len(vecs) < NVEC:
                vecs.append((rng.randrange(1 << IW), rng.randrange(1 << IW), rng.randrange(1 << PW)))
            got = vsim.run_pipeline(m, [dict(i_xval=x, i_yval=y, i_phase=p) for x, y, p in vecs], ["o_xval", "o_yval"])
            assert len(got) == len(vecs)
            out[name] = {"kind": "p2r", "derive": derive, "params": {k: m.consts[k] for k in ("IW", "OW", "WW", "PW", "NSTAGES")},
                         "in": [list(v) for v in vecs], "out": [list(g) for g in got]}
        for name, (gen_args, checked_in, derive) in R2P.items():
            m = vsim.Module(rtl_file(td, name, "r2p", gen_args, checked_in, "topolar.v"))
            IW = m.consts["IW"]
            vecs = [(x, y) for x in corners(IW) for y in corners(IW)]
            while len(vecs) < NVEC:
                vecs.append((rng.randrange(1 << IW), rng.randrange(1 << IW)))
            got = vsim.run_pipeline(m, [dict(i_xval=x, i_yval=y) for x, y in vecs], ["o_mag", "o_phase"])
            assert len(got) == len(vecs)
            out[name] = {"kind": "r2p", "derive": derive, "params": {k: m.consts[k] for k in ("IW", "OW", "WW", "PW", "NSTAGES")},
                         "in": [list(v) for v in vecs], "out": [list(g) for g in got]}
        for name, (gen_args, checked_in, derive) in QTBL.items():
            m = vsim.Module(rtl_file(td, name, "qtbl", gen_args, checked_in, "quadtbl.v"))
            PW = m.consts["PW"]
            vecs = [p & ((1 << PW) - 1) for q in range(4) for p in ((q << (PW - 2)) - 1, q << (PW - 2), (q << (PW - 2)) + 1)]
            while len(vecs) < NVEC:
                vecs.append(rng.randrange(1 << PW))
            got = vsim.run_pipeline(m, [dict(i_phase=p) for p in vecs], ["o_sin"])
            assert len(got) == len(vecs)
            out[name] = {"kind": "qtbl", "derive": derive, "params": {k: m.consts[k] for k in ("PW", "OW", "XTRA", "LGTBL", "CBITS", "LBITS", "QBITS")},
                         "in": vecs, "out": [g[0] for g in got]}
        for name, (gen_args, checked_in, derive) in LUT.items():
            mode = "qtr" if name.startswith("qtr") else "tbl"
            fname = "quarterwav.v" if mode == "qtr" else "sintable.v"
            m = vsim.Module(rtl_file(td, name, mode, gen_args, checked_in, fname))
            PW = m.consts["PW"]
            vecs = [p & ((1 << PW) - 1) for q in range(4) for p in ((q << (PW - 2)) - 1, q << (PW - 2), (q << (PW - 2)) + 1)]
            while len(vecs) < NVEC:
                vecs.append(rng.randrange(1 << PW))
            got = vsim.run_pipeline(m, [dict(i_phase=p) for p in vecs], ["o_val"])
            assert len(got) == len(vecs)
            out[name] = {"kind": mode, "derive": derive, "params": {"PW": PW, "OW": m.consts["OW"]},
                         "in": vecs, "out": [g[0] for g in got]}
        # sequential cores last, so the vectors above keep their random draws
        for table, mode, fname, ins, outs in ((SP2R, "sp2r", "seqcordic.v", ("i_xval", "i_yval", "i_phase"), ("o_xval", "o_yval")),
                                              (SR2P, "sr2p", "seqpolar.v", ("i_xval", "i_yval"), ("o_mag", "o_phase"))):
            for name, (gen_args, checked_in, derive) in table.items():
                path = rtl_file(td, name, mode, gen_args, checked_in, fname)
                hdr = open(path[:-2] + ".h").read()
                cpo = int(re.search(r"#define\s+CLOCKS_PER_OUTPUT\s+(\d+)", hdr).group(1))
                try:
                    m = vsim.Module(path)
                except (SyntaxError, KeyError) as e:
                    # WW == OW+1: sw/seqcordic.cpp's "no rounding" branch tests an i_ce the sequential core does
                    # not have; the emitted Verilog references an undeclared net and cannot be elaborated.
                    out[name] = {"kind": mode, "derive": derive, "unusable_rtl": repr(e)}
                    continue
                IW, PW = m.consts["IW"], m.consts["PW"]
                widths = [IW, IW, PW][:len(ins)]
                vecs = [tuple(v) for v in ([(x, y, p) for x in corners(IW)[:4] for y in corners(IW)[:4]
                                            for p in (0, (1 << PW) - 1, 1 << (PW - 1), (3 << (PW - 3)) - 1)]
                                           if len(ins) == 3 else [(x, y) for x in corners(IW) for y in corners(IW)])]
                while len(vecs) < NSEQ:
                    vecs.append(tuple(rng.randrange(1 << w) for w in widths))
                got = vsim.run_handshake(m, [dict(zip(ins, v)) for v in vecs], list(outs), cpo)
                entry = {"kind": mode, "derive": derive, "clocks_per_output": cpo,
                         "params": {k: m.consts[k] for k in ("IW", "OW", "WW", "PW")}}
                if got is None:        # the core broke the TB's protocol (o_done never rose within CLOCKS_PER_OUTPUT)
                    entry["never_done"] = True
                else:
                    entry["in"] = [list(v) for v in vecs]
                    entry["out"] = [list(g) for g in got]
                out[name] = entry
    with open(os.path.join(HERE, "rtl_vectors.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"), sort_keys=True)
    print("wrote", sum(len(v.get("in", [])) for k, v in out.items() if not k.startswith("_")), "vectors for", len(out) - 1, "cores")


if __name__ == "__main__":
    main()
