#!/usr/bin/env python3
"""Regenerate tests/golden/*.json from the REAL reference generator.

Runs oracle/_ref/gencordic (the reference's sw/*.cpp compiled by oracle/Makefile from
/root/reference, never copied) over a matrix of command lines and records what it
emits: the header constants (rtl/X.h format, sw/basiccordic.cpp:465-498,
sw/topolar.cpp:428-446), the cordic_angle table and pre-rotation constants printed
into the Verilog (sw/cordiclib.cpp:157-200, sw/basiccordic.cpp:203-284,
sw/topolar.cpp:208-251), and digests + strided samples of the $readmemh LUT files
(sw/hexfile.cpp:78-89).  The JSON files are committed; this script only runs in the
build container (the GPU box has no /root/reference and uses the committed JSON).

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GEN = os.path.join(ROOT, "oracle", "_ref", "gencordic")

# (name, mode, iw, ow, xtra, pw, nstages) ; None = flag not given
CORDIC_MATRIX = [
    ("p2r_shipped", "p2r", 13, 13, 2, None, None),
    ("p2r_cfg0", "p2r", 16, 16, 2, 16, None),
    ("p2r_cfg1", "p2r", 18, 18, 2, 24, 20),
    ("p2r_cfg1_nauto", "p2r", 18, 18, 2, 24, None),
    ("p2r_16_16_auto", "p2r", 16, 16, 2, None, None),
    ("p2r_8_8_x0", "p2r", 8, 8, 0, None, None),
    ("p2r_12_16_x1", "p2r", 12, 16, 1, None, None),
    ("p2r_16_12_x3", "p2r", 16, 12, 3, None, None),
    ("p2r_24_24", "p2r", 24, 24, 2, None, None),
    ("p2r_default", "p2r", None, None, 2, None, None),
    ("p2r_only_o", "p2r", None, 14, 2, None, None),
    ("p2r_negx", "p2r", 10, 10, -5, None, None),
    ("p2r_manystages", "p2r", 6, 6, 2, 10, 30),
    ("p2r_x_none", "p2r", 12, 12, None, None, None),
    ("p2r_20_10", "p2r", 20, 10, 2, 18, None),
    ("r2p_shipped", "r2p", 13, 13, 2, None, None),
    ("r2p_cfg2", "r2p", 16, 16, 2, None, None),
    ("r2p_8_8_x0", "r2p", 8, 8, 0, None, None),
    ("r2p_12_16_x1", "r2p", 12, 16, 1, None, None),
    ("r2p_16_12_x3", "r2p", 16, 12, 3, None, None),
    ("r2p_20_20", "r2p", 20, 20, 2, None, None),
    ("r2p_10_10_p14_n20", "r2p", 10, 10, 2, 14, 20),
    ("r2p_negx", "r2p", 10, 10, -5, None, None),
    ("r2p_only_o", "r2p", None, 12, 2, None, None),
]

# (name, mode, iw, pw, ow)
LUT_MATRIX = [
    ("tbl_shipped", "tbl", None, None, 13),          # sw/Makefile:158-163
    ("qtr_shipped", "qtr", None, 18, None),          # sw/Makefile:165-171 (OW defaults to 24)
    ("tbl_p10_o8", "tbl", None, 10, 8),
    ("tbl_i12", "tbl", 12, None, None),
    ("tbl_p16", "tbl", None, 16, None),
    ("tbl_p20_o16", "tbl", None, 20, 16),
    ("qtr_p12_o12", "qtr", None, 12, 12),
    ("qtr_i14", "qtr", 14, None, None),
    ("qtr_p20_o16", "qtr", None, 20, 16),
    ("qtr_p4_o6", "qtr", None, 4, 6),
]


# (name, iw, ow, xtra, pw) for -t qtbl
QTBL_MATRIX = [
    ("qtbl_shipped", None, 13, None, 18),          # sw/Makefile:173-178 (NB=13, PB=18)
    ("qtbl_o13_pauto", None, 13, None, None),
    ("qtbl_o16_p20", None, 16, None, 20),
    ("qtbl_o10_p14_x1", None, 10, 1, 14),
    ("qtbl_o20_p24", None, 20, None, 24),
    ("qtbl_i12_o14_x3", 12, 14, 3, 22),
    ("qtbl_o8_p12", None, 8, None, 12),
]


# (name, mode, iw, ow, xtra, pw, nstages) for the sequential cores (-t sp2r / -t sr2p)
SEQ_MATRIX = [
    ("sp2r_shipped", "sp2r", 13, 13, 2, None, None),          # sw/Makefile:142-144
    ("sp2r_cfg0", "sp2r", 16, 16, 2, 16, None),
    ("sp2r_cfg1", "sp2r", 18, 18, 2, 24, 20),
    ("sp2r_12_16_x1", "sp2r", 12, 16, 1, None, None),
    ("sp2r_manystages", "sp2r", 6, 6, 2, 10, 30),
    ("sp2r_24_24", "sp2r", 24, 24, 2, None, None),
    ("sr2p_shipped", "sr2p", 13, 13, 2, None, None),          # sw/Makefile:122-124
    ("sr2p_cfg2", "sr2p", 16, 16, 2, None, None),
    ("sr2p_8_8_x0", "sr2p", 8, 8, 0, None, None),
    ("sr2p_10_10_p14_n20", "sr2p", 10, 10, 2, 14, 20),
    ("sr2p_20_20", "sr2p", 20, 20, 2, None, None),
]


def run_gen(args, cwd):
    r = subprocess.run([GEN] + args, cwd=cwd, capture_output=True, text=True)
    return r.returncode, r.stdout + r.stderr


def parse_header(path):
    out = {}
    for line in open(path):
        m = re.match(r"const\s+(int|double|bool)\s+(\w+)\s*=\s*([^;]+);", line)
        if m:
            out[m.group(2)] = m.group(3).strip()   # keep the printed text verbatim
    return out


def parse_verilog(path):
    txt = open(path).read()
    angles = []
    for m in re.finditer(r"cordic_angle\[\s*(\d+)\]\s*=\s*\d+'h([0-9a-f_]+);", txt):
        angles.append(int(m.group(2).replace("_", ""), 16))
    prerot = [int(m.group(1), 16) for m in
              re.finditer(r"ph\[0\]\s*<=\s*(?:i_phase\s*-\s*)?\d+'h([0-9a-f]+);", txt)]
    return angles, prerot


def load_hex(path):
    words, addr = {}, 0
    for tok in open(path).read().split():
        if tok.startswith("@"):
            addr = int(tok[1:], 16)
        else:
            words[addr] = int(tok, 16)
            addr += 1
    n = max(words) + 1
    return [words[k] for k in range(n)]


def main():
    if not os.path.exists(GEN):
        sys.exit("build oracle/_ref/gencordic first: make -C oracle")
    params = {}
    with tempfile.TemporaryDirectory() as td:
        for name, mode, iw, ow, x, pw, n in CORDIC_MATRIX:
            d = os.path.join(td, name)
            os.makedirs(d)
            fname = "cordic.v" if mode == "p2r" else "topolar.v"
            args = ["-vca", "-t", mode, "-f", fname, "-c"]
            if iw is not None: args += ["-i", str(iw)]
            if ow is not None: args += ["-o", str(ow)]
            if x is not None: args += ["-x", str(x)]
            if pw is not None: args += ["-p", str(pw)]
            if n is not None: args += ["-n", str(n)]
            rc, log = run_gen(args, d)
            assert rc == 0, (name, log)
            hdr = parse_header(os.path.join(d, fname[:-2] + ".h"))
            angles, prerot = parse_verilog(os.path.join(d, fname))
            params[name] = {
                "mode": mode,
                "args": {"iw": iw, "ow": ow, "xtra": x, "pw": pw, "nstages": n},
                "cmdline": " ".join(args),
                "header": hdr,
                "angles": angles,
                "prerot": prerot,
            }
        luts = {}
        for name, mode, iw, pw, ow in LUT_MATRIX:
            d = os.path.join(td, name)
            os.makedirs(d)
            fname = "sintable.v" if mode == "tbl" else "quarterwav.v"
            args = ["-vca", "-t", mode, "-f", fname]
            if iw is not None: args += ["-i", str(iw)]
            if pw is not None: args += ["-p", str(pw)]
            if ow is not None: args += ["-o", str(ow)]
            rc, log = run_gen(args, d)
            assert rc == 0, (name, log)
            txt = open(os.path.join(d, fname)).read()
            m = re.search(r"PW\s*=\s*(\d+),.*?\n\s*OW\s*=\s*(\d+)", txt, re.S)
            rpw, row = int(m.group(1)), int(m.group(2))
            words = load_hex(os.path.join(d, fname[:-2] + ".hex"))
            import numpy as np
            arr = np.asarray(words, dtype="<u4")
            luts[name] = {
                "mode": mode,
                "args": {"iw": iw, "pw": pw, "ow": ow},
                "cmdline": " ".join(args),
                "pw": rpw, "ow": row, "nwords": len(words),
                "sha256_le_u32": hashlib.sha256(arr.tobytes()).hexdigest(),
                "stride": 257,
                "samples": [int(v) for v in arr[::257]],
                "head": [int(v) for v in arr[:16]],
                "tail": [int(v) for v in arr[-16:]],
            }
        qtbls = {}
        for name, iw, ow, x, pw in QTBL_MATRIX:
            d = os.path.join(td, name)
            os.makedirs(d)
            args = ["-vca", "-t", "qtbl", "-f", "quadtbl.v", "-c"]
            if iw is not None: args += ["-i", str(iw)]
            if ow is not None: args += ["-o", str(ow)]
            if x is not None: args += ["-x", str(x)]
            if pw is not None: args += ["-p", str(pw)]
            rc, log = run_gen(args, d)
            assert rc == 0, (name, log)
            hdr = {}
            for line in open(os.path.join(d, "quadtbl.h")):
                m = re.match(r"const\s+(int|long|double|bool)\s+(\w+)\s*=\s*([^;]+);", line)
                if m:
                    hdr[m.group(2)] = m.group(3).strip()
            txt = open(os.path.join(d, "quadtbl.v")).read()
            lp = {k: int(re.search(k + r"\s*=\s*(\d+)", txt).group(1)) for k in ("LGTBL", "QBITS", "LBITS", "CBITS")}
            lp["XTRA"] = int(re.search(r"XTRA=\s*(\d+)", txt).group(1))
            qtbls[name] = {
                "args": {"iw": iw, "ow": ow, "xtra": x, "pw": pw}, "cmdline": " ".join(args),
                "header": hdr, "localparams": lp,
                "ctbl": load_hex(os.path.join(d, "quadtbl_ctbl.hex")),
                "ltbl": load_hex(os.path.join(d, "quadtbl_ltbl.hex")),
                "qtbl": load_hex(os.path.join(d, "quadtbl_qtbl.hex")),
            }
        seqs = {}
        for name, mode, iw, ow, x, pw, n in SEQ_MATRIX:
            d = os.path.join(td, name)
            os.makedirs(d)
            fname = "seqcordic.v" if mode == "sp2r" else "seqpolar.v"
            args = ["-vca", "-t", mode, "-f", fname, "-c"]
            if iw is not None: args += ["-i", str(iw)]
            if ow is not None: args += ["-o", str(ow)]
            if x is not None: args += ["-x", str(x)]
            if pw is not None: args += ["-p", str(pw)]
            if n is not None: args += ["-n", str(n)]
            rc, log = run_gen(args, d)
            assert rc == 0, (name, log)
            hpath = os.path.join(d, fname[:-2] + ".h")
            hdr = parse_header(hpath)
            cpo = int(re.search(r"#define\s+CLOCKS_PER_OUTPUT\s+(\d+)", open(hpath).read()).group(1))
            angles, _ = parse_verilog(os.path.join(d, fname))
            seqs[name] = {"mode": mode, "args": {"iw": iw, "ow": ow, "xtra": x, "pw": pw, "nstages": n},
                          "cmdline": " ".join(args), "header": hdr, "clocks_per_output": cpo, "angles": angles}
    with open(os.path.join(HERE, "gen_seq.json"), "w") as f:
        json.dump(seqs, f, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "gen_quadtbl.json"), "w") as f:
        json.dump(qtbls, f, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "gen_params.json"), "w") as f:
        json.dump(params, f, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "gen_luts.json"), "w") as f:
        json.dump(luts, f, indent=1, sort_keys=True)
    print("wrote", len(params), "cordic configs,", len(luts), "LUT configs,", len(qtbls), "quadtbl configs and", len(seqs), "sequential configs")


if __name__ == "__main__":
    main()
