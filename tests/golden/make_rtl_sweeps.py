#!/usr/bin/env python3
"""Regenerate tests/golden/rtl_sweeps.json: checksums of LARGE output sets obtained by EXECUTING the reference's RTL.

tests/golden/rtl_vectors.json pins the oracle and the CUDA path on 21,504 individual vectors.  This file goes for volume:
every phase of the three rotation cores the benchmark cares about (rtl/cordic.v as shipped: 2^20 phases; BASELINE
configs[0]: 2^16; configs[1], the headline core: all 2^24), every phase of the shipped sintable / quarterwav / quadtbl
cores, and millions of seeded inputs for per-sample rotation and for the vectoring cores (rtl/topolar.v as shipped and
BASELINE configs[2]) -- 30.2 million samples (the two shipped sequential cores included, through their handshake), each clocked through the reference's Verilog text by oracle/vsim.py the
way bench/cpp/cordic_tb.cpp:127-200 drives the Verilated model.  Only digests are committed: per case the SHA-256 of
the little-endian int32 output array plus one CRC-32 per block of 2^16 samples (so a mismatch can be localised).
The inputs are regenerated from the seeds below by tests/rtl_sweeps.py on any machine; the reference tree is needed
only here.  ~10 minutes on 8 cores.

    python tests/golden/make_rtl_sweeps.py [case ...]
"""
import hashlib
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import vsim  # noqa: E402
from tests.rtl_sweeps import CASES, BLOCK, case_inputs  # noqa: E402

GEN = os.path.join(ROOT, "oracle", "_ref", "gencordic")
REF_RTL = "/root/reference/rtl"
_mods = {}


def _module(path):
    if path not in _mods:
        _mods[path] = vsim.Module(path)
    return _mods[path]


def _work(job):
    path, ins, outs, cols, cpo = job
    m = _module(path)
    vecs = [dict(zip(ins, row)) for row in zip(*[c.tolist() for c in cols])]
    if cpo:        # sequential core: i_stb for one clock, CLOCKS_PER_OUTPUT ticks per sample (cordic_tb.cpp:146-158)
        got = vsim.run_handshake(m, vecs, list(outs), cpo)
        assert got is not None, "handshake broken"
    else:
        got = vsim.run_pipeline(m, vecs, list(outs))
    assert len(got) == len(vecs), (path, len(got), len(vecs))
    return np.array(got, dtype=np.uint64).astype(np.uint32)


def main():
    if not (os.path.exists(GEN) and os.path.exists(REF_RTL)):
        sys.exit("needs the reference tree and oracle/_ref/gencordic (make -C oracle)")
    want = sys.argv[1:] or list(CASES)
    dst = os.path.join(HERE, "rtl_sweeps.json")
    out = json.load(open(dst)) if os.path.exists(dst) else {}
    out["_comment"] = "digests of reference-RTL outputs (oracle/vsim.py executing the Verilog text); see make_rtl_sweeps.py"
    with tempfile.TemporaryDirectory() as td, mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        for name in want:
            c = CASES[name]
            if c["gen"] is None:
                path = os.path.join(REF_RTL, c["file"])
            else:
                d = os.path.join(td, name)
                os.makedirs(d)
                r = subprocess.run([GEN, "-a", "-c"] + c["gen"].split() + ["-f", c["file"]], cwd=d, capture_output=True, text=True)
                assert r.returncode == 0, (name, r.stderr)
                path = os.path.join(d, c["file"])
            cols = case_inputs(name)                      # tuple of uint32 arrays, one per input port (raw port words)
            n = len(cols[0])
            m = _module(path)
            params = {k: m.consts[k] for k in ("IW", "OW", "WW", "PW", "NSTAGES") if k in m.consts}
            cpo = 0
            if c.get("seq"):
                import re
                cpo = int(re.search(r"#define\s+CLOCKS_PER_OUTPUT\s+(\d+)", open(path[:-2] + ".h").read()).group(1))
            step = BLOCK // 16 if cpo else BLOCK      # smaller jobs for the slow handshake (digests stay per BLOCK)
            jobs = [(path, c["ins"], c["outs"], tuple(col[i:i + step] for col in cols), cpo) for i in range(0, n, step)]
            parts = pool.map(_work, jobs, chunksize=1)
            words = np.concatenate(parts)                 # [n, len(outs)] raw (zero-extended) port words
            h = hashlib.sha256(words.astype("<u4").tobytes()).hexdigest()
            crcs = [zlib.crc32(words[i:i + BLOCK].astype("<u4").tobytes()) for i in range(0, n, BLOCK)]
            out[name] = {"n": n, "params": params, "sha256": h, "crc32": crcs,
                         "first": words[:4].tolist(), "last": words[-4:].tolist()}
            print(name, n, h[:16], flush=True)
            with open(dst, "w") as f:                     # checkpoint after every case
                json.dump(out, f, separators=(",", ":"), sort_keys=True)


if __name__ == "__main__":
    main()
