"""Builds libzcordic.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzcordic.so")
SOURCES = ["zc_api.cu", "zc_params.cpp"]
HEADERS = ["zc_internal.h", "zc_kernels.cuh", "zc_seeded.cuh", "zc_quadtbl.cuh", os.path.join(ROOT, "include", "zcordic.h")]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: experiment builds (kernel variants selected by -D macros) next to the product library."""
    if out is None and not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-o", out or LIB] + list(defines) + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libzcordic.so")
    if verbose:
        print(r.stdout + r.stderr)
    return out or LIB


if __name__ == "__main__":
    _out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
