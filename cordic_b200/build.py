"""Builds libzcordic.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).  The kernel families are
separate translation units compiled in parallel, then linked into one shared library."""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzcordic.so")
OBJDIR = os.path.join(HERE, "build")
SOURCES = ["zc_api.cu", "zc_params.cpp", "zc_seedplan.cu", "zc_rot_const.cu", "zc_rot_nco.cu", "zc_rot_dirs.cu",
           "zc_rot_plain.cu", "zc_topolar.cu", "zc_multi.cpp"]
HEADERS = ["zc_internal.h", "zc_kernels.cuh", "zc_generic.cuh", "zc_seeded.cuh", "zc_seedplan.h", "zc_quadtbl.cuh",
           os.path.join(ROOT, "include", "zcordic.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(nvcc, src, obj, defines, verbose):
    cmd = [nvcc, "-O3", "-std=c++17"] + ARCH + ["-lineinfo", "-Xcompiler", "-fPIC,-O2,-Wall", "-I", os.path.join(ROOT, "include"),
                                                 "-I", CSRC, "-c", "-o", obj] + list(defines) + [src]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    return r.returncode, " ".join(cmd), r.stdout + r.stderr


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: experiment builds (kernel variants selected by -D macros) next to the product library."""
    if out is None and not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = OBJDIR if out is None else OBJDIR + "_" + os.path.basename(out)
    os.makedirs(objdir, exist_ok=True)
    jobs = [(os.path.join(CSRC, s), os.path.join(objdir, os.path.splitext(s)[0] + ".o")) for s in SOURCES]
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda j: _compile(nvcc, j[0], j[1], defines, verbose), jobs))
    for rc, cmd, log in results:
        if verbose:
            print(cmd)
            print(log)
        if rc != 0:
            sys.stderr.write(cmd + "\n" + log)
            raise RuntimeError("nvcc failed building libzcordic.so")
    cmd = [nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", out or LIB] + [o for _, o in jobs] + ["-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libzcordic.so")
    return out or LIB


if __name__ == "__main__":
    _out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
