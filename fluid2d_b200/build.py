"""Build libf2d_b200.so (and the -fmad=false twin used by the bit-exactness tests) with
nvcc for sm_100a, in-tree (fluid2d_b200/lib/), so the .so travels with the repo.

    python -m fluid2d_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SOURCES = ["f2d_operators.cu", "f2d_advection.cu", "f2d_diag.cu", "f2d_multigrid.cu", "f2d_comm.cu"]
HEADERS = ["f2d_common.cuh", os.path.join("..", "..", "include", "f2d_b200.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
BASE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--extended-lambda", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]

VARIANTS = {
    # product library: FMA contraction allowed (fp64 pipe is the second bound after HBM)
    "libf2d_b200.so": [],
    # same sources, no FMA contraction: must agree bit for bit with the CPU oracle
    # (oracle is built -ffp-contract=off); used by tests only
    "libf2d_b200_strict.so": ["-fmad=false"],
}


def _newest_source():
    t = 0.
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(CSRC, f) for f in HEADERS]
    for f in files:
        t = max(t, os.path.getmtime(f))
    return max(t, os.path.getmtime(os.path.abspath(__file__)))


def build(force=False, verbose=False, variants=None):
    os.makedirs(LIBDIR, exist_ok=True)
    newest = _newest_source()
    built = []
    jobs = []
    for name, extra in VARIANTS.items():
        if variants and name not in variants:
            continue
        out = os.path.join(LIBDIR, name)
        if not force and os.path.exists(out) and os.path.getmtime(out) >= newest:
            continue
        objs = []
        tag = name.replace(".so", "")
        procs = []
        for src in SOURCES:
            obj = os.path.join(LIBDIR, "%s_%s.o" % (tag, src.replace(".cu", "")))
            cmd = [NVCC] + BASE_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
            objs.append(obj)
        jobs.append((name, out, tag, procs, objs))
    # every translation unit of every variant compiles at the same time
    for name, out, tag, procs, objs in jobs:
        log = []
        for src, p in procs:
            o = p.communicate()[0].decode()
            log.append("== %s\n%s" % (src, o))
            if p.returncode != 0:
                sys.stderr.write(o)
                raise RuntimeError("nvcc failed on %s" % src)
        with open(os.path.join(LIBDIR, tag + ".ptxas.log"), "w") as f:
            f.write("\n".join(log))
        subprocess.check_call([NVCC, "-shared", "-o", out] + objs + ["-lcudart"])
        for o in objs:
            os.remove(o)
        built.append(out)
        if verbose:
            print("built", out)
    return built


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
