"""Builds dsrc_b200/libdsrc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = ["api.cu", "parse.cu", "tags.cu", "rc_model.cu", "q0.cu", "finish.cu", "decode.cu", "synth.cu", "crc.cu", "container.cpp"]
LIB = os.path.join(HERE, "libdsrc_b200.so")


def build(force=False, verbose=False):
    srcs = [os.path.join(HERE, "csrc", s) for s in SRCS]
    deps = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + \
        [os.path.join(os.path.dirname(HERE), "include", h) for h in ("dsrc_b200.h", "dsrc_b200_bench.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps if os.path.exists(d)):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    flags += os.environ.get("DSRC_NVCC_FLAGS", "").split()
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        procs.append(subprocess.Popen([nvcc] + flags + ["-c", s, "-o", o]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
