"""Developer tool: A/B builds of the library. Recompiles the named sources with extra nvcc flags and links them with the stock objects
of dsrc_b200/build into build_variants/libdsrc_<name>.so (git-ignored; select it with DSRC_B200_LIB=<path>).
    python tools/variant.py NAME parse.cu rc_model.cu -- -DPRE_U=4 -DFOO=1"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsrc_b200 import build as B  # noqa: E402

name = sys.argv[1]
rest = sys.argv[2:]
cut = rest.index("--") if "--" in rest else len(rest)
files, extra = rest[:cut], rest[cut + 1:]
B.build()
out_dir = os.path.join(ROOT, "build_variants")
os.makedirs(out_dir, exist_ok=True)
nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets"] + extra
objs, procs = [], []
for s in B.SRCS:
    stock = os.path.join(B.HERE, "build", s + ".o")
    if s in files:
        o = os.path.join(out_dir, "%s_%s.o" % (name, s))
        procs.append(subprocess.Popen([nvcc] + flags + ["-c", os.path.join(B.HERE, "csrc", s), "-o", o]))
        objs.append(o)
    else:
        objs.append(stock)
for p in procs:
    if p.wait() != 0:
        raise SystemExit("nvcc failed")
lib = os.path.join(out_dir, "libdsrc_%s.so" % name)
subprocess.check_call([nvcc, "-shared", "-o", lib] + objs)
print(lib)
