"""Developer tool: timeline of one host-buffer encode call (dsrcgpu_encode_blocks): every payload copy and kernel with its start / end
on its slot's stream, relative to the start of the call (DSRCGPU_TIMELINE).
    python tools/e2e_timeline.py [reads] [profile]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TL = os.path.join(ROOT, "gpurun_out", "timeline.csv")
os.makedirs(os.path.dirname(TL), exist_ok=True)
os.environ["DSRCGPU_TIMELINE"] = TL
from dsrc_b200 import _lib  # noqa: E402
import torch  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
profile = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = _lib.lib()
ctx = C.c_void_p()
assert L.dsrcgpu_create(C.byref(ctx), 0, C.byref(_lib.Dataset(33, 0, 0)), C.byref(_lib.Settings(6, 2, 0, 0, 0)), 256 << 10, 8192) == 0
nbytes = reads * 372
d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
nb = C.c_uint64()
assert L.dsrcgpu_synth_fastq_device(ctx, profile, 99, 0, reads, C.c_void_p(d_in.data_ptr()), nbytes, C.byref(nb)) == 0
h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
h_in.copy_(d_in)
del d_in
h = h_in.numpy()
n = L.dsrcgpu_cut_blocks(h.ctypes.data_as(C.c_void_p), nbytes, 256 << 10, None, None, 0)
offs = np.zeros(n, dtype=np.uint64)
lens = np.zeros(n, dtype=np.uint32)
L.dsrcgpu_cut_blocks(h.ctypes.data_as(C.c_void_p), nbytes, 256 << 10, offs.ctypes.data_as(_lib.u64p), lens.ctypes.data_as(_lib.u32p), n)
h_out = torch.empty(nbytes // 2, dtype=torch.uint8, pin_memory=True)
sizes = np.zeros(n, dtype=np.uint32)


def run():
    t0 = time.perf_counter()
    rc = L.dsrcgpu_encode_blocks(ctx, C.c_void_p(h_in.data_ptr()), offs.ctypes.data_as(_lib.u64p), lens.ctypes.data_as(_lib.u32p), None, n,
                                 C.c_void_p(h_out.data_ptr()), nbytes // 2, sizes.ctypes.data_as(_lib.u32p), None, None)
    assert rc == 0, L.dsrcgpu_last_error(ctx)
    return time.perf_counter() - t0


run()
run()
open(TL, "w").close()
dt = run()
print("call wall %.1f ms -> %.2f GB/s; device-timed %.1f ms" % (dt * 1e3, nbytes / dt / 1e9, L.dsrcgpu_last_call_ms(ctx)))
rows = [l.strip().split(",") for l in open(TL, "rb").read().replace(b"\x00", b"").decode().splitlines() if l.count(",") == 3]
rows = [(int(r[0]), r[1], float(r[2]), float(r[3])) for r in rows]
rows.sort(key=lambda r: r[2])
cur = None
for slot, name, a, b in rows:
    if name == "copy_h2d":
        print()
    print("  slot %d %-14s %8.1f -> %8.1f  (%6.1f ms)" % (slot, name, a, b, b - a))
