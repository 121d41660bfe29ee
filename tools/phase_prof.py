"""Developer tool: per-kernel CUDA-event times and in-kernel phase cycle counters of the bench workload at a reduced size.
    python tools/phase_prof.py [reads] [profile] [inflight]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsrc_b200 import _lib  # noqa: E402
import torch  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
profile = int(sys.argv[2]) if len(sys.argv) > 2 else 0
inflight = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
d_order = int(sys.argv[4]) if len(sys.argv) > 4 else 6
q_order = int(sys.argv[5]) if len(sys.argv) > 5 else 2
L = _lib.lib()
ctx = C.c_void_p()
assert L.dsrcgpu_create(C.byref(ctx), 0, C.byref(_lib.Dataset(33, 0, 0)), C.byref(_lib.Settings(d_order, q_order, 0, 0, 0)), 256 << 10, inflight) == 0
nbytes = reads * 372
d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
nb = C.c_uint64()
assert L.dsrcgpu_synth_fastq_device(ctx, profile, 99, 0, reads, C.c_void_p(d_in.data_ptr()), nbytes, C.byref(nb)) == 0
h = d_in.cpu().numpy()
n = L.dsrcgpu_cut_blocks(h.ctypes.data_as(C.c_void_p), nbytes, 256 << 10, None, None, 0)
offs = np.zeros(n, dtype=np.uint64)
lens = np.zeros(n, dtype=np.uint32)
L.dsrcgpu_cut_blocks(h.ctypes.data_as(C.c_void_p), nbytes, 256 << 10, offs.ctypes.data_as(_lib.u64p), lens.ctypes.data_as(_lib.u32p), n)
d_out = torch.empty(nbytes // 2, dtype=torch.uint8, device="cuda")
sizes = np.zeros(n, dtype=np.uint32)


def run():
    rc = L.dsrcgpu_encode_blocks_device(ctx, C.c_void_p(d_in.data_ptr()), offs.ctypes.data_as(_lib.u64p), lens.ctypes.data_as(_lib.u32p), None, n,
                                        C.c_void_p(d_out.data_ptr()), nbytes // 2, sizes.ctypes.data_as(_lib.u32p), None, None)
    assert rc == 0, L.dsrcgpu_last_error(ctx)


run()
L.dsrcgpu_set_profiling(ctx, 3)
run()
ph = (C.c_uint64 * 64)()
L.dsrcgpu_phase_cycles(ctx, ph, 1)
run()
L.dsrcgpu_phase_cycles(ctx, ph, 1)
names = (C.c_char_p * 16)()
ms = (C.c_float * 16)()
ln = (C.c_uint32 * 16)()
k = L.dsrcgpu_last_kernel_times(ctx, names, ms, ln, 16)
tot = L.dsrcgpu_last_call_ms(ctx)
print("blocks %d  bytes %.2f GB  call %.1f ms  -> %.2f GB/s" % (n, nbytes / 1e9, tot, nbytes / tot / 1e6))
for i in range(k):
    print("  %-14s %8.2f ms  %3d launches" % (names[i].decode(), ms[i], ln[i]))
lab = {0: "q.pass1", 1: "q.pass2+", 2: "q.heads", 3: "q.scan_short", 4: "q.scan_long", 8: "d.pass1", 9: "d.pass2+", 10: "d.heads", 11: "d.scan_short", 12: "d.scan_long", 16: "q.tab ctx", 17: "q.tab sort", 18: "q.tab heads", 19: "q.tab short", 20: "q.tab long", 21: "q.tab store", 22: "q.tab clean", 24: "d.tab ctx", 25: "d.tab sort", 26: "d.tab heads", 27: "d.tab short", 28: "d.tab long", 29: "d.tab store", 30: "d.tab clean", 14: "d.direct", 7: "q.direct", 32: "q.part count", 33: "q.part scatter", 34: "q.part load+sort", 35: "q.part rows", 36: "q.part store", 38: "q.part oversize", 43: "q.part over short", 44: "q.part over long"}
for i in range(64):
    if ph[i]:
        print("  phase %-14s %10.1f Mcycles (sum over CTAs)" % (lab.get(i, str(i)), ph[i] / 1e6))
