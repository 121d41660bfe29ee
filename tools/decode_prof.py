"""Developer tool: decode throughput of the bench workload at a reduced size (device-resident), verified against the input.
    python tools/decode_prof.py [reads] [profile] [inflight]"""
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
inflight = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
L = _lib.lib()
ctx = C.c_void_p()
assert L.dsrcgpu_create(C.byref(ctx), 0, C.byref(_lib.Dataset(33, 0, 0)), C.byref(_lib.Settings(6, 2, 0, 0, 0)), 256 << 10, 8192) == 0     # encoder
nbytes = reads * 372
d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
nb = C.c_uint64()
assert L.dsrcgpu_synth_fastq_device(ctx, profile, 99, 0, reads, C.c_void_p(d_in.data_ptr()), nbytes, C.byref(nb)) == 0
h = d_in.cpu().numpy()
n = L.dsrcgpu_cut_blocks(h.ctypes.data_as(C.c_void_p), nbytes, 256 << 10, None, None, 0)
offs = np.zeros(n, dtype=np.uint64)
lens = np.zeros(n, dtype=np.uint32)
L.dsrcgpu_cut_blocks(h.ctypes.data_as(C.c_void_p), nbytes, 256 << 10, offs.ctypes.data_as(_lib.u64p), lens.ctypes.data_as(_lib.u32p), n)
d_arc = torch.empty(nbytes // 2, dtype=torch.uint8, device="cuda")
sizes = np.zeros(n, dtype=np.uint32)
rc = L.dsrcgpu_encode_blocks_device(ctx, C.c_void_p(d_in.data_ptr()), offs.ctypes.data_as(_lib.u64p), lens.ctypes.data_as(_lib.u32p), None, n,
                                    C.c_void_p(d_arc.data_ptr()), nbytes // 2, sizes.ctypes.data_as(_lib.u32p), None, None)
assert rc == 0, L.dsrcgpu_last_error(ctx)
L.dsrcgpu_destroy(ctx)
ctx = C.c_void_p()
os.environ.setdefault("DSRCGPU_DEC_BATCH", str(inflight))
assert L.dsrcgpu_create(C.byref(ctx), 0, C.byref(_lib.Dataset(33, 0, 0)), C.byref(_lib.Settings(6, 2, 0, 0, 0)), 256 << 10, inflight) == 0  # decoder
coffs = np.concatenate([[0], np.cumsum(sizes.astype(np.uint64))[:-1]]).astype(np.uint64)
d_out = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
osz = np.zeros(n, dtype=np.uint64)
for it in range(3):
    rc = L.dsrcgpu_decode_blocks_device(ctx, C.c_void_p(d_arc.data_ptr()), coffs.ctypes.data_as(_lib.u64p), sizes.ctypes.data_as(_lib.u32p), n,
                                        C.c_void_p(d_out.data_ptr()), nbytes + 64, osz.ctypes.data_as(_lib.u64p))
    assert rc == 0, L.dsrcgpu_last_error(ctx)
    ms = L.dsrcgpu_last_call_ms(ctx)
    print("decode: %d blocks, %.2f GB FASTQ, %.1f ms -> %.2f GB/s" % (n, nbytes / 1e9, ms, nbytes / ms / 1e6))
assert int(osz.sum()) == nbytes
assert torch.equal(d_out[:nbytes], d_in), "decode mismatch"
print("decoded bytes identical to the input")
names = (C.c_char_p * 16)()
ms_ = (C.c_float * 16)()
ln = (C.c_uint32 * 16)()
k = L.dsrcgpu_last_kernel_times(ctx, names, ms_, ln, 16)
for i in range(k):
    print("  %-14s %8.2f ms  %3d launches" % (names[i].decode(), ms_[i], ln[i]))
