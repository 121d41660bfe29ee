"""Host-side mirror of the reference's per-block interface, over the C ABI.

comp::BlockCompressor (src/BlockCompressor.h:66-73): one instance per worker, reused for successive blocks,
carrying the one piece of cross-block state that changes bytes (capacity of TagStats::fields).
`store` / `read` take one chunk like Store / Read; `store_many` / `read_many` hand a whole block queue to the GPU
(what DsrcCompressorST's loop, src/DsrcOperator.cpp:105-130, does one block at a time).
"""
import ctypes as C

import numpy as np

from . import _lib


class DsrcGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dsrc_b200: %s (%s)" % (msg, _lib.ERRORS.get(code, code)))
        self.code = code


class BlockCompressor:
    def __init__(self, quality_offset=33, plus_repetition=False, dna_order=0, quality_order=0,
                 max_block_bytes=8 << 20, max_inflight_blocks=0, device=0, calc_crc32=False):
        self.L = _lib.lib()
        ds = _lib.Dataset(quality_offset, int(plus_repetition), 0)
        cs = _lib.Settings(dna_order, quality_order, 0, 0, int(calc_crc32))
        self.h = C.c_void_p()
        rc = self.L.dsrcgpu_create(C.byref(self.h), device, C.byref(ds), C.byref(cs), max_block_bytes, max_inflight_blocks)
        if rc != 0:
            raise DsrcGpuError(rc, "dsrcgpu_create failed (no CUDA device? there is no CPU fallback)")
        self.tag_capacity = 0          # std::vector<Field>::capacity() of a fresh TagStats
        self.max_block_bytes = max_block_bytes

    def close(self):
        if self.h:
            self.L.dsrcgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, rc):
        return DsrcGpuError(rc, self.L.dsrcgpu_last_error(self.h).decode())

    def _tagcaps(self, data, offs, lens):
        """Q1: emulate the growth of the reference's field vector over the blocks in file order."""
        caps = np.empty(len(offs), dtype=np.uint32)
        mv = memoryview(data)
        for i, (o, l) in enumerate(zip(offs, lens)):
            caps[i] = self.tag_capacity
            end = o
            lim = o + min(l, 65536)
            chunk = bytes(mv[o:lim])
            nl = chunk.find(b"\n")
            cr = chunk.find(b"\r")
            cand = [x for x in (nl, cr) if x >= 0]
            tl = min(cand) if cand else len(chunk)
            nf = self.L.dsrcgpu_tag_field_count(chunk[:tl], tl)
            self.tag_capacity = self.L.dsrcgpu_tag_capacity_after(int(self.tag_capacity), nf)
        return caps

    def store_many(self, data, offs, lens, warm=False, caps=None):
        """data: bytes-like holding the blocks; returns (list of compressed blocks, raw sizes [n,4], comp sizes [n,4]).
        caps: explicit TagStats::fields capacities per block; default: this instance's running state."""
        n = len(offs)
        offs_a = np.ascontiguousarray(offs, dtype=np.uint64)
        lens_a = np.ascontiguousarray(lens, dtype=np.uint32)
        if caps is not None:
            caps = np.ascontiguousarray(caps, dtype=np.uint32)
        elif not warm:
            caps = self._tagcaps(data, offs, lens)
        buf = np.frombuffer(data, dtype=np.uint8)
        cap = int(lens_a.astype(np.uint64).sum()) * 3 // 2 + 8192 * n + (2 << 20)
        sizes = np.zeros(n, dtype=np.uint32)
        raw = np.zeros((n, 4), dtype=np.uint64)
        cmp_ = np.zeros((n, 4), dtype=np.uint64)
        for attempt in range(4):       # tiny blocks with long read-ID fields can expand (many Huffman trees in the tag header)
            out = np.empty(cap, dtype=np.uint8)
            rc = self.L.dsrcgpu_encode_blocks(
                self.h, buf.ctypes.data_as(C.c_void_p), offs_a.ctypes.data_as(_lib.u64p), lens_a.ctypes.data_as(_lib.u32p),
                caps.ctypes.data_as(_lib.u32p) if caps is not None else None, n,
                out.ctypes.data_as(C.c_void_p), cap, sizes.ctypes.data_as(_lib.u32p),
                raw.ctypes.data_as(_lib.u64p), cmp_.ctypes.data_as(_lib.u64p))
            if rc != -3:
                break
            cap *= 4
        if rc != 0:
            raise self._err(rc)
        blocks = []
        p = 0
        for s in sizes:
            blocks.append(out[p:p + int(s)].tobytes())
            p += int(s)
        return blocks, raw, cmp_

    def store(self, chunk):
        """== BlockCompressor::Store on one chunk (no trailing newline)."""
        blocks, raw, cmp_ = self.store_many(chunk, [0], [len(chunk)])
        return blocks[0], [int(x) for x in raw[0]], [int(x) for x in cmp_[0]]

    def read_many(self, data, offs, lens, out_cap):
        n = len(offs)
        offs_a = np.ascontiguousarray(offs, dtype=np.uint64)
        lens_a = np.ascontiguousarray(lens, dtype=np.uint32)
        buf = np.frombuffer(data, dtype=np.uint8)
        out = np.empty(out_cap, dtype=np.uint8)
        sizes = np.zeros(n, dtype=np.uint64)
        rc = self.L.dsrcgpu_decode_blocks(self.h, buf.ctypes.data_as(C.c_void_p), offs_a.ctypes.data_as(_lib.u64p),
                                          lens_a.ctypes.data_as(_lib.u32p), n, out.ctypes.data_as(C.c_void_p), out_cap,
                                          sizes.ctypes.data_as(_lib.u64p))
        if rc != 0:
            raise self._err(rc)
        res = []
        p = 0
        for s in sizes:
            res.append(out[p:p + int(s)].tobytes())
            p += int(s)
        return res

    def read(self, blk, out_cap=None):
        return self.read_many(blk, [0], [len(blk)], out_cap or (len(blk) * 40 + (1 << 20)))[0]

    def kernel_times(self):
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        ln = (C.c_uint32 * 16)()
        k = self.L.dsrcgpu_last_kernel_times(self.h, names, ms, ln, 16)
        return {names[i].decode(): (float(ms[i]), int(ln[i])) for i in range(k)}
