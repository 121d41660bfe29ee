"""ctypes bindings for the two CPU checkers: oracle/liboracle.so (our C restatement) and
oracle/_ref/libdsrcref.so (the unmodified reference, built by oracle/Makefile). Test infra only."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)


def _buf(b):
    return (C.c_uint8 * max(1, len(b))).from_buffer_copy(b if len(b) else b"\0")


class Oracle:
    def __init__(self, qoff=33, plus_rep=0, dna_order=0, qua_order=0, crc=False):
        self.lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        L = self.lib
        L.dsrc_oracle_create.restype = C.c_void_p
        L.dsrc_oracle_create.argtypes = [C.c_uint32, C.c_int, C.c_uint32, C.c_uint32]
        L.dsrc_oracle_destroy.argtypes = [C.c_void_p]
        L.dsrc_oracle_store.restype = C.c_int64
        L.dsrc_oracle_store.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, _u8p, C.c_uint64, _u64p, _u64p]
        L.dsrc_oracle_read.restype = C.c_int64
        L.dsrc_oracle_read.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, _u8p, C.c_uint64]
        L.dsrc_oracle_cut_blocks.restype = C.c_uint64
        L.dsrc_oracle_cut_blocks.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, _u64p, _u64p, C.c_uint64]
        L.dsrc_oracle_compress_mem.restype = C.c_int64
        L.dsrc_oracle_compress_mem.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, _u8p, C.c_uint64]
        L.dsrc_oracle_decompress_mem.restype = C.c_int64
        L.dsrc_oracle_decompress_mem.argtypes = [C.c_char_p, C.c_uint64, _u8p, C.c_uint64]
        L.dsrc_oracle_analyze.restype = C.c_int
        L.dsrc_oracle_analyze.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.dsrc_oracle_set_crc.argtypes = [C.c_void_p, C.c_int]
        L.dsrc_oracle_last_crc_ok.restype = C.c_int
        L.dsrc_oracle_last_crc_ok.argtypes = [C.c_void_p]
        L.dsrc_oracle_compress_mem_crc.restype = C.c_int64
        L.dsrc_oracle_compress_mem_crc.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int, _u8p, C.c_uint64]
        self.h = L.dsrc_oracle_create(qoff, plus_rep, dna_order, qua_order)
        if crc:
            L.dsrc_oracle_set_crc(self.h, 1)

    def crc_ok(self):
        return bool(self.lib.dsrc_oracle_last_crc_ok(self.h))

    def __del__(self):
        try:
            self.lib.dsrc_oracle_destroy(self.h)
        except Exception:
            pass

    def store(self, chunk):
        cap = len(chunk) * 2 + (1 << 16)
        out = (C.c_uint8 * cap)()
        raw = (C.c_uint64 * 4)()
        cmp_ = (C.c_uint64 * 4)()
        n = self.lib.dsrc_oracle_store(self.h, chunk, len(chunk), out, cap, raw, cmp_)
        if n < 0:
            raise RuntimeError("oracle store failed: %d" % n)
        return bytes(out[:n]), list(raw), list(cmp_)

    def read(self, blk, cap=None):
        cap = cap or (len(blk) * 40 + (1 << 20))
        out = (C.c_uint8 * cap)()
        n = self.lib.dsrc_oracle_read(self.h, blk, len(blk), out, cap)
        if n < 0:
            raise RuntimeError("oracle read failed: %d" % n)
        return bytes(out[:n])

    def analyze(self, chunk, qoff=0):
        q = C.c_uint32(qoff)
        pr = C.c_int(0)
        cs = C.c_int(0)
        ok = self.lib.dsrc_oracle_analyze(chunk, len(chunk), C.byref(q), C.byref(pr), C.byref(cs))
        return ok, q.value, bool(pr.value), bool(cs.value)

    def cut(self, data, cbuf):
        n = self.lib.dsrc_oracle_cut_blocks(data, len(data), cbuf, None, None, 0)
        off = (C.c_uint64 * (n + 1))()
        ln = (C.c_uint64 * (n + 1))()
        self.lib.dsrc_oracle_cut_blocks(data, len(data), cbuf, off, ln, n)
        return [(off[i], ln[i]) for i in range(n)]

    def compress(self, data, d, q, buf_bytes, qoff=0, crc=False):
        cap = len(data) * 2 + (1 << 16)
        out = (C.c_uint8 * cap)()
        n = self.lib.dsrc_oracle_compress_mem_crc(data, len(data), d, q, buf_bytes, qoff, int(crc), out, cap)
        if n < 0:
            raise RuntimeError("oracle compress failed: %d" % n)
        return bytes(out[:n])

    def decompress(self, arc, cap):
        out = (C.c_uint8 * cap)()
        n = self.lib.dsrc_oracle_decompress_mem(arc, len(arc), out, cap)
        if n < 0:
            raise RuntimeError("oracle decompress failed: %d" % n)
        return bytes(out[:n])


def ref_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdsrcref.so"))


class Ref:
    """the unmodified reference's BlockCompressor (one instance == one worker's compressor)."""

    def __init__(self, qoff=33, plus_rep=0, dna_order=0, qua_order=0, crc=False):
        self.lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libdsrcref.so"))
        L = self.lib
        L.ref_bc_create.restype = C.c_void_p
        L.ref_bc_create.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.ref_bc_destroy.argtypes = [C.c_void_p]
        L.ref_bc_store.restype = C.c_longlong
        L.ref_bc_store.argtypes = [C.c_void_p, C.c_char_p, C.c_ulonglong, _u8p, C.c_ulonglong, _u64p, _u64p]
        L.ref_bc_read.restype = C.c_longlong
        L.ref_bc_read.argtypes = [C.c_void_p, C.c_char_p, C.c_ulonglong, _u8p, C.c_ulonglong]
        L.ref_compress_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint]
        L.ref_decompress_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.ref_compress_file_crc.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int]
        self.h = L.ref_bc_create(qoff, plus_rep, 0, dna_order, qua_order, 0, int(crc))

    def __del__(self):
        try:
            self.lib.ref_bc_destroy(self.h)
        except Exception:
            pass

    def store(self, chunk):
        cap = len(chunk) * 2 + (1 << 16)
        out = (C.c_uint8 * cap)()
        raw = (C.c_uint64 * 4)()
        cmp_ = (C.c_uint64 * 4)()
        n = self.lib.ref_bc_store(self.h, chunk, len(chunk), out, cap, raw, cmp_)
        if n < 0:
            raise RuntimeError("ref store failed")
        return bytes(out[:n]), list(raw), list(cmp_)

    def read(self, blk, cap=None):
        cap = cap or (len(blk) * 40 + (1 << 20))
        out = (C.c_uint8 * cap)()
        n = self.lib.ref_bc_read(self.h, blk, len(blk), out, cap)
        if n < 0:
            raise RuntimeError("ref read failed")
        return bytes(out[:n])

    def compress_file(self, src, dst, d, q, buf_mb, threads=1, qoff=0, crc=False):
        return self.lib.ref_compress_file_crc(src.encode(), dst.encode(), d, q, buf_mb, threads, qoff, int(crc))

    def decompress_file(self, src, dst, threads=1):
        return self.lib.ref_decompress_file(src.encode(), dst.encode(), threads)


def shim_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdsrcshim.so"))


class Shim:
    """host/BlockCompressorGpu.h -- the C++ binding of INTEGRATION.md, compiled against the reference's own headers
    (oracle/shim_harness.cpp) -- driven exactly like Ref: one instance == one worker's compressor, on the GPU."""

    def __init__(self, qoff=33, plus_rep=0, dna_order=0, qua_order=0, crc=False, max_block=1 << 20):
        self.lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libdsrcshim.so"))
        L = self.lib
        L.shim_bc_create.restype = C.c_void_p
        L.shim_bc_create.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32]
        L.shim_bc_destroy.argtypes = [C.c_void_p]
        L.shim_bc_store.restype = C.c_longlong
        L.shim_bc_store.argtypes = [C.c_void_p, C.c_char_p, C.c_ulonglong, _u8p, C.c_ulonglong, _u64p, _u64p]
        L.shim_bc_read.restype = C.c_longlong
        L.shim_bc_read.argtypes = [C.c_void_p, C.c_char_p, C.c_ulonglong, _u8p, C.c_ulonglong]
        self.h = L.shim_bc_create(qoff, plus_rep, 0, dna_order, qua_order, 0, int(crc), max_block)
        if not self.h:
            raise RuntimeError("BlockCompressorGpu: no CUDA device")

    def __del__(self):
        try:
            if self.h:
                self.lib.shim_bc_destroy(self.h)
        except Exception:
            pass

    def store(self, chunk):
        cap = len(chunk) * 2 + (1 << 16)
        out = (C.c_uint8 * cap)()
        raw = (C.c_uint64 * 4)()
        cmp_ = (C.c_uint64 * 4)()
        n = self.lib.shim_bc_store(self.h, chunk, len(chunk), out, cap, raw, cmp_)
        if n < 0:
            raise RuntimeError("shim store failed: %d" % n)
        return bytes(out[:n]), list(raw), list(cmp_)

    def read(self, blk, cap=None):
        cap = cap or (len(blk) * 40 + (1 << 20))
        out = (C.c_uint8 * cap)()
        n = self.lib.shim_bc_read(self.h, blk, len(blk), out, cap)
        if n < 0:
            raise RuntimeError("shim read failed: %d" % n)
        return bytes(out[:n])

    # host/DsrcOperatorGpu.h: IDsrcOperator::Process over the batch ABI (file to file)
    def compress_file(self, src, dst, d, q, buf_mb, qoff=0, crc=False):
        L = self.lib
        L.shim_compress_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_char_p, C.c_int]
        err = C.create_string_buffer(512)
        rc = L.shim_compress_file(src.encode(), dst.encode(), d, q, buf_mb, qoff, int(crc), err, 512)
        return rc, err.value.decode()

    def decompress_file(self, src, dst):
        L = self.lib
        L.shim_decompress_file.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        err = C.create_string_buffer(512)
        rc = L.shim_decompress_file(src.encode(), dst.encode(), err, 512)
        return rc, err.value.decode()

    # host/DsrcModuleGpu.h: wrap::DsrcModule surface (Configurable setters + Compress/Decompress)
    def module_roundtrip(self, fastq, archive, back, d, q, buf_mb, crc=False):
        L = self.lib
        L.shim_module_roundtrip.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        err = C.create_string_buffer(512)
        rc = L.shim_module_roundtrip(fastq.encode(), archive.encode(), back.encode(), d, q, buf_mb, int(crc), err, 512)
        return rc, err.value.decode()
