"""ctypes binding of the C ABI (include/dsrc_b200.h). Loading fails loudly when the CUDA library is missing:
there is no CPU fallback on the product path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DSRC_B200_LIB") or os.path.join(HERE, "libdsrc_b200.so")   # override: developer A/B builds only

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class Dataset(C.Structure):
    _fields_ = [("quality_offset", C.c_uint32), ("plus_repetition", C.c_uint8), ("color_space", C.c_uint8)]


class Settings(C.Structure):
    _fields_ = [("dna_order", C.c_uint32), ("quality_order", C.c_uint32), ("tag_preserve_flags", C.c_uint64),
                ("lossy", C.c_uint8), ("calc_crc32", C.c_uint8)]


ERRORS = {0: "OK", -1: "CUDA", -2: "ARG", -3: "CAPACITY", -4: "MALFORMED", -5: "UNSUPPORTED", -6: "NOMEM"}

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("dsrc_b200: %s not built (run `python -m dsrc_b200.build`); no CPU fallback exists" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.dsrcgpu_create.restype = C.c_int
    L.dsrcgpu_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Dataset), C.POINTER(Settings), C.c_uint32, C.c_uint32]
    L.dsrcgpu_destroy.argtypes = [vp]
    L.dsrcgpu_last_error.restype = C.c_char_p
    L.dsrcgpu_last_error.argtypes = [vp]
    enc = [vp, vp, u64p, u32p, u32p, C.c_uint32, vp, C.c_uint64, u32p, u64p, u64p]
    L.dsrcgpu_encode_blocks.restype = C.c_int
    L.dsrcgpu_encode_blocks.argtypes = enc
    L.dsrcgpu_encode_blocks_device.restype = C.c_int
    L.dsrcgpu_encode_blocks_device.argtypes = enc
    dec = [vp, vp, u64p, u32p, C.c_uint32, vp, C.c_uint64, u64p]
    L.dsrcgpu_decode_blocks.restype = C.c_int
    L.dsrcgpu_decode_blocks.argtypes = dec
    L.dsrcgpu_decode_blocks_device.restype = C.c_int
    L.dsrcgpu_decode_blocks_device.argtypes = dec
    L.dsrcgpu_tag_field_count.restype = C.c_uint32
    L.dsrcgpu_tag_field_count.argtypes = [C.c_char_p, C.c_uint32]
    L.dsrcgpu_tag_capacity_after.restype = C.c_uint32
    L.dsrcgpu_tag_capacity_after.argtypes = [C.c_uint32, C.c_uint32]
    L.dsrcgpu_last_kernel_times.restype = C.c_int
    L.dsrcgpu_last_kernel_times.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), u32p, C.c_int]
    L.dsrcgpu_set_profiling.argtypes = [vp, C.c_int]
    L.dsrcgpu_phase_cycles.restype = C.c_int
    L.dsrcgpu_phase_cycles.argtypes = [vp, u64p, C.c_int]
    L.dsrcgpu_analyze_first_chunk.restype = C.c_int
    L.dsrcgpu_analyze_first_chunk.argtypes = [vp, C.c_uint64, C.POINTER(Dataset)]
    L.dsrcgpu_archive_footer_size.restype = C.c_uint64
    L.dsrcgpu_archive_footer_size.argtypes = [C.c_uint64]
    L.dsrcgpu_write_archive_header.restype = C.c_int
    L.dsrcgpu_write_archive_header.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.dsrcgpu_write_archive_footer.restype = C.c_int
    L.dsrcgpu_write_archive_footer.argtypes = [vp, C.c_uint64, u32p, C.c_uint64, C.POINTER(Dataset), C.POINTER(Settings)]
    L.dsrcgpu_read_archive_index.restype = C.c_int
    L.dsrcgpu_read_archive_index.argtypes = [vp, C.c_uint64, u64p, u64p, u32p, C.c_uint64, C.POINTER(Dataset), C.POINTER(Settings)]
    L.dsrcgpu_release_workspace.restype = C.c_int
    L.dsrcgpu_release_workspace.argtypes = [vp]
    L.dsrcgpu_device_alloc.restype = C.c_int
    L.dsrcgpu_device_alloc.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    L.dsrcgpu_device_free.restype = C.c_int
    L.dsrcgpu_device_free.argtypes = [vp, vp]
    L.dsrcgpu_memcpy_h2d.restype = C.c_int
    L.dsrcgpu_memcpy_h2d.argtypes = [vp, vp, vp, C.c_uint64]
    L.dsrcgpu_memcpy_d2h.restype = C.c_int
    L.dsrcgpu_memcpy_d2h.argtypes = [vp, vp, vp, C.c_uint64]
    L.dsrcgpu_host_alloc.restype = C.c_int
    L.dsrcgpu_host_alloc.argtypes = [C.c_uint64, C.POINTER(vp)]
    L.dsrcgpu_host_free.restype = C.c_int
    L.dsrcgpu_host_free.argtypes = [vp]
    L.dsrcgpu_synth_fastq_device.restype = C.c_int
    L.dsrcgpu_synth_fastq_device.argtypes = [vp, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, vp, C.c_uint64, u64p]
    L.dsrcgpu_synth_fastq_host.restype = C.c_int
    L.dsrcgpu_synth_fastq_host.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, vp, C.c_uint64, u64p]
    L.dsrcgpu_cut_blocks.restype = C.c_uint64
    L.dsrcgpu_cut_blocks.argtypes = [vp, C.c_uint64, C.c_uint64, u64p, u32p, C.c_uint64]
    L.dsrcgpu_cut_blocks_window.restype = C.c_uint64
    L.dsrcgpu_cut_blocks_window.argtypes = [vp, C.c_uint64, C.c_uint64, u64p, u32p, C.c_uint64, u32p]
    L.dsrcgpu_archive_footer_span.restype = C.c_int
    L.dsrcgpu_archive_footer_span.argtypes = [vp, u64p, u64p]
    L.dsrcgpu_read_archive_footer.restype = C.c_int
    L.dsrcgpu_read_archive_footer.argtypes = [vp, vp, C.c_uint64, C.c_uint64, u64p, u32p, C.c_uint64, C.POINTER(Dataset), C.POINTER(Settings)]
    L.dsrcgpu_last_call_ms.restype = C.c_float
    L.dsrcgpu_last_call_ms.argtypes = [vp]
    _lib = L
    return L
