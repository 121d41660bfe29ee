"""Decode parity: the CUDA Read path (through the C ABI) reproduces the input FASTQ from blocks written by the oracle and by the
CUDA Store path; malformed blocks are reported, not crashed on."""
import hashlib

import numpy as np
import pytest

import cases
import refbind
import synth

pytestmark = pytest.mark.gpu

CASES = cases.all_cases()


def _bc(d, q, pr, max_block):
    from dsrc_b200 import BlockCompressor
    return BlockCompressor(33, bool(pr), d, q, max_block_bytes=max(max_block + 64, 1 << 16))


@pytest.mark.parametrize("name,data,d,q,pr", CASES, ids=[c[0] for c in CASES])
def test_read_matches_input(name, data, d, q, pr):
    ora = refbind.Oracle(33, pr, d, q)
    blk, _, _ = ora.store(data[:-1])
    bc = _bc(d, q, pr, len(data))
    got = bc.read(blk, out_cap=len(data) + 64)
    assert got == data
    assert got == ora.read(blk)
    bc.close()


@pytest.mark.parametrize("d,q", [(6, 2), (0, 0), (9, 1)])
def test_batch_roundtrip_through_gpu_only(d, q):
    """encode on the GPU, decode on the GPU, whole block queue in one call each"""
    big = synth.illumina(6000, seed=41, regime="full")
    ora = refbind.Oracle(33, 0, d, q)
    blocks = ora.cut(big, 1 << 18)
    bc = _bc(d, q, 0, 1 << 18)
    enc, _, _ = bc.store_many(big, [b[0] for b in blocks], [b[1] for b in blocks])
    arc = b"".join(enc)
    offs = np.cumsum([0] + [len(e) for e in enc[:-1]]).tolist()
    dec = bc.read_many(arc, offs, [len(e) for e in enc], out_cap=len(big) + 64)
    assert b"".join(dec) == big
    bc.close()


def test_corrupt_block_is_reported():
    from dsrc_b200 import DsrcGpuError
    data = synth.illumina(200, seed=3)
    blk, _, _ = refbind.Oracle(33, 0, 6, 2).store(data[:-1])
    bc = _bc(6, 2, 0, len(data))
    with pytest.raises(DsrcGpuError):
        bc.read(blk[:40], out_cap=len(data) + 64)               # truncated
    with pytest.raises(DsrcGpuError):
        bc.read(b"\0\0\0\0" + blk[4:], out_cap=len(data) + 64)  # zero records
    assert bc.read(blk, out_cap=len(data) + 64) == data
    bc.close()


def test_bench_shape_roundtrip_sha():
    """size-independent property on the bench workload's generator: decode(encode(x)) == x by SHA-256, 60 blocks of 256 KB"""
    import ctypes as C
    from dsrc_b200 import _lib
    L = _lib.lib()
    n = 42000
    h = np.empty(n * 372, dtype=np.uint8)
    nb = C.c_uint64()
    L.dsrcgpu_synth_fastq_host(0, 99, 0, n, h.ctypes.data_as(C.c_void_p), h.size, C.byref(nb))
    big = h.tobytes()
    blocks = refbind.Oracle().cut(big, 1 << 18)
    bc = _bc(6, 2, 0, 1 << 18)
    enc, _, _ = bc.store_many(big, [b[0] for b in blocks], [b[1] for b in blocks])
    arc = b"".join(enc)
    offs = np.cumsum([0] + [len(e) for e in enc[:-1]]).tolist()
    dec = bc.read_many(arc, offs, [len(e) for e in enc], out_cap=len(big) + 64)
    assert hashlib.sha256(b"".join(dec)).digest() == hashlib.sha256(big).digest()
    bc.close()


def test_crc32_mismatch_is_reported():
    """decode with -c verifies the stored CRC-32 words like BlockCompressor::VerifyChecksum (src/BlockCompressor.cpp:576-594)"""
    from dsrc_b200 import BlockCompressor, DsrcGpuError
    data = synth.illumina(400, seed=61, regime="full")
    blk, _, _ = refbind.Oracle(33, 0, 6, 2, crc=True).store(data[:-1])
    bc = BlockCompressor(33, False, 6, 2, max_block_bytes=1 << 18, calc_crc32=True)
    assert bc.read(blk, out_cap=len(data) + 64) == data
    bad = bytearray(blk)
    bad[17] ^= 0xFF                        # inside the stored title CRC word
    with pytest.raises(DsrcGpuError, match="CRC32"):
        bc.read(bytes(bad), out_cap=len(data) + 64)
    bc.close()
