"""Parity tests proper: the CUDA path (through the C ABI, dsrc_b200.BlockCompressor) against the oracle on the same
inputs, bit-exact per block and per stream size, cold and warm compressor."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases
import refbind
import synth

pytestmark = pytest.mark.gpu

CASES = cases.all_cases()
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "blocks.json")


def _bc(d, q, pr, max_block):
    from dsrc_b200 import BlockCompressor
    return BlockCompressor(33, bool(pr), d, q, max_block_bytes=max(max_block + 64, 1 << 16))


@pytest.mark.parametrize("name,data,d,q,pr", CASES, ids=[c[0] for c in CASES])
def test_store_matches_oracle(name, data, d, q, pr):
    chunk = data[:-1]
    ora = refbind.Oracle(33, pr, d, q)
    bc = _bc(d, q, pr, len(chunk))
    for it in range(2):       # cold, then warm field vector
        exp, eraw, ecmp = ora.store(chunk)
        got, graw, gcmp = bc.store(chunk)
        assert graw == eraw
        assert gcmp == ecmp
        assert got == exp, (name, it)
    bc.close()


def test_store_matches_golden():
    """the committed fixtures (made from the unmodified reference by tests/golden/make_golden.py)"""
    gold = json.load(open(GOLDEN))
    by_name = {c[0]: c for c in CASES}
    for name, g in gold["cases"].items():
        _, data, d, q, pr = by_name[name]
        assert hashlib.sha256(data).hexdigest() == g["input_sha256"], "generator drift: " + name
        bc = _bc(d, q, pr, len(data))
        cold, _, _ = bc.store(data[:-1])
        warm, _, _ = bc.store(data[:-1])
        bc.close()
        assert hashlib.sha256(cold).hexdigest() == g["cold_sha256"], name
        assert hashlib.sha256(warm).hexdigest() == g["warm_sha256"], name


@pytest.mark.parametrize("d,q", [(6, 2), (0, 0), (9, 1)])
def test_batch_of_blocks_matches_oracle(d, q):
    """a block queue in one call: file-order Q1 emulation, dense output, sizes"""
    big = synth.illumina(6000, seed=31, regime="full")
    ora = refbind.Oracle(33, 0, d, q)
    blocks = ora.cut(big, 1 << 18)
    assert len(blocks) > 5
    bc = _bc(d, q, 0, 1 << 18)
    offs = [b[0] for b in blocks]
    lens = [b[1] for b in blocks]
    got, raw, cmp_ = bc.store_many(big, offs, lens)
    for i, (o, l) in enumerate(blocks):
        exp, eraw, ecmp = ora.store(big[o:o + l])
        assert got[i] == exp, i
        assert [int(x) for x in raw[i]] == eraw
        assert [int(x) for x in cmp_[i]] == ecmp
    bc.close()


def test_cutter_matches_oracle():
    from dsrc_b200 import _lib
    import ctypes as C
    L = _lib.lib()
    for seed, cbuf in [(1, 1 << 16), (2, 1 << 18), (3, 100000)]:
        big = synth.illumina(3000, seed=seed, regime="full")
        exp = refbind.Oracle().cut(big, cbuf)
        n = L.dsrcgpu_cut_blocks(big, len(big), cbuf, None, None, 0)
        off = np.zeros(n, dtype=np.uint64)
        ln = np.zeros(n, dtype=np.uint32)
        L.dsrcgpu_cut_blocks(big, len(big), cbuf, off.ctypes.data_as(_lib.u64p), ln.ctypes.data_as(_lib.u32p), n)
        assert [(int(a), int(b)) for a, b in zip(off, ln)] == [(int(a), int(b)) for a, b in exp]


def test_malformed_block_is_reported():
    from dsrc_b200 import DsrcGpuError
    bc = _bc(6, 2, 0, 1 << 16)
    with pytest.raises(DsrcGpuError):
        bc.store(b"@r1\nACGT\n+\nIII")         # quality shorter than sequence
    with pytest.raises(DsrcGpuError):
        bc.store(b"r1\nACGT\n+\nIIII")         # title does not start with '@'
    good, _, _ = bc.store(b"@r1\nACGT\n+\nIIII")
    assert good == refbind.Oracle(33, 0, 6, 2).store(b"@r1\nACGT\n+\nIIII")[0]
    bc.close()


def test_device_synth_matches_host_twin():
    import ctypes as C
    import torch
    from dsrc_b200 import _lib
    L = _lib.lib()
    bc = _bc(6, 2, 0, 1 << 18)
    for profile in (0, 1, 2):
        n = 3000
        cap = n * (372 if profile < 2 else 1300)
        d = torch.empty(cap, dtype=torch.uint8, device="cuda")
        nb = C.c_uint64()
        assert L.dsrcgpu_synth_fastq_device(bc.h, profile, 99, 12345, n, C.c_void_p(d.data_ptr()), cap, C.byref(nb)) == 0
        h = np.empty(cap, dtype=np.uint8)
        nb2 = C.c_uint64()
        assert L.dsrcgpu_synth_fastq_host(profile, 99, 12345, n, h.ctypes.data_as(C.c_void_p), cap, C.byref(nb2)) == 0
        assert nb.value == nb2.value and (profile == 2 or nb.value == cap)
        assert d[:nb.value].cpu().numpy().tobytes() == h[:nb.value].tobytes()
    bc.close()


def test_bench_shape_blocks_match_oracle():
    """the bench workloads' generators at 256 KB blocks, every block of a 40-block sample: Illumina shape (profile 0 and 1) at -d2 -q2,
    454 / Ion shape (profile 2: variable lengths, IUPAC codes -> 8-symbol order-7 DNA + 64-symbol quality) at -d3 -q2, and the
    binned Illumina shape at -d0 -q0 (RLE quality) -- BASELINE configs[1] and configs[2]"""
    import ctypes as C
    from dsrc_b200 import _lib
    L = _lib.lib()
    for profile, d, q in ((0, 6, 2), (1, 6, 2), (2, 9, 2), (0, 0, 0)):
        n = 28000 if profile < 2 else 14000
        h = np.empty(n * (372 if profile < 2 else 1300), dtype=np.uint8)
        nb = C.c_uint64()
        L.dsrcgpu_synth_fastq_host(profile, 99, 0, n, h.ctypes.data_as(C.c_void_p), h.size, C.byref(nb))
        big = h[:nb.value].tobytes()
        ora = refbind.Oracle(33, 0, d, q)
        blocks = ora.cut(big, 1 << 18)
        assert len(blocks) >= 40
        bc = _bc(d, q, 0, 1 << 18)
        got, _, _ = bc.store_many(big, [b[0] for b in blocks], [b[1] for b in blocks])
        for i, (o, l) in enumerate(blocks):
            assert got[i] == ora.store(big[o:o + l])[0], (profile, i)
        dec = bc.read_many(b"".join(got), np.cumsum([0] + [len(g) for g in got[:-1]]).tolist(), [len(g) for g in got], out_cap=len(big) + 64)
        assert b"".join(dec) == big, profile
        bc.close()


@pytest.mark.parametrize("d,q", [(6, 2), (0, 0), (9, 1)])
def test_crc32_blocks_match_oracle(d, q):
    """-c: CRC-32 words of titles / sequences / qualities in the block header"""
    from dsrc_b200 import BlockCompressor
    for data in (synth.illumina(900, seed=51, regime="full"), synth.ion454(200, seed=52, iupac=(d != 0)), synth.illumina(3, seed=53)):
        chunk = data[:-1]
        ora = refbind.Oracle(33, 0, d, q, crc=True)
        bc = BlockCompressor(33, False, d, q, max_block_bytes=max(len(chunk) + 64, 1 << 16), calc_crc32=True)
        for it in range(2):
            exp, _, ecmp = ora.store(chunk)
            got, _, gcmp = bc.store(chunk)
            assert gcmp == ecmp
            assert got == exp
        assert bc.read(got, out_cap=len(data) + 64) == data
        bc.close()


def test_narrow_stream_arenas_are_retried_wide(monkeypatch):
    """the range-coder streams get 1.25 bytes per symbol; a chain that runs out makes the library repeat the call with arenas no chain
    can outgrow (csrc/api.cu encode_retry). DSRCGPU_NARROW_STREAMS=64 shrinks the first attempt to 1/64 byte per symbol, so every
    block of these inputs takes the retry -- the result must not change."""
    monkeypatch.setenv("DSRCGPU_NARROW_STREAMS", "64")
    for data, d, q in [(synth.illumina(900, seed=41, regime="full"), 6, 2), (synth.random_quals(900, seed=42, n_levels=4), 3, 1)]:
        chunk = data[:-1]
        ora = refbind.Oracle(33, 0, d, q)
        bc = _bc(d, q, 0, len(chunk))
        for it in range(2):
            exp, eraw, ecmp = ora.store(chunk)
            got, graw, gcmp = bc.store(chunk)
            assert got == exp and graw == eraw and gcmp == ecmp, it
        bc.close()
