"""CPU-side checks: the oracle against the committed golden fixtures (made from the unmodified reference), and
that the C-ABI library loads and exports every symbol include/dsrc_b200.h declares (no device work)."""
import ctypes as C
import hashlib
import json
import os
import re

import pytest

import cases
import refbind
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "blocks.json")))
CASES = {c[0]: c for c in cases.all_cases()}


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_oracle_matches_golden_blocks(name):
    g = GOLD["cases"][name]
    _, data, d, q, pr = CASES[name]
    assert sha(data) == g["input_sha256"], "generator drift"
    if "input_hex" in g:
        assert data.hex() == g["input_hex"]
    o = refbind.Oracle(33, pr, d, q)
    cold, raw, cmp_ = o.store(data[:-1])
    warm, _, _ = o.store(data[:-1])
    assert (sha(cold), sha(warm)) == (g["cold_sha256"], g["warm_sha256"])
    assert raw == g["raw_streams"] and cmp_ == g["comp_streams"]
    assert o.read(warm) == data


@pytest.mark.parametrize("name", sorted(GOLD["archives"]))
def test_oracle_matches_golden_archives(name):
    g = GOLD["archives"][name]
    big = synth.illumina(8000, seed=7)
    assert sha(big) == g["input_sha256"]
    arc = refbind.Oracle().compress(big, g["dna_level"], g["quality_level"], g["buf_mb"] << 20, 0, crc=g.get("crc", False))
    assert sha(arc) == g["archive_sha256"]
    assert refbind.Oracle().decompress(arc, len(big) + 64) == big


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dsrc_b200.h")).read() + open(os.path.join(ROOT, "include", "dsrc_b200_bench.h")).read()
    names = sorted(set(re.findall(r"\b(dsrcgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15
    lib_path = os.path.join(ROOT, "dsrc_b200", "libdsrc_b200.so")
    if not os.path.exists(lib_path):
        from dsrc_b200 import build
        build.build()
    L = C.CDLL(lib_path)
    for n in names:
        assert hasattr(L, n), n


def test_q1_host_helpers():
    from dsrc_b200 import _lib
    L = _lib.lib()
    t = b"@SIM.1 A00123:45:HXXXXXXX:1:1101:1000:2000 1:N:0:ACGTACGT"
    assert L.dsrcgpu_tag_field_count(t, len(t)) == 13
    cap = 0
    seen = []
    for nf in (13, 13, 20, 3):
        cap = L.dsrcgpu_tag_capacity_after(cap, nf)
        seen.append(cap)
    assert seen == [16, 16, 32, 32]        # libstdc++ doubling growth; capacity never shrinks (TagModeler.h:116-118)


def test_no_device_means_loud_failure():
    """without a CUDA device the product must fail (no CPU fallback)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    from dsrc_b200 import BlockCompressor, DsrcGpuError
    with pytest.raises(DsrcGpuError):
        BlockCompressor(33, False, 6, 2)
