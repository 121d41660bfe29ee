"""Whole-archive parity through the operator mirror on the GPU: == `dsrc c -t1` (oracle compress path / golden fixtures),
and decompress(archive) == input."""
import hashlib
import json
import os

import pytest

import refbind
import synth

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "blocks.json")))


@pytest.mark.parametrize("dl,ql", [(0, 0), (1, 1), (2, 2), (3, 2)])
def test_archive_matches_reference_cli_path(dl, ql):
    from dsrc_b200 import DsrcCompressorMT, DsrcDecompressorMT, InputParameters
    big = synth.illumina(8000, seed=7)
    arc, _ = DsrcCompressorMT().process(InputParameters(dl, ql, 1, 0), big)
    g = GOLD["archives"]["illumina8000_seed7_d%d_q%d_b1" % (dl, ql)]
    assert hashlib.sha256(arc).hexdigest() == g["archive_sha256"]
    assert arc == refbind.Oracle().compress(big, dl, ql, 1 << 20, 0)
    assert DsrcDecompressorMT().process(arc) == big


def test_archive_256k_blocks_and_module_files(tmp_path):
    from dsrc_b200 import DsrcCompressorMT, DsrcDecompressorMT, DsrcModule, InputParameters
    big = synth.illumina(9000, seed=23, regime="full", small_field=True)
    arc, _ = DsrcCompressorMT().process(InputParameters(2, 2, 1, 0, block_bytes=1 << 18), big)
    assert arc == refbind.Oracle().compress(big, 2, 2, 1 << 18, 0)
    assert DsrcDecompressorMT().process(arc) == big
    src = tmp_path / "in.fq"
    src.write_bytes(big)
    m = DsrcModule(dna_level=2, quality_level=2, buffer_mb=1)
    m.compress(str(src), str(tmp_path / "o.dsrc"))
    m.decompress(str(tmp_path / "o.dsrc"), str(tmp_path / "back.fq"))
    assert (tmp_path / "back.fq").read_bytes() == big
    assert (tmp_path / "o.dsrc").read_bytes() == refbind.Oracle().compress(big, 2, 2, 1 << 20, 0)


def test_decompress_reference_archive_454():
    from dsrc_b200 import DsrcDecompressorMT
    data = synth.ion454(400, seed=8)
    arc = refbind.Oracle().compress(data, 3, 2, 1 << 20, 0)
    assert DsrcDecompressorMT().process(arc) == data


def test_archive_with_crc32_matches_reference():
    from dsrc_b200 import DsrcCompressorMT, DsrcDecompressorMT, InputParameters
    big = synth.illumina(8000, seed=7)
    arc, _ = DsrcCompressorMT().process(InputParameters(2, 2, 1, 0, calculateCrc32=True), big)
    assert arc == refbind.Oracle().compress(big, 2, 2, 1 << 20, 0, crc=True)
    assert hashlib.sha256(arc).hexdigest() == GOLD["archives"]["illumina8000_seed7_d2_q2_b1_crc"]["archive_sha256"]
    assert DsrcDecompressorMT().process(arc) == big


def test_archive_8mb_blocks_reference_default():
    """-b8 (the reference's default buffer): one 8 MB block plus a remainder, -d2 -q2 and -d0 -q0"""
    from dsrc_b200 import DsrcCompressorMT, DsrcDecompressorMT, InputParameters
    big = synth.illumina(26000, seed=71)
    assert len(big) > (8 << 20)
    for dl, ql in [(2, 2), (0, 0)]:
        arc, _ = DsrcCompressorMT().process(InputParameters(dl, ql, 8, 0), big)
        assert arc == refbind.Oracle().compress(big, dl, ql, 8 << 20, 0)
        assert DsrcDecompressorMT().process(arc) == big


def test_archive_changing_title_structure():
    """field-vector capacity carried across blocks with different field counts (general form)"""
    from dsrc_b200 import DsrcCompressorMT, DsrcDecompressorMT, InputParameters
    big = synth.changing_titles()
    arc, _ = DsrcCompressorMT().process(InputParameters(2, 2, 1, 0), big)
    assert arc == refbind.Oracle().compress(big, 2, 2, 1 << 20, 0)
    assert DsrcDecompressorMT().process(arc) == big
