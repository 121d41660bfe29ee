"""Pins the C restatement (oracle/dsrc_oracle.c) to the unmodified reference (oracle/_ref): byte-equal
blocks for every case, cold and warm compressor, and byte-equal decode. Skipped where the reference
library is not built (it is prebuilt and travels to the GPU box)."""
import hashlib

import pytest

import cases
import refbind

pytestmark = pytest.mark.skipif(not refbind.ref_available(), reason="oracle/_ref not built")

CASES = cases.all_cases()


@pytest.mark.parametrize("name,data,d,q,pr", CASES, ids=[c[0] for c in CASES])
def test_store_matches_reference(name, data, d, q, pr):
    chunk = data[:-1]
    o = refbind.Oracle(33, pr, d, q)
    r = refbind.Ref(33, pr, d, q)
    for it in range(2):  # first call: cold TagStats::fields vector, second: warm
        a, ra, ca = o.store(chunk)
        b, rb, cb = r.store(chunk)
        assert ra == rb
        assert ca == cb
        assert hashlib.sha256(a).digest() == hashlib.sha256(b).digest(), (name, it, len(a), len(b))


@pytest.mark.parametrize("name,data,d,q,pr", CASES, ids=[c[0] for c in CASES])
def test_read_roundtrip(name, data, d, q, pr):
    chunk = data[:-1]
    o = refbind.Oracle(33, pr, d, q)
    r = refbind.Ref(33, pr, d, q)
    blk, _, _ = r.store(chunk)
    out_ref = r.read(blk)
    out_orc = o.read(blk)
    assert out_orc == out_ref
    assert out_orc == data


def test_crlf_store_matches_reference():
    import synth
    data = synth.illumina(300, seed=4, crlf=True)
    chunk = data[:-2]
    for d, q in [(0, 0), (6, 2)]:
        a, _, _ = refbind.Oracle(33, 0, d, q).store(chunk)
        b, _, _ = refbind.Ref(33, 0, d, q).store(chunk)
        assert a == b


def test_whole_file_matches_reference_cli_path(tmp_path):
    """oracle cutter + container + codec == DsrcCompressorST (`dsrc c -t1`), 3 blocks at -b1."""
    import synth
    big = synth.illumina(8000, seed=7)
    src = tmp_path / "in.fq"
    src.write_bytes(big)
    R = refbind.Ref()
    for dl, ql in [(0, 0), (2, 2)]:
        dst = tmp_path / ("o%d%d.dsrc" % (dl, ql))
        assert R.compress_file(str(src), str(dst), dl, ql, 1, 1, 0) == 0
        ref = dst.read_bytes()
        orc = refbind.Oracle().compress(big, dl, ql, 1 << 20, 0)
        assert ref == orc
        assert refbind.Oracle().decompress(ref, len(big) + 64) == big


def test_crc32_blocks_and_archive_match_reference(tmp_path):
    """-c: three CRC-32 words in the block header, crc flag in the footer; verify-by-decode"""
    import synth
    for data, d, q in [(synth.illumina(500, seed=31), 6, 2), (synth.ion454(150, seed=32), 9, 2), (synth.illumina(300, seed=33, regime="full"), 0, 0)]:
        chunk = data[:-1]
        o = refbind.Oracle(33, 0, d, q, crc=True)
        r = refbind.Ref(33, 0, d, q, crc=True)
        for it in range(2):
            a, _, ca = o.store(chunk)
            b, _, cb = r.store(chunk)
            assert a == b and ca == cb
        assert o.read(a) == data and o.crc_ok()
        assert r.read(a) == data
        bad = bytearray(a)
        bad[len(bad) * 2 // 3] ^= 0x10
        try:
            o.read(bytes(bad))
            assert not o.crc_ok()
        except RuntimeError:
            pass
    big = synth.illumina(8000, seed=7)
    src = tmp_path / "in.fq"
    src.write_bytes(big)
    dst = tmp_path / "c.dsrc"
    assert refbind.Ref().compress_file(str(src), str(dst), 2, 2, 1, 1, 0, crc=True) == 0
    arc = dst.read_bytes()
    assert arc == refbind.Oracle().compress(big, 2, 2, 1 << 20, 0, crc=True)
    assert refbind.Oracle().decompress(arc, len(big) + 64) == big


def test_changing_title_structure_matches_reference_cli_path(tmp_path):
    """Q1 in general form: the capacity of TagStats::fields evolves across blocks whose first titles have different field counts"""
    import synth
    big = synth.changing_titles()
    src = tmp_path / "in.fq"
    src.write_bytes(big)
    dst = tmp_path / "o.dsrc"
    assert refbind.Ref().compress_file(str(src), str(dst), 2, 2, 1, 1, 0) == 0
    assert dst.read_bytes() == refbind.Oracle().compress(big, 2, 2, 1 << 20, 0)


@pytest.mark.parametrize("crlf", [False, True])
def test_cutter_window_ending_exactly_at_eof_matches_reference_cli_path(tmp_path, crlf):
    """IFastqStreamReader::ReadNextChunk, the `r <= 0` branch (src/FastqStream.cpp:66-69): when a full chunk window ends exactly at
    EOF the last chunk is the bare carry-over -- its size keeps the final newline (and CR). First window and second window."""
    import synth
    cbuf = 1 << 20
    first = synth.exact_size(cbuf, crlf=crlf)
    second_start = refbind.Oracle().cut(synth.exact_size(3 * cbuf, crlf=crlf), cbuf)[1][0]
    for data in (first, synth.exact_size(second_start + cbuf, crlf=crlf), synth.exact_size(cbuf + 1, crlf=crlf), synth.exact_size(cbuf - 1, crlf=crlf)):
        src = tmp_path / "in.fq"
        src.write_bytes(data)
        dst = tmp_path / "o.dsrc"
        assert refbind.Ref().compress_file(str(src), str(dst), 2, 2, 1, 1, 0) == 0
        assert dst.read_bytes() == refbind.Oracle().compress(data, 2, 2, cbuf, 0), len(data)
    blocks = refbind.Oracle().cut(first, cbuf)
    assert len(blocks) == 2 and blocks[1][0] + blocks[1][1] == len(first)       # the tail chunk keeps the file's last byte
