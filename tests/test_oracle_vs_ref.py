"""Pins the C restatement (oracle/dsrc_oracle.c) to the unmodified reference (oracle/_ref): byte-equal
blocks for every case, cold and warm compressor, and byte-equal decode. Skipped where the reference
library is not built (it is prebuilt and travels to the GPU box)."""
import hashlib

import pytest

import cases
import refbind

pytestmark = pytest.mark.skipif(not refbind.ref_available(), reason="oracle/_ref not built")

CASES = cases.small_cases()


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
