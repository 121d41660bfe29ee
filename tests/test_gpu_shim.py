"""The drop-in itself: host/BlockCompressorGpu.h (the C++ class a DSRC maintainer puts in place of comp::BlockCompressor,
INTEGRATION.md binding 1), compiled against the reference's own headers by oracle/Makefile (target `shim`), against the
UNMODIFIED reference's BlockCompressor driven through identical calls -- first ("cold") and second ("warm") Store of one
instance (SURVEY 8-Q1 lives in the shim's tagCapacity), StreamsInfo, and Read in both directions (each side decodes the
other's block)."""
import pytest

import cases
import refbind

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (refbind.ref_available() and refbind.shim_available()), reason="oracle/_ref harnesses not built")]

NAMES = ["ill_binned_d6_q2", "ill_full_d6_q2", "ion454_d9_q2", "ill_binned_d0_q0", "ill_plusrep_d6_q2", "mixed_titles_d6_q2",
         "q1_valuevar_d6_q2", "ill_mid12_d3_q1", "tiny2_d6_q2"]
CASES = [c for c in cases.small_cases() if c[0] in NAMES]


@pytest.mark.parametrize("name,data,d,q,pr", CASES, ids=[c[0] for c in CASES])
def test_shim_is_a_drop_in_for_block_compressor(name, data, d, q, pr):
    chunk = data[:-1]
    ref = refbind.Ref(33, pr, d, q)
    gpu = refbind.Shim(33, pr, d, q, max_block=len(chunk) + 64)
    for call in ("cold", "warm"):
        exp, eraw, ecmp = ref.store(chunk)
        got, graw, gcmp = gpu.store(chunk)
        assert got == exp, "%s Store differs" % call
        assert (graw, gcmp) == (eraw, ecmp)
    assert gpu.read(exp) == data          # the GPU decodes the reference's block
    assert ref.read(got) == data          # the reference decodes the GPU's block


def test_shim_crc_blocks():
    name, data, d, q, pr = [c for c in cases.small_cases() if c[0] == "ill_binned_d6_q2"][0]
    chunk = data[:-1]
    ref = refbind.Ref(33, pr, d, q, crc=True)
    gpu = refbind.Shim(33, pr, d, q, crc=True, max_block=len(chunk) + 64)
    exp, _, _ = ref.store(chunk)
    got, _, _ = gpu.store(chunk)
    assert got == exp
    assert gpu.read(exp) == data


@pytest.mark.parametrize("d,q,buf_mb,crc", [(2, 2, 1, False), (0, 0, 1, False), (3, 1, 1, True)])
def test_operator_archive_is_byte_identical_to_dsrc_c_t1(tmp_path, d, q, buf_mb, crc):
    """host/DsrcOperatorGpu.h: DsrcCompressorGpu / DsrcDecompressorGpu (IDsrcOperator::Process, file to file) against the
    reference's single-thread operators (`dsrc c -t1` / `dsrc d`), in both directions."""
    import synth
    big = synth.illumina(8000, seed=7)
    src = str(tmp_path / "in.fastq")
    open(src, "wb").write(big)
    ref = refbind.Ref()
    gpu = refbind.Shim()
    a_ref, a_gpu = str(tmp_path / "ref.dsrc"), str(tmp_path / "gpu.dsrc")
    assert ref.compress_file(src, a_ref, d, q, buf_mb, threads=1, crc=crc) == 0
    rc, err = gpu.compress_file(src, a_gpu, d, q, buf_mb, crc=crc)
    assert rc == 0, err
    assert open(a_gpu, "rb").read() == open(a_ref, "rb").read()
    back_gpu, back_ref = str(tmp_path / "gpu.fastq"), str(tmp_path / "ref.fastq")
    rc, err = gpu.decompress_file(a_ref, back_gpu)          # the GPU operator reads the reference's archive
    assert rc == 0, err
    assert open(back_gpu, "rb").read() == big
    assert ref.decompress_file(a_gpu, back_ref, threads=1) == 0   # and the reference reads the GPU operator's
    assert open(back_ref, "rb").read() == big


def test_operator_errors_follow_the_reference(tmp_path):
    gpu = refbind.Shim()
    rc, err = gpu.compress_file(str(tmp_path / "missing.fastq"), str(tmp_path / "x.dsrc"), 2, 2, 1)
    assert rc != 0 and "Cannot open file to read:" in err
    bad = str(tmp_path / "bad.fastq")
    open(bad, "wb").write(b"this is not FASTQ\n" * 10)
    rc, err = gpu.compress_file(bad, str(tmp_path / "y.dsrc"), 2, 2, 1)
    assert rc != 0 and err
    junk = str(tmp_path / "junk.dsrc")
    open(junk, "wb").write(b"\x00" * 100)
    rc, err = gpu.decompress_file(junk, str(tmp_path / "z.fastq"))
    assert rc != 0 and "Invalid archive" in err


def test_module_compress_decompress(tmp_path):
    """host/DsrcModuleGpu.h: wrap::DsrcModule's surface (Configurable setters, Compress, Decompress, DsrcException on error)."""
    import synth
    big = synth.illumina(6000, seed=9, regime="full")
    src, arc, back, a_ref = (str(tmp_path / n) for n in ("in.fastq", "m.dsrc", "m.fastq", "ref.dsrc"))
    open(src, "wb").write(big)
    gpu = refbind.Shim()
    rc, err = gpu.module_roundtrip(src, arc, back, 2, 2, 1)
    assert rc == 0, err
    assert open(back, "rb").read() == big
    assert refbind.Ref().compress_file(src, a_ref, 2, 2, 1, threads=1) == 0
    assert open(arc, "rb").read() == open(a_ref, "rb").read()
    rc, err = gpu.module_roundtrip(str(tmp_path / "missing"), arc, back, 2, 2, 1)
    assert rc != 0 and "Cannot open file to read:" in err


@pytest.mark.parametrize("crlf", [False, True])
def test_operator_streams_the_file_through_a_bounded_window(tmp_path, monkeypatch, crlf):
    """host/DsrcOperatorGpu.h is windowed (bounded host RAM whatever the file size): with a 4 MiB window a 13 MB file goes through in
    four windows of 1 MB chunks -- the blocks at the window seams, the carried field-vector capacity and the reader's
    CRLF state must come out exactly as `dsrc c -t1` writes them; the decompressor is windowed too."""
    import synth
    monkeypatch.setenv("DSRCGPU_WINDOW_MB", "4")
    big = synth.illumina(30000, seed=61, small_field=True, crlf=crlf) + synth.exact_size(3 << 20, seed=62, crlf=crlf)
    assert len(big) > (12 << 20)
    src, a_ref, a_gpu, back = (str(tmp_path / n) for n in ("in.fastq", "ref.dsrc", "gpu.dsrc", "gpu.fastq"))
    open(src, "wb").write(big)
    assert refbind.Ref().compress_file(src, a_ref, 2, 2, 1, threads=1) == 0
    gpu = refbind.Shim()
    rc, err = gpu.compress_file(src, a_gpu, 2, 2, 1)
    assert rc == 0, err
    assert open(a_gpu, "rb").read() == open(a_ref, "rb").read()
    monkeypatch.setenv("DSRCGPU_WINDOW_MB", "2")          # 512 KiB of compressed blocks per window on the way back
    rc, err = gpu.decompress_file(a_ref, back)
    assert rc == 0, err
    assert open(back, "rb").read() == big.replace(b"\r\n", b"\n")
