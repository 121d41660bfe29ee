"""Host logic of the operator layer (no device work): first-chunk analysis, container header/footer, block sharding, and the
two-rank path over gloo with the oracle standing in for the device encoder."""
import os
import socket
import sys

import numpy as np
import pytest

import refbind
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_encoder(data, offs, lens, caps, st):
    """test stand-in for the device call: one oracle instance per block, its field-vector capacity forced through a warm-up title"""
    out = []
    for o, l, c in zip(offs, lens, caps):
        ora = refbind.Oracle(st["quality_offset"], int(st["plus_repetition"]), st["dna_order"], st["quality_order"], crc=st.get("calc_crc32", False))
        chunk = bytes(memoryview(data)[int(o):int(o) + int(l)])
        if c:        # bring TagStats::fields to capacity c: a block whose first title has c fields (2 records so Analyze-free Store works)
            title = b"@" + b" ".join([b"x"] * int(c))
            ora.store(title + b"\nA\n+\nI\n" + title + b"\nA\n+\nI")
        out.append(ora.store(chunk)[0])
    return out


def test_analyze_first_chunk_matches_oracle():
    from dsrc_b200.operators import analyze_first_chunk
    o = refbind.Oracle()
    for data in [synth.illumina(50, seed=1), synth.illumina(40, seed=2, plus_rep=True), synth.ion454(30, seed=3), synth.illumina(30, seed=4, crlf=True)]:
        chunk = data[:-1]
        ok, q, pr, cs = o.analyze(chunk)
        assert ok == 1
        assert analyze_first_chunk(chunk) == (q, pr, cs)
    # Illumina 1.3+ offset 64
    hi = synth.illumina(30, seed=5).replace(b"\n+\n", b"\n+\n")
    recs = hi.split(b"\n")
    for i in range(3, len(recs), 4):
        recs[i] = bytes(min(126, c + 31) for c in recs[i])
    hi = b"\n".join(recs)
    ok, q, pr, cs = o.analyze(hi[:-1])
    assert ok == 1 and q == 64
    assert analyze_first_chunk(hi[:-1]) == (64, pr, cs)


def test_header_footer_match_oracle_archive():
    from dsrc_b200 import operators as op
    big = synth.illumina(8000, seed=7)
    arc = refbind.Oracle().compress(big, 2, 2, 1 << 20, 0)
    offs, sizes, st = op.read_archive_index(arc)
    assert st == dict(quality_offset=33, plus_repetition=False, dna_order=6, quality_order=2, calc_crc32=False)
    footer = op.write_footer(sizes, 33, False, 6, 2)
    total = int(sizes.astype(np.uint64).sum())
    assert op.write_header(len(footer), 40 + total, len(sizes)) == arc[:40]
    assert footer == arc[40 + total:]
    assert int(offs[0]) == 40
    arc_c = refbind.Oracle().compress(big, 2, 2, 1 << 20, 0, crc=True)
    offs, sizes, st = op.read_archive_index(arc_c)
    assert st["calc_crc32"] is True
    total = int(sizes.astype(np.uint64).sum())
    assert op.write_footer(sizes, 33, False, 6, 2, True) == arc_c[40 + total:]


def test_shard_ranges_cover_and_balance():
    from dsrc_b200.operators import shard_ranges
    lens = np.array([100] * 37 + [5000] + [100] * 10, dtype=np.uint32)
    for world in (1, 2, 3, 8, 64):
        rs = shard_ranges(lens, world)
        assert rs[0][0] == 0 and rs[-1][1] == len(lens)
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert all(a <= b for a, b in rs)


def test_single_rank_archive_with_standin_encoder_matches_oracle():
    from dsrc_b200.operators import DsrcCompressorMT, InputParameters
    big = synth.illumina(6000, seed=21, small_field=True)
    args = InputParameters(2, 2, 1, 0, block_bytes=1 << 18)
    arc, _ = DsrcCompressorMT(encoder=_oracle_encoder).process(args, big)
    assert arc == refbind.Oracle().compress(big, 2, 2, 1 << 18, 0)


def _rank_main(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from dsrc_b200.operators import DsrcCompressorMT, InputParameters
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)

    def gather(sizes):
        out = [None] * world
        dist.all_gather_object(out, sizes)
        return out

    big = synth.illumina(6000, seed=21, small_field=True)
    args = InputParameters(2, 2, 1, 0, block_bytes=1 << 18)
    arc, (off, payload) = DsrcCompressorMT(encoder=_oracle_encoder, gather=gather, rank=rank, world=world).process(args, big)
    q.put((rank, arc, off, payload))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_over_gloo_assemble_the_reference_archive():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
    header, footer = res[0][1]
    assert res[1][1] is None
    total = sum(len(r[3]) for r in res)
    out = bytearray(40 + total + len(footer))
    out[:40] = header
    for _, _, off, payload in res:           # every rank pwrites its slice at 40 + exclusive scan of the gathered sizes
        out[off:off + len(payload)] = payload
    out[40 + total:] = footer
    big = synth.illumina(6000, seed=21, small_field=True)
    assert bytes(out) == refbind.Oracle().compress(big, 2, 2, 1 << 18, 0)
    assert res[1][2] > 40 and len(res[1][3]) > 0


def test_c_abi_format_helpers_match_oracle():
    """dsrcgpu_analyze_first_chunk / write_archive_header / write_archive_footer / read_archive_index (pure host C)"""
    import ctypes as C
    from dsrc_b200 import _lib
    L = _lib.lib()
    o = refbind.Oracle()
    for data in [synth.illumina(50, seed=1), synth.illumina(40, seed=2, plus_rep=True), synth.ion454(30, seed=3)]:
        ok, q, pr, cs = o.analyze(data[:-1])
        ds = _lib.Dataset(0, 0, 0)
        assert L.dsrcgpu_analyze_first_chunk(data[:-1], len(data) - 1, C.byref(ds)) == 0
        assert (ds.quality_offset, bool(ds.plus_repetition), bool(ds.color_space)) == (q, pr, cs)
    assert L.dsrcgpu_analyze_first_chunk(b"@a\nAC\n+\nII", 10, C.byref(_lib.Dataset(0, 0, 0))) != 0      # a single record: reference refuses
    big = synth.illumina(8000, seed=7)
    for crc in (False, True):
        arc = refbind.Oracle().compress(big, 2, 2, 1 << 20, 0, crc=crc)
        n = C.c_uint64()
        off = np.zeros(16, dtype=np.uint64)
        ln = np.zeros(16, dtype=np.uint32)
        ds = _lib.Dataset(0, 0, 0)
        cs = _lib.Settings(0, 0, 0, 0, 0)
        assert L.dsrcgpu_read_archive_index(arc, len(arc), C.byref(n), off.ctypes.data_as(_lib.u64p), ln.ctypes.data_as(_lib.u32p), 16, C.byref(ds), C.byref(cs)) == 0
        nb = n.value
        assert (ds.quality_offset, cs.dna_order, cs.quality_order, bool(cs.calc_crc32)) == (33, 6, 2, crc)
        total = int(ln[:nb].astype(np.uint64).sum())
        hdr = (C.c_uint8 * 40)()
        assert L.dsrcgpu_write_archive_header(hdr, nb, total) == 0
        assert bytes(hdr) == arc[:40]
        fs = L.dsrcgpu_archive_footer_size(nb)
        foot = (C.c_uint8 * fs)()
        assert L.dsrcgpu_write_archive_footer(foot, fs, ln.ctypes.data_as(_lib.u32p), nb, C.byref(ds), C.byref(cs)) == 0
        assert bytes(foot) == arc[40 + total:]
        assert int(off[0]) == 40 and int(off[nb - 1]) + int(ln[nb - 1]) == 40 + total
    assert L.dsrcgpu_read_archive_index(b"\0" * 64, 64, C.byref(n), None, None, 0, None, None) != 0


def test_cpp_operator_host_paths_need_no_device(tmp_path):
    """host/DsrcOperatorGpu.h through the shim harness: the failures the reference reports before any block is coded (missing input,
    not FASTQ, not an archive) come back through IDsrcOperator::GetError with the reference's texts -- no device involved."""
    import ctypes as C
    import os
    import refbind
    if not refbind.shim_available():
        import pytest
        pytest.skip("oracle/_ref/libdsrcshim.so not built")
    L = C.CDLL(os.path.join(refbind.ROOT, "oracle", "_ref", "libdsrcshim.so"))
    for sym in ("shim_bc_create", "shim_bc_store", "shim_bc_read", "shim_compress_file", "shim_decompress_file", "shim_module_roundtrip"):
        assert hasattr(L, sym)
    L.shim_compress_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_char_p, C.c_int]
    L.shim_decompress_file.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    err = C.create_string_buffer(512)
    out = str(tmp_path / "o").encode()
    assert L.shim_compress_file(str(tmp_path / "missing").encode(), out, 2, 2, 1, 0, 0, err, 512) != 0
    assert b"Cannot open file to read:" in err.value
    bad = tmp_path / "bad.fastq"
    bad.write_bytes(b"this is not FASTQ\n" * 10)
    assert L.shim_compress_file(str(bad).encode(), out, 2, 2, 1, 0, 0, err, 512) != 0
    assert b"Error analyzing FASTQ dataset" in err.value
    junk = tmp_path / "junk.dsrc"
    junk.write_bytes(b"\0" * 100)
    assert L.shim_decompress_file(str(junk).encode(), out, err, 512) != 0
    assert b"Invalid archive" in err.value


def test_product_cutter_matches_oracle_including_exact_window_eof():
    """dsrcgpu_cut_blocks (pure host) == the oracle's restatement of IFastqStreamReader::ReadNextChunk, which
    test_oracle_vs_ref.py pins to `dsrc c -t1` -- including a chunk window that ends exactly at EOF (src/FastqStream.cpp:66-69)"""
    from dsrc_b200.operators import cut_blocks
    cbuf = 1 << 20
    o = refbind.Oracle()
    second_start = o.cut(synth.exact_size(3 * cbuf), cbuf)[1][0]
    files = [synth.exact_size(cbuf), synth.exact_size(cbuf, crlf=True), synth.exact_size(second_start + cbuf), synth.exact_size(cbuf + 1),
             synth.exact_size(cbuf - 1), synth.illumina(3000, seed=2, regime="full"), b"@r\nA\n+\nI\n", b"@r\r\nA\r\n+\r\nI\r\n"]
    for data in files:
        for cb in (cbuf, 1 << 18, 100000):
            off, ln = cut_blocks(data, cb)
            assert [(int(a), int(b)) for a, b in zip(off, ln)] == [(int(a), int(b)) for a, b in o.cut(data, cb)], (len(data), cb)


def test_windowed_cutter_matches_whole_file_cutter():
    """dsrcgpu_cut_blocks_window (what the streaming C++ operator uses, host/DsrcOperatorGpu.h): cutting a file window by window -- dropping
    every window's last block and restarting there, carrying the reader's CRLF state -- gives the whole-file block queue"""
    import ctypes as C
    from dsrc_b200 import _lib
    L = _lib.lib()
    for crlf in (False, True):
        big = synth.illumina(12000, seed=61, small_field=True, crlf=crlf) + synth.exact_size(3 << 20, seed=62, crlf=crlf)
        buf = np.frombuffer(big, dtype=np.uint8)
        cb = 1 << 20
        whole = [(int(a), int(b)) for a, b in refbind.Oracle().cut(big, cb)]
        for win in (4 << 20, 5 << 20, (3 << 20) + 12345):
            pos, got = 0, []
            st = np.zeros(1, dtype=np.uint32)
            while pos < len(big):
                w = min(win, len(big) - pos)
                last = pos + w == len(big)
                sub = buf[pos:pos + w]
                st2 = st.copy()
                k = L.dsrcgpu_cut_blocks_window(sub.ctypes.data_as(C.c_void_p), w, cb, None, None, 0, st2.ctypes.data_as(_lib.u32p))
                o = np.zeros(k, dtype=np.uint64)
                ln = np.zeros(k, dtype=np.uint32)
                L.dsrcgpu_cut_blocks_window(sub.ctypes.data_as(C.c_void_p), w, cb, o.ctypes.data_as(_lib.u64p), ln.ctypes.data_as(_lib.u32p), k, st.ctypes.data_as(_lib.u32p))
                take = k if last else k - 1
                assert take > 0
                got += [(int(a) + pos, int(b)) for a, b in zip(o[:take], ln[:take])]
                pos = len(big) if last else pos + int(o[take])
            assert got == whole, (crlf, win)
