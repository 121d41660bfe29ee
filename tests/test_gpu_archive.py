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


def _gpu_rank_main(rank, world, port, path, back, q):
    """one rank of the multi-GPU operator path with the REAL device encoder/decoder: rank r uses GPU r when the box has that many,
    else every rank shares GPU 0 (the exchange then runs over gloo -- NCCL refuses two ranks on one device; bench.py uses NCCL)"""
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from dsrc_b200 import operators as op
    ndev = torch.cuda.device_count()
    dev = rank if ndev >= world else 0
    nccl = ndev >= world
    if nccl:
        torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if nccl else "gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    gdev = torch.device("cuda", dev) if nccl else "cpu"
    big = synth.illumina(9000, seed=23, regime="full", small_field=True)
    args = op.InputParameters(2, 2, 1, 0, block_bytes=1 << 18)
    arc, sl = op.DsrcCompressorMT(device=dev, gather=lambda s: op.dist_gather_sizes(s, gdev), rank=rank, world=world).process(args, big)
    op.write_archive_sharded(path, arc, sl, rank, barrier=dist.barrier)
    archive = open(path, "rb").read()
    off, part = op.DsrcDecompressorMT(device=dev, rank=rank, world=world).process(archive)
    fd = os.open(back, os.O_RDWR | os.O_CREAT, 0o644)
    os.pwrite(fd, part, off)
    os.close(fd)
    dist.barrier()
    q.put((rank, off, len(part), len(sl[1])))
    dist.destroy_process_group()


def test_two_ranks_write_and_read_one_archive(tmp_path):
    """BASELINE configs[3]/[4] in miniature, with the device codec on every rank: two ranks encode contiguous block ranges, gather the
    block sizes, rank 0 writes header + footer and each rank pwrites its slice (src/DsrcFile.cpp:112-170); the file equals the 1-rank
    archive (== `dsrc c -t1`) byte for byte; then both ranks decode their share of it and pwrite it into one FASTQ == the input."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    path, back = str(tmp_path / "two.dsrc"), str(tmp_path / "back.fq")
    ps = [ctx.Process(target=_gpu_rank_main, args=(r, 2, port, path, back, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in ps:
        p.join(timeout=120)
    big = synth.illumina(9000, seed=23, regime="full", small_field=True)
    assert all(r[3] > 0 for r in res) and res[1][1] > 0          # both ranks coded blocks, the second one's FASTQ part starts past 0
    arc = open(path, "rb").read()
    assert hashlib.sha256(arc).hexdigest() == hashlib.sha256(refbind.Oracle().compress(big, 2, 2, 1 << 18, 0)).hexdigest()
    assert open(back, "rb").read() == big
