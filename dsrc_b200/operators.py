"""Host-side mirror of the reference's operator layer for the block path, over the C ABI.

    DsrcCompressorMT / DsrcDecompressorMT   src/DsrcOperator.cpp:230-394, 397-525  (`Process(InputParameters)`)
    DsrcModule.Compress / Decompress        src/DsrcModule.cpp:47-100

The reference's reader thread + worker pool + ordered writer become: cut the file into the blocks the reference's reader would
produce (dsrcgpu_cut_blocks == IFastqStreamReader::ReadNextChunk, src/FastqStream.cpp:18-98), hand the whole block queue to the
GPU (BlockCompressor.store_many), write header / blocks / footer (DsrcFileWriter, src/DsrcFile.cpp:75-170). The archive equals
`dsrc c -t1` byte for byte. With several ranks (one per GPU) the block queue is split into contiguous ranges; the only exchange
is the gather of the compressed block sizes for the single footer.

Everything here is host glue (format, sharding); the codec itself runs on the device -- there is no CPU encode/decode path.
"""
import ctypes as C
import struct
from dataclasses import dataclass

import numpy as np

from . import _lib
from .block_compressor import BlockCompressor, DsrcGpuError


@dataclass
class InputParameters:
    """src/Common.h:149-193 (the fields the lossless block path uses)"""
    dnaCompressionLevel: int = 0          # -d 0..3
    qualityCompressionLevel: int = 0      # -q 0..2
    fastqBufferSizeMB: int = 8            # -b 1..1024
    qualityOffset: int = 0                # -o, 0 = auto (src/FastqParser.cpp:27-138)
    calculateCrc32: bool = False          # -c
    block_bytes: int = 0                  # non-reference extension: chunk buffer in bytes (e.g. 256 KiB); 0 = fastqBufferSizeMB << 20
    max_inflight_blocks: int = 0

    def chunk_buffer(self):
        return self.block_bytes if self.block_bytes else self.fastqBufferSizeMB << 20


def _lines(buf, pos, size):
    """SkipLine (src/FastqParser.h:93-115): returns (start, length, next position)"""
    s = pos
    while pos < size and buf[pos] not in (10, 13):
        pos += 1
    ln = pos - s
    if pos + 1 < size and buf[pos] == 13 and buf[pos + 1] == 10:
        pos += 1
    return s, ln, pos + 1


def analyze_first_chunk(chunk, quality_offset=0):
    """FastqParser::Analyze (src/FastqParser.cpp:27-138): quality offset auto-detection, '+' repetition, colour space.
    Returns (quality_offset, plus_repetition, color_space); raises ValueError like the reference's 'Error analyzing FASTQ dataset'."""
    buf = memoryview(chunk)
    size = len(buf)
    pos = 0
    recs = 0
    minq, maxq = 255, 0
    estimate = quality_offset == 0
    plus_rep = cs = False
    while pos < size:
        t, tl, pos = _lines(buf, pos, size)
        if tl == 0 or buf[t] != 64:
            break
        s, sl, pos = _lines(buf, pos, size)
        if sl == 0:
            break
        p, pl, pos = _lines(buf, pos, size)
        if p >= size or buf[p] != 43:
            break
        q, ql, pos = _lines(buf, pos, size)
        if estimate:
            if ql:
                qa = np.frombuffer(buf[q:q + ql], dtype=np.uint8)
                minq = min(minq, int(qa.min()))
                maxq = max(maxq, int(qa.max()))
        elif ql == 0:
            break
        cenc = sl > 1 and (48 <= buf[s + 1] <= 51 or buf[s + 1] == 46)
        prep = pl > 1
        if recs:
            if cs != cenc or (cs and 48 <= buf[s] <= 51) or plus_rep != prep:
                raise ValueError("Error analyzing FASTQ dataset")
        else:
            plus_rep, cs = prep, cenc
        recs += 1
    if estimate:
        if maxq <= 74:
            if minq >= 33:
                quality_offset = 33
        elif maxq <= 105:
            if minq >= 64:
                quality_offset = 64
            elif minq >= 59:
                quality_offset = 59
        if quality_offset == 0:
            if minq >= 33:
                quality_offset = 33
            else:
                raise ValueError("Error analyzing FASTQ dataset")
    if recs <= 1:
        raise ValueError("Error analyzing FASTQ dataset")
    return quality_offset, plus_rep, cs


def cut_blocks(data, cbuf):
    L = _lib.lib()
    buf = np.frombuffer(data, dtype=np.uint8)
    n = L.dsrcgpu_cut_blocks(buf.ctypes.data_as(C.c_void_p), buf.size, cbuf, None, None, 0)
    off = np.zeros(n, dtype=np.uint64)
    ln = np.zeros(n, dtype=np.uint32)
    L.dsrcgpu_cut_blocks(buf.ctypes.data_as(C.c_void_p), buf.size, cbuf, off.ctypes.data_as(_lib.u64p), ln.ctypes.data_as(_lib.u32p), n)
    return off, ln


def tag_capacities(data, offs, lens):
    """capacity of the reference's TagStats::fields vector before every block, one compressor, file order."""
    L = _lib.lib()
    mv = memoryview(data)
    caps = np.zeros(len(offs), dtype=np.uint32)
    cap = 0
    for i, (o, l) in enumerate(zip(offs, lens)):
        caps[i] = cap
        o = int(o)
        head = bytes(mv[o:o + min(int(l), 65536)])
        ends = [x for x in (head.find(b"\n"), head.find(b"\r")) if x >= 0]
        tl = min(ends) if ends else len(head)
        cap = L.dsrcgpu_tag_capacity_after(cap, L.dsrcgpu_tag_field_count(head[:tl], tl))
    return caps


def shard_ranges(lens, world):
    """contiguous block ranges per rank, balanced by bytes"""
    total = int(np.asarray(lens, dtype=np.uint64).sum())
    csum = np.concatenate([[0], np.cumsum(np.asarray(lens, dtype=np.uint64))])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        k = int(np.searchsorted(csum, target, side="left"))
        cuts.append(min(max(k, cuts[-1]), len(lens)))
    cuts.append(len(lens))
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def write_header(footer_size, footer_offset, n_blocks):
    """DsrcFileWriter::WriteFileHeader (src/DsrcFile.cpp:112-131, src/DsrcFile.h:26-47): 40 bytes"""
    return bytes([0xAA, 2, 0, 2]) + struct.pack(">IQQQ", footer_size, footer_offset, 0, n_blocks) + b"\xAA" * 8


def write_footer(sizes, quality_offset, plus_rep, dna_order, quality_order, calc_crc32=False):
    """DsrcFileWriter::WriteFileFooter (src/DsrcFile.cpp:133-170): block sizes are host-endian (LE) u32; compFlags bit 1 = CRC32"""
    return (b"\xCC" + np.asarray(sizes, dtype="<u4").tobytes() + bytes([1 if plus_rep else 0, quality_offset, 2 if calc_crc32 else 0, dna_order, quality_order])
            + struct.pack(">Q", 0))


def read_archive_index(arc):
    """DsrcFileReader::ReadFileHeader / ReadFileFooter (src/DsrcFile.cpp:186-314) -> (offsets, sizes, dataset/settings dict)"""
    if len(arc) < 40 or arc[0] != 0xAA or arc[1] != 2:
        raise ValueError("Invalid archive header")
    footer_size, footer_off, _, n = struct.unpack(">IQQQ", arc[4:32])
    if n == 0 or footer_off + footer_size > len(arc) or footer_size < 1 + 4 * n + 13 or arc[footer_off] != 0xCC:
        raise ValueError("Invalid archive footer")
    sizes = np.frombuffer(arc[footer_off + 1:footer_off + 1 + 4 * n], dtype="<u4").astype(np.uint32)
    p = footer_off + 1 + 4 * n
    flags, qoff, cflags, dna_order, qua_order = arc[p], arc[p + 1], arc[p + 2], arc[p + 3], arc[p + 4]
    if flags & 2 or cflags & 1:
        raise ValueError("colour-space / lossy archives are outside the supported envelope")
    offs = 40 + np.concatenate([[0], np.cumsum(sizes.astype(np.uint64))[:-1]]).astype(np.uint64)
    if int(offs[-1]) + int(sizes[-1]) > footer_off:
        raise ValueError("Invalid archive footer")
    return offs, sizes, dict(quality_offset=qoff, plus_repetition=bool(flags & 1), dna_order=dna_order, quality_order=qua_order, calc_crc32=bool(cflags & 2))


class DsrcCompressorMT:
    """IDsrcOperator::Process for compression (src/DsrcOperator.cpp:230-394) on in-memory FASTQ."""

    def __init__(self, device=0, encoder=None, gather=None, rank=0, world=1):
        self.device, self.rank, self.world = device, rank, world
        self.encoder = encoder        # tests inject a stand-in for the device call; None = the CUDA path
        self.gather = gather          # list-of-arrays all-gather across ranks; None = single rank
        self.error = ""

    def _encode(self, data, offs, lens, caps, settings, args):
        if self.encoder is not None:
            return self.encoder(data, offs, lens, caps, settings)
        bc = BlockCompressor(settings["quality_offset"], settings["plus_repetition"], settings["dna_order"], settings["quality_order"],
                             max_block_bytes=max(int(lens.max()) + 64, 1 << 16), max_inflight_blocks=args.max_inflight_blocks, device=self.device,
                             calc_crc32=settings["calc_crc32"])
        try:
            blocks, _, _ = bc.store_many(data, offs, lens, caps=caps)
        finally:
            bc.close()
        return blocks

    def process(self, args, fastq):
        """returns (archive bytes on rank 0 | None, this rank's (offset, bytes) slice of the block area)"""
        offs, lens = cut_blocks(fastq, args.chunk_buffer())
        if len(offs) == 0:
            raise ValueError("Error analyzing FASTQ dataset")
        o0, l0 = int(offs[0]), int(lens[0])
        qoff, plus_rep, cs = analyze_first_chunk(memoryview(fastq)[o0:o0 + l0], args.qualityOffset)   # FastqReader::AnalyzeFirstChunk
        if cs:
            raise ValueError("colour-space FASTQ is outside the supported envelope")
        settings = dict(quality_offset=qoff, plus_repetition=plus_rep, dna_order=args.dnaCompressionLevel * 3,      # DsrcOperator.h:74-90
                        quality_order=args.qualityCompressionLevel, calc_crc32=bool(args.calculateCrc32))
        caps = tag_capacities(fastq, offs, lens)
        b0, b1 = shard_ranges(lens, self.world)[self.rank]
        mine = self._encode(fastq, offs[b0:b1], lens[b0:b1], caps[b0:b1], settings, args) if b1 > b0 else []
        my_sizes = np.array([len(b) for b in mine], dtype=np.uint32)
        all_sizes = self.gather(my_sizes) if self.gather is not None else [my_sizes]
        sizes = np.concatenate(all_sizes) if len(all_sizes) else my_sizes
        assert len(sizes) == len(offs)
        my_off = 40 + int(np.concatenate(all_sizes[:self.rank]).astype(np.uint64).sum()) if self.rank else 40
        payload = b"".join(mine)
        archive = None
        if self.world == 1:
            footer = write_footer(sizes, qoff, plus_rep, settings["dna_order"], settings["quality_order"], settings["calc_crc32"])
            archive = write_header(len(footer), 40 + len(payload), len(sizes)) + payload + footer
        elif self.rank == 0:
            footer = write_footer(sizes, qoff, plus_rep, settings["dna_order"], settings["quality_order"], settings["calc_crc32"])
            total = int(sizes.astype(np.uint64).sum())
            archive = (write_header(len(footer), 40 + total, len(sizes)), footer)      # header + footer; ranks pwrite their slices
        return archive, (my_off, payload)


def dist_gather_sizes(my_sizes, device):
    """the one exchange of the multi-rank path (src/DsrcFile.cpp:142: one footer with every block's size): all-gather of the ranks'
    uint32 block sizes over the default torch.distributed group -- NCCL when `device` is the rank's GPU, gloo when it is "cpu"."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    n = torch.tensor([len(my_sizes)], dtype=torch.int64, device=device)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    mx = max(1, max(int(x) for x in ns))
    pad = torch.zeros(mx, dtype=torch.int64, device=device)
    if len(my_sizes):
        pad[:len(my_sizes)] = torch.from_numpy(np.asarray(my_sizes).astype(np.int64)).to(device)
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return [o[:int(k)].cpu().numpy().astype(np.uint32) for o, k in zip(outs, ns)]


def write_archive_sharded(path, archive, my_slice, rank, barrier=None):
    """DsrcFileWriter as N ranks see it (src/DsrcFile.cpp:75-170; the ordered writer of src/DsrcIo.cpp:19-89 becomes offsets):
    rank 0 writes the 40-byte header and the footer, EVERY rank pwrites its compressed blocks at 40 + (bytes of the ranks before it).
    `archive` / `my_slice` are what DsrcCompressorMT.process returned on this rank."""
    import os
    my_off, payload = my_slice
    if rank == 0:
        fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
        if isinstance(archive, tuple):
            header, footer = archive
            footer_off = struct.unpack(">Q", header[8:16])[0]
            os.pwrite(fd, header, 0)
            os.pwrite(fd, footer, footer_off)
            os.pwrite(fd, payload, my_off)
        else:
            os.pwrite(fd, archive, 0)
        os.close(fd)
    if barrier is not None:
        barrier()                      # the file exists before the other ranks open it
    if rank != 0:
        fd = os.open(path, os.O_RDWR)
        os.pwrite(fd, payload, my_off)
        os.close(fd)
    if barrier is not None:
        barrier()


class DsrcDecompressorMT:
    """IDsrcOperator::Process for decompression (src/DsrcOperator.cpp:397-525) on an in-memory archive."""

    def __init__(self, device=0, rank=0, world=1):
        self.device, self.rank, self.world = device, rank, world

    def process(self, archive, max_inflight_blocks=0):
        """single rank: the whole FASTQ. With several ranks (BASELINE configs[4]) every rank reads the footer, takes a contiguous block
        range balanced by compressed bytes and returns (offset of its part in the FASTQ file, bytes): the output offsets need nothing
        but every block's chunkSize (bytes 12..15 of its header, src/BlockCompressor.cpp:302-308), so there is no exchange at all."""
        offs, sizes, st = read_archive_index(archive)
        mv = memoryview(archive)
        out_sizes = np.array([struct.unpack(">I", mv[int(o) + 12:int(o) + 16])[0] + 1 for o in offs], dtype=np.uint64)
        b0, b1 = shard_ranges(sizes, self.world)[self.rank]
        my_off = int(out_sizes[:b0].sum())
        if b1 <= b0:
            return b"" if self.world == 1 else (my_off, b"")
        bc = BlockCompressor(st["quality_offset"], st["plus_repetition"], st["dna_order"], st["quality_order"],
                             max_block_bytes=1 << 20, max_inflight_blocks=max_inflight_blocks, device=self.device, calc_crc32=st["calc_crc32"])
        try:
            parts = bc.read_many(archive, offs[b0:b1], sizes[b0:b1], int(out_sizes[b0:b1].sum()) + 64)
        finally:
            bc.close()
        data = b"".join(parts)
        return data if self.world == 1 else (my_off, data)


class DsrcModule:
    """wrap::DsrcModule (include/dsrc/DsrcModule.h:22-40): file-to-file Compress / Decompress with Configurable-style settings."""

    def __init__(self, dna_level=0, quality_level=0, buffer_mb=8, quality_offset=0, device=0):
        self.args = InputParameters(dna_level, quality_level, buffer_mb, quality_offset)
        self.device = device

    def compress(self, in_path, out_path):
        data = open(in_path, "rb").read()
        arc, _ = DsrcCompressorMT(self.device).process(self.args, data)
        open(out_path, "wb").write(arc)

    def decompress(self, in_path, out_path):
        arc = open(in_path, "rb").read()
        open(out_path, "wb").write(DsrcDecompressorMT(self.device).process(arc))
