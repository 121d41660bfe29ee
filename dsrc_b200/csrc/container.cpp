// Host-side format helpers of the C ABI (no device work): first-chunk analysis and the .dsrc container header / footer, so that a
// C or C++ host can produce and read whole archives around dsrcgpu_encode_blocks / dsrcgpu_decode_blocks.
//   FastqParser::Analyze                      src/FastqParser.cpp:27-138
//   DsrcFileWriter::WriteFileHeader / Footer  src/DsrcFile.cpp:112-170, src/DsrcFile.h:26-47
//   DsrcFileReader::ReadFileHeader / Footer   src/DsrcFile.cpp:264-314
#include "../../include/dsrc_b200.h"
#include <string.h>

typedef uint8_t u8; typedef uint32_t u32; typedef uint64_t u64;

static u32 skip_line(const u8* m, u64 size, u64* pos)      // SkipLine (src/FastqParser.h:93-115): returns the line length
{
    u64 s = *pos, p = s;
    while (p < size && m[p] != '\n' && m[p] != '\r') ++p;
    const u32 len = (u32)(p - s);
    if (p + 1 < size && m[p] == '\r' && m[p + 1] == '\n') ++p;
    *pos = p + 1;
    return len;
}

extern "C" int dsrcgpu_analyze_first_chunk(const uint8_t* m, uint64_t size, dsrcgpu_dataset_t* ds)
{
    if (!m || !ds) return DSRCGPU_E_ARG;
    u64 pos = 0; u32 recs = 0; u8 minq = 255, maxq = 0;
    const bool estimate = ds->quality_offset == 0;
    int plus_rep = 0, cs = 0;
    while (pos < size) {
        const u64 t = pos; const u32 tl = skip_line(m, size, &pos);
        if (tl == 0 || m[t] != '@') break;
        const u64 s = pos; const u32 sl = skip_line(m, size, &pos);
        if (sl == 0) break;
        const u64 p = pos; const u32 pl = skip_line(m, size, &pos);
        if (p >= size || m[p] != '+') break;
        const u64 q = pos; const u32 ql = skip_line(m, size, &pos);
        if (estimate) { for (u32 i = 0; i < ql; ++i) { if (m[q + i] < minq) minq = m[q + i]; if (m[q + i] > maxq) maxq = m[q + i]; } }
        else if (ql == 0) break;
        const int cenc = sl > 1 && ((m[s + 1] >= '0' && m[s + 1] <= '3') || m[s + 1] == '.');
        const int prep = pl > 1;
        if (recs) { if (cs != cenc || (cs && m[s] >= '0' && m[s] <= '3') || plus_rep != prep) return DSRCGPU_E_MALFORMED; }
        else { plus_rep = prep; cs = cenc; }
        ++recs;
    }
    u32 qoff = ds->quality_offset;
    if (estimate) {
        if (maxq <= 74) { if (minq >= 33) qoff = 33; }
        else if (maxq <= 105) { if (minq >= 64) qoff = 64; else if (minq >= 59) qoff = 59; }
        if (qoff == 0) { if (minq >= 33) qoff = 33; else return DSRCGPU_E_MALFORMED; }
    }
    if (recs <= 1) return DSRCGPU_E_MALFORMED;             // reference: "Error analyzing FASTQ dataset"
    ds->quality_offset = qoff; ds->plus_repetition = (u8)plus_rep; ds->color_space = (u8)cs;
    return DSRCGPU_OK;
}

static void be32(u8* p, u32 v) { p[0] = (u8)(v >> 24); p[1] = (u8)(v >> 16); p[2] = (u8)(v >> 8); p[3] = (u8)v; }
static void be64(u8* p, u64 v) { be32(p, (u32)(v >> 32)); be32(p + 4, (u32)v); }
static u32 rd32(const u8* p) { return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | p[3]; }
static u64 rd64(const u8* p) { return ((u64)rd32(p) << 32) | rd32(p + 4); }

extern "C" uint64_t dsrcgpu_archive_footer_size(uint64_t n_blocks) { return 1 + 4 * n_blocks + 2 + 3 + 8; }

extern "C" int dsrcgpu_write_archive_header(uint8_t* out40, uint64_t n_blocks, uint64_t blocks_total_bytes)
{
    if (!out40) return DSRCGPU_E_ARG;
    out40[0] = 0xAA; out40[1] = 2; out40[2] = 0; out40[3] = 2;
    be32(out40 + 4, (u32)dsrcgpu_archive_footer_size(n_blocks)); be64(out40 + 8, 40 + blocks_total_bytes); be64(out40 + 16, 0); be64(out40 + 24, n_blocks);
    memset(out40 + 32, 0xAA, 8);
    return DSRCGPU_OK;
}

extern "C" int dsrcgpu_write_archive_footer(uint8_t* out, uint64_t out_cap, const uint32_t* block_sizes, uint64_t n_blocks,
                                            const dsrcgpu_dataset_t* ds, const dsrcgpu_settings_t* cs)
{
    if (!out || !block_sizes || !ds || !cs) return DSRCGPU_E_ARG;
    if (out_cap < dsrcgpu_archive_footer_size(n_blocks)) return DSRCGPU_E_CAPACITY;
    u64 p = 0;
    out[p++] = 0xCC;
    for (u64 i = 0; i < n_blocks; ++i) { const u32 v = block_sizes[i]; out[p++] = (u8)v; out[p++] = (u8)(v >> 8); out[p++] = (u8)(v >> 16); out[p++] = (u8)(v >> 24); }   // host-endian (LE) in the reference
    out[p++] = (u8)((ds->plus_repetition ? 1 : 0) | (ds->color_space ? 2 : 0));
    out[p++] = (u8)ds->quality_offset;
    out[p++] = (u8)((cs->lossy ? 1 : 0) | (cs->calc_crc32 ? 2 : 0));
    out[p++] = (u8)cs->dna_order; out[p++] = (u8)cs->quality_order;
    be64(out + p, cs->tag_preserve_flags);
    return DSRCGPU_OK;
}

extern "C" int dsrcgpu_read_archive_index(const uint8_t* arc, uint64_t size, uint64_t* n_blocks, uint64_t* blk_off, uint32_t* blk_len,
                                          uint64_t max_blocks, dsrcgpu_dataset_t* ds, dsrcgpu_settings_t* cs)
{
    if (!arc || !n_blocks) return DSRCGPU_E_ARG;
    if (size < 40 || arc[0] != 0xAA || arc[1] != 2) return DSRCGPU_E_MALFORMED;
    const u32 fsize = rd32(arc + 4); const u64 foff = rd64(arc + 8), n = rd64(arc + 24);
    // the header is untrusted: bound the block count by what the footer and the block area can hold BEFORE it enters any arithmetic
    // (4 * n wraps in 64 bits), every block is at least the 16-byte meta header
    if (foff < 40 || foff > size || fsize > size - foff || fsize < 14) return DSRCGPU_E_MALFORMED;
    if (n == 0 || n > ((u64)fsize - 14) / 4 || n > (foff - 40) / 16 + 1 || arc[foff] != 0xCC) return DSRCGPU_E_MALFORMED;
    *n_blocks = n;
    const u8* f = arc + foff;
    u64 p = 40;
    for (u64 i = 0; i < n; ++i) {
        const u32 v = (u32)f[1 + 4 * i] | ((u32)f[2 + 4 * i] << 8) | ((u32)f[3 + 4 * i] << 16) | ((u32)f[4 + 4 * i] << 24);
        if (v > foff - p) return DSRCGPU_E_MALFORMED;
        if (i < max_blocks) { if (blk_off) blk_off[i] = p; if (blk_len) blk_len[i] = v; }
        p += v;
    }
    const u8* t = f + 1 + 4 * n;
    if (ds) { ds->plus_repetition = t[0] & 1; ds->color_space = (t[0] >> 1) & 1; ds->quality_offset = t[1]; }
    if (cs) { cs->lossy = t[2] & 1; cs->calc_crc32 = (t[2] >> 1) & 1; cs->dna_order = t[3]; cs->quality_order = t[4]; cs->tag_preserve_flags = rd64(t + 5); }
    return DSRCGPU_OK;
}

// the same for a host that streams: it has read the 40-byte header and then the footer (at the offset / size the header names), not the
// blocks. blk_len[i] = size of block i; block i starts at 40 + sum of the sizes before it.
extern "C" int dsrcgpu_read_archive_footer(const uint8_t* header40, const uint8_t* footer, uint64_t footer_bytes, uint64_t file_size,
                                           uint64_t* n_blocks, uint32_t* blk_len, uint64_t max_blocks, dsrcgpu_dataset_t* ds, dsrcgpu_settings_t* cs)
{
    if (!header40 || !n_blocks) return DSRCGPU_E_ARG;
    if (header40[0] != 0xAA || header40[1] != 2) return DSRCGPU_E_MALFORMED;
    const u32 fsize = rd32(header40 + 4); const u64 foff = rd64(header40 + 8), n = rd64(header40 + 24);
    if (foff < 40 || foff > file_size || fsize > file_size - foff || fsize < 14) return DSRCGPU_E_MALFORMED;
    if (n == 0 || n > ((u64)fsize - 14) / 4 || n > (foff - 40) / 16 + 1) return DSRCGPU_E_MALFORMED;
    *n_blocks = n;
    if (!footer) return DSRCGPU_OK;                        // header only: the caller learns where the footer is (dsrcgpu_archive_footer_span)
    if (footer_bytes < fsize || footer[0] != 0xCC) return DSRCGPU_E_MALFORMED;
    u64 p = 40;
    for (u64 i = 0; i < n; ++i) {
        const u32 v = (u32)footer[1 + 4 * i] | ((u32)footer[2 + 4 * i] << 8) | ((u32)footer[3 + 4 * i] << 16) | ((u32)footer[4 + 4 * i] << 24);
        if (v > foff - p) return DSRCGPU_E_MALFORMED;
        if (i < max_blocks && blk_len) blk_len[i] = v;
        p += v;
    }
    const u8* t = footer + 1 + 4 * n;
    if (ds) { ds->plus_repetition = t[0] & 1; ds->color_space = (t[0] >> 1) & 1; ds->quality_offset = t[1]; }
    if (cs) { cs->lossy = t[2] & 1; cs->calc_crc32 = (t[2] >> 1) & 1; cs->dna_order = t[3]; cs->quality_order = t[4]; cs->tag_preserve_flags = rd64(t + 5); }
    return DSRCGPU_OK;
}
extern "C" int dsrcgpu_archive_footer_span(const uint8_t* header40, uint64_t* footer_offset, uint64_t* footer_bytes)
{
    if (!header40 || !footer_offset || !footer_bytes) return DSRCGPU_E_ARG;
    if (header40[0] != 0xAA || header40[1] != 2) return DSRCGPU_E_MALFORMED;
    *footer_bytes = rd32(header40 + 4); *footer_offset = rd64(header40 + 8);
    return DSRCGPU_OK;
}
