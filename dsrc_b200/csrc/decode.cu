// Decode kernels: BlockCompressor::Read (src/BlockCompressor.cpp:262-356, 491-570) and what it calls.
//
// A compressed block is ONE sequential stream -- meta | tags | quality | dna with no stored sub-stream sizes -- whose every
// part is bit-serial (Huffman codes) or an adaptive range-coder chain, so a block decodes as a chain of dependent steps:
//   k_dec_probe     meta (ReadMetaData :300-356): record count and chunk size, so the host can lay the batch out exactly
//   k_dec_tags      one thread per block: TagTokenizerDecoder / TagRawDecoder (src/TagModeler.cpp:887-1213, 1288-1347) +
//                   the read-length bits of ReadTags (BlockCompressor.cpp:503-570) -> titles, lengths, where quality starts
//   k_dec_quality   one thread per block: QualityOrder (RangeDecoder src/RangeCoder.h:98-134 + TSymbolCoderRC::DecodeSymbol
//                   src/SymbolCoderRC.h:50-63), positional / truncated / RLE Huffman decoders -> processed quality bytes
//   k_dec_dna       one thread per block: order-k range decoder, 2-bit, Huffman -> retained base indices
//   k_dec_assemble  one CTA per block: LosslessRecordsProcessor::ProcessBackward (src/RecordsProcessor.cpp:269-295) and the
//                   FASTQ layout of ReadTags, parallel over records with coalesced stores
// Throughput comes from thousands of blocks in flight. The adaptive rows of a chain live in a per-chain arena in HBM: a direct
// table when it fits the arena, else an open-addressing hash of (context -> row) slots; a chain that fills its hash reports
// ST_RETRY and the host decodes that block again with a full-size table.
#include "common.cuh"
#include "kernels.h"

#define ST_RETRY 4u
#define DEC_CTA 32
#ifndef DEC_DNA_PF
#define DEC_DNA_PF 2                 // rows prefetched ahead of a DNA chain: 1 = the next base's 4 rows (L1), 2 = also the 16 rows two bases ahead (L2)
#endif

struct BitR {                     // BitMemoryReader (src/BitMemory.h:28-212): bytes, MSB-first bits through an 8-bit window
    const u8* p; u32 size, pos, cur, ncur; bool ovr;
    __device__ void init(const u8* p_, u32 size_, u32 pos_) { p = p_; size = size_; pos = pos_; cur = 0; ncur = 0; ovr = false; }
    __device__ u32 byte() { if (pos >= size) { ovr = true; ++pos; return 0; } return p[pos++]; }
    __device__ u32 bit() { if (ncur == 0) { cur = byte(); ncur = 8; } return (cur >> (--ncur)) & 1u; }
    __device__ u32 bits(u32 n) { u32 v = 0; for (u32 i = 0; i < n; ++i) v = (v << 1) | bit(); return v; }
    __device__ void flush() { ncur = 0; }
    __device__ u32 be32() { u32 v = byte(); v = (v << 8) | byte(); v = (v << 8) | byte(); return (v << 8) | byte(); }
};

// Huffman decode trees (HuffmanEncoder::LoadTree, src/huffman.cpp:225-260) in a per-block node pool.
// node.y == 0xFFFFFFFF: leaf with id node.x; else children node.x (bit 0) / node.y (bit 1).
struct NodePool {
    uint2* nodes; u32 n, cap; bool ovf;
    __device__ u32 load(BitR& r)
    {
        r.flush();
        (void)r.be32(); (void)r.be32();
        const u32 ns = r.be32(); (void)r.byte();
        const u32 bpi = dsrc_ilog2(ns) + ((ns & (ns - 1)) ? 1u : 0u);
        const u32 root = n;
        u32 stk[600]; u32 sp = 0;                       // incomplete internal nodes on the current path
        for (;;) {
            if (n >= cap || r.ovr) { ovf = true; break; }
            const u32 id = n++;
            if (sp) {
                uint2& par = nodes[stk[sp - 1]];
                if (par.x == 0xFFFFFFFEu) par.x = id; else { par.y = id; --sp; }
            }
            if (r.bit()) nodes[id] = make_uint2(r.bits(bpi), 0xFFFFFFFFu);
            else {
                nodes[id] = make_uint2(0xFFFFFFFEu, 0xFFFFFFFEu);
                if (sp >= 600) { ovf = true; break; }
                stk[sp++] = id;
            }
            if (sp == 0) break;
        }
        r.flush();
        return root;
    }
    __device__ u32 get(u32 root, BitR& r) const
    {
        u32 id = root;
        uint2 nd = nodes[id];
        while (nd.y != 0xFFFFFFFFu && !r.ovr) {
            id = r.bit() ? nd.y : nd.x;
            if (id >= n) return 0;
            nd = nodes[id];
        }
        return nd.x;
    }
    __device__ u32* words(u32 count)                     // plain u32 storage carved from the pool
    {
        const u32 need = (count + 1) / 2;
        if (n + need > cap) { ovf = true; return nullptr; }
        u32* w = (u32*)(nodes + n); n += need; return w;
    }
};

__global__ void k_dec_probe(Workspace ws)
{
    const u32 blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= ws.n_blocks) return;
    const BlockDesc& d = ws.desc[blk];
    BlockState& st = ws.state[blk];
    BitR r; r.init(ws.in + d.in_off, d.in_len, 0);
    const u32 n = r.be32(), max_len = r.be32(), flags = r.be32(), chunk = r.be32();
    const u32 min_len = (flags & 2u) ? r.be32() : max_len;
    if (ws.calc_crc) { st.crc_expected[0] = r.be32(); st.crc_expected[1] = r.be32(); st.crc_expected[2] = r.be32(); }   // ReadMetaData :340-355
    st.status = ST_OK;
    if (r.ovr || n == 0 || flags >= 256 || (flags & 1u) || max_len > 65535 || min_len > max_len || chunk >= 0xFFFFFFF0u || n > chunk) st.status = ST_MALFORMED;
    st.n_rec = n; st.max_len = max_len; st.min_len = min_len; st.flags = flags; st.chunk_size = chunk + 1;
    st.stream_size[0] = r.pos;                            // where the tag stream starts
    ws.probe[blk].n_lines = st.status == ST_OK ? n : 0;
    ws.probe[blk].n_fields = st.status == ST_OK ? chunk + 1 : 0;
}

// utils.h:52 to_string
__device__ __forceinline__ u32 dec_num_to_str(u8* s, u32 v)
{
    u8 tmp[10]; u32 n = 0;
    do { tmp[n++] = (u8)('0' + v % 10); v /= 10; } while (v);
    for (u32 i = 0; i < n; ++i) s[i] = tmp[n - 1 - i];
    return n;
}

struct DField {
    u32 len, max_len, min_len, bits_value, bits_num, bits_len, data_pos, ham_pos, hg, hl, rle_len;
    i32 min_value, min_delta, rle_sym, prev;
    u8 sep, is_constant, is_numeric, is_len_constant, scheme, var_stat;
};

__global__ void __launch_bounds__(DEC_CTA) k_dec_tags(Workspace ws, uint2* pool_base, u32 pool_stride)
{
    const u32 blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= ws.n_blocks) return;
    const BlockDesc& d = ws.desc[blk];
    BlockState& st = ws.state[blk];
    if (st.status != ST_OK) return;
    const u8* in = ws.in + d.in_off;
    BitR r; r.init(in, d.in_len, st.stream_size[0]);
    NodePool np; np.nodes = pool_base + (u64)blk * pool_stride; np.n = 0; np.cap = pool_stride; np.ovf = false;
    const RecArrays& R = ws.rec;
    const u32 n = st.n_rec, rb = d.rec_base, flags = st.flags, max_len = st.max_len, min_len = st.min_len;
    const u32 len_bits = dsrc_bit_length((u64)(max_len - min_len));
    u8* tb = ws.streams + d.stream_base;                  // title arena of the block
    const u32 tcap = d.stream_cap[1];
    u32 tpos = 0, status = ST_OK;
    u64 out_bytes = 0; u32 q_total = 0;
    DField fl[TAG_MAX_FIELDS];
    u32 nf = 0, min_title = 0, max_title = 0, tl_bits = 0, raw_root = 0;
    u8 symbols[128];
    if (n > d.rec_cap) { st.status = ST_OVERFLOW; return; }

    if (!(flags & 4u)) {            // TagTokenizerDecoder::ReadFields (TagModeler.cpp:893-1003)
        nf = r.byte();
        if (nf > TAG_MAX_FIELDS) { st.status = ST_UNSUPPORTED; return; }
        for (u32 i = 0; i < nf && !r.ovr && status == ST_OK; ++i) {
            DField& f = fl[i];
            f.hg = 0xFFFFFFFFu; f.hl = 0xFFFFFFFFu; f.rle_len = 0; f.rle_sym = 0; f.prev = 0; f.var_stat = 0; f.is_numeric = 0; f.is_len_constant = 0;
            f.sep = (u8)r.byte(); f.is_constant = r.byte() != 0;
            if (f.is_constant) { f.len = r.be32(); if (f.len >= 4096 || (u64)r.pos + f.len > d.in_len) { status = ST_MALFORMED; break; } f.data_pos = r.pos; r.pos += f.len; continue; }
            f.is_numeric = r.byte() != 0;
            if (f.is_numeric) {
                f.scheme = (u8)r.byte(); f.min_value = (i32)r.be32();
                const i32 maxv = (i32)r.be32();
                f.bits_value = dsrc_bit_length((u64)(i64)(i32)((u32)maxv - (u32)f.min_value)); f.bits_num = f.bits_value;
                if (f.scheme >= 3 && f.scheme <= 5) {
                    f.min_delta = (i32)r.be32();
                    const i32 maxd = (i32)r.be32();
                    f.bits_num = dsrc_bit_length((u64)(i64)(i32)((u32)maxd - (u32)f.min_delta));
                } else if (f.scheme != 1 && f.scheme != 2) { status = ST_MALFORMED; break; }
                if (f.bits_value > 32 || f.bits_num > 32) { status = ST_UNSUPPORTED; break; }
                if (f.scheme == 3 || f.scheme == 1) { f.var_stat = (u8)r.byte(); if (f.var_stat) f.hg = np.load(r); }
                continue;
            }
            f.is_len_constant = r.byte() != 0;
            f.len = r.be32(); f.max_len = r.be32(); f.min_len = r.be32();
            if (f.len >= 4096 || f.max_len >= 4096 || f.min_len > f.max_len) { status = ST_MALFORMED; break; }
            f.bits_len = dsrc_bit_length((u64)(f.max_len - f.min_len));
            if ((u64)r.pos + f.len + (f.len + 7) / 8 > d.in_len) { status = ST_MALFORMED; break; }   // a corrupt length must not send the copies below past the block
            f.data_pos = r.pos; r.pos += f.len;
            f.ham_pos = r.pos; r.pos += (f.len + 7) / 8;     // len mask bits, MSB first, then a flush
            r.flush();
            const u32 lim = f.max_len < TAG_STAT_LEN ? f.max_len : TAG_STAT_LEN;
            u32* roots = np.words(TAG_STAT_LEN + 1);
            if (!roots) { status = ST_RETRY; break; }
            f.hl = (u32)(roots - (u32*)np.nodes);
            for (u32 j = 0; j <= TAG_STAT_LEN; ++j) roots[j] = 0xFFFFFFFFu;
            for (u32 j = 0; j < lim; ++j) {
                const bool ham = j < f.len && ((in[f.ham_pos + (j >> 3)] >> (7 - (j & 7))) & 1u);
                if (!ham) roots[j] = np.load(r);
            }
            if (f.max_len >= TAG_STAT_LEN) roots[TAG_STAT_LEN] = np.load(r);
        }
    } else {                        // TagRawDecoder::StartDecoding (TagModeler.cpp:1288-1312)
        min_title = r.be32(); max_title = r.be32(); tl_bits = dsrc_bit_length((u64)(max_title - min_title));
        u32 nsym = 0;
        for (u32 i = 0; i < 128; ++i) symbols[i] = 255;
        for (u32 i = 0; i < 128; ++i) if (r.bit()) symbols[nsym++] = (u8)i;
        raw_root = np.load(r);
    }
    if (np.ovf && (status == ST_OK || status == ST_MALFORMED)) status = r.ovr ? ST_MALFORMED : ST_RETRY;
    if (r.ovr && status == ST_OK) status = ST_MALFORMED;

    const u32* pw = (const u32*)np.nodes;
    for (u32 k = 0; k < n && status == ST_OK && !r.ovr; ++k) {
        u8* t = tb + tpos;
        u32 tl = 0;
        if (!(flags & 4u)) {        // DecodeNextFields :1006-1064, ReadNumericField :1066-1166
            for (u32 j = 0; j < nf; ++j) {
                DField& f = fl[j];
                if (f.is_constant) {
                    if (tpos + tl + f.len + 1 > tcap) { status = ST_OVERFLOW; break; }
                    for (u32 x = 0; x < f.len; ++x) t[tl + x] = in[f.data_pos + x];
                    tl += f.len; t[tl++] = f.sep; continue;
                }
                if (f.is_numeric) {
                    u32 v = 0;
                    if (k == 0) { v = r.bits(f.bits_value); if (f.scheme == 2) { f.rle_len = r.bits(8); f.rle_sym = (i32)v; } v += (u32)f.min_value; }
                    else switch (f.scheme) {
                        case 5: v = (u32)f.prev + (u32)f.min_delta; break;
                        case 4: case 2:
                            if ((f.scheme == 4 && k == 1) || f.rle_len == 0) { v = r.bits(f.bits_num); f.rle_sym = (i32)v; f.rle_len = r.bits(8); }
                            else { f.rle_len--; v = (u32)f.rle_sym; }
                            v += f.scheme == 4 ? (u32)f.prev + (u32)f.min_delta : (u32)f.min_value;
                            break;
                        default:
                            v = f.hg != 0xFFFFFFFFu ? np.get(f.hg, r) : r.bits(f.bits_num);
                            v += f.scheme == 3 ? (u32)f.prev + (u32)f.min_delta : (u32)f.min_value;
                    }
                    if (tpos + tl + 12 > tcap) { status = ST_OVERFLOW; break; }
                    tl += dec_num_to_str(t + tl, v); f.prev = (i32)v; t[tl++] = f.sep; continue;
                }
                const u32 fl_len = f.is_len_constant ? f.len : r.bits(f.bits_len) + f.min_len;
                if (fl_len > f.max_len) { status = ST_MALFORMED; break; }
                if (tpos + tl + fl_len + 1 > tcap) { status = ST_OVERFLOW; break; }
                const u32* roots = pw + f.hl;
                for (u32 x = 0; x < fl_len; ++x) {
                    if (x < f.len && ((in[f.ham_pos + (x >> 3)] >> (7 - (x & 7))) & 1u)) t[tl++] = in[f.data_pos + x];
                    else {
                        const u32 root = roots[x < TAG_STAT_LEN ? x : TAG_STAT_LEN];
                        if (root == 0xFFFFFFFFu) { status = ST_MALFORMED; break; }
                        t[tl++] = (u8)np.get(root, r);
                    }
                }
                if (status != ST_OK) break;
                t[tl++] = f.sep;
            }
            if (status != ST_OK) break;
            if (tl == 0) { status = ST_MALFORMED; break; }
            tl--;
        } else {                    // TagRawDecoder::DecodeNextFields :1314-1333
            tl = tl_bits > 0 ? r.bits(tl_bits) + min_title : max_title;
            if (tl > 65535 || tpos + tl + 1 > tcap) { status = ST_OVERFLOW; break; }
            for (u32 i = 0; i < tl; ++i) t[i] = symbols[np.get(raw_root, r) & 127];
        }
        if (tl > 65535) { status = ST_MALFORMED; break; }
        const u32 ql = len_bits > 0 ? r.bits(len_bits) + min_len : max_len;
        if (ql > 65535) { status = ST_MALFORMED; break; }
        const u32 g = rb + k;
        R.title_off[g] = tpos; R.title_len[g] = (u16)tl; R.qua_len[g] = (u16)ql; R.qcat_off[g] = q_total;
        tpos += tl; q_total += ql;
        out_bytes += (u64)tl + 1 + ql + 2 + (ws.plus_rep ? tl - 1 : 0) + 1 + ql + 1;
    }
    r.flush();
    if (status == ST_OK && r.ovr) status = ST_MALFORMED;
    if (status == ST_OK && out_bytes != st.chunk_size) status = ST_MALFORMED;
    if (status == ST_OK && q_total > d.sym_cap) status = ST_OVERFLOW;
    st.status = status;
    st.q_total = q_total;
    st.stream_size[1] = r.pos;                            // where the quality stream starts
}

// ---- adaptive rows of a decoding chain ----
struct RowStore {
    u8* base; u32 slot_bytes, row_off, cap, used, limit; bool direct, fail;
    __device__ void setup(u8* arena, u64 arena_bytes, u32 N, u32 key_bits)
    {
        base = arena; used = 0; fail = false;
        const u64 table = ((u64)2 * N) << key_bits;
        if (table <= arena_bytes) { direct = true; slot_bytes = 2 * N; row_off = 0; cap = 0; limit = 0; return; }
        direct = false;
        row_off = 16;                                        // slot: u32 key in a 16-byte header, then the 16-byte aligned row
        slot_bytes = 16 + ((2 * N + 15) & ~15u);
        cap = (u32)(arena_bytes / slot_bytes);
        limit = (u32)((u64)cap * 7 / 8);
    }
    // returns the row of ctx; fresh = the row has not been touched by this block (contents undefined: treat as all ones)
    __device__ __forceinline__ u16* row(u32 ctx, bool& fresh)
    {
        if (direct) {
            u16* p = (u16*)(base + (u64)ctx * slot_bytes);
            fresh = p[0] == 0;                              // counters never drop below 1; the arena is zeroed per batch
            return p;
        }
        u32 h = __umulhi(ctx * 0x9E3779B1u, cap);           // multiplicative hash, range-reduced to [0, cap)
        for (;;) {
            u8* s = base + (u64)h * slot_bytes;
            if (slot_bytes > 32) asm volatile("prefetch.global.L1 [%0];" :: "l"(s + 32));   // the rest of the row in flight together with the key
            const u32 key = *(u32*)s;
            if (key == ctx + 1) { fresh = false; return (u16*)(s + row_off); }
            if (key == 0) {
                if (used >= limit) fail = true; else { *(u32*)s = ctx + 1; ++used; }
                fresh = true;
                return (u16*)(s + row_off);
            }
            if (++h == cap) h = 0;
        }
    }
};

struct RcDec {                     // RangeDecoder (src/RangeCoder.h:98-134)
    u64 low, buffer; u32 range; BitR* r;
    __device__ void start(BitR* r_) { r = r_; buffer = 0; for (u32 i = 1; i <= 8; ++i) buffer |= (u64)r->byte() << (64 - i * 8); low = 0; range = 0xFFFFFFFFu; }
    // GetCumulativeFreq: a valid stream keeps buffer below the old range, so the quotient is a 32-bit division
    __device__ __forceinline__ u32 cum(u32 tot)
    {
        range /= tot;
        return (buffer >> 32) ? (u32)(buffer / range) : (u32)buffer / range;
    }
    __device__ __forceinline__ void update(u32 f, u32 lo)
    {
        const u32 rr = lo * range;
        buffer -= rr; low += rr; range *= f;
        while (range <= 0x00FFFFFFu) {
            if ((low ^ (low + range)) & 0xFF00000000000000ull) { const u32 q = (u32)low; range = (q | 0x00FFFFFFu) - q; }
            buffer = (buffer << 8) + r->byte();
            low <<= 8; range <<= 8;
        }
    }
};

// TSymbolCoderRC<N>::DecodeSymbol (src/SymbolCoderRC.h:50-63) on the row at p. Small rows are held in registers as packed
// u16 pairs (vector load, one 2-byte store back unless the row was fresh or rescaled); larger rows are walked in memory.
template <int N>
__device__ __forceinline__ u32 rc_decode_row(RcDec& rc, u16* p, bool fresh)
{
    u32 c[N / 2];
    if (fresh) {
#pragma unroll
        for (int k = 0; k < N / 2; ++k) c[k] = 0x00010001u;
    } else if (N == 4) { const uint2 v = *(const uint2*)p; c[0] = v.x; c[1] = v.y; }
    else {
#pragma unroll
        for (int k = 0; k < N / 8; ++k) { const uint4 v = ((const uint4*)p)[k]; c[4 * k] = v.x; c[4 * k + 1] = v.y; c[4 * k + 2] = v.z; c[4 * k + 3] = v.w; }
    }
    u32 tot = 0;
#pragma unroll
    for (int k = 0; k < N / 2; ++k) tot += (c[k] & 0xFFFFu) + (c[k] >> 16);
    bool whole = fresh;
    if (tot >= (1u << 16) - 2 * N) {
        tot = 0; whole = true;
#pragma unroll
        for (int k = 0; k < N / 2; ++k) { u32 lo = c[k] & 0xFFFFu, hi = c[k] >> 16; lo -= lo >> 1; hi -= hi >> 1; c[k] = lo | (hi << 16); tot += lo + hi; }
    }
    const u32 cul = rc.cum(tot);
    u32 idx = N - 1, f = 0, hi = 0, acc = 0; bool found = false;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const u32 v = (c[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
        acc += v;
        const bool hit = !found && (acc > cul || k == N - 1);
        if (hit) { idx = k; f = v; hi = acc - v; found = true; }
    }
    rc.update(f, hi);
    if (whole) {
#pragma unroll
        for (int k = 0; k < N / 2; ++k) c[k] += (u32)k == (idx >> 1) ? (2u << ((idx & 1) * 16)) : 0u;
        if (N == 4) *(uint2*)p = make_uint2(c[0], c[1]);
        else {
#pragma unroll
            for (int k = 0; k < N / 8; ++k) ((uint4*)p)[k] = make_uint4(c[4 * k], c[4 * k + 1], c[4 * k + 2], c[4 * k + 3]);
        }
    } else p[idx] = (u16)(f + 2);
    return idx;
}
// rows of 32 / 64 / 128 counters stay in memory (L1): one pass of 16-byte loads gives the sums of the groups of 8 counters (kept in
// registers) and the total; the group holding the cumulative target is reloaded and searched.
__device__ __forceinline__ u32 hsum8(const uint4 v)
{
    const u32 a = (v.x & 0xFFFFu) + (v.x >> 16) + (v.y & 0xFFFFu) + (v.y >> 16);
    return a + (v.z & 0xFFFFu) + (v.z >> 16) + (v.w & 0xFFFFu) + (v.w >> 16);
}
template <int N>
__device__ u32 rc_decode_row_big(RcDec& rc, u16* stt, bool fresh)
{
    uint4* row = (uint4*)stt;
    u32 g[N / 8];
    u32 tot = 0;
    if (fresh) {
        const uint4 ones = make_uint4(0x00010001u, 0x00010001u, 0x00010001u, 0x00010001u);
#pragma unroll
        for (int k = 0; k < N / 8; ++k) { row[k] = ones; g[k] = 8; }
        tot = N;
    } else {
#pragma unroll
        for (int k = 0; k < N / 8; ++k) { g[k] = hsum8(row[k]); tot += g[k]; }
    }
    if (tot >= (1u << 16) - 2 * N) {
        tot = 0;
#pragma unroll
        for (int k = 0; k < N / 8; ++k) {
            uint4 v = row[k];
            u32* w = (u32*)&v;
#pragma unroll
            for (int j = 0; j < 4; ++j) { u32 lo = w[j] & 0xFFFFu, hi = w[j] >> 16; lo -= lo >> 1; hi -= hi >> 1; w[j] = lo | (hi << 16); }
            row[k] = v; g[k] = hsum8(v); tot += g[k];
        }
    }
    const u32 cul = rc.cum(tot);
    u32 gi = N / 8 - 1, before = 0, acc = 0; bool found = false;
#pragma unroll
    for (int k = 0; k < N / 8; ++k) {
        const bool hit = !found && (acc + g[k] > cul || k == N / 8 - 1);
        if (hit) { gi = k; before = acc; found = true; }
        acc += g[k];
    }
    const uint4 v = row[gi];
    const u32 w[4] = {v.x, v.y, v.z, v.w};
    u32 j = 7, f = 0, hi = 0; acc = before; found = false;
    const bool last_group = gi == N / 8 - 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const u32 c = (w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
        acc += c;
        const bool hit = !found && (acc > cul || (k == 7 && last_group));
        if (hit) { j = k; f = c; hi = acc - c; found = true; }
    }
    if (!found) { j = 7; f = (w[3] >> 16) & 0xFFFFu; hi = acc - f; }     // cannot happen for a consistent row; keeps the state defined
    rc.update(f, hi);
    const u32 idx = gi * 8 + j;
    stt[idx] = (u16)(f + 2);
    return idx;
}
__device__ __forceinline__ u32 rc_decode_any(RcDec& rc, u16* p, u32 N, bool fresh)
{
    switch (N) {
    case 4: return rc_decode_row<4>(rc, p, fresh);
    case 8: return rc_decode_row<8>(rc, p, fresh);
    case 16: return rc_decode_row<16>(rc, p, fresh);
    case 32: return rc_decode_row_big<32>(rc, p, fresh);
    case 64: return rc_decode_row_big<64>(rc, p, fresh);
    default: return rc_decode_row_big<128>(rc, p, fresh);
    }
}

// template arguments per scheme (src/QualityModelerProxy.h:231-254)
__device__ __forceinline__ bool dec_quality_cfg(u32 order, u32 scheme, u32& alpha, u32& bits, u32& so, u32& rescale)
{
    if (scheme > 7 || order < 1 || order > 2) return false;
    const u32 k = scheme & 3;
    alpha = 16u << k; bits = 4 + k;
    so = order == 1 ? (k == 0 ? 3 : k == 1 ? 2 : 1) : (k == 0 ? 4 : k == 1 ? 3 : k == 2 ? 2 : 1);
    rescale = (scheme & 4) ? alpha : 8;
    return true;
}

#ifndef DEC_Q_MINB
#define DEC_Q_MINB 1
#endif
__global__ void __launch_bounds__(DEC_CTA, DEC_Q_MINB) k_dec_quality(Workspace ws, uint2* pool_base, u32 pool_stride, u8* arena, u64 arena_stride, u64 arena_bytes, u32 first_chain)
{
    const u32 blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= ws.n_blocks) return;
    const BlockDesc& d = ws.desc[blk];
    BlockState& st = ws.state[blk];
    if (st.status != ST_OK) return;
    const u8* in = ws.in + d.in_off;
    BitR r; r.init(in, d.in_len, st.stream_size[1]);
    const RecArrays& R = ws.rec;
    const u32 n = st.n_rec, rb = d.rec_base;
    u8* qc = ws.qcat + d.sym_base;
    u32 status = ST_OK, d_total = 0;
    const u32 scheme = r.byte();
    st.q_scheme = (u8)scheme;
    if (scheme == 255) {
        // SchemeNone: no quality data was stored -- the reference leaves the buffer as is; only legal with zero-length reads
        for (u32 k = 0; k < n; ++k) { R.dna_len[rb + k] = R.qua_len[rb + k]; R.dcat_off[rb + k] = d_total; d_total += R.qua_len[rb + k]; }
        if (st.q_total != 0) status = ST_MALFORMED;
    } else if (ws.qua_order > 0) {
        // TQualityOrderModeler::Decode + TTranslationalQualityEncoder (src/QualityOrderModeler.h:53-70, QualityEncoder.h:312-367)
        u32 alpha, bits, so, rescale;
        if (!dec_quality_cfg(ws.qua_order, scheme, alpha, bits, so, rescale)) { st.status = ST_MALFORMED; return; }
        u8 symbols[256]; u32 nsym = 0;
        r.flush();
        for (u32 i = 0; i < 256; ++i) symbols[i] = 255;
        for (u32 i = 0; i < 256; ++i) if (r.bit()) symbols[nsym++] = (u8)i;
        r.flush();
        RowStore rs; rs.setup(arena + (u64)(blk - first_chain) * arena_stride, arena_bytes, alpha, bits * (so + 1));
        // TQualityModelBase::UpdateHash (QualityEncoder.h:77-89)
        const u32 bits_lo = (so / 2) * bits, bits_hi = (so / 2 + 1) * bits;
        const u64 sym_mask = (1ull << bits) - 1, swap_mask = ((1ull << bits_lo) - 1) | ~((1ull << bits_hi) - 1), hash_mask = (1ull << (so * bits)) - 1;
        u64 hash = 0, sym_buf = 0;
        RcDec rc; rc.start(&r);
        for (u32 k = 0; k < n && !r.ovr && !rs.fail; ++k) {
            const u32 len = R.qua_len[rb + k];
            u8* q = qc + R.qcat_off[rb + k];
            u32 nc = 0, acc = 0, pctx = 0;                 // pctx = j * rescale / len, kept incrementally
            for (u32 j = 0; j < len; ++j) {
                const u32 ctx = (u32)(((hash & hash_mask) << bits) | pctx);
                bool fresh;
                u16* row = rs.row(ctx, fresh);
                if (rs.fail) break;
                const u32 sym = rc_decode_any(rc, row, alpha, fresh);
                hash <<= bits;
                const u64 next = (hash >> bits_lo) & sym_mask;
                const u64 swp = (next + sym_buf) / 2;
                hash &= swap_mask; hash |= swp << bits_lo; hash |= sym;
                sym_buf = next;
                const u8 v = symbols[sym];
                q[j] = v; nc += v >= 128;
                acc += rescale; while (acc >= len) { acc -= len; ++pctx; }
            }
            R.dna_len[rb + k] = (u16)(len - nc); R.dcat_off[rb + k] = d_total; d_total += len - nc;
        }
        if (rs.fail) status = ST_RETRY;
    } else if (scheme < 2) {
        // IQualityPositionModeler::Decode, Plain / Truncated (src/QualityPositionModeler.cpp:74-105, 189-224, 289-340)
        NodePool np; np.nodes = pool_base + (u64)blk * pool_stride; np.n = 0; np.cap = pool_stride; np.ovf = false;
        const bool truncated = scheme == 1;
        r.flush();
        const u32 L = r.be32();
        u8 symbols[256]; u32 nsym = 0;
        for (u32 i = 0; i < 256; ++i) symbols[i] = 255;
        for (u32 i = 0; i < 256; ++i) if (r.bit()) symbols[nsym++] = (u8)i;
        if (L > 65535) { st.status = ST_MALFORMED; return; }
        u32* roots = np.words(L ? L : 1);
        if (!roots) { st.status = ST_RETRY; return; }
        for (u32 j = 0; j < L && !np.ovf; ++j) roots[j] = np.load(r);
        if (np.ovf) { st.status = r.ovr ? ST_MALFORMED : ST_RETRY; return; }
        const u32 max_bits = dsrc_bit_length((u64)L);
        const bool variable = truncated ? r.bit() != 0 : false;
        for (u32 k = 0; k < n && !r.ovr; ++k) {
            const u32 len = R.qua_len[rb + k];
            u8* q = qc + R.qcat_off[rb + k];
            u32 th = len, nc = 0;
            if (truncated && r.bit()) th = r.bits(variable ? dsrc_bit_length((u64)len) : max_bits);
            if (th > len || th > L) { status = ST_MALFORMED; break; }
            for (u32 j = 0; j < th; ++j) { const u8 v = symbols[np.get(roots[j], r) & 255]; q[j] = v; nc += v >= 128; }
            for (u32 j = th; j < len; ++j) q[j] = 2;
            R.dna_len[rb + k] = (u16)(len - nc); R.dcat_off[rb + k] = d_total; d_total += len - nc;
        }
        r.flush();
    } else if (scheme == 2) {
        // QualityRLEModeler::Decode (src/QualityRLEModeler.cpp:375-489): runs over the whole block
        NodePool np; np.nodes = pool_base + (u64)blk * pool_stride; np.n = 0; np.cap = pool_stride; np.ovf = false;
        r.flush();
        const u32 runs = r.be32();
        u8 qs[256], ls[256]; u32 nq = 0, nl = 0;
        for (u32 i = 0; i < 256; ++i) if (r.bit()) qs[nq++] = (u8)i;
        for (u32 i = 0; i < 256; ++i) if (r.bit()) ls[nl++] = (u8)i;
        r.flush();
        if (r.ovr || runs == 0 || nq == 0 || nl == 0) { st.status = ST_MALFORMED; return; }
        u32* qroot = nullptr; u32* lroot = nullptr;
        u8 lb = 0, le = 0;
        if (nq > 1) {
            qroot = np.words(nq); lroot = np.words(nq);
            if (!qroot || !lroot) { st.status = ST_RETRY; return; }
            for (u32 i = 0; i < nq && !np.ovf; ++i) { qroot[i] = np.load(r); lroot[i] = np.load(r); }
            if (np.ovf) { st.status = r.ovr ? ST_MALFORMED : ST_RETRY; return; }
            r.flush();
        } else {
            r.flush();
            if (nl > 1) { r.flush(); lb = ls[r.byte() % nl]; le = ls[0]; if (le == lb) le = ls[1]; } else { lb = ls[0]; le = lb; }
        }
        u32 idx = 0, cur_len = 0, prev = 0; u8 cur_q = 0;
        for (u32 k = 0; k < n && status == ST_OK && !r.ovr; ++k) {
            const u32 len = R.qua_len[rb + k];
            u8* q = qc + R.qcat_off[rb + k];
            u32 nc = 0;
            for (u32 j = 0; j < len; ++j) {
                if (cur_len == 0) {
                    if (idx >= runs) { status = ST_MALFORMED; break; }
                    if (nq > 1) {
                        const u32 qi = np.get(qroot[prev], r) % nq;
                        cur_q = qs[qi]; prev = qi;
                        cur_len = (u32)ls[np.get(lroot[prev], r) % nl] + 1;
                    } else { cur_q = qs[0]; cur_len = (u32)(idx + 1 == runs ? le : lb) + 1; }
                    ++idx;
                }
                q[j] = cur_q; --cur_len; nc += cur_q >= 128;
            }
            R.dna_len[rb + k] = (u16)(len - nc); R.dcat_off[rb + k] = d_total; d_total += len - nc;
        }
        r.flush();
    } else status = ST_MALFORMED;
    if (status == ST_OK && r.ovr) status = ST_MALFORMED;
    if (status == ST_OK && d_total > d.sym_cap) status = ST_OVERFLOW;
    st.status = status;
    st.d_total = d_total;
    st.stream_size[3] = r.pos;                            // where the DNA stream starts
}

__global__ void __launch_bounds__(DEC_CTA) k_dec_dna(Workspace ws, uint2* pool_base, u32 pool_stride, u8* arena, u64 arena_stride, u64 arena_bytes, u32 first_chain)
{
    const u32 blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= ws.n_blocks) return;
    const BlockDesc& d = ws.desc[blk];
    BlockState& st = ws.state[blk];
    if (st.status != ST_OK) return;
    const u8* in = ws.in + d.in_off;
    BitR r; r.init(in, d.in_len, st.stream_size[3]);
    u8* dc = ws.dcat + d.sym_base;
    const u32 M = st.d_total;
    u32 status = ST_OK;
    const u32 scheme = r.byte();
    st.d_scheme = (u8)scheme;
    if (scheme == 255) { if (M != 0) status = ST_MALFORMED; }
    else if (scheme > 1) status = ST_MALFORMED;
    else if (ws.dna_order == 0) {
        if (scheme == 0) {          // DnaModelerBasicB2::Decode (src/DnaModelerBasicB2.h:48-62)
            for (u32 i = 0; i < M; ++i) dc[i] = (u8)r.bits(2);
            r.flush();
        } else {                    // DnaModelerHuffman::Decode (src/DnaModelerHuffman.cpp:75-118)
            NodePool np; np.nodes = pool_base + (u64)blk * pool_stride; np.n = 0; np.cap = pool_stride; np.ovf = false;
            u8 symbols[20]; u32 ns = 0;
            for (u32 i = 0; i < 20; ++i) symbols[i] = 255;
            for (u32 i = 0; i < 20; ++i) if (r.bit()) symbols[ns++] = (u8)i;
            const u32 root = np.load(r);
            if (np.ovf) status = r.ovr ? ST_MALFORMED : ST_RETRY;
            for (u32 i = 0; i < M && !r.ovr && status == ST_OK; ++i) dc[i] = symbols[np.get(root, r) % 20];
            r.flush();
        }
    } else {                        // TDnaRCOrderModeler::Decode (src/DnaModelerRCO.h:64-81)
        const u32 alpha = scheme == 0 ? 4 : 8, bits = scheme == 0 ? 2 : 3;
        const u32 ord = (scheme == 1 && ws.dna_order > 7) ? 7 : ws.dna_order;
        const u32 mask = (1u << (ord * bits)) - 1;
        RowStore rs; rs.setup(arena + (u64)(blk - first_chain) * arena_stride, arena_bytes, alpha, ord * bits);
        RcDec rc; rc.start(&r);
        u32 hash = 0;
        for (u32 i = 0; i < M && !r.ovr; ++i) {
            if (rs.direct) {
                // the rows the next two symbols can meet are contiguous (the new base enters the low bits of the context):
                // start fetching them now, the chain is otherwise one DRAM round trip per base
                const u8* nx1 = rs.base + (u64)((hash << bits) & mask) * rs.slot_bytes;
                const u8* nx2 = rs.base + (u64)((hash << (2 * bits)) & mask) * rs.slot_bytes;
#if DEC_DNA_PF >= 1
                asm volatile("prefetch.global.L1 [%0];" :: "l"(nx1));
#endif
#if DEC_DNA_PF >= 2
                asm volatile("prefetch.global.L2 [%0];" :: "l"(nx2));
                if (alpha == 8) asm volatile("prefetch.global.L2 [%0];" :: "l"(nx2 + 512));
#endif
            }
            bool fresh;
            u16* row = rs.row(hash, fresh);
            if (rs.fail) break;
            const u32 sym = rc_decode_any(rc, row, alpha, fresh);
            dc[i] = (u8)sym;
            hash = ((hash << bits) | sym) & mask;
        }
        if (rs.fail) status = ST_RETRY;
    }
    if (status == ST_OK && r.ovr) status = ST_MALFORMED;
    st.status = status;
    st.stream_size[2] = r.pos;
}

// FASTQ layout of ReadTags (BlockCompressor.cpp:521-570) + ProcessBackward (RecordsProcessor.cpp:269-295): one CTA per block
__global__ void __launch_bounds__(DSRC_CTA) k_dec_assemble(Workspace ws)
{
    const u32 blk = blockIdx.x;
    const BlockDesc& d = ws.desc[blk];
    const BlockState& st = ws.state[blk];
    BlockResult& res = ws.result[blk];
    if (threadIdx.x == 0) { res.status = st.status; res.total_size = st.status == ST_OK ? st.chunk_size : 0; res.out_off = d.out_off; }
    if (st.status != ST_OK) return;
    const RecArrays& R = ws.rec;
    const u32 n = st.n_rec, rb = d.rec_base;
    const u8* tb = ws.streams + d.stream_base;
    const u8* qc = ws.qcat + d.sym_base;
    const u8* dc = ws.dcat + d.sym_base;
    u8* out = ws.out + d.out_off;
    __shared__ u32 sm[DSRC_WARPS + 1];
    __shared__ u32 s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    // output offset of every record (exclusive scan of record sizes) -> seq_off
    for (u32 base = 0; base < n; base += DSRC_CTA) {
        const u32 r = base + threadIdx.x;
        u32 sz = 0;
        if (r < n) { const u32 tl = R.title_len[rb + r], ql = R.qua_len[rb + r]; sz = tl + 1 + ql + 2 + (ws.plus_rep ? tl - 1 : 0) + 1 + ql + 1; }
        u32 total, ex = block_excl_sum(sz, sm, &total);
        const u32 carry = s_carry;
        if (r < n) R.seq_off[rb + r] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    const char* alphabet = "AGCTNRWSKMDVHBYXU.-";
    const u32 w = warp_id(), ln = lane_id();
    for (u32 r = w; r < n; r += DSRC_WARPS) {
        const u32 g = rb + r;
        const u32 tl = R.title_len[g], ql = R.qua_len[g];
        const u8* t = tb + R.title_off[g];
        const u8* q = qc + R.qcat_off[g];
        const u8* s = dc + R.dcat_off[g];
        u8* o = out + R.seq_off[g];
        for (u32 i = ln; i < tl; i += 32) o[i] = t[i];
        if (ln == 0) o[tl] = '\n';
        u8* os = o + tl + 1;
        u8* op = os + ql + 1;                              // '+' line
        if (ln == 0) { os[ql] = '\n'; op[0] = '+'; }
        u32 pl = 1;
        if (ws.plus_rep) { for (u32 i = ln; i + 1 < tl; i += 32) op[1 + i] = t[1 + i]; pl = tl; }
        if (ln == 0) op[pl] = '\n';
        u8* oq = op + pl + 1;
        u32 base = 0;
        for (u32 j0 = 0; j0 < ql; j0 += 32) {
            const u32 j = j0 + ln; const bool in = j < ql;
            u32 qv = in ? q[j] : 0;
            const bool kept = in && qv < 128;
            const u32 m = __ballot_sync(0xFFFFFFFFu, kept);
            u32 sv;
            if (kept) sv = s[base + __popc(m & ((1u << ln) - 1))];
            else { sv = (qv - 128 + 16) / 8 + 3 - 1; qv &= 7; }
            if (in) { os[j] = sv < 19 ? (u8)alphabet[sv] : (u8)255; oq[j] = (u8)(ws.qoff + qv); }
            base += __popc(m);
        }
        if (ln == 0) oq[ql] = '\n';
    }
}

void launch_dec_probe(const Workspace& ws, cudaStream_t s) { k_dec_probe<<<(ws.n_blocks + 127) / 128, 128, 0, s>>>(ws); }
void launch_dec_tags(const Workspace& ws, cudaStream_t s, void* pool, u32 pool_stride)
{
    k_dec_tags<<<(ws.n_blocks + DEC_CTA - 1) / DEC_CTA, DEC_CTA, 0, s>>>(ws, (uint2*)pool, pool_stride);
}
void launch_dec_quality(const Workspace& ws, cudaStream_t s, void* pool, u32 pool_stride, u8* arena, u64 arena_stride, u64 arena_bytes)
{
    k_dec_quality<<<(ws.n_blocks + DEC_CTA - 1) / DEC_CTA, DEC_CTA, 0, s>>>(ws, (uint2*)pool, pool_stride, arena, arena_stride, arena_bytes, 0);
}
void launch_dec_dna(const Workspace& ws, cudaStream_t s, void* pool, u32 pool_stride, u8* arena, u64 arena_stride, u64 arena_bytes)
{
    k_dec_dna<<<(ws.n_blocks + DEC_CTA - 1) / DEC_CTA, DEC_CTA, 0, s>>>(ws, (uint2*)pool, pool_stride, arena, arena_stride, arena_bytes, 0);
}
void launch_dec_assemble(const Workspace& ws, cudaStream_t s) { k_dec_assemble<<<ws.n_blocks, DSRC_CTA, 0, s>>>(ws); }
