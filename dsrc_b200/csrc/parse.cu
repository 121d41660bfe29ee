// Block parsing and record preprocessing kernels (one CTA owns one block).
//   k_count_lines : counts line terminators per block so the host can size the record arrays exactly
//   k_parse       : FastqParser::ParseFrom / ReadNextRecord / SkipLine   (src/FastqParser.cpp:140-164, FastqParser.h:40-115)
//   k_preprocess  : LosslessRecordsProcessor::ProcessForward + Initialize/FinalizeStats
//                   (src/RecordsProcessor.cpp:104-133, 180-267) and BlockCompressor::AnalyzeMetaData
//                   (src/BlockCompressor.cpp:184-205), re-expressed as scans + warp ballots:
//                   the reference compacts in place; we emit two dense per-block symbol arrays
//                   (qcat: processed quality bytes, dcat: retained DNA indices) that every later
//                   kernel indexes with coalesced accesses.
#include "common.cuh"
#include "kernels.h"

// a byte terminates a line iff it is '\r', or '\n' not directly after '\r' (SkipLine: CR LF counts once)
__device__ __forceinline__ bool is_term(const u8* b, u32 p)
{
    u8 c = b[p];
    return c == '\r' || (c == '\n' && !(p > 0 && b[p - 1] == '\r'));
}

// Terminator scan over 16-byte aligned chunks. Chunk c of a block covers block positions [16c - mis, 16c - mis + 16) where
// mis = misalignment of the block start; term_mask() returns one bit per byte of the chunk (bit k = byte k terminates a line)
// and the CR bytes among them, using the byte-wise SIMD compares. prev_cr: the byte before the chunk is '\r'.
struct ChunkScan {
    const u8* base; u32 mis, len;                    // base = block start rounded down to 16 bytes
    __device__ __forceinline__ void init(const u8* b, u32 in_len) { mis = (u32)((uintptr_t)b & 15u); base = b - mis; len = in_len; }
    __device__ __forceinline__ u32 n_chunks() const { return (len + mis + 15) / 16; }
    // valid-byte mask of chunk c (bits of positions inside [0, len))
    __device__ __forceinline__ u32 valid(u32 c) const
    {
        const i32 lo = (i32)(16 * c) - (i32)mis;       // block position of byte 0
        u32 m = 0xFFFFu;
        if (lo < 0) m &= 0xFFFFu << (u32)(-lo);
        if ((u32)(lo + 16) > len) m &= 0xFFFFu >> (u32)(lo + 16 - (i32)len);
        return m;
    }
    __device__ __forceinline__ void masks(u32 c, u32& term, u32& cr) const
    {
        const uint4 v = *(const uint4*)(base + 16 * (u64)c);
        const u32 w[4] = {v.x, v.y, v.z, v.w};
        u32 crm = 0, nlm = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u32 a = __vcmpeq4(w[k], 0x0D0D0D0Du) & 0x01010101u, n = __vcmpeq4(w[k], 0x0A0A0A0Au) & 0x01010101u;
            // gather the 4 flag bits (bit 0 of every byte) into a nibble
            crm |= (((a * 0x00204081u) >> 21) & 0xFu) << (4 * k);   // bits 0,8,16,24 -> 0..3
            nlm |= (((n * 0x00204081u) >> 21) & 0xFu) << (4 * k);
        }
        const u32 vm = valid(c);
        crm &= vm; nlm &= vm;
        u32 prev = (crm << 1) & 0xFFFFu;
        const i32 lo = (i32)(16 * c) - (i32)mis;
        if (lo > 0 && base[16 * (u64)c - 1] == '\r') prev |= 1u;
        term = crm | (nlm & ~prev); cr = crm;
    }
};

__global__ void __launch_bounds__(DSRC_CTA) k_count_lines(Workspace ws)
{
    const BlockDesc& d = ws.desc[blockIdx.x];
    const u8* b = ws.in + d.in_off;
    ChunkScan cs; cs.init(b, d.in_len);
    const u32 nc = cs.n_chunks();
    u32 cnt = 0;
    for (u32 c = threadIdx.x; c < nc; c += DSRC_CTA) { u32 t, cr; cs.masks(c, t, cr); cnt += __popc(t); }
    cnt = warp_red_sum(cnt);
    __shared__ u32 sm[DSRC_WARPS];
    if (lane_id() == 0) sm[warp_id()] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
        for (int w = 0; w < DSRC_WARPS; ++w) t += sm[w];
        BlockState& st = ws.state[blockIdx.x];
        st.n_lines = t + 1;     // the last line has no terminator (the chunk is cut before its final '\n')
        st.status = ST_OK;
        // field count of the first title (TagAnalyzer::InitializeFieldsStats, src/TagModeler.cpp:159-224): sizes the field table
        u32 nf = 0, i = 0;
        for (; i < d.in_len && b[i] != '\n' && b[i] != '\r'; ++i) {
            u8 c = b[i];
            nf += (c == ' ' || c == '.' || c == '_' || c == ',' || c == '=' || c == ':' || c == '/' || c == '-' || c == '#' || c == 0);
        }
        ws.probe[blockIdx.x].n_lines = t + 1;
        ws.probe[blockIdx.x].n_fields = nf + 1;
    }
}

#define PARSE_CH 4                                   // consecutive 16-byte chunks per thread and step
__global__ void __launch_bounds__(DSRC_CTA) k_parse(Workspace ws)
{
    const BlockDesc& d = ws.desc[blockIdx.x];
    BlockState& st = ws.state[blockIdx.x];
    const u8* b = ws.in + d.in_off;
    u32* lines = ws.lines + d.line_base;
    __shared__ u32 sm[DSRC_WARPS + 1];
    __shared__ u32 s_carry, s_skipped, s_bad;
    __shared__ unsigned long long s_raw[3];
    if (threadIdx.x == 0) { s_carry = 0; s_skipped = 0; s_bad = 0xFFFFFFFFu; s_raw[0] = s_raw[1] = s_raw[2] = 0; }
    __syncthreads();
    // phase A: positions of all terminators, in order. lines[k] = position | (CRLF ? 1<<31 : 0)
    u32 skipped = 0;
    ChunkScan cs; cs.init(b, d.in_len);
    const u32 nc = cs.n_chunks();
    for (u32 c0 = 0; c0 < nc; c0 += DSRC_CTA * PARSE_CH) {
        u32 tm[PARSE_CH], crm[PARSE_CH], cnt = 0;
        const u32 cb = c0 + threadIdx.x * PARSE_CH;
#pragma unroll
        for (int k = 0; k < PARSE_CH; ++k) { tm[k] = 0; crm[k] = 0; if (cb + k < nc) cs.masks(cb + k, tm[k], crm[k]); cnt += __popc(tm[k]); }
        u32 total, ex = block_excl_sum(cnt, sm, &total);
        u32 kq = s_carry + ex;
#pragma unroll
        for (int k = 0; k < PARSE_CH; ++k) {
            u32 m = tm[k];
            while (m) {
                const u32 bit = __ffs(m) - 1; m &= m - 1;
                const u32 p = 16 * (cb + k) + bit - cs.mis;
                const bool crlf = ((crm[k] >> bit) & 1u) && p + 1 < d.in_len && b[p + 1] == '\n';
                skipped += crlf;
                if (kq < d.line_cap) lines[kq] = p | (crlf ? 0x80000000u : 0u);
                ++kq;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    skipped = warp_red_sum(skipped);
    if (lane_id() == 0 && skipped) atomicAdd(&s_skipped, skipped);
    __syncthreads();
    const u32 n_term = s_carry;
    // the final line ends at in_len unless the chunk ends with a terminator
    u32 last_end = n_term ? ((lines[n_term - 1] & 0x7FFFFFFFu) + 1 + (lines[n_term - 1] >> 31)) : 0;
    u32 n_lines = n_term + (last_end < d.in_len ? 1 : 0);
    u32 n_rec = n_lines / 4;
    if (n_lines % 4 != 0 || n_rec == 0 || n_rec > d.rec_cap) {   // the reference would silently drop the tail record
        if (threadIdx.x == 0) { st.status = ST_MALFORMED; st.n_rec = 0; }
        return;
    }
    // phase B: one thread per record
    const RecArrays& R = ws.rec;
    unsigned long long rt = 0, rs = 0, rq = 0;
    for (u32 r = threadIdx.x; r < n_rec; r += DSRC_CTA) {
        u32 e[4], s[4];
        for (int l = 0; l < 4; ++l) {
            u32 k = 4 * r + l;
            u32 prev = k ? lines[k - 1] : 0;
            s[l] = k ? ((prev & 0x7FFFFFFFu) + 1 + (prev >> 31)) : 0;
            e[l] = k < n_term ? (lines[k] & 0x7FFFFFFFu) : d.in_len;
        }
        u32 tl = e[0] - s[0], sl = e[1] - s[1], pl = e[2] - s[2], ql = e[3] - s[3];
        bool ok = tl > 0 && tl <= 65535 && b[s[0]] == '@' && pl > 0 && sl == ql && sl <= 65535;
        if (!ok) atomicMin(&s_bad, r);
        u32 g = d.rec_base + r;
        R.title_off[g] = s[0]; R.title_len[g] = (u16)tl;
        R.seq_off[g] = s[1]; R.qua_off[g] = s[3]; R.qua_len[g] = (u16)ql;
        rt += tl; rs += sl; rq += ql;
    }
    for (int o = 16; o; o >>= 1) {
        rt += __shfl_xor_sync(0xFFFFFFFFu, rt, o); rs += __shfl_xor_sync(0xFFFFFFFFu, rs, o); rq += __shfl_xor_sync(0xFFFFFFFFu, rq, o);
    }
    if (lane_id() == 0) { atomicAdd(&s_raw[0], rt); atomicAdd(&s_raw[1], rs); atomicAdd(&s_raw[2], rq); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_bad != 0xFFFFFFFFu) { st.status = ST_MALFORMED; st.n_rec = 0; return; }
        st.n_rec = n_rec;
        st.chunk_size = d.in_len - s_skipped;
        st.raw[0] = 0; st.raw[1] = s_raw[0]; st.raw[2] = s_raw[1]; st.raw[3] = s_raw[2];
    }
}

// dnaToIndexTable (src/RecordsProcessor.cpp:185-205): A0 G1 C2 T3 N4 R5 W6 S7 K8 M9 D10 V11 H12 B13 Y14 X15 U16 .17 -18
__device__ __forceinline__ u32 dna_index(u8 c)
{
    switch (c) {
    case 'A': return 0; case 'G': return 1; case 'C': return 2; case 'T': return 3; case 'N': return 4;
    case 'R': return 5; case 'W': return 6; case 'S': return 7; case 'K': return 8; case 'M': return 9;
    case 'D': return 10; case 'V': return 11; case 'H': return 12; case 'B': return 13; case 'Y': return 14;
    case 'X': return 15; case 'U': return 16; case '.': return 17; case '-': return 18;
    default: return 255;
    }
}

__global__ void __launch_bounds__(DSRC_CTA) k_preprocess(Workspace ws)
{
    const BlockDesc& d = ws.desc[blockIdx.x];
    BlockState& st = ws.state[blockIdx.x];
    if (st.status != ST_OK || st.pre_flat) return;    // pre_flat: k_preprocess_flat has done the block
    const u8* b = ws.in + d.in_off;
    const RecArrays& R = ws.rec;
    const u32 n_rec = st.n_rec, rb = d.rec_base;
    u8* qcat = ws.qcat + d.sym_base;
    u8* dcat = ws.dcat + d.sym_base;

    __shared__ u32 sm[DSRC_WARPS + 1];
    __shared__ u32 s_carry;
    __shared__ u32 s_qf[DSRC_WARPS][256];
    __shared__ u32 s_df[DSRC_WARPS][20];
    __shared__ u32 s_min, s_max, s_th, s_rle, s_bad;
    __shared__ u8 s_lut[256];                          // dnaToIndexTable, once per CTA
    s_lut[threadIdx.x] = (u8)dna_index((u8)threadIdx.x);
    for (u32 i = threadIdx.x; i < DSRC_WARPS * 256; i += DSRC_CTA) (&s_qf[0][0])[i] = 0;
    for (u32 i = threadIdx.x; i < DSRC_WARPS * 20; i += DSRC_CTA) (&s_df[0][0])[i] = 0;
    if (threadIdx.x == 0) { s_carry = 0; s_min = 0xFFFFFFFFu; s_max = 0; s_th = 0; s_rle = 0; s_bad = 0; }
    __syncthreads();

    // A: exclusive prefix of quality lengths -> qcat offsets
    for (u32 base = 0; base < n_rec; base += DSRC_CTA) {
        u32 r = base + threadIdx.x;
        u32 len = r < n_rec ? R.qua_len[rb + r] : 0;
        u32 total, ex = block_excl_sum(len, sm, &total);
        u32 carry = s_carry;
        if (r < n_rec) R.qcat_off[rb + r] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    const u32 q_total = s_carry;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = 0;
    if (q_total > d.sym_cap) { if (threadIdx.x == 0) st.status = ST_OVERFLOW; return; }

    // B: one warp per record: LUT, offset removal, ambiguity transfer, statistics
    const u32 w = warp_id(), ln = lane_id();
    u32 mn = 0xFFFFFFFFu, mx = 0, th_sum = 0, rle_sum = 0, bad = 0;
    for (u32 r = w; r < n_rec; r += DSRC_WARPS) {
        const u32 len = R.qua_len[rb + r];
        const u8* seq = b + R.seq_off[rb + r];
        const u8* qua = b + R.qua_off[rb + r];
        u8* qo = qcat + R.qcat_off[rb + r];
        u32 kept = 0, th = 0, rle = 0; u32 last_q = 255;
        for (u32 j0 = 0; j0 < len; j0 += 32) {
            u32 j = j0 + ln; bool in = j < len;
            u32 s = in ? s_lut[seq[j]] : 0;
            u32 q = in ? (u8)(qua[j] - ws.qoff) : 0;
            bool moved = in && s > 3 && q < 7;                    // RecordsProcessor.cpp:228-233
            if (moved) q = (u8)(q + (128 + ((s - 3 + 1) << 3) - 16));
            bool keep = in && !moved;
            if (in && s == 255) bad = 1;                         // not a DNA symbol the reference knows
            if (in) { qo[j] = (u8)q; atomicAdd(&s_qf[w][q], 1u); }
            if (keep && s < 20) atomicAdd(&s_df[w][s], 1u);
            kept += __popc(__ballot_sync(0xFFFFFFFFu, keep));
            u32 pq = __shfl_up_sync(0xFFFFFFFFu, q, 1);
            if (ln == 0) pq = last_q;
            rle += __popc(__ballot_sync(0xFFFFFFFFu, in && q != pq));
            u32 not2 = __ballot_sync(0xFFFFFFFFu, in && q != 2);
            if (not2) th = j0 + 31 - __clz(not2);
            u32 cnt_in = min(32u, len - j0);
            last_q = __shfl_sync(0xFFFFFFFFu, q, cnt_in - 1);
        }
        if (len > 0 && last_q == 2 && rle > 0) rle -= 1;          // :259-260 (per record; the block counter is > 0 whenever it matters)
        if (ln == 0) { R.dna_len[rb + r] = (u16)kept; R.trunc_len[rb + r] = (u16)(th + (len > 0)); }
        mn = min(mn, len); mx = max(mx, len); th_sum += th; rle_sum += rle;
    }
    if (ln == 0) { atomicMin(&s_min, mn); atomicMax(&s_max, mx); atomicAdd(&s_th, th_sum); atomicAdd(&s_rle, rle_sum); if (bad) s_bad = 1; }
    if (bad && ln != 0) s_bad = 1;
    __syncthreads();

    // C: exclusive prefix of retained-base counts -> dcat offsets
    for (u32 base = 0; base < n_rec; base += DSRC_CTA) {
        u32 r = base + threadIdx.x;
        u32 len = r < n_rec ? R.dna_len[rb + r] : 0;
        u32 total, ex = block_excl_sum(len, sm, &total);
        u32 carry = s_carry;
        if (r < n_rec) R.dcat_off[rb + r] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    const u32 d_total = s_carry;

    // D: compaction of the retained bases (ballot prefix inside the warp)
    for (u32 r = w; r < n_rec; r += DSRC_WARPS) {
        const u32 len = R.qua_len[rb + r];
        const u8* seq = b + R.seq_off[rb + r];
        const u8* qua = b + R.qua_off[rb + r];
        u8* dout = dcat + R.dcat_off[rb + r];
        u32 base = 0;
        for (u32 j0 = 0; j0 < len; j0 += 32) {
            u32 j = j0 + ln; bool in = j < len;
            u32 s = in ? s_lut[seq[j]] : 0;
            u32 q = in ? (u8)(qua[j] - ws.qoff) : 0;
            bool keep = in && !(s > 3 && q < 7);
            u32 m = __ballot_sync(0xFFFFFFFFu, keep);
            if (keep) dout[base + __popc(m & ((1u << ln) - 1))] = (u8)s;
            base += __popc(m);
        }
    }
    __syncthreads();

    // E: FinalizeStats (dense symbol ranks) + AnalyzeMetaData
    if (threadIdx.x < 256) {
        u32 f = 0;
        for (int k = 0; k < DSRC_WARPS; ++k) f += s_qf[k][threadIdx.x];
        st.qfreq[threadIdx.x] = f;
        s_qf[0][threadIdx.x] = f;
    }
    if (threadIdx.x < 20) {
        u32 f = 0;
        for (int k = 0; k < DSRC_WARPS; ++k) f += s_df[k][threadIdx.x];
        st.dfreq[threadIdx.x] = f;
        s_df[0][threadIdx.x] = f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 qc = 0, dc = 0;
        for (u32 i = 0; i < 256; ++i) st.qrank[i] = s_qf[0][i] ? (u8)qc++ : (u8)255;
        for (u32 i = 0; i < 20; ++i) st.drank[i] = s_df[0][i] ? (u8)dc++ : (u8)255;
        st.q_count = qc; st.d_count = dc;
        st.q_total = q_total; st.d_total = d_total;
        st.min_len = s_min; st.max_len = s_max; st.raw_len = q_total; st.th_len = s_th; st.rle_len = s_rle;
        st.flags = (s_min != s_max) ? 2u : 0u;
        if (s_bad) st.status = ST_UNSUPPORTED;
    }
}

// 0x80 in every non-zero byte of x
__device__ __forceinline__ u32 nz_bytes(u32 x) { return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }

// ---- k_preprocess_flat: the same work for blocks whose reads all have one length (Illumina) under a range-coded quality model, laid
// out by POSITION instead of by record. The byte-per-lane kernel above spends 134 warp instructions per 32 symbols; here a thread takes
// 8 consecutive positions of the block's quality string (its record and offset by one multiply-high), fetches the 8 quality and 8
// base bytes with aligned 8-byte loads (two pieces where the positions straddle two records), and works on them four at a time
// inside 32-bit words: A/C/G/T -> index by two byte permutes (any other letter sends the thread's 8 symbols down a per-byte path:
// ambiguity transfer, RecordsProcessor.cpp:228-233), byte-wise offset removal, change / presence / base counts by byte masks and
// popcounts. The processed qualities leave as one aligned 8-byte store per thread; the retained bases are compacted across the CTA
// (exclusive scan of the threads' counts) through a shared staging buffer that keeps the ragged tail for the next 2048 positions,
// so they leave as aligned 8-byte stores too. Of the per-record arrays only qcat_off is filled: dcat_off, dna_len and trunc_len
// are never needed for these blocks (their users are the -q0 coders).
#define FLAT_SYMS (DSRC_CTA * 8)
#ifndef FLAT_PF
#define FLAT_PF 3                                    // rounds the L2 prefetch runs ahead
#endif
__device__ __forceinline__ u64 flat_load8(const u8* b, u32 off, u32 in_len)
{
    if (off + 16 <= in_len) {                          // two aligned words hold the 8 bytes
        const u8* p = b + off;
        const u32 a = (u32)((uintptr_t)p & 7u), sh = 8 * (a & 3u);
        const uint2* B = (const uint2*)(p - a);
        const uint2 B0 = B[0], B1 = B[1];                // (both inside the block: off + 16 <= in_len)
        const u32 lo = a < 4 ? __funnelshift_r(B0.x, B0.y, sh) : __funnelshift_r(B0.y, B1.x, sh);
        const u32 hi = a < 4 ? __funnelshift_r(B0.y, B1.x, sh) : __funnelshift_r(B1.x, B1.y, sh);
        return ((u64)hi << 32) | lo;
    }
    u64 v = 0;                                         // the block's last bytes: nothing is read beyond them
    for (u32 j = 0; j < 8 && off + j < in_len; ++j) v |= (u64)b[off + j] << (8 * j);
    return v;
}
#ifndef FLAT_MINB
#define FLAT_MINB 4
#endif
__global__ void __launch_bounds__(DSRC_CTA, FLAT_MINB) k_preprocess_flat(Workspace ws)
{
    const BlockDesc& d = ws.desc[blockIdx.x];
    BlockState& st = ws.state[blockIdx.x];
    if (st.status != ST_OK) return;
    const u8* b = ws.in + d.in_off;
    const RecArrays& R = ws.rec;
    const u32 n_rec = st.n_rec, rb = d.rec_base, tid = threadIdx.x, w = warp_id(), ln = lane_id();
    __shared__ u32 sm[DSRC_WARPS + 1];
    __shared__ u32 s_min, s_max, s_rle, s_bad, s_ends2;
    __shared__ u8 s_qp[256];
    __shared__ u32 s_df[DSRC_WARPS][20];
    __shared__ u8 s_lut[256];
    __shared__ u8 s_last[2][DSRC_CTA + 1];             // [round parity][thread + 1]: last processed quality byte of every thread
    __shared__ __align__(16) u8 s_stage[2][FLAT_SYMS + 16];
    s_lut[tid] = (u8)dna_index((u8)tid);
    s_qp[tid] = 0;
    for (u32 i = tid; i < DSRC_WARPS * 20; i += DSRC_CTA) (&s_df[0][0])[i] = 0;
    if (tid == 0) { s_min = 0xFFFFFFFFu; s_max = 0; s_rle = 0; s_bad = 0; s_ends2 = 0; s_last[1][DSRC_CTA] = 255; }
    __syncthreads();
    {   // one read length?
        u32 mn = 0xFFFFFFFFu, mx = 0;
        for (u32 r = tid; r < n_rec; r += DSRC_CTA) { const u32 l = R.qua_len[rb + r]; mn = min(mn, l); mx = max(mx, l); }
        mn = __reduce_min_sync(0xFFFFFFFFu, mn); mx = __reduce_max_sync(0xFFFFFFFFu, mx);
        if (ln == 0) { atomicMin(&s_min, mn); atomicMax(&s_max, mx); }
        __syncthreads();
    }
    const u32 L = s_max;
    const u64 total64 = (u64)n_rec * L;
    if (s_min != L || L < 8 || ws.qua_order == 0 || total64 > d.sym_cap || total64 >= (1u << 22)) { if (tid == 0) st.pre_flat = 0; return; }
    const u32 q_total = (u32)total64;
    for (u32 r = tid; r < n_rec; r += DSRC_CTA) R.qcat_off[rb + r] = r * L;      // (the sort engine's position buckets walk the records, rc_model.cu)
    u8* qcat = ws.qcat + d.sym_base;
    u8* dcat = ws.dcat + d.sym_base;
    const u32 magic_L = 0xFFFFFFFFu / L + 1u;          // p / L = umulhi(p, magic) for p < 2^22, L < 2^10 .. (p * (magic * L - 2^32) < 2^32)
    const u32 qoff4 = ws.qoff * 0x01010101u;
    u32 cnt1 = 0, cnt2 = 0, cnt3 = 0, cntv = 0, chg = 0, ends2 = 0, bad = 0, sawff = 0;
    u32 carry = 0;                                     // retained bases written so far
    // the 8 base and 8 quality bytes of positions [p0, p0 + 8), fetched one round ahead of their use (the round is otherwise two
    // dependent memory round trips -- record offsets, then bytes -- in front of three barriers)
    auto fetch = [&](u32 p0, u64& xs, u64& xq, u32& off) {
        xs = 0; xq = 0; off = 0;
        if (p0 >= q_total) return;
        const u32 n = min(8u, q_total - p0);
        const u32 r = L < 1024 ? __umulhi(p0, magic_L) : p0 / L;
        off = p0 - r * L;
        const u32 n1 = min(8u, L - off);               // symbols of record r; the rest (n - n1, if any) open record r + 1
        xs = flat_load8(b, R.seq_off[rb + r] + off, d.in_len); xq = flat_load8(b, R.qua_off[rb + r] + off, d.in_len);
        if (n1 < n) {
            const u64 m1 = (1ull << (8 * n1)) - 1;
            xs = (xs & m1) | (flat_load8(b, R.seq_off[rb + r + 1], d.in_len) << (8 * n1));
            xq = (xq & m1) | (flat_load8(b, R.qua_off[rb + r + 1], d.in_len) << (8 * n1));
        }
    };
    u64 nxs, nxq; u32 noff;
    fetch(8 * tid, nxs, nxq, noff);
    // the block's bytes are consumed front to back, in_len / rounds per round: the lines of the round after next but one are pulled into L2
    const u32 rounds = (q_total + FLAT_SYMS - 1) / FLAT_SYMS, per_round = d.in_len / rounds + 1;
    for (u32 c0 = 0, it = 0; c0 < q_total; c0 += FLAT_SYMS, ++it) {
        {
            const u32 o = (it + FLAT_PF) * per_round + 128 * tid;
            if (128 * tid < per_round + 128 && o < d.in_len) asm volatile("prefetch.global.L2 [%0];" :: "l"(b + o));
        }
        const u32 p0 = c0 + 8 * tid;
        const u32 n = p0 < q_total ? min(8u, q_total - p0) : 0u;
        u64 q8 = 0, k8 = 0; u32 kept = 0;
        const u64 xs = nxs, xq = nxq; const u32 off = noff;
        fetch(p0 + FLAT_SYMS, nxs, nxq, noff);
        if (n) {
            const u64 vm = n >= 8 ? ~0ull : (1ull << (8 * n)) - 1;
            u32 idx[2], q4[2], amb[2];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const u32 ws_ = (u32)(xs >> (32 * hf)), qw = (u32)(xq >> (32 * hf)), v = (u32)(vm >> (32 * hf));
                // A 0x41, C 0x43, G 0x47, T 0x54: bits 1..2 tell them apart (A 0, C 1, T 2, G 3); two byte permutes give the letter that code
                // stands for (compared with the input) and the index A0 G1 C2 T3
                const u32 c = (ws_ >> 1) & 0x03030303u;
                const u32 sel = (c & 3u) | ((c >> 4) & 0x30u) | ((c >> 8) & 0x300u) | ((c >> 12) & 0x3000u);
                amb[hf] = nz_bytes((__byte_perm(0x47544341u, 0u, sel) ^ ws_) & v);               // 0x80 in every byte that is not A/C/G/T
                idx[hf] = __byte_perm(0x01030200u, 0u, sel) & v;
                q4[hf] = (((qw | 0x80808080u) - (qoff4 & 0x7F7F7F7Fu)) ^ ((qw ^ ~qoff4) & 0x80808080u)) & v;    // byte-wise qua - offset (wrapping)
            }
            kept = n;
            u32 drop[2] = {0u, 0u};                      // 0x80 in every byte whose base moves into its quality byte
            if (amb[0] | amb[1]) {
                // letters that are not A/C/G/T, one by one: LUT; a low-quality one moves into its quality byte and leaves the base
                // string (ambiguity transfer, RecordsProcessor.cpp:228-233), the others stay with their index
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    for (u32 m = amb[hf]; m; m &= m - 1) {
                        const u32 sh = (u32)__ffs((int)m) - 8;                                     // bit offset of the byte
                        const u32 s = s_lut[((u32)(xs >> (32 * hf)) >> sh) & 255u], q = (q4[hf] >> sh) & 255u;
                        if (s == 255) bad = 1;
                        if (s > 3 && q < 7) { q4[hf] = (q4[hf] & ~(255u << sh)) | (((q + (128 + ((s - 3 + 1) << 3) - 16)) & 255u) << sh); drop[hf] |= 0x80u << sh; --kept; }
                        else { idx[hf] = (idx[hf] & ~(255u << sh)) | ((s & 255u) << sh); if (s < 20) atomicAdd(&s_df[w][s], 1u); else { drop[hf] |= 0x80u << sh; --kept; } }
                    }
                }
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {             // counts of the plain bases (the others were counted above)
                const u32 vb = (((u32)(vm >> (32 * hf)) & 0x80808080u) & ~amb[hf]) >> 7, b0 = idx[hf], b1 = idx[hf] >> 1;
                cnt1 += __popc(b0 & ~b1 & vb); cnt2 += __popc(~b0 & b1 & vb); cnt3 += __popc(b0 & b1 & vb); cntv += __popc(vb);
            }
            q8 = ((u64)q4[1] << 32) | q4[0]; k8 = ((u64)idx[1] << 32) | idx[0];
            for (u64 m = ((u64)drop[1] << 32) | drop[0]; m; ) {          // the moved bases leave the string: close the gaps from the top
                const u32 sh = 63 - __clzll((long long)m) - 7;
                const u64 low = sh ? (1ull << sh) - 1 : 0ull;
                k8 = (k8 & low) | ((k8 >> 8) & ~low);
                m &= ~(0x80ull << sh);
            }
            *(u64*)(qcat + p0) = q8;                   // aligned; the arena has slack behind q_total
        }
        s_last[it & 1][tid + 1] = n ? (u8)(q8 >> (8 * (n - 1))) : (u8)255;
        // position of every thread's retained bases (the scan's barriers also publish s_last)
        u32 total, ex = block_excl_sum(kept, sm, &total);
        if (n) {
            // changes against the previous symbol (255 before a record's first), presence of the quality bytes, records ending with a 2
            const u32 before = tid ? s_last[it & 1][tid] : s_last[(it + 1) & 1][DSRC_CTA];          // thread 0: the previous round's last thread
            const u32 qlo = (u32)q8, qhi = (u32)(q8 >> 32);
            u32 plo = __byte_perm(qlo, before, 0x2104), phi = __byte_perm(qhi, qlo, 0x2107);         // the symbols one position earlier
            const u32 js = off ? L - off : 0u;            // first record start inside these 8 positions (>= 8: none)
            if (js < 4) plo |= 255u << (8 * js); else if (js < 8) phi |= 255u << (8 * (js - 4));
            u32 m0 = nz_bytes(qlo ^ plo), m1 = nz_bytes(qhi ^ phi);
            if (n < 8) { const u64 vmz = ((1ull << (8 * n)) - 1) & 0x8080808080808080ull; m0 &= (u32)vmz; m1 &= (u32)(vmz >> 32); }
            chg += __popc(m0) + __popc(m1);
            for (u32 m = m0; m; m &= m - 1) s_qp[(qlo >> ((u32)__ffs((int)m) - 8)) & 255u] = 1;
            for (u32 m = m1; m; m &= m - 1) s_qp[(qhi >> ((u32)__ffs((int)m) - 8)) & 255u] = 1;
            const u32 je = L - 1 - off;                   // the record's last symbol, if inside
            if (je < n) { const u32 ql = (u32)(q8 >> (8 * je)) & 255u; ends2 += ql == 2; }
            if (nz_bytes(~qlo) != 0x80808080u || nz_bytes(~qhi) != 0x80808080u) sawff = 1;       // a symbol 255: leave the block to k_preprocess
        }
        {
            u8* stage = s_stage[it & 1];
            const u32 pad = carry & 7u;
            u8* mine = stage + pad + ex;
            const u32 klo = (u32)k8, khi = (u32)(k8 >> 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) { if ((u32)j < kept) mine[j] = (u8)(klo >> (8 * j)); if ((u32)j + 4 < kept) mine[j + 4] = (u8)(khi >> (8 * j)); }
            __syncthreads();
            const u32 nq = (pad + total + 7) / 8;
            for (u32 i = tid; i < nq; i += DSRC_CTA) *(u64*)(dcat + (carry - pad) + 8 * i) = *(const u64*)(stage + 8 * i);
            const u32 npad = (pad + total) & 7u;         // ragged tail: first bytes of the next round's staging buffer
            if (tid < npad) s_stage[(it + 1) & 1][tid] = stage[pad + total - npad + tid];
            carry += total;
        }
    }
    __syncthreads();
    {
        cnt1 = __reduce_add_sync(0xFFFFFFFFu, cnt1); cnt2 = __reduce_add_sync(0xFFFFFFFFu, cnt2); cnt3 = __reduce_add_sync(0xFFFFFFFFu, cnt3); cntv = __reduce_add_sync(0xFFFFFFFFu, cntv);
        chg = __reduce_add_sync(0xFFFFFFFFu, chg); ends2 = __reduce_add_sync(0xFFFFFFFFu, ends2);
        bad = __any_sync(0xFFFFFFFFu, bad); sawff = __any_sync(0xFFFFFFFFu, sawff);
        if (ln == 0) {
            atomicAdd(&s_df[w][0], cntv - cnt1 - cnt2 - cnt3); atomicAdd(&s_df[w][1], cnt1); atomicAdd(&s_df[w][2], cnt2); atomicAdd(&s_df[w][3], cnt3);
            atomicAdd(&s_rle, chg); atomicAdd(&s_ends2, ends2);
            if (bad) s_bad |= 1; if (sawff) s_bad |= 2;
        }
    }
    __syncthreads();
    if (s_bad & 2) { if (tid == 0) st.pre_flat = 0; return; }            // (a quality byte one below the offset: the per-record rule for it lives in k_preprocess)
    if (tid < 256) st.qfreq[tid] = s_qp[tid];                            // presence only (nothing downstream uses the counts)
    if (tid < 20) {
        u32 f = 0;
        for (int k = 0; k < DSRC_WARPS; ++k) f += s_df[k][tid];
        st.dfreq[tid] = f;
        s_df[0][tid] = f;
    }
    __syncthreads();
    if (tid == 0) {
        u32 qc = 0, dc = 0;
        for (u32 i = 0; i < 256; ++i) st.qrank[i] = s_qp[i] ? (u8)qc++ : (u8)255;
        for (u32 i = 0; i < 20; ++i) st.drank[i] = s_df[0][i] ? (u8)dc++ : (u8)255;
        st.q_count = qc; st.d_count = dc;
        st.q_total = q_total; st.d_total = carry;
        // every record's first symbol counts as a change; a record ending with symbol 2 counts one less (RecordsProcessor.cpp:259-260)
        st.min_len = L; st.max_len = L; st.raw_len = q_total; st.th_len = 0; st.rle_len = s_rle - s_ends2;
        st.flags = 0;
        if (s_bad & 1) st.status = ST_UNSUPPORTED;
        st.pre_flat = 1;
    }
}

void launch_count_lines(const Workspace& ws, cudaStream_t s) { k_count_lines<<<ws.n_blocks, DSRC_CTA, 0, s>>>(ws); }
void launch_parse(const Workspace& ws, cudaStream_t s) { k_parse<<<ws.n_blocks, DSRC_CTA, 0, s>>>(ws); }
void launch_preprocess(const Workspace& ws, cudaStream_t s)
{
    k_preprocess_flat<<<ws.n_blocks, DSRC_CTA, 0, s>>>(ws);      // blocks with one read length under -q1 / -q2 (sets pre_flat)
    k_preprocess<<<ws.n_blocks, DSRC_CTA, 0, s>>>(ws);           // every other block
}
