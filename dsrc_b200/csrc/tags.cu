// Read-ID (tag) modeler: TagAnalyzer + TagTokenizerEncoder + TagRawEncoder (src/TagModeler.cpp:159-884, 1217-1284)
// and BlockCompressor::AnalyzeTags / StoreTags (src/BlockCompressor.cpp:359-401, 458-488).
//
// The reference walks the records serially, updating std::map histograms and run lists. Every
// statistic it gathers is an order-independent reduction over records or a run-length structure, so here
// one CTA owns one block and
//   P1  one thread per record tokenises its title against the field template of record 0 -> field table
//   P2  per field: min/max/AND reductions over the table (warp shuffles + shared atomics)
//   P3  per numeric field: run structure via a max-scan of run heads (value runs and delta runs, cap 256)
//   P4  histograms for the Huffman-coded fields (shared / global atomics)
//   P5  one warp per Huffman tree
//   P6  header serialisation (one thread; a few hundred bytes)
//   P7  per-record bit lengths -> exclusive scan -> every thread ORs its record's bits into the stream
// The one history-dependent quirk arrives as BlockDesc::tag_cap.
#include "common.cuh"
#include "huff.cuh"
#include "kernels.h"

struct Slot256 { u32 hist[256]; u32 code[256]; u8 len[256]; u32 ser_size; u8 ser[380]; };
struct Slot512 { u32 hist[512]; u32 code[512]; u8 len[512]; u32 ser_size; u8 ser[764]; };
struct TagPool {
    HufWork work[DSRC_WARPS];
    Slot256 text[TAG_TEXT_SLOTS];
    Slot512 num[TAG_NUM_SLOTS];
};
u64 tagpool_bytes_per_block() { return (sizeof(TagPool) + 255) & ~(u64)255; }

// field table entry: value:32 | start:12 | len:12 | isnum:1 | vstart:1 | dstart:1
#define FE_VAL(e) ((u32)((e) >> 32))
#define FE_START(e) ((u32)((e) >> 20) & 0xFFFu)
#define FE_LEN(e) ((u32)((e) >> 8) & 0xFFFu)
#define FE_ISNUM(e) ((u32)((e) >> 7) & 1u)
#define FE_VSTART(e) ((u32)((e) >> 6) & 1u)
#define FE_DSTART(e) ((u32)((e) >> 5) & 1u)
#define FE_MAKE(val, start, len, isnum) (((u64)(val) << 32) | ((u64)(start) << 20) | ((u64)(len) << 8) | ((u64)(isnum) << 7))

enum { SCH_NONE = 0, SCH_VALUE_VAR, SCH_VALUE_RLE, SCH_DELTA_VAR, SCH_DELTA_RLE, SCH_DELTA_CONST };   // TagModeler.h:73

struct FieldD {
    u32 start0, len0, v0;
    u32 min_len, max_len;
    i32 min_v, max_v, min_d, max_d;
    u32 bits_num, bits_value, bits_len;
    u32 neq, lenneq, nonnum;
    u32 ham[8];                 // bit p = 1: position p equals the template in every record that has it
    u32 need[4];                // text: positions < 128 that carry a Huffman tree
    u32 slot;                   // first Huffman slot
    u32 v_runs, d_runs;
    u8 sep, is_const, is_len_const, is_num, scheme, var_stat, wiped, pooled;
};

struct TagShared {
    FieldD f[TAG_MAX_FIELDS];
    u32 scan[DSRC_WARPS + 1];
    u32 hist[TAG_NUM_HUF];
    u32 t0w[64];                                  // P2: the template field of the pass, as words
    u32 nf, mixed, min_title, max_title, status;
    u32 carry, carry2, n_text_slots, n_num_slots;
    u32 hdr_bytes;
    u32 jobs[TAG_TEXT_SLOTS + TAG_NUM_SLOTS + 1]; u32 n_jobs;
    unsigned long long total_bits;
};

__device__ __forceinline__ bool tag_is_sep(u8 c)
{
    return c == ' ' || c == '.' || c == '_' || c == ',' || c == '=' || c == ':' || c == '/' || c == '-' || c == '#' || c == 0;
}
// utils.h:163 is_num
__device__ __forceinline__ bool tag_parse_num(const u8* s, u32 len, u32* val)
{
    u32 v = 0, i;
    for (i = 0; i < len; ++i) { u8 c = s[i]; if (c < '0' || c > '9') break; v = v * 10 + (u32)(c - '0'); }
    *val = v;
    return i == len && (len == 1 || s[0] != '0');
}

// inclusive max-scan over the CTA of v (tile of DSRC_CTA values), with carry-in; two barriers
__device__ __forceinline__ u32 block_incl_max(u32 v, u32 carry_in, u32* sm)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane_id() >= (u32)o) v = max(v, t); }
    __syncthreads();
    if (lane_id() == 31) sm[warp_id()] = v;
    __syncthreads();
    u32 m = carry_in;
    for (u32 w = 0; w < warp_id(); ++w) m = max(m, sm[w]);
    return max(v, m);
}

struct BitSink {
    u32* words; u64 pos; bool write;
    __device__ __forceinline__ void put(u32 v, u32 n) { if (write) bits_or(words, pos, v, n); pos += n; }
};

// numeric token of record r for field f (TagTokenizerEncoder::StoreNumericField, TagModeler.cpp:753-874)
__device__ void emit_numeric(BitSink& bs, const FieldD& F, const Slot512* numslots, const u64* col, u32 r, u32 n_rec)
{
    const u64 e = col[r];
    const u32 v = FE_VAL(e);
    if (r == 0) {
        bs.put(v - (u32)F.min_v, F.bits_value);
        if (F.scheme == SCH_VALUE_RLE) {
            u32 sz = 1;
            if (bs.write) while (r + sz < n_rec && sz < 256 && FE_VAL(col[r + sz]) == v) ++sz;
            bs.put(sz - 1, 8);
        }
        return;
    }
    switch (F.scheme) {
    case SCH_DELTA_RLE:
        if (FE_DSTART(e)) {
            u32 dlt = v - FE_VAL(col[r - 1]);
            u32 sz = 1;
            if (bs.write) while (r + sz < n_rec && sz < 256 && (FE_VAL(col[r + sz]) - FE_VAL(col[r + sz - 1])) == dlt) ++sz;
            bs.put(dlt - (u32)F.min_d, F.bits_num); bs.put(sz - 1, 8);
        }
        break;
    case SCH_DELTA_VAR: {
        u32 s = v - FE_VAL(col[r - 1]) - (u32)F.min_d;
        if (F.var_stat) bs.put(numslots[F.slot].code[s & 511], numslots[F.slot].len[s & 511]); else bs.put(s, F.bits_num);
        break; }
    case SCH_VALUE_RLE:
        if (FE_VSTART(e)) {
            u32 sz = 1;
            if (bs.write) while (r + sz < n_rec && sz < 256 && FE_VAL(col[r + sz]) == v) ++sz;
            bs.put(v - (u32)F.min_v, F.bits_value); bs.put(sz - 1, 8);
        }
        break;
    case SCH_VALUE_VAR: {
        u32 s = v - (u32)F.min_v;
        if (F.var_stat) bs.put(numslots[F.slot].code[s & 511], numslots[F.slot].len[s & 511]); else bs.put(s, F.bits_num);
        break; }
    default: break;   // DeltaConst: nothing
    }
}

__device__ __forceinline__ u32 text_slot(const FieldD& F, u32 j)
{
    if (j >= TAG_STAT_LEN) return F.slot + __popc(F.need[0]) + __popc(F.need[1]) + __popc(F.need[2]) + __popc(F.need[3]);
    u32 s = F.slot;
    for (u32 w = 0; w < (j >> 5); ++w) s += __popc(F.need[w]);
    return s + __popc(F.need[j >> 5] & ((1u << (j & 31)) - 1));
}

__global__ void __launch_bounds__(DSRC_CTA, 4) k_tags(Workspace ws)
{
    __shared__ TagShared S;
    TagPool* pool = (TagPool*)(ws.tagpool + (u64)blockIdx.x * ws.tagpool_stride);
    const u32 tid = threadIdx.x;

    for (u32 blk = blockIdx.x; blk < ws.n_blocks; blk += gridDim.x) {
        const BlockDesc& d = ws.desc[blk];
        BlockState& st = ws.state[blk];
        __syncthreads();
        if (st.status != ST_OK) continue;
        const u8* b = ws.in + d.in_off;
        const RecArrays& R = ws.rec;
        const u32 n_rec = st.n_rec, rb = d.rec_base;
        u64* ftab = ws.ftab + d.ftab_base;
        const u32 fcap = d.rec_cap;                      // column stride of the field table
        u8* out = ws.streams + d.stream_base + stream_offset(d, 1);
        const u32 out_cap = d.stream_cap[1];
        const u32 len_bits = dsrc_bit_length((u64)(st.max_len - st.min_len));
        const u32 min_qlen = st.min_len;

        // the titles' lines on their way into L2 while thread 0 builds the template (the batch's input has left L2 since preprocessing)
        for (u32 r = tid; r < n_rec; r += DSRC_CTA) {
            const u8* t = b + R.title_off[rb + r];
            asm volatile("prefetch.global.L2 [%0];" :: "l"(t));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(t + R.title_len[rb + r]));
        }
        // ---- P0: field template from record 0 (InitializeFieldsStats, TagModeler.cpp:159-224)
        if (tid == 0) {
            const u8* t = b + R.title_off[rb]; const u32 tl = R.title_len[rb];
            u32 nf = 0, start = 0; S.status = ST_OK;
            for (u32 i = 0; i <= tl; ++i) {
                u8 c = t[i];                              // t[tl] is the line terminator: always inside the block
                if (!tag_is_sep(c) && i != tl) continue;
                if (nf >= TAG_MAX_FIELDS) { S.status = ST_UNSUPPORTED; break; }
                FieldD& F = S.f[nf];
                F.start0 = start; F.len0 = i - start; F.sep = c;
                u32 v; F.is_num = tag_parse_num(t + start, i - start, &v); F.v0 = v;
                F.min_len = F.max_len = F.len0;
                F.min_v = F.max_v = F.is_num ? (i32)v : 0; F.min_d = 0x7FFFFFFF; F.max_d = (i32)0x80000000;
                F.neq = F.lenneq = F.nonnum = 0;
                for (int k = 0; k < 8; ++k) F.ham[k] = 0xFFFFFFFFu;
                F.v_runs = F.d_runs = 0; F.scheme = SCH_NONE; F.var_stat = 0; F.wiped = 0; F.pooled = 0; F.slot = 0;
                start = i + 1; ++nf;
            }
            if (tl > 4095 || nf != d.n_fields) S.status = ST_UNSUPPORTED;
            // std::vector<Field> growth wipes num_values of the fields pushed before the last reallocation
            if (d.tag_cap != 0xFFFFFFFFu) {
                u32 cap = d.tag_cap, last = 0;
                for (u32 k = 0; k < nf; ++k) if (k == cap) { cap = cap ? cap * 2 : 1; last = k; }
                for (u32 k = 0; k < last && k < nf; ++k) S.f[k].wiped = 1;
            }
            S.nf = nf; S.mixed = 0; S.min_title = 0xFFFFFFFFu; S.max_title = 0;
            S.n_text_slots = 0; S.n_num_slots = 0; S.n_jobs = 0; S.total_bits = 0;
        }
        __syncthreads();
        if (S.status != ST_OK) { if (tid == 0) st.status = S.status; continue; }
        const u32 nf = S.nf;

        // ---- P1: tokenise every title against the template (UpdateFieldsStats walk, :248-331)
        {
            u32 mn = 0xFFFFFFFFu, mx = 0, mixed = 0, toolong = 0;
            for (u32 r = tid; r < n_rec; r += DSRC_CTA) {
                const u8* t = b + R.title_off[rb + r]; const u32 tl = R.title_len[rb + r];
                mn = min(mn, tl); mx = max(mx, tl);
                if (tl > 4095) { toolong = 1; continue; }
                // one pass over the title, four bytes per (aligned) load -- a byte load per lane touches 32 different lines per
                // instruction -- with the field's number parsed on the way (is_num, utils.h:163: the value of the leading digits; a
                // number iff all digits and no leading zero)
                u32 c = 0, start = 0, k;
                const u32 al = (u32)((uintptr_t)t & 3u);
                const u32* wp = (const u32*)(t - al);
                u32 wcur = *wp >> (8 * al), wleft = 4 - al;
                u32 v = 0, first = 0; bool digits = true;
                u32 sep = S.f[0].sep;
                for (k = 0; k <= tl && c < nf; ++k) {
                    u32 ch = 0;
                    if (k < tl) {
                        ch = wcur & 255u; wcur >>= 8;
                        if (--wleft == 0) { wcur = *++wp; wleft = 4; }       // (the word after the title's last byte is inside the block: the read follows)
                        if (k == start) first = ch;
                        if (ch != sep) {
                            if (digits) { if (ch >= '0' && ch <= '9') v = v * 10 + (ch - '0'); else digits = false; }
                            continue;
                        }
                    }
                    const u32 len = k - start;
                    if (len == 0) first = t[start];                          // is_num looks at s[0] whatever the length
                    const bool isn = digits && (len == 1 || first != '0');
                    ftab[(u64)c * fcap + r] = FE_MAKE(v, start, len, isn ? 1 : 0);
                    start = k + 1; ++c;
                    v = 0; digits = true;
                    if (c < nf) sep = S.f[c].sep;
                }
                if (c != nf || k != tl + 1) mixed = 1;
            }
            mn = warp_red_min(mn); mx = warp_red_max(mx);
            if (lane_id() == 0) { atomicMin(&S.min_title, mn); atomicMax(&S.max_title, mx); }
            if (mixed) S.mixed = 1;
            if (toolong) S.status = ST_UNSUPPORTED;
        }
        __syncthreads();
        if (S.status != ST_OK) { if (tid == 0) st.status = S.status; continue; }

        BitW hw;                                            // header writer (thread 0 only)
        if (S.mixed) {
            // ================= TagRawEncoder (TagModeler.cpp:1217-1284) =================
            for (u32 i = tid; i < 256; i += DSRC_CTA) S.hist[i] = 0;
            __syncthreads();
            for (u32 r = tid; r < n_rec; r += DSRC_CTA) {
                const u8* t = b + R.title_off[rb + r]; const u32 tl = R.title_len[rb + r];
                for (u32 k = 0; k < tl; ++k) atomicAdd(&S.hist[t[k] & 127], 1u);
            }
            if (tid == 0) {   // record 0 is counted twice (Initialize + Update)
                const u8* t = b + R.title_off[rb]; const u32 tl = R.title_len[rb];
                for (u32 k = 0; k < tl; ++k) atomicAdd(&S.hist[t[k] & 127], 1u);
            }
            __syncthreads();
            Slot256& sl = pool->text[0];
            // dense ranks of the present symbols; S.hist[128 + i] = rank
            if (tid == 0) {
                u32 ns = 0;
                for (u32 i = 0; i < 128; ++i) { if (S.hist[i]) { sl.hist[ns] = S.hist[i]; S.hist[128 + i] = ns++; } else S.hist[128 + i] = 255; }
                S.carry = ns;
            }
            __syncthreads();
            if (warp_id() == 0) {
                u32 sz = huf_build_warp(sl.hist, S.carry, &pool->work[0], sl.code, sl.len, sl.ser, sizeof(sl.ser));
                if (lane_id() == 0) sl.ser_size = sz;
            }
            __syncthreads();
            const u32 tl_bits = dsrc_bit_length((u64)(S.max_title - S.min_title));
            if (tid == 0) {
                hw.init(out, out_cap);
                hw.be32(S.min_title); hw.be32(S.max_title);
                for (u32 i = 0; i < 128; ++i) hw.bit(S.hist[128 + i] != 255);
                hw.flush();
                if (sl.ser_size == 0xFFFFFFFFu) S.status = ST_OVERFLOW;
                else for (u32 i = 0; i < sl.ser_size; ++i) hw.byte(sl.ser[i]);
                if (hw.ovf) S.status = ST_OVERFLOW;
                S.hdr_bytes = hw.pos; S.carry = 0;
            }
            __syncthreads();
            if (S.status != ST_OK) { if (tid == 0) st.status = S.status; continue; }
            // two passes: count bits, scan, write
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 1) {
                    u64 nbytes = (S.total_bits + 7) / 8;
                    if ((u64)S.hdr_bytes + nbytes + 8 > out_cap) { if (tid == 0) st.status = ST_OVERFLOW; break; }
                    for (u64 i = tid; i < nbytes + 8; i += DSRC_CTA) out[S.hdr_bytes + i] = 0;
                    __syncthreads();
                    if (tid == 0) S.carry = 0;
                    __syncthreads();
                }
                unsigned long long run = 0;
                for (u32 base = 0; base < n_rec; base += DSRC_CTA) {
                    u32 r = base + tid; u32 nb = 0;
                    const u8* t = nullptr; u32 tl = 0;
                    if (r < n_rec) {
                        t = b + R.title_off[rb + r]; tl = R.title_len[rb + r];
                        nb = tl_bits + len_bits;
                        for (u32 k = 0; k < tl; ++k) nb += sl.len[S.hist[128 + (t[k] & 127)]];
                    }
                    u32 total, ex = block_excl_sum(nb, S.scan, &total);
                    if (pass == 1 && r < n_rec) {
                        BitSink bs; bs.words = (u32*)out; bs.write = true;
                        bs.pos = (u64)S.hdr_bytes * 8 + run + ex;
                        if (tl_bits) bs.put(tl - S.min_title, tl_bits);
                        for (u32 k = 0; k < tl; ++k) { u32 s = S.hist[128 + (t[k] & 127)]; bs.put(sl.code[s], sl.len[s]); }
                        if (len_bits) bs.put(R.qua_len[rb + r] - min_qlen, len_bits);
                    }
                    run += total;
                    __syncthreads();
                }
                if (pass == 0) { if (tid == 0) S.total_bits = run; __syncthreads(); }
            }
            __syncthreads();
            if (tid == 0 && st.status == ST_OK) { st.stream_size[1] = S.hdr_bytes + (u32)((S.total_bits + 7) / 8); st.flags |= 4u; }
            continue;
        }

        // ---- P2: per-field reductions
        for (u32 f = 0; f < nf; ++f) {
            const u64* col = ftab + (u64)f * fcap;
            const u32 start0 = S.f[f].start0, len0 = S.f[f].len0;
            const u8* t0 = b + R.title_off[rb] + start0;
            u32 mnl = 0xFFFFFFFFu, mxl = 0, neq = 0, lenneq = 0, nonnum = 0;
            i32 mnv = 0x7FFFFFFF, mxv = (i32)0x80000000, mnd = 0x7FFFFFFF, mxd = (i32)0x80000000;
            u32 hamclr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            // the template field as words in shared memory (zero beyond its end): records are compared with it four bytes at a time
            __syncthreads();                             // (the previous field's comparisons are done)
            for (u32 i = tid; i < 64; i += DSRC_CTA) {
                u32 wv = 0;
                for (u32 j = 0; j < 4; ++j) if (4 * i + j < len0) wv |= (u32)t0[4 * i + j] << (8 * j);
                S.t0w[i] = wv;
            }
            __syncthreads();
            for (u32 r = tid; r < n_rec; r += DSRC_CTA) {
                const u64 e = col[r];
                const u32 len = FE_LEN(e); const u8* t = b + R.title_off[rb + r] + FE_START(e);
                mnl = min(mnl, len); mxl = max(mxl, len);
                if (len != len0) { lenneq = 1; neq = 1; }
                const u32 cmp = min(min(len, len0), 256u);
                if (cmp) {
                    const u32 al = (u32)((uintptr_t)t & 3u);
                    const u32* wp = (const u32*)(t - al);
                    u32 lo = wp[0];
#pragma unroll
                    for (int hk = 0; hk < 8; ++hk) {
                        if (32u * hk >= cmp) break;
                        u32 acc = 0;
                        for (u32 p = 32u * hk; p < cmp && p < 32u * hk + 32u; p += 4) {
                            const u32 hi = wp[(p >> 2) + 1];                 // (at most 4 bytes beyond the field: inside the block)
                            u32 dwd = __funnelshift_r(lo, hi, 8 * al) ^ S.t0w[p >> 2];
                            lo = hi;
                            if (cmp - p < 4) dwd &= (1u << (8 * (cmp - p))) - 1;
                            if (dwd) {
                                const u32 m = (((dwd & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | dwd) & 0x80808080u;      // 0x80 in every differing byte
                                acc |= (((m >> 7) & 1u) | ((m >> 14) & 2u) | ((m >> 21) & 4u) | ((m >> 28) & 8u)) << (p & 31u);
                            }
                        }
                        if (acc) { hamclr[hk] |= acc; neq = 1; }
                    }
                }
                if (!FE_ISNUM(e)) nonnum = 1;
                const i32 v = (i32)FE_VAL(e);
                mnv = min(mnv, v); mxv = max(mxv, v);
                if (r >= 1) { i32 dl = (i32)(FE_VAL(e) - FE_VAL(col[r - 1])); mnd = min(mnd, dl); mxd = max(mxd, dl); }
            }
            mnl = warp_red_min(mnl); mxl = warp_red_max(mxl);
            mnv = warp_red_imin(mnv); mxv = warp_red_imax(mxv); mnd = warp_red_imin(mnd); mxd = warp_red_imax(mxd);
            FieldD& F = S.f[f];
            if (lane_id() == 0) {
                atomicMin(&F.min_len, mnl); atomicMax(&F.max_len, mxl);
                if (n_rec > 0) { atomicMin(&F.min_v, mnv); atomicMax(&F.max_v, mxv); }
                if (mnd <= mxd) { atomicMin(&F.min_d, mnd); atomicMax(&F.max_d, mxd); }
            }
            if (neq) F.neq = 1;
            if (lenneq) F.lenneq = 1;
            if (nonnum) F.nonnum = 1;
#pragma unroll
            for (int k = 0; k < 8; ++k) if (hamclr[k]) atomicAnd(&F.ham[k], ~hamclr[k]);
        }
        __syncthreads();
        if (tid < nf) {
            FieldD& F = S.f[tid];
            F.is_const = !F.neq; F.is_len_const = !F.lenneq; F.is_num = F.is_num && !F.nonnum;
            if (n_rec < 2) { F.min_d = 1 << 30; F.max_d = -(1 << 30); }
        }
        __syncthreads();

        // ---- P3: run structure of numeric fields (UpdateNumericField RLE bookkeeping, :351-370, :405-424)
        for (u32 f = 0; f < nf; ++f) {
            if (!S.f[f].is_num || S.f[f].is_const) continue;
            u64* col = ftab + (u64)f * fcap;
            if (tid == 0) { S.carry = 0; S.carry2 = 0; }
            u32 vcount = 0, dcount = 0;
            __syncthreads();
            for (u32 base = 0; base < n_rec; base += DSRC_CTA) {
                const u32 r = base + tid; const bool in = r < n_rec;
                u64 e = in ? col[r] : 0;
                const u32 v = FE_VAL(e);
                const u32 pv = (in && r >= 1) ? FE_VAL(col[r - 1]) : 0, ppv = (in && r >= 2) ? FE_VAL(col[r - 2]) : 0;
                const bool vhead = in && (r == 0 || v != pv);
                const bool dhead = in && r >= 1 && (r == 1 || (v - pv) != (pv - ppv));
                const u32 vh = block_incl_max(vhead ? r : 0u, S.carry, S.scan);
                const u32 dh = block_incl_max(dhead ? r : 0u, S.carry2, S.scan);
                const bool vs = in && (((r - vh) & 255u) == 0);
                const bool ds = in && r >= 1 && (((r - dh) & 255u) == 0);
                if (in) { col[r] = (e & ~(u64)0x60) | ((u64)vs << 6) | ((u64)ds << 5); }
                vcount += vs; dcount += ds;
                __syncthreads();
                if (tid == DSRC_CTA - 1) { S.carry = vh; S.carry2 = dh; }   // inclusive value of the last lane = running max
                if (in && r == n_rec - 1) { S.f[f].v_runs = (r - vh) & 255u; S.f[f].d_runs = (r >= 1) ? ((r - dh) & 255u) : 0; }
                __syncthreads();
            }
            vcount = warp_red_sum(vcount); dcount = warp_red_sum(dcount);
            __syncthreads();
            if (tid == 0) {
                // v_runs/d_runs currently hold the position of the last record inside its sub-run (== final cur_len)
                u32 lastv = S.f[f].v_runs, lastd = S.f[f].d_runs;
                S.f[f].v_runs = (lastv > 0); S.f[f].d_runs = (lastd > 0);
                S.hist[0] = 0; S.hist[1] = 0;
            }
            __syncthreads();
            if (lane_id() == 0) { atomicAdd(&S.hist[0], vcount); atomicAdd(&S.hist[1], dcount); }
            __syncthreads();
            if (tid == 0) {
                S.f[f].v_runs += S.hist[0] - 1;                                    // run_len == sub-runs - 1 (+1 if the last one is open)
                S.f[f].d_runs = (n_rec >= 2) ? S.f[f].d_runs + S.hist[1] - 1 : 0;
            }
            __syncthreads();
        }

        // ---- FinalizeFieldsStats (:461-551) + Huffman slot assignment
        if (tid == 0) {
            u32 status = ST_OK;
            for (u32 f = 0; f < nf; ++f) {
                FieldD& F = S.f[f];
                if (F.is_const) continue;
                if (!F.is_num) {
                    F.bits_len = dsrc_bit_length((u64)(F.max_len - F.min_len));
                    if (F.max_len > 255 || F.len0 > 255) { status = ST_UNSUPPORTED; break; }
                    const u32 lim = min(F.max_len, (u32)TAG_STAT_LEN);
                    u32 cnt = 0;
                    for (u32 w = 0; w < 4; ++w) {
                        u32 m = 0;
                        for (u32 j = w * 32; j < w * 32 + 32 && j < lim; ++j)
                            if (j >= F.len0 || !((F.ham[j >> 5] >> (j & 31)) & 1u)) m |= 1u << (j & 31);
                        F.need[w] = m; cnt += __popc(m);
                    }
                    F.pooled = F.max_len >= TAG_STAT_LEN;
                    F.slot = S.n_text_slots;
                    S.n_text_slots += cnt + F.pooled;
                    if (S.n_text_slots > TAG_TEXT_SLOTS) { status = ST_UNSUPPORTED; break; }
                    continue;
                }
                const i32 dv = (i32)((u32)F.max_v - (u32)F.min_v), dd = (i32)((u32)F.max_d - (u32)F.min_d);
                bool is_delta; i32 diff;
                if (dv < dd) { is_delta = false; diff = dv; } else { is_delta = true; diff = dd; }
                const bool try_val = __fdiv_rn((float)n_rec, (float)F.v_runs) > 1.25f;
                bool delta_const = false, try_delta = false;
                if (is_delta) {
                    delta_const = diff == 0;
                    if (!delta_const) try_delta = __fdiv_rn((float)n_rec, (float)F.d_runs) > 1.25f;
                }
                if (is_delta && delta_const) F.scheme = SCH_DELTA_CONST;
                else if (is_delta && try_delta) F.scheme = SCH_DELTA_RLE;
                else if (try_val) F.scheme = SCH_VALUE_RLE;
                else if (is_delta) { F.scheme = SCH_DELTA_VAR; F.var_stat = ((u32)dd + 1u) <= TAG_NUM_HUF && n_rec >= 2; }
                else { F.scheme = SCH_VALUE_VAR; F.var_stat = ((u32)dv + 1u) <= TAG_NUM_HUF; }
                F.bits_num = dsrc_bit_length((u64)(i64)diff);
                F.bits_value = dsrc_bit_length((u64)(i64)dv);
                if (F.bits_value > 32 || (F.bits_num > 32 && n_rec > 1)) { status = ST_UNSUPPORTED; break; }
                if (F.var_stat) {
                    F.slot = S.n_num_slots++;
                    if (S.n_num_slots > TAG_NUM_SLOTS) { status = ST_UNSUPPORTED; break; }
                }
            }
            S.status = status;
        }
        __syncthreads();
        if (S.status != ST_OK) { if (tid == 0) st.status = S.status; continue; }

        // ---- P4: histograms
        for (u32 i = tid; i < S.n_text_slots * 256; i += DSRC_CTA) pool->text[i >> 8].hist[i & 255] = 0;
        __syncthreads();
        for (u32 f = 0; f < nf; ++f) {
            const FieldD& F = S.f[f];
            if (F.is_const) continue;
            const u64* col = ftab + (u64)f * fcap;
            if (F.is_num) {
                if (!F.var_stat) continue;
                for (u32 i = tid; i < TAG_NUM_HUF; i += DSRC_CTA) S.hist[i] = 0;
                __syncthreads();
                if (F.scheme == SCH_VALUE_VAR) {
                    for (u32 r = tid; r < n_rec; r += DSRC_CTA) atomicAdd(&S.hist[(FE_VAL(col[r]) - (u32)F.min_v) & 511], 1u);
                    if (tid == 0 && !F.wiped) atomicAdd(&S.hist[(F.v0 - (u32)F.min_v) & 511], 1u);   // Q2 double count, unless Q1 wiped it
                } else {
                    for (u32 r = tid + 1; r < n_rec; r += DSRC_CTA) atomicAdd(&S.hist[(FE_VAL(col[r]) - FE_VAL(col[r - 1]) - (u32)F.min_d) & 511], 1u);
                }
                __syncthreads();
                for (u32 i = tid; i < TAG_NUM_HUF; i += DSRC_CTA) pool->num[F.slot].hist[i] = S.hist[i];
                if (tid == 0) S.jobs[S.n_jobs++] = 0x80000000u | (F.slot << 16) | (F.scheme == SCH_VALUE_VAR ? ((u32)F.max_v - (u32)F.min_v + 1u) : ((u32)F.max_d - (u32)F.min_d + 1u));
                __syncthreads();
            } else {
                for (u32 r = tid; r < n_rec; r += DSRC_CTA) {
                    const u64 e = col[r]; const u32 len = FE_LEN(e); const u8* t = b + R.title_off[rb + r] + FE_START(e);
                    for (u32 j = 0; j < len; ++j) {
                        if (j < TAG_STAT_LEN) { if (!((F.need[j >> 5] >> (j & 31)) & 1u)) continue; }
                        atomicAdd(&pool->text[text_slot(F, j)].hist[t[j]], 1u);
                    }
                }
                if (tid == 0) {
                    u32 cnt = __popc(F.need[0]) + __popc(F.need[1]) + __popc(F.need[2]) + __popc(F.need[3]) + F.pooled;
                    for (u32 k = 0; k < cnt; ++k) S.jobs[S.n_jobs++] = F.slot + k;
                }
            }
        }
        __syncthreads();

        // ---- P5: one warp per tree
        for (u32 j = warp_id(); j < S.n_jobs; j += DSRC_WARPS) {
            const u32 job = S.jobs[j];
            if (job & 0x80000000u) {
                Slot512& sl = pool->num[(job >> 16) & 0x7FFF];
                u32 sz = huf_build_warp(sl.hist, job & 0xFFFF, &pool->work[warp_id()], sl.code, sl.len, sl.ser, sizeof(sl.ser));
                if (lane_id() == 0) sl.ser_size = sz;
            } else {
                Slot256& sl = pool->text[job];
                u32 sz = huf_build_warp(sl.hist, 256, &pool->work[warp_id()], sl.code, sl.len, sl.ser, sizeof(sl.ser));
                if (lane_id() == 0) sl.ser_size = sz;
            }
        }
        __syncthreads();

        // ---- P6: field dictionary (TagTokenizerEncoder::StoreFields, :569-693)
        if (tid == 0) {
            hw.init(out, out_cap);
            hw.byte((u8)nf);
            const u8* t0 = b + R.title_off[rb];
            bool bad = false;
            for (u32 f = 0; f < nf; ++f) {
                const FieldD& F = S.f[f];
                hw.byte(F.sep); hw.byte(F.is_const);
                if (F.is_const) { hw.be32(F.len0); for (u32 i = 0; i < F.len0; ++i) hw.byte(t0[F.start0 + i]); continue; }
                hw.byte(F.is_num);
                if (F.is_num) {
                    hw.byte(F.scheme); hw.be32((u32)F.min_v); hw.be32((u32)F.max_v);
                    if (F.scheme >= SCH_DELTA_VAR) { hw.be32((u32)F.min_d); hw.be32((u32)F.max_d); }
                    if (F.scheme == SCH_DELTA_VAR || F.scheme == SCH_VALUE_VAR) {
                        hw.byte(F.var_stat);
                        if (F.var_stat) {
                            const Slot512& sl = pool->num[F.slot];
                            if (sl.ser_size == 0xFFFFFFFFu) { bad = true; break; }
                            for (u32 i = 0; i < sl.ser_size; ++i) hw.byte(sl.ser[i]);
                        }
                    }
                    continue;
                }
                hw.byte(F.is_len_const); hw.be32(F.len0); hw.be32(F.max_len); hw.be32(F.min_len);
                for (u32 i = 0; i < F.len0; ++i) hw.byte(t0[F.start0 + i]);
                for (u32 j = 0; j < F.len0; ++j) hw.bit((F.ham[j >> 5] >> (j & 31)) & 1u);
                hw.flush();
                u32 cnt = __popc(F.need[0]) + __popc(F.need[1]) + __popc(F.need[2]) + __popc(F.need[3]) + F.pooled;
                for (u32 k = 0; k < cnt && !bad; ++k) {
                    const Slot256& sl = pool->text[F.slot + k];
                    if (sl.ser_size == 0xFFFFFFFFu) { bad = true; break; }
                    for (u32 i = 0; i < sl.ser_size; ++i) hw.byte(sl.ser[i]);
                }
                if (bad) break;
            }
            if (bad || hw.ovf) S.status = ST_OVERFLOW;
            S.hdr_bytes = hw.pos;
        }
        __syncthreads();
        if (S.status != ST_OK) { if (tid == 0) st.status = S.status; continue; }

        // ---- P7: records (EncodeNextFields :695-751 + the read-length bits of StoreTags, BlockCompressor.cpp:481-484)
        bool failed = false;
        for (int pass = 0; pass < 2 && !failed; ++pass) {
            if (pass == 1) {
                const u64 nbytes = (S.total_bits + 7) / 8;
                if ((u64)S.hdr_bytes + nbytes + 8 > out_cap) { if (tid == 0) st.status = ST_OVERFLOW; failed = true; break; }
                for (u64 i = tid; i < nbytes + 8; i += DSRC_CTA) out[S.hdr_bytes + i] = 0;
                __syncthreads();
            }
            unsigned long long run = 0;
            for (u32 base = 0; base < n_rec; base += DSRC_CTA) {
                const u32 r = base + tid;
                // pass 0 counts, pass 1 needs the exclusive offset first: count again (cheap) then emit
                u32 nb = 0;
                if (r < n_rec) {
                    BitSink cs; cs.words = nullptr; cs.write = false; cs.pos = 0;
                    for (u32 f = 0; f < nf; ++f) {
                        const FieldD& F = S.f[f];
                        if (F.is_const) continue;
                        const u64* col = ftab + (u64)f * fcap;
                        if (F.is_num) { emit_numeric(cs, F, pool->num, col, r, n_rec); continue; }
                        const u64 e = col[r]; const u32 len = FE_LEN(e); const u8* t = b + R.title_off[rb + r] + FE_START(e);
                        if (!F.is_len_const) cs.pos += F.bits_len;
                        for (u32 j = 0; j < len; ++j)
                            if (j >= F.len0 || !((F.ham[j >> 5] >> (j & 31)) & 1u)) cs.pos += pool->text[text_slot(F, j)].len[t[j]];
                    }
                    nb = (u32)cs.pos + len_bits;
                }
                u32 total, ex = block_excl_sum(nb, S.scan, &total);
                if (pass == 1 && r < n_rec) {
                    BitSink bs; bs.words = (u32*)out; bs.write = true; bs.pos = (u64)S.hdr_bytes * 8 + run + ex;
                    for (u32 f = 0; f < nf; ++f) {
                        const FieldD& F = S.f[f];
                        if (F.is_const) continue;
                        const u64* col = ftab + (u64)f * fcap;
                        if (F.is_num) { emit_numeric(bs, F, pool->num, col, r, n_rec); continue; }
                        const u64 e = col[r]; const u32 len = FE_LEN(e); const u8* t = b + R.title_off[rb + r] + FE_START(e);
                        if (!F.is_len_const) bs.put(len - F.min_len, F.bits_len);
                        for (u32 j = 0; j < len; ++j)
                            if (j >= F.len0 || !((F.ham[j >> 5] >> (j & 31)) & 1u)) { const Slot256& sl = pool->text[text_slot(F, j)]; bs.put(sl.code[t[j]], sl.len[t[j]]); }
                    }
                    if (len_bits) bs.put(R.qua_len[rb + r] - min_qlen, len_bits);
                }
                run += total;
                __syncthreads();
            }
            if (pass == 0) { if (tid == 0) S.total_bits = run; __syncthreads(); }
        }
        __syncthreads();
        if (tid == 0 && !failed && st.status == ST_OK) st.stream_size[1] = S.hdr_bytes + (u32)((S.total_bits + 7) / 8);
    }
}

void launch_tags(const Workspace& ws, cudaStream_t s, u32 ctas)
{
    u32 grid = ctas < ws.n_blocks ? ctas : ws.n_blocks;
    k_tags<<<grid ? grid : 1, DSRC_CTA, 0, s>>>(ws);
}
