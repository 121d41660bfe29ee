/*
 * TEST INFRASTRUCTURE ONLY -- see dsrc_oracle.h. Plain-C restatement of the DSRC 2.02 block codec.
 * Parity: pinned by execution against oracle/_ref (the unmodified reference), no golden vectors upstream.
 * Citations are file:line relative to /root/reference/.
 *
 * Scope: lossless, non-colour-space, no -c CRC, no -f field filtering.
 */
#include "dsrc_oracle.h"
#include <stdlib.h>
#include <string.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint64_t u64;

#define ERR_CAP (-1)
#define ERR_FORMAT (-2)
#define ERR_UNSUPPORTED (-3)

/* ------------------------------------------------------------------------------------------------
 * utils  (src/utils.h:138-185)
 * ---------------------------------------------------------------------------------------------- */
static u32 bit_length64(u64 x) /* utils.h:177: i for x < 2^i (i<32), else 64 */
{
    for (u32 i = 0; i < 32; ++i)
        if (x < (1ull << i))
            return i;
    return 64;
}
static u32 ilog2_floor(u32 x) /* utils.h:138 int_log(x, 2) */
{
    u32 r = 0;
    for (u64 t = 2; t <= x; t *= 2)
        ++r;
    return r;
}
static int parse_num(const u8* s, u32 len, u32* val) /* utils.h:163 is_num */
{
    u32 v = 0, i;
    for (i = 0; i < len; ++i) {
        if (s[i] < '0' || s[i] > '9')
            break;
        v = v * 10 + (u32)(s[i] - '0');
    }
    *val = v;
    return i == len && (len == 1 || s[0] != '0');
}
static u32 to_num(const u8* s, u32 len) /* utils.h:154 */
{
    u32 r = 0;
    for (u32 i = 0; i < len; ++i)
        r = r * 10 + (u32)(s[i] - '0');
    return r;
}
static u32 num_to_str(u8* s, u32 value) /* utils.h:52 to_string */
{
    u8 tmp[12];
    u32 n = 0;
    if (value == 0) {
        s[0] = '0';
        return 1;
    }
    while (value) {
        tmp[n++] = (u8)('0' + value % 10);
        value /= 10;
    }
    for (u32 i = 0; i < n; ++i)
        s[i] = tmp[n - 1 - i];
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * bit writer / reader  (src/BitMemory.h). The reference keeps a lazily flushed 32-bit accumulator;
 * because every byte-level put is preceded by a flush, the byte stream equals a plain MSB-first
 * bit concatenation padded to a byte at each flush. That is what is implemented.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    u8* buf;
    u64 cap, pos;
    u64 acc;
    u32 nacc;
    int overflow;
} bitw_t;

static void bw_init(bitw_t* w, u8* buf, u64 cap)
{
    w->buf = buf; w->cap = cap; w->pos = 0; w->acc = 0; w->nacc = 0; w->overflow = 0;
}
static void bw_byte_raw(bitw_t* w, u8 b)
{
    if (w->pos < w->cap)
        w->buf[w->pos] = b;
    else
        w->overflow = 1;
    w->pos++;
}
static void bw_bits(bitw_t* w, u32 v, u32 n) /* BitMemory.h:318 PutBits; n==0 is a no-op */
{
    if (n == 0)
        return;
    if (n < 32)
        v &= (1u << n) - 1;
    w->acc = (w->acc << n) | v;
    w->nacc += n;
    while (w->nacc >= 8) {
        bw_byte_raw(w, (u8)(w->acc >> (w->nacc - 8)));
        w->nacc -= 8;
    }
}
static void bw_bit(bitw_t* w, u32 b) { bw_bits(w, b & 1, 1); }
static void bw_flush(bitw_t* w) /* BitMemory.h:394 FlushPartialWordBuffer */
{
    if (w->nacc)
        bw_bits(w, 0, 8 - w->nacc);
}
static void bw_byte(bitw_t* w, u8 b) { bw_byte_raw(w, b); }                     /* :340 */
static void bw_u32(bitw_t* w, u32 v)                                             /* :378 PutWord, big endian */
{
    bw_byte_raw(w, (u8)(v >> 24)); bw_byte_raw(w, (u8)(v >> 16)); bw_byte_raw(w, (u8)(v >> 8)); bw_byte_raw(w, (u8)v);
}
static void bw_bytes(bitw_t* w, const u8* p, u32 n) { for (u32 i = 0; i < n; ++i) bw_byte_raw(w, p[i]); }

typedef struct {
    const u8* buf;
    u64 size, pos;
    u32 cur, ncur; /* 8-bit window, BitMemory.h:198 */
    int overrun;
} bitr_t;

static void br_init(bitr_t* r, const u8* buf, u64 size) { r->buf = buf; r->size = size; r->pos = 0; r->cur = 0; r->ncur = 0; r->overrun = 0; }
static u8 br_byte(bitr_t* r)
{
    if (r->pos >= r->size) { r->overrun = 1; r->pos++; return 0; }
    return r->buf[r->pos++];
}
static u32 br_bit(bitr_t* r) /* :55 */
{
    if (r->ncur == 0) { r->cur = br_byte(r); r->ncur = 8; }
    return (r->cur >> (--r->ncur)) & 1;
}
static u32 br_bits(bitr_t* r, u32 n) /* :89 */
{
    u32 v = 0;
    for (u32 i = 0; i < n; ++i)
        v = (v << 1) | br_bit(r);
    return v;
}
static void br_flush(bitr_t* r) { r->ncur = 0; } /* :171 FlushInputWordBuffer */
static u32 br_u32(bitr_t* r) { u32 v = br_byte(r); v = (v << 8) | br_byte(r); v = (v << 8) | br_byte(r); return (v << 8) | br_byte(r); }

/* ------------------------------------------------------------------------------------------------
 * Huffman  (src/huffman.cpp:94-221, src/huffman.h:67-70,109-122)
 * The reference builds with std::*_heap over a strict total order (frequency asc, symbol asc), so the
 * sequence of extracted minima does not depend on heap layout; we simply extract minima.
 * ---------------------------------------------------------------------------------------------- */
#define HUF_MAX 512
typedef struct {
    u32 n;            /* n_symbols after the "<2 -> 2" rule (huffman.cpp:101) */
    u32 root;         /* root_id */
    u32 min_len, bits_per_id;
    u32 code[2 * HUF_MAX], len[2 * HUF_MAX];
    i32 left[2 * HUF_MAX], right[2 * HUF_MAX];
} huf_t;

static void huf_build(huf_t* h, const u32* freq_in, u32 n_in) /* Restart+Insert*+Complete */
{
    static u32 hs[HUF_MAX + 2], hf[HUF_MAX + 2];
    u32 n = n_in, cnt, i;
    for (i = 0; i < n_in; ++i) { hs[i] = i; hf[i] = freq_in[i]; }
    if (n < 2) { /* huffman.cpp:101: heap[1] is a default Frequency{0,0} when capacity allows */
        for (i = n_in; i < 2; ++i) { hs[i] = 0; hf[i] = 0; }
        n = 2;
    }
    h->n = n;
    for (i = 0; i < 2 * n; ++i) { h->code[i] = 0; h->len[i] = 0; h->left[i] = -1; h->right[i] = -1; }
    cnt = n;
#define HUF_MIN_IDX(res)                                                                        \
    do { u32 m_ = 0; for (u32 k_ = 1; k_ < cnt; ++k_)                                           \
        if (hf[k_] < hf[m_] || (hf[k_] == hf[m_] && hs[k_] < hs[m_])) m_ = k_; (res) = m_; } while (0)
#define HUF_POP(idx) do { hs[idx] = hs[cnt - 1]; hf[idx] = hf[cnt - 1]; --cnt; } while (0)
    {
        u32 m;
        HUF_MIN_IDX(m);
        if (cnt == 2 && hf[m] == 0) { /* huffman.cpp:128-133: both frequencies are forced to >= 1 IN PLACE, without re-heapifying,
                                       * so the element that was the minimum stays heap[0] and becomes the LEFT child even when the
                                       * other symbol is smaller and also has frequency 1. Keeping the minimum's key at 0 here
                                       * reproduces that order (with two symbols the merged frequency is never used). */
            if (hf[1 - m] == 0) hf[1 - m] = 1;
        } else {
            for (;;) {           /* :136 drop zero-frequency symbols while more than two remain */
                HUF_MIN_IDX(m);
                if (!(cnt > 2 && hf[m] == 0)) break;
                HUF_POP(m);
            }
        }
    }
    {
        u32 present = cnt;
        for (i = 0; i + 1 < present; ++i) { /* :146-158 */
            u32 m, ls, lf, rs, rf;
            HUF_MIN_IDX(m); ls = hs[m]; lf = hf[m]; HUF_POP(m);
            HUF_MIN_IDX(m); rs = hs[m]; rf = hf[m]; HUF_POP(m);
            hs[cnt] = n + i; hf[cnt] = lf + rf; ++cnt;
            h->left[n + i] = (i32)ls; h->right[n + i] = (i32)rs;
        }
        for (i32 k = (i32)(n + present) - 2; k >= (i32)n; --k) { /* :161-168 */
            h->len[h->left[k]] = h->len[k] + 1;  h->code[h->left[k]] = h->code[k] << 1;
            h->len[h->right[k]] = h->len[k] + 1; h->code[h->right[k]] = (h->code[k] << 1) | 1;
        }
        h->root = n + present - 2;
    }
#undef HUF_MIN_IDX
#undef HUF_POP
}

static void huf_store_node(const huf_t* h, bitw_t* w, i32 id) /* huffman.h:109 EncodeProcess */
{
    if (h->left[id] == -1) { bw_bit(w, 1); bw_bits(w, (u32)id, h->bits_per_id); }
    else { bw_bit(w, 0); huf_store_node(h, w, h->left[id]); huf_store_node(h, w, h->right[id]); }
}
static void huf_store(huf_t* h, bitw_t* w) /* huffman.cpp:177 StoreTree */
{
    u64 p0; u32 sz;
    bw_flush(w);
    p0 = w->pos;
    bw_u32(w, 0);
    h->bits_per_id = ilog2_floor(h->n) + ((h->n & (h->n - 1)) ? 1 : 0);
    h->min_len = h->n;
    for (u32 i = 0; i < h->n; ++i)
        if (h->len[i] < h->min_len && h->len[i] > 0) h->min_len = h->len[i];
    bw_u32(w, h->root); bw_u32(w, h->n); bw_byte(w, (u8)h->min_len);
    huf_store_node(h, w, (i32)h->root);
    bw_flush(w);
    sz = (u32)(w->pos - p0);
    if (p0 + 4 <= w->cap) { w->buf[p0] = (u8)(sz >> 24); w->buf[p0 + 1] = (u8)(sz >> 16); w->buf[p0 + 2] = (u8)(sz >> 8); w->buf[p0 + 3] = (u8)sz; }
}
static void huf_put(const huf_t* h, bitw_t* w, u32 sym) { bw_bits(w, h->code[sym], h->len[sym]); }

/* decoder side: rebuild the code tree from the pre-order serialisation (huffman.cpp:225-260).
 * Node numbering is private to the decoder; only the (bit path -> leaf id) map matters. */
typedef struct { u32 n_nodes; i32 child[2 * HUF_MAX + 2][2]; i32 leaf[2 * HUF_MAX + 2]; } hufd_t;
static i32 hufd_load_node(hufd_t* d, bitr_t* r, u32 bits_per_id, u32 depth)
{
    i32 id = (i32)d->n_nodes++;
    if (d->n_nodes > 2 * HUF_MAX || depth > 600) { r->overrun = 1; d->n_nodes--; return 0; }
    if (br_bit(r)) { d->leaf[id] = (i32)br_bits(r, bits_per_id); d->child[id][0] = d->child[id][1] = -1; }
    else { d->leaf[id] = -1; d->child[id][0] = hufd_load_node(d, r, bits_per_id, depth + 1); d->child[id][1] = hufd_load_node(d, r, bits_per_id, depth + 1); }
    return id;
}
static void hufd_load(hufd_t* d, bitr_t* r)
{
    u32 n, bpi;
    br_flush(r);
    (void)br_u32(r); (void)br_u32(r); n = br_u32(r); (void)br_byte(r);
    bpi = ilog2_floor(n) + ((n & (n - 1)) ? 1 : 0);
    d->n_nodes = 0;
    hufd_load_node(d, r, bpi, 0);
    br_flush(r);
}
static u32 hufd_get(const hufd_t* d, bitr_t* r)
{
    i32 id = 0;
    while (d->leaf[id] < 0 && !r->overrun)
        id = d->child[id][br_bit(r)];
    return (u32)d->leaf[id];
}

/* ------------------------------------------------------------------------------------------------
 * range coder + adaptive model  (src/RangeCoder.h:51-134, src/SymbolCoderRC.h:24-93)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { u64 low; u32 range; bitw_t* w; } rce_t;
static void rce_start(rce_t* e, bitw_t* w) { e->low = 0; e->range = 0xFFFFFFFFu; e->w = w; }
static void rce_encode(rce_t* e, u32 freq, u32 cum, u32 tot)
{
    e->range /= tot;
    e->low += (u32)(e->range * cum);
    e->range *= freq;
    while (e->range <= 0x00FFFFFFu) {
        if ((e->low ^ (e->low + e->range)) & 0xFF00000000000000ull) {
            u32 r = (u32)e->low;
            e->range = (r | 0x00FFFFFFu) - r;
        }
        bw_byte(e->w, (u8)(e->low >> 56));
        e->low <<= 8; e->range <<= 8;
    }
}
static void rce_end(rce_t* e) { for (int i = 0; i < 8; ++i) { bw_byte(e->w, (u8)(e->low >> 56)); e->low <<= 8; } }

typedef struct { u64 low, buffer; u32 range; bitr_t* r; } rcd_t;
static void rcd_start(rcd_t* d, bitr_t* r)
{
    d->r = r; d->buffer = 0;
    for (u32 i = 1; i <= 8; ++i) d->buffer |= (u64)br_byte(r) << (64 - i * 8);
    d->low = 0; d->range = 0xFFFFFFFFu;
}
static u32 rcd_cum(rcd_t* d, u32 tot) { d->range /= tot; return (u32)(d->buffer / d->range); }
static void rcd_update(rcd_t* d, u32 freq, u32 lo)
{
    u32 r = lo * d->range;
    d->buffer -= r; d->low += r; d->range *= freq;
    while (d->range <= 0x00FFFFFFu) {
        if ((d->low ^ (d->low + d->range)) & 0xFF00000000000000ull) {
            u32 q = (u32)d->low;
            d->range = (q | 0x00FFFFFFu) - q;
        }
        d->buffer = (d->buffer << 8) + br_byte(d->r);
        d->low <<= 8; d->range <<= 8;
    }
}

static u32 model_total(u16* st, u32 n) /* SymbolCoderRC.h:75 Accumulate (+Rescale :69) */
{
    u32 acc = 0, i;
    for (i = 0; i < n; ++i) acc += st[i];
    if (acc >= (1u << 16) - n * 2) {
        acc = 0;
        for (i = 0; i < n; ++i) { st[i] = (u16)(st[i] - (st[i] >> 1)); acc += st[i]; }
    }
    return acc;
}
static void model_encode(u16* st, u32 n, rce_t* e, u32 sym) /* :35 */
{
    u32 tot = model_total(st, n), lo = 0;
    for (u32 i = 0; i < sym; ++i) lo += st[i];
    rce_encode(e, st[sym], lo, tot);
    st[sym] = (u16)(st[sym] + 2);
}
static u32 model_decode(u16* st, u32 n, rcd_t* d) /* :50 */
{
    u32 tot = model_total(st, n), cul = rcd_cum(d, tot), idx = 0, hi = 0;
    for (idx = 0, hi = 0; (hi += st[idx]) <= cul; ++idx)
        if (idx + 1 >= n) { break; }
    hi -= st[idx];
    rcd_update(d, st[idx], hi);
    st[idx] = (u16)(st[idx] + 2);
    return idx;
}

/* ------------------------------------------------------------------------------------------------
 * records
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    u32 title, seq, qua;                 /* offsets into the work buffer */
    u16 title_len, seq_len, qua_len, trunc_len; /* uint16 as in src/Fastq.h:37-40 */
} rec_t;

typedef struct { u32 count; u32 freq[20]; u8 rank[20]; } dna_stats_t;           /* src/Stats.h:44 */
typedef struct { u32 count; u32 freq[256]; u8 rank[256]; u32 min_len, max_len, raw_len, th_len, rle_len; } qua_stats_t; /* :69 */

struct dsrc_oracle {
    u32 qoff; int plus_rep; u32 dna_order, qua_order;
    int calc_crc;                        /* CompressionSettings::calculateCrc32 (-c) */
    u32 last_crc[3]; int last_crc_ok;    /* checksums read / verified by the last dsrc_oracle_read */
    u32 tag_cap;                         /* capacity of TagStats::fields (Q1) */
    /* per-block scratch */
    u8* work; u64 work_cap;
    rec_t* recs; u64 recs_cap;
    u16* model; u64 model_cap;
};

dsrc_oracle_t* dsrc_oracle_create(u32 qoff, int plus_rep, u32 dna_order, u32 qua_order)
{
    dsrc_oracle_t* o = (dsrc_oracle_t*)calloc(1, sizeof(*o));
    o->qoff = qoff; o->plus_rep = plus_rep; o->dna_order = dna_order; o->qua_order = qua_order;
    return o;
}
void dsrc_oracle_destroy(dsrc_oracle_t* o)
{
    if (!o) return;
    free(o->work); free(o->recs); free(o->model); free(o);
}
u32 dsrc_oracle_tag_capacity(const dsrc_oracle_t* o) { return o->tag_cap; }
void dsrc_oracle_set_crc(dsrc_oracle_t* o, int on) { o->calc_crc = on != 0; }
int dsrc_oracle_last_crc_ok(const dsrc_oracle_t* o) { return o->last_crc_ok; }

/* core::Crc32Hasher (src/Crc32.h:24-92): reflected CRC-32, polynomial 0xEDB88320, seed ~0, final xor ~0 */
static u32 crc32_update(u32 crc, const u8* p, u32 n)
{
    static u32 table[256]; static int ready = 0;
    if (!ready) {
        for (u32 i = 0; i < 256; ++i) { u32 h = i; for (int j = 0; j < 8; ++j) h = (h & 1) ? 0xEDB88320u ^ (h >> 1) : h >> 1; table[i] = h; }
        ready = 1;
    }
    for (u32 i = 0; i < n; ++i) crc = (crc >> 8) ^ table[(p[i] ^ crc) & 0xFF];
    return crc;
}

static u16* get_model(dsrc_oracle_t* o, u64 entries)
{
    if (o->model_cap < entries) { free(o->model); o->model = (u16*)malloc(entries * 2); o->model_cap = entries; }
    for (u64 i = 0; i < entries; ++i) o->model[i] = 1; /* Clear(): DnaModelerRCO.h:111, QualityEncoder.h:45 */
    return o->model;
}

/* FastqParser::SkipLine (src/FastqParser.h:93-115) */
static u32 skip_line(const u8* m, u64 size, u64* pos, u64* skipped)
{
    u32 len = 0;
    for (;;) {
        if (*pos == size) break;
        u8 c = m[(*pos)++];
        if (c != '\n' && c != '\r') { len++; }
        else {
            if (c == '\r' && *pos < size && m[*pos] == '\n') { (*pos)++; (*skipped)++; }
            break;
        }
    }
    return len;
}

/* FastqParser::ParseFrom + ReadNextRecord (src/FastqParser.cpp:140-164, FastqParser.h:40-60) */
static i32 parse_records(dsrc_oracle_t* o, u64 size, u64* n_out, u64* chunk_size, u64* raw4)
{
    const u8* m = o->work;
    u64 pos = 0, skipped = 0, n = 0;
    raw4[0] = raw4[1] = raw4[2] = raw4[3] = 0;
    while (pos < size) {
        rec_t r; u32 l, plus;
        if (pos == size) break;
        r.title = (u32)pos; l = skip_line(m, size, &pos, &skipped); if (l > 65535) return ERR_UNSUPPORTED; r.title_len = (u16)l;
        if (r.title_len == 0 || m[r.title] != '@') break;
        r.seq = (u32)pos; l = skip_line(m, size, &pos, &skipped); if (l > 65535) return ERR_UNSUPPORTED; r.seq_len = (u16)l;
        plus = skip_line(m, size, &pos, &skipped);
        r.qua = (u32)pos; l = skip_line(m, size, &pos, &skipped); if (l > 65535) return ERR_UNSUPPORTED; r.qua_len = (u16)l;
        r.trunc_len = 0;
        if (n + 1 > o->recs_cap) { o->recs_cap = o->recs_cap ? o->recs_cap * 2 : 8192; o->recs = (rec_t*)realloc(o->recs, o->recs_cap * sizeof(rec_t)); }
        o->recs[n] = r;
        if (!(plus > 0 && r.seq_len == r.qua_len)) break;
        raw4[1] += r.title_len; raw4[2] += r.seq_len; raw4[3] += r.qua_len;
        n++;
    }
    *n_out = n; *chunk_size = size - skipped;
    return 0;
}

static const char DNA_ALPHABET[] = "AGCTNRWSKMDVHBYXU.-"; /* RecordsProcessor.cpp:185-205 */

/* LosslessRecordsProcessor::ProcessForward + Initialize/FinalizeStats (RecordsProcessor.cpp:104-133, 209-267) */
static void preprocess(dsrc_oracle_t* o, u64 n, dna_stats_t* ds, qua_stats_t* qs)
{
    u8 lut[256];
    memset(lut, 255, sizeof(lut));
    for (u32 i = 0; DNA_ALPHABET[i]; ++i) lut[(u8)DNA_ALPHABET[i]] = (u8)i;
    memset(ds, 0, sizeof(*ds)); memset(qs, 0, sizeof(*qs));
    memset(ds->rank, 255, sizeof(ds->rank)); memset(qs->rank, 255, sizeof(qs->rank));
    qs->min_len = 0xFFFFFFFFu;
    for (u64 k = 0; k < n; ++k) {
        rec_t* r = &o->recs[k];
        u8* seq = o->work + r->seq; u8* qua = o->work + r->qua;
        u32 kept = 0, th = 0; u8 prev = 255;
        for (u32 i = 0; i < r->seq_len; ++i) {
            u8 s = lut[seq[i]]; u8 q = (u8)(qua[i] - o->qoff);
            if (s > 3 && q < 7) { q = (u8)(q + (128 + (((u32)s - 3 + 1) << 3) - 16)); }
            else { seq[kept++] = s; if (s < 20) ds->freq[s]++; }
            qua[i] = q;
            qs->freq[q]++;
            if (q != prev) qs->rle_len++;
            if (q != 2) th = i;
            prev = q;
        }
        r->seq_len = (u16)kept;
        r->trunc_len = (u16)(th + (r->qua_len > 0));
        if (prev == 2 && qs->rle_len > 0) qs->rle_len--;
        qs->raw_len += r->qua_len; qs->th_len += th;
        if (r->qua_len < qs->min_len) qs->min_len = r->qua_len;
        if (r->qua_len > qs->max_len) qs->max_len = r->qua_len;
    }
    for (u32 i = 0; i < 20; ++i) if (ds->freq[i]) ds->rank[i] = (u8)ds->count++;
    for (u32 i = 0; i < 256; ++i) if (qs->freq[i]) qs->rank[i] = (u8)qs->count++;
}

/* ------------------------------------------------------------------------------------------------
 * tag modeler  (src/TagModeler.cpp)
 * ---------------------------------------------------------------------------------------------- */
#define TAG_MAX_FIELDS 255
#define TAG_STAT_LEN 128
#define TAG_NUM_HUF 512

typedef struct { i32 key[TAG_NUM_HUF + 2]; u32 cnt[TAG_NUM_HUF + 2]; u32 n; } vmap_t; /* std::map<int32,int32> capped (TagModeler.cpp:375-382) */
static void vmap_inc(vmap_t* m, i32 k)
{
    u32 lo = 0, hi = m->n;
    while (lo < hi) { u32 mid = (lo + hi) / 2; if (m->key[mid] < k) lo = mid + 1; else hi = mid; }
    if (lo < m->n && m->key[lo] == k) { m->cnt[lo]++; return; }
    memmove(&m->key[lo + 1], &m->key[lo], (m->n - lo) * sizeof(i32));
    memmove(&m->cnt[lo + 1], &m->cnt[lo], (m->n - lo) * sizeof(u32));
    m->key[lo] = k; m->cnt[lo] = 1; m->n++;
}
static u32 vmap_get(const vmap_t* m, i32 k)
{
    for (u32 i = 0; i < m->n; ++i) if (m->key[i] == k) return m->cnt[i];
    return 0;
}
typedef struct { i32 cur_sym; u32 cur_len, run_len; u8* lens; u32 n_lens, cap_lens; } rle_t;
static void rle_push(rle_t* r, u32 v)
{
    if (r->n_lens == r->cap_lens) { r->cap_lens = r->cap_lens ? r->cap_lens * 2 : 64; r->lens = (u8*)realloc(r->lens, r->cap_lens); }
    r->lens[r->n_lens++] = (u8)v;
}
typedef struct {
    u32 len, min_len, max_len; u8 sep;
    int is_constant, is_len_constant, is_numeric;
    i32 min_value, max_value, min_delta, max_delta;
    u32 bits_num, bits_value, bits_len;
    int is_delta_coding, try_rle_val, try_rle_delta, is_delta_const, var_stat_encode;
    u8 scheme;  /* 1 ValueVar 2 ValueRle 3 DeltaVar 4 DeltaRle 5 DeltaConst (TagModeler.h:73) */
    u8* data; u8* ham;
    rle_t rle_val, rle_delta;
    vmap_t* num_values; vmap_t* delta_values;
    u32* chars;      /* [TAG_STAT_LEN+1][256], lazily allocated */
    huf_t* huf_global; huf_t** huf_local;
} field_t;

typedef struct {
    field_t* f; u32 n;
    u32 min_title, max_title, sym_freq[256];
    int mixed;
    i32 prev[TAG_MAX_FIELDS + 1];
    u32 rec_counter;
} tags_t;

static void field_free(field_t* f)
{
    free(f->data); free(f->ham); free(f->rle_val.lens); free(f->rle_delta.lens);
    free(f->num_values); free(f->delta_values); free(f->chars); free(f->huf_global);
    if (f->huf_local) { for (u32 i = 0; i <= TAG_STAT_LEN; ++i) free(f->huf_local[i]); free(f->huf_local); }
}
static void tags_free(tags_t* t) { for (u32 i = 0; i < t->n; ++i) field_free(&t->f[i]); free(t->f); t->f = NULL; t->n = 0; }

static int is_sep(u8 c) { return c == ' ' || c == '.' || c == '_' || c == ',' || c == '=' || c == ':' || c == '/' || c == '-' || c == '#' || c == 0; }

/* TagAnalyzer::InitializeFieldsStats (TagModeler.cpp:159-224), including the vector-growth side effect (Q1) */
static i32 tags_init(dsrc_oracle_t* o, tags_t* t, const u8* title, u32 title_len)
{
    u32 start = 0;
    memset(t, 0, sizeof(*t));
    t->min_title = 0xFFFFFFFFu;
    t->f = (field_t*)calloc(TAG_MAX_FIELDS + 1, sizeof(field_t));
    for (u32 i = 0; i <= title_len; ++i) {
        field_t* f; u32 v;
        if (title[i] < 128) t->sym_freq[title[i]] += (i != title_len);
        if (!is_sep(title[i]) && i != title_len) continue;
        if (t->n >= TAG_MAX_FIELDS) return ERR_UNSUPPORTED;
        /* push_back(Field()): when size == capacity libstdc++ doubles and copy-constructs the old
         * elements; Field's copy-ctor (TagModeler.cpp:54-130) does not copy num_values */
        if (t->n == o->tag_cap) {
            o->tag_cap = o->tag_cap ? o->tag_cap * 2 : 1;
            for (u32 k = 0; k < t->n; ++k) if (t->f[k].num_values) t->f[k].num_values->n = 0;
        }
        f = &t->f[t->n];
        f->len = f->min_len = f->max_len = i - start;
        f->data = (u8*)malloc(f->len + 1); memcpy(f->data, title + start, f->len); f->data[f->len] = 0;
        f->sep = title[i];
        f->is_constant = 1; f->is_len_constant = 1;
        f->is_numeric = parse_num(f->data, f->len, &v);
        f->ham = (u8*)malloc(f->len + 1); memset(f->ham, 1, f->len + 1);
        f->min_value = 1 << 30; f->max_value = -(1 << 30); f->min_delta = 1 << 30; f->max_delta = -(1 << 30);
        f->num_values = (vmap_t*)calloc(1, sizeof(vmap_t)); f->delta_values = (vmap_t*)calloc(1, sizeof(vmap_t));
        if (f->is_numeric) { f->min_value = f->max_value = (i32)v; vmap_inc(f->num_values, (i32)v); }
        start = i + 1;
        t->n++;
    }
    t->rec_counter = 0;
    return 0;
}

/* TagAnalyzer::UpdateNumericField (TagModeler.cpp:335-459) */
static void rle_step(rle_t* r, i32 v)
{
    if (r->cur_sym != v) { r->run_len++; r->cur_sym = v; rle_push(r, r->cur_len); r->cur_len = 0; }
    else { r->cur_len++; if (r->cur_len > 255) { rle_push(r, 255); r->cur_len = 0; r->run_len++; } }
}
static void tags_update_numeric(tags_t* t, field_t* f, i32 cur, i32 prev)
{
    if (cur < f->min_value) f->min_value = cur; else if (cur > f->max_value) f->max_value = cur;
    if (t->rec_counter > 0) {
        rle_step(&f->rle_val, cur);
        if (f->num_values->n) { vmap_inc(f->num_values, cur); if (f->num_values->n > TAG_NUM_HUF) f->num_values->n = 0; }
    } else {
        f->rle_val.cur_sym = cur; f->rle_val.cur_len = 0; f->rle_val.run_len = 0; f->rle_val.n_lens = 0;
        vmap_inc(f->num_values, cur);
    }
    if (t->rec_counter >= 1) {
        i32 d = (i32)((u32)cur - (u32)prev);
        if (t->rec_counter > 1) {
            if (d > f->max_delta) f->max_delta = d; else if (d < f->min_delta) f->min_delta = d;
            rle_step(&f->rle_delta, d);
            if (f->delta_values->n) { vmap_inc(f->delta_values, d); if (f->delta_values->n > TAG_NUM_HUF) f->delta_values->n = 0; }
        } else {
            f->max_delta = f->min_delta = d;
            f->rle_delta.cur_sym = d; f->rle_delta.cur_len = 0; f->rle_delta.run_len = 0; f->rle_delta.n_lens = 0;
            vmap_inc(f->delta_values, d);
        }
    }
}

/* TagAnalyzer::UpdateFieldsStats (TagModeler.cpp:226-333). title[title_len] must be readable. */
static void tags_update(tags_t* t, const u8* title, u32 title_len)
{
    u32 c = 0, start = 0, k;
    if (title_len < t->min_title) t->min_title = title_len;
    if (title_len > t->max_title) t->max_title = title_len;
    if (t->mixed) { for (u32 i = 0; i < title_len; ++i) if (title[i] < 128) t->sym_freq[title[i]]++; return; }
    for (k = 0; k <= title_len && c < t->n; ++k) {
        field_t* f; u32 flen, clen;
        if (title[k] < 128) t->sym_freq[title[k]] += (k != title_len);
        if (title[k] != t->f[c].sep && k < title_len) continue;
        f = &t->f[c]; flen = k - start;
        if (flen > f->max_len) f->max_len = flen; else if (flen < f->min_len) f->min_len = flen;
        if (!f->chars) f->chars = (u32*)calloc((TAG_STAT_LEN + 1) * 256, sizeof(u32));
        clen = flen < TAG_STAT_LEN ? flen : TAG_STAT_LEN;
        for (u32 x = 0; x < clen; ++x) f->chars[x * 256 + title[start + x]]++;
        for (u32 x = TAG_STAT_LEN; x < flen; ++x) f->chars[TAG_STAT_LEN * 256 + title[start + x]]++;
        if (f->is_constant) f->is_constant = (flen == f->len) && memcmp(f->data, title + start, f->len) == 0;
        if (f->is_len_constant) f->is_len_constant = f->len == flen;
        if (f->is_numeric) {
            u32 v;
            f->is_numeric = parse_num(title + start, flen, &v);
            if (f->is_numeric) { tags_update_numeric(t, f, (i32)v, t->prev[c]); t->prev[c] = (i32)v; }
        }
        if (!f->is_constant)
            for (u32 p = 0; p < flen && p < f->len; ++p) f->ham[p] &= (u8)(f->data[p] == title[p + start]);
        start = k + 1; c++;
    }
    if (c != t->n || k != title_len + 1u) t->mixed = 1;
    t->rec_counter++;
}

/* TagAnalyzer::FinalizeFieldsStats (TagModeler.cpp:461-551) */
static void tags_finalize(tags_t* t)
{
    if (t->mixed) return;
    for (u32 i = 0; i < t->n; ++i) {
        field_t* f = &t->f[i]; i32 diff;
        if (!f->is_numeric) { if (!f->is_constant) f->bits_len = bit_length64((u64)(f->max_len - f->min_len)); continue; }
        {   /* int32 arithmetic wraps exactly as the reference's does on x86-64 */
            i32 dv = (i32)((u32)f->max_value - (u32)f->min_value), dd = (i32)((u32)f->max_delta - (u32)f->min_delta);
            if (dv < dd) { f->is_delta_coding = 0; diff = dv; } else { f->is_delta_coding = 1; diff = dd; }
        }
        rle_push(&f->rle_val, f->rle_val.cur_len);
        if (f->rle_val.cur_len > 0) { f->rle_val.cur_len = 0; f->rle_val.run_len++; }
        if ((float)t->rec_counter / (float)f->rle_val.run_len > 1.25f) f->try_rle_val = 1;
        if (f->is_delta_coding) {
            f->is_delta_const = diff == 0;
            if (!f->is_delta_const) {
                rle_push(&f->rle_delta, f->rle_delta.cur_len);
                if (f->rle_delta.cur_len > 0) { f->rle_delta.cur_len = 0; f->rle_delta.run_len++; }
                if ((float)t->rec_counter / (float)f->rle_delta.run_len > 1.25f) f->try_rle_delta = 1;
            }
        }
        if (f->is_delta_coding && f->is_delta_const) f->scheme = 5;
        else if (f->is_delta_coding && f->try_rle_delta) f->scheme = 4;
        else if (f->try_rle_val) f->scheme = 2;
        else if (f->is_delta_coding) { u32 d = (u32)(f->max_delta - f->min_delta) + 1; f->scheme = 3; f->var_stat_encode = d <= TAG_NUM_HUF && f->delta_values->n; }
        else { u32 d = (u32)(f->max_value - f->min_value) + 1; f->scheme = 1; f->var_stat_encode = d <= TAG_NUM_HUF && f->num_values->n; }
        f->bits_num = bit_length64((u64)(int64_t)diff);
        diff = (i32)((u32)f->max_value - (u32)f->min_value);
        f->bits_value = bit_length64((u64)(int64_t)diff);
    }
}

/* TagTokenizerEncoder::StoreFields (TagModeler.cpp:569-693) */
static void tags_store_fields(tags_t* t, bitw_t* w)
{
    u32 tmp[TAG_NUM_HUF];
    bw_byte(w, (u8)t->n);
    for (u32 i = 0; i < t->n; ++i) {
        field_t* f = &t->f[i];
        bw_byte(w, f->sep); bw_byte(w, (u8)f->is_constant);
        if (f->is_constant) { bw_u32(w, f->len); bw_bytes(w, f->data, f->len); continue; }
        bw_byte(w, (u8)f->is_numeric);
        if (f->is_numeric) {
            bw_byte(w, f->scheme); bw_u32(w, (u32)f->min_value); bw_u32(w, (u32)f->max_value);
            if (f->scheme >= 3) {
                bw_u32(w, (u32)f->min_delta); bw_u32(w, (u32)f->max_delta);
                if (f->scheme == 3) {
                    bw_byte(w, (u8)f->var_stat_encode);
                    if (f->var_stat_encode) {
                        u32 d = (u32)(f->max_delta - f->min_delta) + 1;
                        for (u32 j = 0; j < d; ++j) tmp[j] = vmap_get(f->delta_values, f->min_delta + (i32)j);
                        f->huf_global = (huf_t*)malloc(sizeof(huf_t)); huf_build(f->huf_global, tmp, d); huf_store(f->huf_global, w);
                    }
                }
            } else if (f->scheme == 1) {
                bw_byte(w, (u8)f->var_stat_encode);
                if (f->var_stat_encode) {
                    u32 d = (u32)(f->max_value - f->min_value) + 1;
                    for (u32 j = 0; j < d; ++j) tmp[j] = vmap_get(f->num_values, f->min_value + (i32)j);
                    f->huf_global = (huf_t*)malloc(sizeof(huf_t)); huf_build(f->huf_global, tmp, d); huf_store(f->huf_global, w);
                }
            }
            continue;
        }
        bw_byte(w, (u8)f->is_len_constant); bw_u32(w, f->len); bw_u32(w, f->max_len); bw_u32(w, f->min_len);
        bw_bytes(w, f->data, f->len);
        for (u32 j = 0; j < f->len; ++j) bw_bit(w, f->ham[j]);
        bw_flush(w);
        f->huf_local = (huf_t**)calloc(TAG_STAT_LEN + 1, sizeof(huf_t*));
        if (!f->chars) f->chars = (u32*)calloc((TAG_STAT_LEN + 1) * 256, sizeof(u32));
        for (u32 j = 0; j < (f->max_len < TAG_STAT_LEN ? f->max_len : TAG_STAT_LEN); ++j) {
            if (j >= f->len || !f->ham[j]) {
                f->huf_local[j] = (huf_t*)malloc(sizeof(huf_t));
                huf_build(f->huf_local[j], &f->chars[j * 256], 256); huf_store(f->huf_local[j], w);
            }
        }
        if (f->max_len >= TAG_STAT_LEN) { /* max_len == 128 exactly reads chars[128] out of bounds upstream; we read zeros */
            f->huf_local[TAG_STAT_LEN] = (huf_t*)malloc(sizeof(huf_t));
            huf_build(f->huf_local[TAG_STAT_LEN], &f->chars[TAG_STAT_LEN * 256], 256); huf_store(f->huf_local[TAG_STAT_LEN], w);
        }
    }
}

/* TagTokenizerEncoder::StoreNumericField (TagModeler.cpp:753-874) */
static void tags_store_numeric(bitw_t* w, field_t* f, u32 rec_counter, i32 cur, i32 prev)
{
    if (rec_counter == 0) {
        i32 d = (i32)((u32)cur - (u32)f->min_value);
        bw_bits(w, (u32)d, f->bits_value);
        if (f->scheme == 2) { f->rle_val.run_len = 0; f->rle_val.cur_len = f->rle_val.lens[0]; f->rle_val.cur_sym = d; bw_bits(w, f->rle_val.cur_len, 8); }
        return;
    }
    switch (f->scheme) {
    case 5: break;
    case 4: {
        i32 d = (i32)((u32)cur - (u32)prev - (u32)f->min_delta);
        if (rec_counter == 1 || f->rle_delta.cur_len == 0) {
            if (rec_counter == 1) f->rle_delta.run_len = 0; else f->rle_delta.run_len++;
            f->rle_delta.cur_len = f->rle_delta.lens[f->rle_delta.run_len]; f->rle_delta.cur_sym = d;
            bw_bits(w, (u32)d, f->bits_num); bw_bits(w, f->rle_delta.cur_len, 8);
        } else f->rle_delta.cur_len--;
        break; }
    case 3: {
        i32 s = (i32)((u32)cur - (u32)prev - (u32)f->min_delta);
        if (f->huf_global) huf_put(f->huf_global, w, (u32)s); else bw_bits(w, (u32)s, f->bits_num);
        break; }
    case 2: {
        i32 d = (i32)((u32)cur - (u32)f->min_value);
        if (f->rle_val.cur_len == 0) {
            f->rle_val.run_len++; f->rle_val.cur_len = f->rle_val.lens[f->rle_val.run_len]; f->rle_val.cur_sym = d;
            bw_bits(w, (u32)d, f->bits_value); bw_bits(w, f->rle_val.cur_len, 8);
        } else f->rle_val.cur_len--;
        break; }
    case 1: {
        i32 s = (i32)((u32)cur - (u32)f->min_value);
        if (f->huf_global) huf_put(f->huf_global, w, (u32)s); else bw_bits(w, (u32)s, f->bits_num);
        break; }
    default: break;
    }
}

/* TagTokenizerEncoder::EncodeNextFields (TagModeler.cpp:695-751) */
static void tags_encode_record(tags_t* t, bitw_t* w, const u8* title, u32 title_len, u32 rec_counter, i32* prev)
{
    u32 c = 0, start = 0;
    for (u32 k = 0; k <= title_len && c < t->n; ++k) {
        field_t* f = &t->f[c]; u32 flen;
        if (title[k] != f->sep && k < title_len) continue;
        flen = k - start;
        if (f->is_constant) { start = k + 1; c++; continue; }
        if (f->is_numeric) {
            i32 v = (i32)to_num(title + start, flen);
            tags_store_numeric(w, f, rec_counter, v, prev[c]); prev[c] = v;
            start = k + 1; c++; continue;
        }
        if (!f->is_len_constant) bw_bits(w, flen - f->min_len, f->bits_len);
        for (u32 j = 0; j < flen; ++j)
            if (j >= f->len || !f->ham[j]) huf_put(f->huf_local[j < TAG_STAT_LEN ? j : TAG_STAT_LEN], w, title[start + j]);
        start = k + 1; c++;
    }
}

/* ------------------------------------------------------------------------------------------------
 * quality modelers
 * ---------------------------------------------------------------------------------------------- */
/* QualityNormalModelerProxy::SelectSchemeId (QualityModelerProxy.h:113-122) */
static u8 q0_scheme(const qua_stats_t* s)
{
    if ((float)s->th_len / (float)s->rle_len > 1.25f) return 2;
    if ((float)s->raw_len / (float)s->th_len > 1.10f) return 1;
    return 0;
}
/* QualityOrderModelerProxyLossless::SelectSchemeId (QualityModelerProxy.h:261-282) */
static u8 qo_scheme(const qua_stats_t* s, u32 order)
{
    u8 sc = 255;
    for (u32 i = 0; i < 8; ++i) if ((16u << i) >= s->count) { sc = (u8)i; break; }
    if (sc != 255 && order == 2) {
        double ratio = (double)s->raw_len / (double)s->rle_len;
        if (s->max_len == s->min_len && ratio > 1.175) sc = (u8)(sc + 4);
    }
    return sc;
}
/* template parameters <SymbolCount, SymbolOrder, Rescale> per scheme (QualityModelerProxy.h:231-254) */
static int qo_params(u32 order, u8 scheme, u32* alpha, u32* sym_order, u32* rescale)
{
    static const u32 A[4] = {16, 32, 64, 128};
    static const u32 O1[4] = {3, 2, 1, 1}, O2[4] = {4, 3, 2, 1};
    if (scheme > 7) return 0;
    *alpha = A[scheme & 3];
    *sym_order = (order == 1 ? O1 : O2)[scheme & 3];
    *rescale = (scheme & 4) ? *alpha : 8;
    return 1;
}
static u32 log2u(u32 x) { u32 r = 0; while (x > 1) { x >>= 1; ++r; } return r; }

/* TQualityModelBase::UpdateHash (QualityEncoder.h:77-89) */
typedef struct { u64 hash, sym_buf; u32 bits, bits_lo, bits_hi; u64 sym_mask, swap_mask, hash_mask; } qhash_t;
static void qhash_init(qhash_t* h, u32 alpha, u32 sym_order)
{
    h->hash = 0; h->sym_buf = 0; h->bits = log2u(alpha);
    h->bits_lo = (sym_order / 2) * h->bits; h->bits_hi = (sym_order / 2 + 1) * h->bits;
    h->sym_mask = (1ull << h->bits) - 1;
    h->swap_mask = ((1ull << h->bits_lo) - 1) | ~((1ull << h->bits_hi) - 1);
    h->hash_mask = (1ull << (sym_order * h->bits)) - 1;
}
static void qhash_update(qhash_t* h, u32 sym)
{
    u64 next, swp;
    h->hash <<= h->bits;
    next = (h->hash >> h->bits_lo) & h->sym_mask;
    swp = (next + h->sym_buf) / 2;
    h->hash &= h->swap_mask; h->hash |= swp << h->bits_lo; h->hash |= sym;
    h->sym_buf = next;
}

/* TQualityOrderModeler::Encode (QualityOrderModeler.h:36-51) + TTranslationalQualityEncoder (QualityEncoder.h:281-367) */
static void qua_order_encode(dsrc_oracle_t* o, bitw_t* w, u64 n, const qua_stats_t* qs, u8 scheme)
{
    u32 alpha, so, rescale; qhash_t h; rce_t e; u16* model;
    qo_params(o->qua_order, scheme, &alpha, &so, &rescale);
    bw_flush(w);
    for (u32 i = 0; i < 256; ++i) bw_bit(w, qs->rank[i] != 255);
    bw_flush(w);
    model = get_model(o, ((u64)1 << (log2u(alpha) * (so + 1))) * alpha);
    qhash_init(&h, alpha, so);
    rce_start(&e, w);
    for (u64 k = 0; k < n; ++k) {
        const rec_t* r = &o->recs[k]; const u8* q = o->work + r->qua;
        for (u32 j = 0; j < r->qua_len; ++j) {
            u32 sym = qs->rank[q[j]], pctx = j * rescale / r->qua_len;
            u64 ctx = ((h.hash & h.hash_mask) << h.bits) | pctx;
            model_encode(model + ctx * alpha, alpha, &e, sym);
            qhash_update(&h, sym);
        }
    }
    rce_end(&e);
}
static i32 qua_order_decode(dsrc_oracle_t* o, bitr_t* r, u64 n, u8 scheme)
{
    u32 alpha, so, rescale, nsym = 0; qhash_t h; rcd_t d; u16* model; u8 symbols[256];
    if (!qo_params(o->qua_order, scheme, &alpha, &so, &rescale)) return ERR_FORMAT;
    memset(symbols, 255, sizeof(symbols));
    br_flush(r);
    for (u32 i = 0; i < 256; ++i) if (br_bit(r)) symbols[nsym++] = (u8)i;
    br_flush(r);
    model = get_model(o, ((u64)1 << (log2u(alpha) * (so + 1))) * alpha);
    qhash_init(&h, alpha, so);
    rcd_start(&d, r);
    for (u64 k = 0; k < n; ++k) {
        rec_t* rc = &o->recs[k]; u8* q = o->work + rc->qua; u32 nc = 0;
        for (u32 j = 0; j < rc->qua_len; ++j) {
            u32 pctx = j * rescale / rc->qua_len, sym;
            u64 ctx = ((h.hash & h.hash_mask) << h.bits) | pctx;
            sym = model_decode(model + ctx * alpha, alpha, &d);
            qhash_update(&h, sym);
            q[j] = symbols[sym]; nc += q[j] >= 128;
        }
        rc->seq_len = (u16)(rc->qua_len - nc);
        if (r->overrun) return ERR_FORMAT;
    }
    return 0;
}

/* IQualityPositionModeler::Encode (QualityPositionModeler.cpp:57-72) Plain (:160-187) / Truncated (:240-287) */
static void qua_position_encode(dsrc_oracle_t* o, bitw_t* w, u64 n, const qua_stats_t* qs, int truncated)
{
    u32 L = qs->max_len, S = qs->count;
    u32* st = (u32*)calloc((u64)L * S + 1, sizeof(u32));
    huf_t* hf = (huf_t*)malloc(sizeof(huf_t) * (L ? L : 1));
    for (u64 k = 0; k < n; ++k) {
        const rec_t* r = &o->recs[k]; const u8* q = o->work + r->qua; u32 lim = truncated ? r->trunc_len : r->qua_len;
        for (u32 j = 0; j < lim; ++j) st[(u64)j * S + qs->rank[q[j]]]++;
    }
    for (u32 j = 0; j < L; ++j) huf_build(&hf[j], &st[(u64)j * S], S);
    bw_flush(w);
    bw_u32(w, L);
    for (u32 i = 0; i < 256; ++i) bw_bit(w, qs->rank[i] != 255);
    for (u32 j = 0; j < L; ++j) huf_store(&hf[j], w);
    if (truncated) {
        int variable = qs->min_len != qs->max_len; u32 max_bits = bit_length64(L);
        bw_bit(w, (u32)variable);
        for (u64 k = 0; k < n; ++k) {
            const rec_t* r = &o->recs[k]; const u8* q = o->work + r->qua;
            bw_bit(w, r->qua_len != r->trunc_len);
            if (r->qua_len != r->trunc_len) bw_bits(w, r->trunc_len, variable ? bit_length64(r->qua_len) : max_bits);
            for (u32 j = 0; j < r->trunc_len; ++j) huf_put(&hf[j], w, qs->rank[q[j]]);
        }
    } else {
        for (u64 k = 0; k < n; ++k) {
            const rec_t* r = &o->recs[k]; const u8* q = o->work + r->qua;
            for (u32 j = 0; j < r->qua_len; ++j) huf_put(&hf[j], w, qs->rank[q[j]]);
        }
    }
    bw_flush(w);
    free(st); free(hf);
}
static i32 qua_position_decode(dsrc_oracle_t* o, bitr_t* r, u64 n, int truncated)
{
    u32 L, nsym = 0, max_bits; u8 symbols[256]; hufd_t* hd; int variable = 0;
    br_flush(r);
    L = br_u32(r);
    if (L > 65535) return ERR_FORMAT;
    memset(symbols, 255, sizeof(symbols));
    for (u32 i = 0; i < 256; ++i) if (br_bit(r)) symbols[nsym++] = (u8)i;
    hd = (hufd_t*)malloc(sizeof(hufd_t) * (L ? L : 1));
    for (u32 j = 0; j < L; ++j) hufd_load(&hd[j], r);
    max_bits = bit_length64(L);
    if (truncated) variable = (int)br_bit(r);
    for (u64 k = 0; k < n && !r->overrun; ++k) {
        rec_t* rc = &o->recs[k]; u8* q = o->work + rc->qua; u32 th = rc->qua_len, nc = 0;
        if (truncated && br_bit(r)) th = br_bits(r, variable ? bit_length64(rc->qua_len) : max_bits);
        if (th > rc->qua_len || th > L) { free(hd); return ERR_FORMAT; }
        for (u32 j = 0; j < th; ++j) { q[j] = symbols[hufd_get(&hd[j], r) & 255]; nc += q[j] >= 128; }
        for (u32 j = th; j < rc->qua_len; ++j) q[j] = 2;
        rc->seq_len = (u16)(rc->qua_len - nc);
    }
    br_flush(r);
    free(hd);
    return r->overrun ? ERR_FORMAT : 0;
}

/* QualityRLEModeler::Encode (QualityRLEModeler.cpp:121-373) */
static void qua_rle_encode(dsrc_oracle_t* o, bitw_t* w, u64 n, const qua_stats_t* qs)
{
    u8* sym_run = (u8*)malloc(qs->raw_len + 1); u8* len_run = (u8*)malloc(qs->raw_len + 1);
    u32 qf[256] = {0}, lf[256] = {0}, runs = 0, nq = 0, nl = 0;
    u8 qrank[256], lrank[256], prev = 255, cur_len = 0;
    for (u64 k = 0; k < n; ++k) {                       /* EncodeRecords :142-205, runs span record boundaries */
        const rec_t* r = &o->recs[k]; const u8* q = o->work + r->qua;
        for (u32 j = 0; j < r->qua_len; ++j) {
            if (q[j] == prev && cur_len < 254) cur_len++;
            else {
                if (prev != 255) { sym_run[runs] = prev; len_run[runs++] = cur_len; qf[prev]++; lf[cur_len]++; }
                cur_len = 0; prev = q[j];
            }
        }
    }
    sym_run[runs] = prev; len_run[runs++] = cur_len; qf[prev]++; lf[cur_len]++;
    memset(qrank, 255, 256); memset(lrank, 255, 256);    /* CalculateSymbolIndices :207-231 */
    for (u32 i = 0; i < 256; ++i) { if (qf[i]) qrank[i] = (u8)nq++; if (lf[i]) lrank[i] = (u8)nl++; }
    bw_flush(w);
    bw_u32(w, runs);
    for (u32 i = 0; i < 256; ++i) bw_bit(w, qrank[i] != 255);
    for (u32 i = 0; i < 256; ++i) bw_bit(w, lrank[i] != 255);
    if (nq > 1) {                                        /* ComputeHuffmanContext :233-310 */
        u32* QF = (u32*)calloc((u64)nq * nq, 4); u32* LF = (u32*)calloc((u64)nq * nl, 4);
        huf_t* qh = (huf_t*)malloc(sizeof(huf_t) * nq); huf_t* lh = (huf_t*)malloc(sizeof(huf_t) * nq);
        u32 p = 0;
        for (u32 i = 0; i < runs; ++i) { u32 q = qrank[sym_run[i]], l = lrank[len_run[i]]; QF[p * nq + q]++; LF[q * nl + l]++; p = q; }
        for (u32 i = 0; i < nq; ++i) { huf_build(&qh[i], &QF[i * nq], nq); huf_build(&lh[i], &LF[i * nl], nl); }
        for (u32 i = 0; i < nq; ++i) { huf_store(&qh[i], w); huf_store(&lh[i], w); }
        p = 0;
        for (u32 i = 0; i < runs; ++i) { u32 q = qrank[sym_run[i]], l = lrank[len_run[i]]; huf_put(&qh[p], w, q); huf_put(&lh[q], w, l); p = q; }
        free(QF); free(LF); free(qh); free(lh);
    } else if (nl > 1) { bw_flush(w); bw_byte(w, lrank[len_run[0]]); }
    bw_flush(w);
    free(sym_run); free(len_run);
}
static i32 qua_rle_decode(dsrc_oracle_t* o, bitr_t* r, u64 n)
{
    u32 runs = br_u32(r), nq = 0, nl = 0, idx = 0, cur_len = 0; u8 qs[256], ls[256], cur_q = 0; u8 *sr, *lr;
    for (u32 i = 0; i < 256; ++i) if (br_bit(r)) qs[nq++] = (u8)i;
    for (u32 i = 0; i < 256; ++i) if (br_bit(r)) ls[nl++] = (u8)i;
    br_flush(r);
    if (r->overrun || runs == 0 || nq == 0 || nl == 0 || runs > (1u << 30)) return ERR_FORMAT;
    sr = (u8*)malloc(runs); lr = (u8*)malloc(runs);
    if (nq > 1) {
        hufd_t* qh = (hufd_t*)malloc(sizeof(hufd_t) * nq); hufd_t* lh = (hufd_t*)malloc(sizeof(hufd_t) * nq); u32 p = 0;
        for (u32 i = 0; i < nq; ++i) { hufd_load(&qh[i], r); hufd_load(&lh[i], r); }
        br_flush(r);
        for (u32 i = 0; i < runs && !r->overrun; ++i) { u32 q = hufd_get(&qh[p], r) % nq; sr[i] = qs[q]; p = q; lr[i] = ls[hufd_get(&lh[p], r) % nl]; }
        free(qh); free(lh);
    } else {
        u8 lb, le;
        br_flush(r);
        if (nl > 1) { br_flush(r); lb = ls[br_byte(r) % nl]; le = ls[0]; if (le == lb) le = ls[1]; } else { lb = ls[0]; le = lb; }
        memset(sr, qs[0], runs); memset(lr, lb, runs); lr[runs - 1] = le;
    }
    for (u64 k = 0; k < n; ++k) {
        rec_t* rc = &o->recs[k]; u8* q = o->work + rc->qua; u32 nc = 0;
        for (u32 j = 0; j < rc->qua_len; ++j) {
            if (cur_len == 0) { if (idx >= runs) { free(sr); free(lr); return ERR_FORMAT; } cur_q = sr[idx]; cur_len = (u32)lr[idx] + 1; idx++; }
            q[j] = cur_q; --cur_len; nc += cur_q >= 128;
        }
        rc->seq_len = (u16)(rc->qua_len - nc);
    }
    br_flush(r);
    free(sr); free(lr);
    return r->overrun ? ERR_FORMAT : 0;
}

/* ------------------------------------------------------------------------------------------------
 * DNA modelers
 * ---------------------------------------------------------------------------------------------- */
static void dna_order_params(u32 order, u8 scheme, u32* alpha, u32* eff_order) /* DnaModelerProxy.h:192-227 */
{
    *alpha = scheme == 0 ? 4 : 8;
    *eff_order = (scheme == 1 && order > 7) ? 7 : order;
}
static void dna_encode(dsrc_oracle_t* o, bitw_t* w, u64 n, const dna_stats_t* ds)
{
    u8 scheme;
    if (ds->count == 0) { bw_byte(w, 255); return; }       /* DnaModelerProxy.h:50-60 */
    scheme = ds->count <= 4 ? 0 : 1;
    bw_byte(w, scheme);
    if (o->dna_order == 0) {
        if (scheme == 0) {                                  /* DnaModelerBasicB2.h:34-46 */
            for (u64 k = 0; k < n; ++k) { const rec_t* r = &o->recs[k]; const u8* s = o->work + r->seq; for (u32 j = 0; j < r->seq_len; ++j) bw_bits(w, s[j], 2); }
            bw_flush(w);
        } else {                                            /* DnaModelerHuffman.cpp:21-73 */
            huf_t h; u32 fr[20];
            /* reference indexes symbolFreqs[symbols[i]] (:36): exact only when the present symbols are the
             * prefix 0..k-1; outside that it reads out of bounds -- we use 0 there */
            for (u32 i = 0; i < ds->count; ++i) fr[i] = ds->rank[i] < 20 ? ds->freq[ds->rank[i]] : 0;
            huf_build(&h, fr, ds->count);
            for (u32 i = 0; i < 20; ++i) bw_bit(w, ds->rank[i] != 255);
            bw_flush(w);
            huf_store(&h, w);
            for (u64 k = 0; k < n; ++k) { const rec_t* r = &o->recs[k]; const u8* s = o->work + r->seq; for (u32 j = 0; j < r->seq_len; ++j) huf_put(&h, w, ds->rank[s[j]]); }
            bw_flush(w);
        }
    } else {                                                /* DnaModelerRCO.h:45-62, 94-133 */
        u32 alpha, ord, bits; u64 mask, hash = 0; rce_t e; u16* model;
        dna_order_params(o->dna_order, scheme, &alpha, &ord);
        bits = log2u(alpha); mask = ((u64)1 << (ord * bits)) - 1;
        model = get_model(o, ((u64)1 << (ord * bits)) * alpha);
        rce_start(&e, w);
        for (u64 k = 0; k < n; ++k) {
            const rec_t* r = &o->recs[k]; const u8* s = o->work + r->seq;
            for (u32 j = 0; j < r->seq_len; ++j) { model_encode(model + hash * alpha, alpha, &e, s[j]); hash = ((hash << bits) | s[j]) & mask; }
        }
        rce_end(&e);
    }
}
static i32 dna_decode(dsrc_oracle_t* o, bitr_t* r, u64 n)
{
    u8 scheme = br_byte(r);
    if (scheme == 255) return 0;
    if (scheme > 1) return ERR_FORMAT;
    if (o->dna_order == 0) {
        if (scheme == 0) {
            for (u64 k = 0; k < n; ++k) { rec_t* rc = &o->recs[k]; u8* s = o->work + rc->seq; for (u32 j = 0; j < rc->seq_len; ++j) s[j] = (u8)br_bits(r, 2); }
            br_flush(r);
        } else {
            u8 symbols[20]; u32 ns = 0; hufd_t hd;
            memset(symbols, 255, sizeof(symbols));
            for (u32 i = 0; i < 20; ++i) if (br_bit(r)) symbols[ns++] = (u8)i;
            hufd_load(&hd, r);
            for (u64 k = 0; k < n && !r->overrun; ++k) { rec_t* rc = &o->recs[k]; u8* s = o->work + rc->seq; for (u32 j = 0; j < rc->seq_len; ++j) s[j] = symbols[hufd_get(&hd, r) % 20]; }
            br_flush(r);
        }
    } else {
        u32 alpha, ord, bits; u64 mask, hash = 0; rcd_t d; u16* model;
        dna_order_params(o->dna_order, scheme, &alpha, &ord);
        bits = log2u(alpha); mask = ((u64)1 << (ord * bits)) - 1;
        model = get_model(o, ((u64)1 << (ord * bits)) * alpha);
        rcd_start(&d, r);
        for (u64 k = 0; k < n && !r->overrun; ++k) {
            rec_t* rc = &o->recs[k]; u8* s = o->work + rc->seq;
            for (u32 j = 0; j < rc->seq_len; ++j) { u32 sym = model_decode(model + hash * alpha, alpha, &d); s[j] = (u8)sym; hash = ((hash << bits) | sym) & mask; }
        }
    }
    return r->overrun ? ERR_FORMAT : 0;
}

/* ------------------------------------------------------------------------------------------------
 * BlockCompressor::Store  (src/BlockCompressor.cpp:208-259, 359-488)
 * ---------------------------------------------------------------------------------------------- */
int64_t dsrc_oracle_store(dsrc_oracle_t* o, const u8* fastq, u64 size, u8* out, u64 cap, u64* raw4, u64* comp4)
{
    u64 n = 0, chunk_size = 0, raw[4], pos; i32 rc; dna_stats_t ds; qua_stats_t qs; tags_t tg; bitw_t w;
    u32 flags = 0, len_bits; u64 cmp[4];
    if (size >= (1ull << 31)) return ERR_UNSUPPORTED;
    if (o->work_cap < size + 8) { free(o->work); o->work_cap = size + 8 + size / 4; o->work = (u8*)malloc(o->work_cap); }
    memcpy(o->work, fastq, size);
    o->work[size] = '\n'; /* the byte the reference reads as title[titleLen] of the last record */
    rc = parse_records(o, size, &n, &chunk_size, raw);
    if (rc < 0) return rc;
    if (n == 0) return ERR_FORMAT;
    u32 crc[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    if (o->calc_crc) { /* IRecordsProcessor::ProcessForward(records, n, flags) (RecordsProcessor.cpp:135-152): hash the raw fields first */
        for (u64 k = 0; k < n; ++k) {
            const rec_t* r = &o->recs[k];
            crc[0] = crc32_update(crc[0], o->work + r->title, r->title_len);
            crc[1] = crc32_update(crc[1], o->work + r->seq, r->seq_len);
            crc[2] = crc32_update(crc[2], o->work + r->qua, r->qua_len);
        }
    }
    preprocess(o, n, &ds, &qs);
    /* AnalyzeMetaData :184-205 */
    if (qs.max_len != qs.min_len) flags |= 2;
    /* AnalyzeTags :359-401 */
    rc = tags_init(o, &tg, o->work + o->recs[0].title, o->recs[0].title_len);
    if (rc < 0) { tags_free(&tg); return rc; }
    for (u64 k = 0; k < n; ++k) tags_update(&tg, o->work + o->recs[k].title, o->recs[k].title_len);
    tags_finalize(&tg);
    if (tg.mixed) flags |= 4;

    bw_init(&w, out, cap);
    /* StoreMetaData :403-443 */
    bw_u32(&w, (u32)n); bw_u32(&w, qs.max_len); bw_u32(&w, flags); bw_u32(&w, (u32)chunk_size);
    if (flags & 2) bw_u32(&w, qs.min_len);
    if (o->calc_crc) { bw_u32(&w, ~crc[0]); bw_u32(&w, ~crc[1]); bw_u32(&w, ~crc[2]); }   /* :424-440 */
    bw_flush(&w);
    cmp[0] = w.pos; pos = w.pos;
    /* StoreTags :458-488 */
    len_bits = bit_length64((u64)(qs.max_len - qs.min_len));
    if (!tg.mixed) {
        i32 prev[TAG_MAX_FIELDS + 1];
        memset(prev, 0, sizeof(prev));
        tags_store_fields(&tg, &w);
        for (u64 k = 0; k < n; ++k) {
            tags_encode_record(&tg, &w, o->work + o->recs[k].title, o->recs[k].title_len, (u32)k, prev);
            if (len_bits > 0) bw_bits(&w, o->recs[k].qua_len - qs.min_len, len_bits);
        }
        bw_flush(&w);
    } else { /* TagRawEncoder :1217-1284 */
        u32 tl_bits = bit_length64((u64)(tg.max_title - tg.min_title)), fr[128], ns = 0; u8 rank[128]; huf_t h;
        bw_u32(&w, tg.min_title); bw_u32(&w, tg.max_title);
        memset(rank, 255, sizeof(rank));
        for (u32 i = 0; i < 128; ++i) if (tg.sym_freq[i] > 0) { rank[i] = (u8)ns; fr[ns++] = tg.sym_freq[i]; }
        huf_build(&h, fr, ns);
        for (u32 i = 0; i < 128; ++i) bw_bit(&w, rank[i] != 255);
        bw_flush(&w);
        huf_store(&h, &w);
        for (u64 k = 0; k < n; ++k) {
            const rec_t* r = &o->recs[k]; const u8* t = o->work + r->title;
            if (tl_bits > 0) bw_bits(&w, r->title_len - tg.min_title, tl_bits);
            for (u32 i = 0; i < r->title_len; ++i) huf_put(&h, &w, rank[t[i] & 127]);
            if (len_bits > 0) bw_bits(&w, r->qua_len - qs.min_len, len_bits);
        }
        bw_flush(&w);
    }
    tags_free(&tg);
    cmp[1] = w.pos - pos; pos = w.pos;
    /* StoreQuality :452 -> IQualityModelerProxy::Encode (QualityModelerProxy.h:48-58) */
    {
        u8 scheme = o->qua_order == 0 ? q0_scheme(&qs) : qo_scheme(&qs, o->qua_order);
        bw_byte(&w, scheme);
        if (scheme != 255) {
            if (o->qua_order == 0) {
                if (scheme == 2) qua_rle_encode(o, &w, n, &qs);
                else qua_position_encode(o, &w, n, &qs, scheme == 1);
            } else qua_order_encode(o, &w, n, &qs, scheme);
        }
    }
    cmp[3] = w.pos - pos; pos = w.pos;
    /* StoreDNA :446 */
    dna_encode(o, &w, n, &ds);
    cmp[2] = w.pos - pos;
    bw_flush(&w);
    if (raw4) memcpy(raw4, raw, sizeof(raw));
    if (comp4) memcpy(comp4, cmp, sizeof(cmp));
    return w.overflow ? ERR_CAP : (int64_t)w.pos;
}

/* ------------------------------------------------------------------------------------------------
 * BlockCompressor::Read  (src/BlockCompressor.cpp:262-356, 491-570) + ProcessBackward (RecordsProcessor.cpp:269-295)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    u8 sep; int is_constant, is_numeric, is_len_constant; u8 scheme; int var_stat;
    u32 len, max_len, min_len, bits_value, bits_num, bits_len; i32 min_value, min_delta;
    u8* data; u8* ham; hufd_t* hg; hufd_t** hl; u32 rle_len; i32 rle_sym;
} dfield_t;

static void dfields_free(dfield_t* f, u32 n)
{
    for (u32 i = 0; i < n; ++i) {
        free(f[i].data); free(f[i].ham); free(f[i].hg);
        if (f[i].hl) { for (u32 j = 0; j <= TAG_STAT_LEN; ++j) free(f[i].hl[j]); free(f[i].hl); }
    }
    free(f);
}

int64_t dsrc_oracle_read(dsrc_oracle_t* o, const u8* blk, u64 size, u8* out, u64 cap)
{
    bitr_t r; u32 n, max_len, min_len, flags, chunk_size, len_bits; u64 pos = 0; i32 rc = 0;
    br_init(&r, blk, size);
    n = br_u32(&r); max_len = br_u32(&r); flags = br_u32(&r); chunk_size = br_u32(&r);   /* ReadMetaData :300-356 */
    min_len = (flags & 2) ? br_u32(&r) : max_len;
    o->last_crc_ok = 1;
    if (o->calc_crc) { o->last_crc[0] = br_u32(&r); o->last_crc[1] = br_u32(&r); o->last_crc[2] = br_u32(&r); }   /* :340-355 */
    br_flush(&r);
    if (r.overrun || n == 0 || flags >= 256 || max_len > 65535 || min_len > max_len) return ERR_FORMAT;
    chunk_size += 1;
    if (chunk_size > cap) return ERR_CAP;
    if (o->work_cap < (u64)chunk_size + 8) { free(o->work); o->work_cap = (u64)chunk_size + 8; o->work = (u8*)malloc(o->work_cap); }
    if (o->recs_cap < n) { o->recs_cap = n; o->recs = (rec_t*)realloc(o->recs, (u64)n * sizeof(rec_t)); }
    len_bits = bit_length64((u64)(max_len - min_len));
#define NEED(k) do { if (pos + (u64)(k) > chunk_size) { rc = ERR_FORMAT; goto tags_done; } } while (0)
    {   /* ReadTags :503-570 */
        u8* ob = o->work; dfield_t* fl = NULL; u32 nf = 0;
        u32 min_title = 0, max_title = 0, tl_bits = 0, nsym = 0; u8 symbols[128]; hufd_t* raw_h = NULL;
        i32 prev[TAG_MAX_FIELDS + 1];
        memset(prev, 0, sizeof(prev));
        if (!(flags & 4)) {  /* TagTokenizerDecoder::ReadFields (TagModeler.cpp:893-1003) */
            nf = br_byte(&r);
            fl = (dfield_t*)calloc(nf + 1, sizeof(dfield_t));
            for (u32 i = 0; i < nf && !r.overrun; ++i) {
                dfield_t* f = &fl[i];
                f->sep = br_byte(&r); f->is_constant = br_byte(&r) != 0;
                if (f->is_constant) { f->len = br_u32(&r); if (f->len >= 1024) { rc = ERR_FORMAT; break; } f->data = (u8*)malloc(f->len + 1); for (u32 j = 0; j < f->len; ++j) f->data[j] = br_byte(&r); continue; }
                f->is_numeric = br_byte(&r) != 0;
                if (f->is_numeric) {
                    i32 maxv, maxd = 0;
                    f->scheme = br_byte(&r); f->min_value = (i32)br_u32(&r); maxv = (i32)br_u32(&r);
                    f->bits_value = bit_length64((u64)(int64_t)(i32)((u32)maxv - (u32)f->min_value)); f->bits_num = f->bits_value;
                    if (f->scheme >= 3 && f->scheme <= 5) {
                        f->min_delta = (i32)br_u32(&r); maxd = (i32)br_u32(&r);
                        f->bits_num = bit_length64((u64)(int64_t)(i32)((u32)maxd - (u32)f->min_delta));
                    } else if (f->scheme != 1 && f->scheme != 2) { rc = ERR_FORMAT; break; }
                    if (f->scheme == 3 || f->scheme == 1) { f->var_stat = br_byte(&r); if (f->var_stat) { f->hg = (hufd_t*)malloc(sizeof(hufd_t)); hufd_load(f->hg, &r); } }
                    continue;
                }
                f->is_len_constant = br_byte(&r) != 0;
                f->len = br_u32(&r); f->max_len = br_u32(&r); f->min_len = br_u32(&r);
                if (f->len >= 1024 || f->max_len >= 1024 || f->min_len > f->max_len) { rc = ERR_FORMAT; break; }
                f->bits_len = bit_length64((u64)(f->max_len - f->min_len));
                f->data = (u8*)malloc(f->len + 1); for (u32 j = 0; j < f->len; ++j) f->data[j] = br_byte(&r);
                f->ham = (u8*)malloc(f->len + 1); for (u32 j = 0; j < f->len; ++j) f->ham[j] = (u8)br_bit(&r);
                br_flush(&r);
                f->hl = (hufd_t**)calloc(TAG_STAT_LEN + 1, sizeof(hufd_t*));
                for (u32 j = 0; j < (f->max_len < TAG_STAT_LEN ? f->max_len : TAG_STAT_LEN); ++j)
                    if (j >= f->len || !f->ham[j]) { f->hl[j] = (hufd_t*)malloc(sizeof(hufd_t)); hufd_load(f->hl[j], &r); }
                if (f->max_len >= TAG_STAT_LEN) { f->hl[TAG_STAT_LEN] = (hufd_t*)malloc(sizeof(hufd_t)); hufd_load(f->hl[TAG_STAT_LEN], &r); }
            }
        } else {             /* TagRawDecoder::StartDecoding (TagModeler.cpp:1288-1312) */
            min_title = br_u32(&r); max_title = br_u32(&r); tl_bits = bit_length64((u64)(max_title - min_title));
            memset(symbols, 255, sizeof(symbols));
            for (u32 i = 0; i < 128; ++i) if (br_bit(&r)) symbols[nsym++] = (u8)i;
            raw_h = (hufd_t*)malloc(sizeof(hufd_t)); hufd_load(raw_h, &r);
        }
        for (u32 k = 0; k < n && rc == 0 && !r.overrun; ++k) {
            rec_t* rec = &o->recs[k]; u32 tl = 0; u8* t = ob + pos;
            rec->title = (u32)pos;
            if (!(flags & 4)) { /* DecodeNextFields :1006-1064, ReadNumericField :1066-1166 */
                for (u32 j = 0; j < nf; ++j) {
                    dfield_t* f = &fl[j];
                    if (f->is_constant) { NEED(tl + f->len + 1); memcpy(t + tl, f->data, f->len); tl += f->len; t[tl++] = f->sep; continue; }
                    if (f->is_numeric) {
                        u32 v = 0;
                        if (k == 0) { v = br_bits(&r, f->bits_value); if (f->scheme == 2) { f->rle_len = br_bits(&r, 8); f->rle_sym = (i32)v; } v += (u32)f->min_value; }
                        else switch (f->scheme) {
                            case 5: v = (u32)prev[j] + (u32)f->min_delta; break;
                            case 4: case 2:
                                if ((f->scheme == 4 && k == 1) || f->rle_len == 0) { v = br_bits(&r, f->bits_num); f->rle_sym = (i32)v; f->rle_len = br_bits(&r, 8); }
                                else { f->rle_len--; v = (u32)f->rle_sym; }
                                v += f->scheme == 4 ? (u32)prev[j] + (u32)f->min_delta : (u32)f->min_value;
                                break;
                            default:
                                v = f->hg ? hufd_get(f->hg, &r) : br_bits(&r, f->bits_num);
                                v += f->scheme == 3 ? (u32)prev[j] + (u32)f->min_delta : (u32)f->min_value;
                        }
                        { u8 num[12]; u32 nl = num_to_str(num, v); NEED(tl + nl + 1); memcpy(t + tl, num, nl); tl += nl; }
                        prev[j] = (i32)v; t[tl++] = f->sep; continue;
                    }
                    {
                        u32 fl_len = f->is_len_constant ? f->len : br_bits(&r, f->bits_len) + f->min_len;
                        if (fl_len > f->max_len) { rc = ERR_FORMAT; break; }
                        NEED(tl + fl_len + 1);
                        for (u32 x = 0; x < fl_len; ++x) {
                            if (x < f->len && f->ham[x]) t[tl++] = f->data[x];
                            else { hufd_t* h = f->hl[x < TAG_STAT_LEN ? x : TAG_STAT_LEN]; if (!h) { rc = ERR_FORMAT; break; } t[tl++] = (u8)hufd_get(h, &r); }
                        }
                        t[tl++] = f->sep;
                    }
                }
                if (rc) break;
                tl--;
            } else {            /* TagRawDecoder::DecodeNextFields :1314-1333 */
                tl = tl_bits > 0 ? br_bits(&r, tl_bits) + min_title : max_title;
                NEED(tl + 1);
                for (u32 i = 0; i < tl; ++i) t[i] = symbols[hufd_get(raw_h, &r) & 127];
            }
            rec->title_len = (u16)tl;
            pos += tl; NEED(1); ob[pos++] = '\n';
            rec->qua_len = (u16)(len_bits > 0 ? br_bits(&r, len_bits) + min_len : max_len);
            rec->seq_len = rec->qua_len; rec->seq = (u32)pos; pos += rec->seq_len;
            NEED(2); ob[pos++] = '\n'; ob[pos++] = '+';
            if (o->plus_rep) { NEED(tl); memcpy(ob + pos, ob + rec->title + 1, tl - 1); pos += tl - 1; }
            NEED(1); ob[pos++] = '\n';
            rec->qua = (u32)pos; pos += rec->qua_len;
            NEED(1); ob[pos++] = '\n';
        }
        br_flush(&r);
tags_done:
        if (fl) dfields_free(fl, nf);
        free(raw_h);
        if (rc) return rc;
        if (r.overrun) return ERR_FORMAT;
    }
#undef NEED
    {   /* ReadQuality :497 */
        u8 scheme = br_byte(&r);
        if (scheme != 255) {
            if (o->qua_order == 0) {
                if (scheme == 2) rc = qua_rle_decode(o, &r, n);
                else if (scheme < 2) rc = qua_position_decode(o, &r, n, scheme == 1);
                else rc = ERR_FORMAT;
            } else rc = qua_order_decode(o, &r, n, scheme);
            if (rc) return rc;
        }
    }
    rc = dna_decode(o, &r, n);                                /* ReadDNA :491 */
    if (rc) return rc;
    for (u32 k = 0; k < n; ++k) {                             /* ProcessBackward */
        rec_t* rec = &o->recs[k]; u8* s = o->work + rec->seq; u8* q = o->work + rec->qua; i32 si = (i32)rec->seq_len - 1;
        for (i32 i = (i32)rec->qua_len - 1; i >= 0; --i) {
            u32 qv = q[i], sv;
            if (qv >= 128) { sv = (qv - 128 + 16) / 8 + 3 - 1; qv &= 7; }
            else { if (si < 0) return ERR_FORMAT; sv = s[si--]; }
            s[i] = sv < 19 ? (u8)DNA_ALPHABET[sv] : 255;
            q[i] = (u8)(o->qoff + qv);
        }
    }
    if (o->calc_crc) { /* BlockCompressor::VerifyChecksum :576-594 */
        u32 crc[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        for (u32 k = 0; k < n; ++k) {
            const rec_t* rec = &o->recs[k];
            crc[0] = crc32_update(crc[0], o->work + rec->title, rec->title_len);
            crc[1] = crc32_update(crc[1], o->work + rec->seq, rec->qua_len);
            crc[2] = crc32_update(crc[2], o->work + rec->qua, rec->qua_len);
        }
        o->last_crc_ok = (~crc[0] == o->last_crc[0]) && (~crc[1] == o->last_crc[1]) && (~crc[2] == o->last_crc[2]);
    }
    memcpy(out, o->work, chunk_size);
    return (int64_t)chunk_size;
}

/* ------------------------------------------------------------------------------------------------
 * block cutter, first-chunk analysis, container
 * ---------------------------------------------------------------------------------------------- */
static void skip_to_eol(const u8* d, u64* pos, u64 size, int* crlf) /* FastqStream.h:74-89 */
{
    while (*pos < size && d[*pos] != '\n' && d[*pos] != '\r') ++*pos;
    if (*pos < size && d[*pos] == '\r' && *pos + 1 < size && d[*pos + 1] == '\n') { *crlf = 1; ++*pos; }
}
static u64 next_record_pos(const u8* d, u64 pos, u64 size, int* crlf) /* FastqStream.cpp:74-98 */
{
    u64 pos0;
    skip_to_eol(d, &pos, size, crlf); ++pos;
    while (pos < size && d[pos] != '@') { skip_to_eol(d, &pos, size, crlf); ++pos; }
    pos0 = pos;
    skip_to_eol(d, &pos, size, crlf); ++pos;
    if (pos < size && d[pos] == '@') return pos;
    return pos0;
}
u64 dsrc_oracle_cut_blocks(const u8* file, u64 size, u64 cbuf, u64* off, u64* len, u64 max_blocks)
{
    /* ReadNextChunk (FastqStream.cpp:18-72): buffer = carry-over + fresh bytes == file[p, p+cbuf) */
    u64 p = 0, nb = 0; int crlf = 0, eof = 0, tail_only = 0;
    while (!eof) {
        u64 avail = size - p, blk_len;
        /* Read() returned r bytes; the carry-over (p..) is part of the cbuf window */
        if (!tail_only && avail >= cbuf && cbuf > 8192) {   /* r == toRead: somewhere before the end */
            u64 end = next_record_pos(file + p, cbuf - 8192, cbuf, &crlf);
            blk_len = end - 1 - (crlf ? 1 : 0);
            if (nb < max_blocks) { off[nb] = p; len[nb] = blk_len; }
            nb++; p += end;
            /* the window ended exactly at EOF: the next Read returns 0 (the `else` of :66-69), the chunk is the
               carry-over as it stands -- size = bufferSize, no "- 1", no CRLF adjustment */
            if (avail == cbuf) tail_only = 1;
            if (p == size) eof = 1;
        } else {                                      /* at the end of file: r < toRead (:57-64), or r == 0 after an exact window */
            u64 drop = tail_only ? 0 : 1 + (crlf ? 1u : 0u);
            if (avail == 0) break;
            blk_len = avail > drop ? avail - drop : 0;
            if (nb < max_blocks) { off[nb] = p; len[nb] = blk_len; }
            nb++; eof = 1;
        }
    }
    return nb;
}

int dsrc_oracle_analyze(const u8* m, u64 size, u32* qoff, int* plus_rep, int* color_space) /* FastqParser.cpp:27-138 */
{
    u64 pos = 0, skipped = 0; u32 recs = 0; u8 minq = 255, maxq = 0; int estimate = *qoff == 0;
    *plus_rep = 0; *color_space = 0;
    while (pos < size) {
        u64 t = pos, s, p, q; u32 tl = skip_line(m, size, &pos, &skipped), sl, pl, ql; int prep, cenc;
        if (tl == 0 || m[t] != '@') break;
        s = pos; sl = skip_line(m, size, &pos, &skipped); if (sl == 0) break;
        p = pos; pl = skip_line(m, size, &pos, &skipped); prep = pl > 1; if (m[p] != '+') break;
        q = pos; ql = skip_line(m, size, &pos, &skipped);
        if (estimate) { for (u32 i = 0; i < ql; ++i) { if (m[q + i] < minq) minq = m[q + i]; if (m[q + i] > maxq) maxq = m[q + i]; } }
        else if (ql == 0) break;
        cenc = (m[s + 1] >= '0' && m[s + 1] <= '3') || m[s + 1] == '.';
        if (recs != 0) { if (*color_space != cenc) return 0; if (*color_space && m[s] >= '0' && m[s] <= '3') return 0; if (*plus_rep != prep) return 0; }
        else { *plus_rep = prep; *color_space = cenc; }
        recs++;
    }
    if (estimate) {
        if (maxq <= 74) { if (minq >= 33) *qoff = 33; }
        else if (maxq <= 105) { if (minq >= 64) *qoff = 64; else if (minq >= 59) *qoff = 59; }
        if (*qoff == 0) { if (minq >= 33) *qoff = 33; else return 0; }
    }
    return recs > 1;
}

static void put_be32(u8* p, u32 v) { p[0] = (u8)(v >> 24); p[1] = (u8)(v >> 16); p[2] = (u8)(v >> 8); p[3] = (u8)v; }
static void put_be64(u8* p, u64 v) { put_be32(p, (u32)(v >> 32)); put_be32(p + 4, (u32)v); }
static u32 get_be32(const u8* p) { return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | p[3]; }
static u64 get_be64(const u8* p) { return ((u64)get_be32(p) << 32) | get_be32(p + 4); }

static int g_archive_crc = 0;   /* set by dsrc_oracle_compress_mem_crc around its call */
int64_t dsrc_oracle_write_archive(const u8* blocks, const u32* sizes, u64 nb, u32 qoff, int plus_rep, int color_space,
                                  u32 dna_order, u32 qua_order, u8* out, u64 cap)
{
    /* DsrcFile.h:26-47, DsrcFile.cpp:75-170 */
    u64 total = 0, foot_off, foot_size = 1 + nb * 4 + 2 + 3 + 8, p;
    for (u64 i = 0; i < nb; ++i) total += sizes[i];
    foot_off = 40 + total;
    if (foot_off + foot_size > cap) return ERR_CAP;
    out[0] = 0xAA; out[1] = 2; out[2] = 0; out[3] = 2;
    put_be32(out + 4, (u32)foot_size); put_be64(out + 8, foot_off); put_be64(out + 16, 0); put_be64(out + 24, nb);
    memset(out + 32, 0xAA, 8);
    memcpy(out + 40, blocks, total);
    p = foot_off;
    out[p++] = 0xCC;
    for (u64 i = 0; i < nb; ++i) { u32 v = sizes[i]; out[p++] = (u8)v; out[p++] = (u8)(v >> 8); out[p++] = (u8)(v >> 16); out[p++] = (u8)(v >> 24); } /* host-endian (LE) */
    out[p++] = (u8)((plus_rep ? 1 : 0) | (color_space ? 2 : 0));
    out[p++] = (u8)qoff;
    out[p++] = (u8)(g_archive_crc ? 2 : 0);   /* compFlags: bit0 lossy, bit1 crc (DsrcFile.cpp:155-165) */
    out[p++] = (u8)dna_order; out[p++] = (u8)qua_order;
    put_be64(out + p, 0); p += 8; /* tagPreserveFlags */
    return (int64_t)p;
}

int64_t dsrc_oracle_compress_mem(const u8* file, u64 size, u32 dna_level, u32 qua_level, u64 buf_bytes, u32 qoff, u8* out, u64 cap)
{
    u64 nb = dsrc_oracle_cut_blocks(file, size, buf_bytes, NULL, NULL, 0), total = 0;
    u64* off = (u64*)malloc((nb + 1) * 8); u64* len = (u64*)malloc((nb + 1) * 8); u32* sizes = (u32*)malloc((nb + 1) * 4);
    u8* tmp = (u8*)malloc(size + size / 2 + 4096); int plus_rep = 0, cs = 0; int64_t res = 0; dsrc_oracle_t* o;
    dsrc_oracle_cut_blocks(file, size, buf_bytes, off, len, nb);
    if (nb == 0 || !dsrc_oracle_analyze(file + off[0], len[0], &qoff, &plus_rep, &cs) || cs) { res = ERR_FORMAT; goto done; }
    o = dsrc_oracle_create(qoff, plus_rep, dna_level * 3, qua_level); /* DsrcOperator.h:74-90 */
    o->calc_crc = g_archive_crc;
    for (u64 b = 0; b < nb; ++b) {
        int64_t s = dsrc_oracle_store(o, file + off[b], len[b], tmp + total, size + size / 2 + 4096 - total, NULL, NULL);
        if (s < 0) { res = s; break; }
        sizes[b] = (u32)s; total += (u64)s;
    }
    dsrc_oracle_destroy(o);
    if (res == 0) res = dsrc_oracle_write_archive(tmp, sizes, nb, qoff, plus_rep, cs, dna_level * 3, qua_level, out, cap);
done:
    free(off); free(len); free(sizes); free(tmp);
    return res;
}

int64_t dsrc_oracle_compress_mem_crc(const u8* file, u64 size, u32 dna_level, u32 qua_level, u64 buf_bytes, u32 qoff, int crc, u8* out, u64 cap)
{
    int64_t r;
    g_archive_crc = crc != 0;
    r = dsrc_oracle_compress_mem(file, size, dna_level, qua_level, buf_bytes, qoff, out, cap);
    g_archive_crc = 0;
    return r;
}

int64_t dsrc_oracle_decompress_mem(const u8* arc, u64 size, u8* out, u64 cap)
{
    /* DsrcFileReader::StartDecompress / ReadFileHeader / ReadFileFooter (DsrcFile.cpp:186-314) */
    u64 foot_off, nb, p, total = 0; u32 foot_size, qoff, dna_order, qua_order; int plus_rep; const u8* f; dsrc_oracle_t* o; int64_t res = 0;
    if (size < 40 || arc[0] != 0xAA || arc[1] != 2 || arc[2] != 0) return ERR_FORMAT;
    foot_size = get_be32(arc + 4); foot_off = get_be64(arc + 8); nb = get_be64(arc + 24);
    if (nb == 0 || foot_off + foot_size > size || foot_size < 1 + nb * 4 + 13) return ERR_FORMAT;
    f = arc + foot_off;
    if (f[0] != 0xCC) return ERR_FORMAT;
    p = 1 + nb * 4;
    plus_rep = f[p] & 1; if (f[p] & 2) return ERR_UNSUPPORTED; qoff = f[p + 1];
    if (f[p + 2] & 1) return ERR_UNSUPPORTED;
    dna_order = f[p + 3]; qua_order = f[p + 4];
    o = dsrc_oracle_create(qoff, plus_rep, dna_order, qua_order);
    o->calc_crc = (f[p + 2] & 2) != 0;
    p = 40;
    for (u64 b = 0; b < nb; ++b) {
        u32 bs = (u32)f[1 + b * 4] | ((u32)f[2 + b * 4] << 8) | ((u32)f[3 + b * 4] << 16) | ((u32)f[4 + b * 4] << 24);
        int64_t s;
        if (p + bs > foot_off) { res = ERR_FORMAT; break; }
        s = dsrc_oracle_read(o, arc + p, bs, out + total, cap - total);
        if (s < 0) { res = s; break; }
        if (!o->last_crc_ok) { res = ERR_FORMAT; break; }
        total += (u64)s; p += bs;
    }
    dsrc_oracle_destroy(o);
    return res < 0 ? res : (int64_t)total;
}
