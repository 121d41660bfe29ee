// Order-k range-coder modelers for DNA and quality, re-designed for the GPU.
//
// Reference: TDnaRCOrderModeler (src/DnaModelerRCO.h:45-62,94-133), TQualityOrderModeler +
// TTranslationalQualityEncoder + TQualityModelExt (src/QualityOrderModeler.h:36-51, src/QualityEncoder.h:77-136,
// 281-342), TSymbolCoderRC (src/SymbolCoderRC.h:24-93), RangeEncoder (src/RangeCoder.h:51-84).
//
// The reference codes symbol after symbol: look the context row up in a 32 KB .. 64 MB table, sum it, emit,
// bump one counter. On a GPU that is one DRAM round trip per symbol. Two observations split the work:
//   (1) the context of symbol i is a pure function of the INPUT (the previous <= 5 symbols and the read
//       position), never of the coder state -- so all contexts of a block are computable in parallel;
//   (2) the adaptive row a symbol sees is determined by the earlier symbols of the SAME context only:
//       freq = 1 + 2*#(same ctx, same sym before), cum/tot likewise, with the halving rescale of
//       TSymbolCoderRC::Rescale applied when tot crosses 2^16 - 2N.
// So the MODEL (the exact (freq, cum, tot) triple each symbol would meet) is computed block-parallel by k_model, one CTA per
// block, with one of three engines -- the shared-memory direct engine for 4-symbol DNA (model_dna.cuh), the tile/table engine
// with the scan engine for rows of <= 16 symbols (model_tab.cuh), the sort engine in this file for larger alphabets
// (stable LSD radix sort by context in HBM scratch, then every context is a contiguous run in original order) -- and what
// remains serial is only RangeEncoder::EncodeFrequency over the triples: k_rc_encode, one thread per (block, stream).
// DESIGN.md section 4 has the full account.
#include "common.cuh"
#include "kernels.h"
#include <cstddef>
#include <cstdlib>

#ifndef MODEL_REGS
#define MODEL_REGS 56
#endif
#define SORT_MAX_BITS 10                      // radix digits of at most 10 bits: 8 warps x 1024 counters = 32 KB
#define LONG_T 48                             // context runs longer than this are walked by a whole warp
#define SCAN_TILE 2048                        // sorted elements staged in shared memory per step of the run walker
#define SCAN_LOOK (LONG_T + 1)
#define FULL 0xFFFFFFFFu

struct ModelCfg {
    u32 alpha, bits, key_bits, sym_order, rescale, ord;   // ord: DNA order; sym_order/rescale: quality
};

// QualityOrderModelerProxyLossless::SelectSchemeId (src/QualityModelerProxy.h:261-282)
__device__ u32 quality_order_scheme(const BlockState& st, u32 order)
{
    u32 sc = 255;
    for (u32 i = 0; i < 8; ++i) if ((16u << i) >= st.q_count) { sc = i; break; }
    if (sc != 255 && order == 2) {
        double ratio = __ddiv_rn((double)st.raw_len, (double)st.rle_len);
        if (st.max_len == st.min_len && ratio > 1.175) sc += 4;
    }
    return sc;
}
// template arguments per scheme (src/QualityModelerProxy.h:231-254)
__device__ bool quality_cfg(u32 order, u32 scheme, ModelCfg& c)
{
    if (scheme > 7) return false;
    const u32 k = scheme & 3;
    c.alpha = 16u << k; c.bits = 4 + k;
    c.sym_order = order == 1 ? (k == 0 ? 3 : k == 1 ? 2 : 1) : (k == 0 ? 4 : k == 1 ? 3 : k == 2 ? 2 : 1);
    c.rescale = (scheme & 4) ? c.alpha : 8;
    c.key_bits = c.bits * (c.sym_order + 1);
    c.ord = 0;
    return true;
}
// DnaOrderModelerProxy (src/DnaModelerProxy.h:160-227)
__device__ void dna_cfg(u32 order, u32 scheme, ModelCfg& c)
{
    c.alpha = scheme == 0 ? 4 : 8; c.bits = scheme == 0 ? 2 : 3;
    c.ord = (scheme == 1 && order > 7) ? 7 : order;
    c.key_bits = c.ord * c.bits; c.sym_order = 0; c.rescale = 0;
}

#define TRIP(f, cum, tot) ((u64)(f) | ((u64)(cum) << 16) | ((u64)(tot) << 32))
#define PROF_MARK(slot) do { if (ws.prof && threadIdx.x == 0) { long long t_ = clock64(); atomicAdd((unsigned long long*)&ws.prof[slot], (unsigned long long)(t_ - prof_t)); prof_t = t_; } } while (0)
#include "model_tab.cuh"
#include "model_dna.cuh"

union ModelSortShared {
    u32 H[DSRC_WARPS << SORT_MAX_BITS];        // sort: per-warp digit counters, then scatter offsets
    union {
        struct { u64 tile[SCAN_TILE + SCAN_LOOK]; u16 heads[SCAN_TILE]; u8 cnt[16384]; } t;   // short-run walker
        struct { u32 B[DSRC_WARPS][128]; u32 P[DSRC_WARPS][128]; } l;                        // long-run walker: per-warp row state
    } g;
};
// small state first, the per-engine scratch last: the quality kernel never uses the DNA direct engine's 52 KiB and is launched
// with the smaller footprint (one more CTA per SM)
struct ModelShared {
    u32 scan[DSRC_WARPS + 1];
    u32 n_long, n_heads;
    u8 rank[256];
    ModelCfg cfg;
    u32 M, ok;
    u32 blk, tab_id;                               // block fetched from the launch's queue; table taken from the context's pool
    union {
        u32 H[DSRC_WARPS << SORT_MAX_BITS];
        union {
            struct { u64 tile[SCAN_TILE + SCAN_LOOK]; u16 heads[SCAN_TILE]; u8 cnt[16384]; } t;
            struct { u32 B[DSRC_WARPS][128]; u32 P[DSRC_WARPS][128]; } l;
        } g;
        TabShared tab;                             // tile/table engine (model_tab.cuh)
        DnaDirectShared dna;                       // shared-memory table engine of the 4-symbol DNA model (model_dna.cuh)
    } u;
};
static_assert(offsetof(TabShared, part) >= sizeof(ModelSortShared), "PartState must survive sort_pass / group_scan on an oversize partition");
#define MODEL_SMEM_QUALITY (offsetof(ModelShared, u) + (sizeof(TabShared) > sizeof(ModelSortShared) ? sizeof(TabShared) : sizeof(ModelSortShared)))
// the launch without the partition engine never touches TabShared's tail (pB, pP, part): it keeps the smaller footprint, which is what
// lets the range-coder CTAs of the other streams' batches stay resident beside four model CTAs per SM
#define MODEL_SMEM_QUALITY_CLASSIC (offsetof(ModelShared, u) + (offsetof(TabShared, pB) > sizeof(ModelSortShared) ? offsetof(TabShared, pB) : sizeof(ModelSortShared)))


// ---- element sources of a sort pass. An element is (ctx << 40) | (sym << 32) | index. A warp walks consecutive rows of
// 32 symbols; the symbols a context needs from before the row come from the neighbouring lanes / the previous row.
struct TabShared;
struct FetchSorted {
    typedef u64 Raw;
    static const bool TILE8 = false;
    __device__ __forceinline__ void tile8(TabShared&, u32, u32) {}
    const u64* src;
    __device__ __forceinline__ void begin(u32) {}
    __device__ __forceinline__ void prefetch(u32, u32) {}
    __device__ __forceinline__ Raw ld(u32 i, bool in) { return in ? src[i] : 0ull; }
    __device__ __forceinline__ u64 mk(Raw r, u32, bool) { return r; }
};
// quality context (TQualityModelBase::UpdateHash/GetHash, QualityEncoder.h:77-94; position bucket :307)
struct FetchQ {
    typedef u32 Raw;
    static const bool TILE8 = true;
    template <bool SCR> __device__ __forceinline__ void tile8x(TabShared& S, u32 t0, u32 n);
    __device__ __forceinline__ void tile8(TabShared& S, u32 t0, u32 n) { tile8x<false>(S, t0, n); }
    u32 kmul, kmask;                                  // tile8x<true>: the row index is scrambled, key' = (key * kmul) & kmask (partition engine)
    const u8* q; const u8* pctx; const u8* rank; u32 so, h, bits, M; u32 prev;
    const u8* plut; u32 fixed_len, jpos;              // fixed_len != 0: position bucket from the shared-memory table, indexed by i % len
    u32 ebits, pbits;                                 // tile8: bits per hash slot / per position bucket of the compact row index
    __device__ __forceinline__ void begin(u32 start)
    {
        const u32 j = start - 32 + lane_id(); prev = (start >= 32 && j < M) ? rank[q[j]] : 0u;
        if (fixed_len) jpos = (start + lane_id()) % fixed_len;
    }
    __device__ __forceinline__ void prefetch(u32 start, u32 t)      // 2 KB of symbols (and of position buckets) = 16 lines each
    {
        const u32 i = start + (t & 15u) * 128;
        if (t < 16) { if (i < M) asm volatile("prefetch.global.L2 [%0];" :: "l"(q + i)); }
        else if (t < 32 && !fixed_len) { if (i < M) asm volatile("prefetch.global.L2 [%0];" :: "l"(pctx + i)); }
    }
    __device__ __forceinline__ Raw ld(u32 i, bool in)
    {
        u32 pc;
        if (fixed_len) { pc = plut[jpos]; jpos += 32; while (jpos >= fixed_len) jpos -= fixed_len; }
        else pc = in ? pctx[i] : 0u;
        return in ? ((u32)q[i] | (pc << 8)) : 0u;
    }
    __device__ __forceinline__ u64 mk(Raw raw, u32 i, bool in)
    {
        const u32 ln = lane_id();
        const u32 r0 = in ? rank[raw & 255u] : 0u;
        u32 y[6]; y[0] = r0;                          // y[k] = symbol k steps back (0 before the block start)
#pragma unroll
        for (int k = 1; k < 6; ++k) {
            const u32 a = __shfl_up_sync(FULL, r0, k), b = __shfl_sync(FULL, prev, (ln + 32 - k) & 31);
            y[k] = ln >= (u32)k ? a : b;
        }
        prev = r0;
        u32 hash = 0;                                 // raw symbols below slot h, pairwise means from slot h on
        if (so == 1) hash = y[1];
        else {
#pragma unroll
            for (int t = 0; t < 4; ++t) if ((u32)t < so) { const u32 v = (u32)t < h ? y[t + 1] : ((y[t + 1] + y[t + 2]) >> 1); hash |= v << (t * bits); }
        }
        if (!in) return 0ull;
        const u32 ctx = (hash << bits) | (raw >> 8);
        return ((u64)ctx << 40) | ((u64)r0 << 32) | i;
    }
};
// DNA context: the previous `ord` symbols (TDnaRCOrderModeler, DnaModelerRCO.h:45-62)
struct FetchD {
    typedef u32 Raw;
    static const bool TILE8 = false;
    __device__ __forceinline__ void tile8(TabShared&, u32, u32) {}
    const u8* sq; u32 ord, bits, M; u32 prev;
    __device__ __forceinline__ void begin(u32 start) { const u32 j = start - 32 + lane_id(); prev = (start >= 32 && j < M) ? sq[j] : 0u; }
    __device__ __forceinline__ void prefetch(u32 start, u32 t) { const u32 i = start + t * 128; if (t < 16 && i < M) asm volatile("prefetch.global.L2 [%0];" :: "l"(sq + i)); }
    __device__ __forceinline__ Raw ld(u32 i, bool in) { return in ? sq[i] : 0u; }
    __device__ __forceinline__ u64 mk(Raw r0, u32 i, bool in)
    {
        const u32 ln = lane_id(), mask = (1u << bits) - 1;
        u32 ctx = 0;
        for (u32 t = 0; t < ord; ++t) {
            const u32 k = t + 1;
            const u32 a = __shfl_up_sync(FULL, r0, k), b = __shfl_sync(FULL, prev, (ln + 32 - k) & 31);
            ctx |= ((ln >= k ? a : b) & mask) << (t * bits);
        }
        prev = r0;
        return in ? (((u64)ctx << 40) | ((u64)r0 << 32) | i) : 0ull;
    }
};

// Contexts of one tile of the tile/table engine for 16-symbol quality rows (bits == 4): every thread takes 8 consecutive
// symbols, the 5 symbols of history they need come with one more 8-byte load, so the hash window slides through registers
// (no shuffles, two loads per thread). Same arithmetic as FetchQ::mk, except that the row index is packed with as many
// bits per hash slot as the block's symbol count needs (symbols are dense ranks, their pairwise means are no larger) and
// log2(rescale) bits of position bucket: the index only has to be injective -- the table is private to the block -- and a
// 5-symbol block then sorts 16-bit keys (two 8-bit passes, 256 bins) instead of 20-bit ones and keeps its rows within 2 MB.
template <bool SCR>
__device__ __forceinline__ void FetchQ::tile8x(TabShared& S, u32 t0, u32 n)
{
    const u32 p = threadIdx.x * 8, i = t0 + p;
    if (p >= n) return;
    const u64 cur = *(const u64*)(q + i);            // the arena has slack behind M
    const u64 prv = i ? *(const u64*)(q + i - 8) : 0ull;
    u32 r[13];
#pragma unroll
    for (int k = 0; k < 5; ++k) r[k] = i ? (u32)rank[(u32)(prv >> (8 * (3 + k))) & 255u] : 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) r[5 + k] = rank[(u32)(cur >> (8 * k)) & 255u];
    u64 pcw = 0; u32 jp = 0;
    if (fixed_len) jp = i % fixed_len; else pcw = *(const u64*)(pctx + i);
    u32 el[8];
    // hash slots (TQualityModelBase::UpdateHash, QualityEncoder.h:77-89): slot t holds the symbol t+1 steps back for t < h, the mean
    // of the symbols t+1 and t+2 steps back from slot h on. The two orders that exist for 16-symbol rows are written out with the
    // slot shifts as multipliers (one multiply-add per slot and symbol); anything else takes the generic loop.
    u32 a[12];                                        // a[i] = mean of r[i], r[i+1]
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = (r[i] + r[i + 1]) >> 1;
    const u32 M1 = 1u << ebits, M2 = 1u << (2 * ebits), M3 = 1u << (3 * ebits), MP = 1u << pbits;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        u32 pc;
        if (fixed_len) { pc = plut[jp]; jp = jp + 1 == fixed_len ? 0u : jp + 1; } else pc = (u32)(pcw >> (8 * k)) & 255u;
        u32 hash;
        if (so == 4 && h == 2) hash = r[4 + k] + r[3 + k] * M1 + a[1 + k] * M2 + a[k] * M3;
        else if (so == 3 && h == 1) hash = r[4 + k] + a[2 + k] * M1 + a[1 + k] * M2;
        else if (so == 2) hash = r[4 + k] + a[2 + k] * M1;
        else if (so == 1) hash = r[4 + k];            // SymbolOrder 1: the swap slot is slot 0 and symBuffer stays 0 (QualityEncoder.h:77-89)
        else {
            hash = 0;                                 // y[t] = r[5 + k - t]: symbol t steps back
#pragma unroll
            for (int t = 0; t < 4; ++t) if ((u32)t < so) {
                const u32 v = (u32)t < h ? r[4 + k - t] : a[3 + k - t];
                hash |= v << (t * ebits);
            }
        }
        u32 key = hash * MP + pc;
        if (SCR) key = (key * kmul) & kmask;
        el[k] = (key << TT_SHIFT) | (p + k);
    }
    ((uint4*)S.el[0])[threadIdx.x * 2] = make_uint4(el[0], el[1], el[2], el[3]);
    ((uint4*)S.el[0])[threadIdx.x * 2 + 1] = make_uint4(el[4], el[5], el[6], el[7]);
    *(uint2*)&S.sym[p] = make_uint2(r[5] | (r[6] << 8) | (r[7] << 16) | (r[8] << 24), r[9] | (r[10] << 8) | (r[11] << 16) | (r[12] << 24));
}

// One stable counting-sort pass on the `bits`-wide digit at `shift`. Every warp owns a contiguous eighth of the input:
// it histograms its part into its own counters, the counters are scanned in (digit, warp) order, and each warp
// scatters its part in order through its own running offsets -- no CTA barrier inside the two loops.
#define SORT_U 4                                // rows in flight per warp (memory-level parallelism)
template <class F>
__device__ void sort_pass(ModelShared& S, F f, u64* dst, u32 M, u32 shift, u32 bits)
{
    const u32 tid = threadIdx.x, w = warp_id(), ln = lane_id(), lt = (1u << ln) - 1;
    const u32 bins = 1u << bits, dmask = bins - 1;
    const u32 seg = (((M + DSRC_WARPS - 1) / DSRC_WARPS) + 31) & ~31u;
    const u32 wb = min(M, w * seg), we = min(M, wb + seg);
    u32* H = S.u.H + w * bins;
    for (u32 i = tid; i < DSRC_WARPS * bins; i += DSRC_CTA) S.u.H[i] = 0;
    __syncthreads();
    f.begin(wb);
    for (u32 r = wb; r < we; r += 32 * SORT_U) {
        typename F::Raw raw[SORT_U];
#pragma unroll
        for (int k = 0; k < SORT_U; ++k) { const u32 i = r + k * 32 + ln; raw[k] = f.ld(i, i < we); }
#pragma unroll
        for (int k = 0; k < SORT_U; ++k) {
            const u32 i = r + k * 32 + ln; const bool in = i < we;
            if (r + k * 32 >= we) break;
            const u64 e = f.mk(raw[k], i, in);
            const u32 d = (u32)(e >> shift) & dmask;
            const u32 peers = match_bits(d, bits, __ballot_sync(FULL, in));
            if (in && (__ffs(peers) - 1) == (int)ln) H[d] += __popc(peers);
            __syncwarp();
        }
    }
    __syncthreads();
    {
        const u32 per = bins > DSRC_CTA ? bins / DSRC_CTA : 1u, d0 = tid * per;
        u32 sum = 0;
        if (d0 < bins) for (u32 k = 0; k < per; ++k) for (u32 ww = 0; ww < DSRC_WARPS; ++ww) sum += S.u.H[ww * bins + d0 + k];
        u32 total, run = block_excl_sum(sum, S.scan, &total);
        if (d0 < bins) for (u32 k = 0; k < per; ++k) for (u32 ww = 0; ww < DSRC_WARPS; ++ww) { const u32 c = S.u.H[ww * bins + d0 + k]; S.u.H[ww * bins + d0 + k] = run; run += c; }
    }
    __syncthreads();
    f.begin(wb);
    for (u32 r = wb; r < we; r += 32 * SORT_U) {
        typename F::Raw raw[SORT_U];
#pragma unroll
        for (int k = 0; k < SORT_U; ++k) { const u32 i = r + k * 32 + ln; raw[k] = f.ld(i, i < we); }
#pragma unroll
        for (int k = 0; k < SORT_U; ++k) {
            const u32 i = r + k * 32 + ln; const bool in = i < we;
            if (r + k * 32 >= we) break;
            const u64 e = f.mk(raw[k], i, in);
            const u32 d = (u32)(e >> shift) & dmask;
            const u32 peers = match_bits(d, bits, __ballot_sync(FULL, in));
            const u32 pos = in ? H[d] + __popc(peers & lt) : 0u;
            __syncwarp();
            if (in) {
                if ((__ffs(peers) - 1) == (int)ln) H[d] += __popc(peers);
                dst[pos] = e;
            }
            __syncwarp();
        }
    }
    __syncthreads();
}

// A whole warp walks ONE long context run starting at sorted[a], 32 symbols per step. Within a row of 32 symbols of the
// same context the adaptive row seen by lane l is the row at the start of the row plus 2 x (the symbols of the lanes
// before l), so freq / cum / tot come from ballots and popcounts; the halving rescale of TSymbolCoderRC::Rescale
// (src/SymbolCoderRC.h:69-73) can only fire once per ~32 K symbols of a context: a row that could contain it is
// replayed symbol by symbol by lane 0.
__device__ void warp_run(ModelShared& S, const u64* sorted, u64* trip, u32 M, u32 a)
{
    const u32 N = S.cfg.alpha, limit = (1u << 16) - 2 * N, ln = lane_id(), lt = (1u << ln) - 1;
    u32* B = S.u.g.l.B[warp_id()]; u32* P = S.u.g.l.P[warp_id()];
    for (u32 s = ln; s < N; s += 32) B[s] = 1;
    u32 T = N;
    const u64 key = sorted[a] >> 40;
    u64 e_nx = a + ln < M ? sorted[a + ln] : ~0ull;
    __syncwarp();
    for (u32 row = a;; row += 32) {
        const u64 e = e_nx;
        e_nx = row + 32 + ln < M ? sorted[row + 32 + ln] : ~0ull;      // next row in flight while this one is processed
        const bool valid = (e >> 40) == key;
        const u32 vm = __ballot_sync(FULL, valid);
        if (!vm) break;
        const u32 nv = __popc(vm);
        const u32 sym = (u32)(e >> 32) & 255u;
        if (T + 2 * (nv - 1) >= limit) {
            if (ln == 0) {
                for (u32 j = 0; j < nv; ++j) {
                    const u64 ej = sorted[row + j]; const u32 s = (u32)(ej >> 32) & 255u;
                    if (T >= limit) { T = 0; for (u32 q = 0; q < N; ++q) { u32 c = B[q]; c -= c >> 1; B[q] = c; T += c; } }
                    const u32 f = B[s]; u32 cum = 0;
                    for (u32 q = 0; q < s; ++q) cum += B[q];
                    trip[(u32)ej] = TRIP(f, cum, T);
                    B[s] = f + 2; T += 2;
                }
            }
            T = __shfl_sync(FULL, T, 0);
            __syncwarp();
        } else {
            if (N <= 32) {
                const u32 v = ln < N ? B[ln] : 0u;
                const u32 inc = warp_incl_sum(v);
                if (ln < N) P[ln] = inc - v;
            } else {
                const u32 per = N / 32, b0 = ln * per; u32 s = 0;
                for (u32 k = 0; k < per; ++k) { const u32 t = B[b0 + k]; P[b0 + k] = s; s += t; }
                const u32 off = warp_incl_sum(s) - s;
                for (u32 k = 0; k < per; ++k) P[b0 + k] += off;
            }
            __syncwarp();
            u32 peers = 0;
            if (valid) peers = __match_any_sync(vm, sym);
            u32 c_lt = 0, rem = vm;
            while (rem) {                              // warp-uniform: one step per distinct symbol of the row
                const int leader = __ffs(rem) - 1;
                const u32 gs = __shfl_sync(FULL, sym, leader);
                const u32 gm = __shfl_sync(FULL, peers, leader);
                if (valid && gs < sym) c_lt += __popc(gm & lt);
                rem &= ~gm;
            }
            if (valid) trip[(u32)e] = TRIP(B[sym] + 2 * __popc(peers & lt), P[sym] + 2 * c_lt, T + 2 * __popc(vm & lt));
            __syncwarp();
            if (valid && (__ffs(peers) - 1) == (int)ln) B[sym] += 2 * __popc(peers);
            T += 2 * nv;
            __syncwarp();
        }
        if (nv < 32) break;
    }
}

// walk the context runs of the sorted array and emit the adaptive-model triple of every symbol
// (TSymbolCoderRC<N>::EncodeSymbol / Accumulate / Rescale, src/SymbolCoderRC.h:35-48, 69-90).
// The sorted array is staged through shared memory in tiles (coalesced loads); per tile the run heads are compacted,
// every short run (<= LONG_T symbols: no rescale possible, counters fit a byte) is walked by one thread out of shared
// memory, and the long runs are queued (in `longq`, global) for the warp-cooperative walker.
__device__ void group_scan(ModelShared& S, const u64* sorted, u32* longq, u64* trip, u32 M, const Workspace& ws, long long& prof_t, int prof_base)
{
    const u32 N = S.cfg.alpha;
    const u32 nthr = N <= 64 ? (u32)DSRC_CTA : (u32)DSRC_CTA / 2;         // 16 KB of byte counters: cnt[s * nthr + tid]
    const u32 tid = threadIdx.x, ln = lane_id(), lt = (1u << ln) - 1;
    u64* tile = S.u.g.t.tile; u16* heads = S.u.g.t.heads; u8* cnt = S.u.g.t.cnt + tid;
    if (tid == 0) S.n_long = 0;
    for (u32 t0 = 0; t0 < M; t0 += SCAN_TILE) {
        const u32 n = min((u32)SCAN_TILE, M - t0), avail = min(n + SCAN_LOOK, M - t0);
        __syncthreads();
        if (tid == 0) S.n_heads = 0;
        for (u32 p = tid; p < avail; p += DSRC_CTA) tile[p] = sorted[t0 + p];
        const u32 prevkey = t0 ? (u32)(sorted[t0 - 1] >> 40) : 0xFFFFFFFFu;
        __syncthreads();
        for (u32 p0 = 0; p0 < n; p0 += DSRC_CTA) {
            const u32 p = p0 + tid; const bool in = p < n;
            const u64 e = in ? tile[p] : ~0ull;
            const u32 key = (u32)(e >> 40);
            bool push = false;
            if (in && key != (p ? (u32)(tile[p - 1] >> 40) : prevkey)) {
                const u32 nk = p + 1 < avail ? (u32)(tile[p + 1] >> 40) : 0xFFFFFFFFu;
                if (nk != key) trip[(u32)e] = TRIP(1, (u32)(e >> 32) & 255u, N);          // run of one: fresh row, all ones
                else if (p + LONG_T < avail && (u32)(tile[p + LONG_T] >> 40) == key) longq[atomicAdd(&S.n_long, 1u)] = t0 + p;
                else push = true;
            }
            const u32 m = __ballot_sync(FULL, push);
            if (m) {
                u32 base = 0;
                const int leader = __ffs(m) - 1;
                if ((int)ln == leader) base = atomicAdd(&S.n_heads, (u32)__popc(m));
                base = __shfl_sync(FULL, base, leader);
                if (push) heads[base + __popc(m & lt)] = (u16)p;
            }
        }
        __syncthreads();
        const u32 nh = S.n_heads;
        if (tid < nthr) {
            for (u32 h = tid; h < nh; h += nthr) {
                u32 p = heads[h];
                u64 e = tile[p];
                const u32 key = (u32)(e >> 40);
                for (u32 s = 0; s < N; ++s) cnt[s * nthr] = 1;
                u32 tot = N;
                do {
                    const u32 s = (u32)(e >> 32) & 255u;
                    const u32 f = cnt[s * nthr];
                    u32 cum = 0;
                    if (s * 2 <= N) { for (u32 q = 0; q < s; ++q) cum += cnt[q * nthr]; }
                    else { u32 hi = 0; for (u32 q = s; q < N; ++q) hi += cnt[q * nthr]; cum = tot - hi; }
                    trip[(u32)e] = TRIP(f, cum, tot);
                    cnt[s * nthr] = (u8)(f + 2); tot += 2;
                    if (++p >= avail) break;
                    e = tile[p];
                } while ((u32)(e >> 40) == key);
            }
        }
    }
    __syncthreads();
    PROF_MARK(prof_base + 3);
    const u32 n_long = S.n_long;
    for (u32 r = warp_id(); r < n_long; r += DSRC_WARPS) warp_run(S, sorted, trip, M, longq[r]);
    __syncthreads();
    PROF_MARK(prof_base + 4);
}

#include "model_part.cuh"
#include "model_walk.cuh"

// Pool of adaptive-row tables shared by every model CTA of the context (all slots): bit set = table in use. A table is
// all-zero whenever it is in the pool. The pool holds one table per CTA that can be resident (4 per SM), so a CTA never
// waits unless tables are held by CTAs of another launch that are still running -- which then finish and release.
__device__ u32 tab_acquire(u32* mask, u32 count)
{
    const u32 words = (count + 31) / 32;
    u32 w = blockIdx.x % words;
    for (;;) {
        for (u32 k = 0; k < words; ++k, w = (w + 1 == words ? 0 : w + 1)) {
            const u32 valid = (w + 1) * 32 <= count ? 0xFFFFFFFFu : ((1u << (count & 31)) - 1);
            u32 free = ~*(volatile u32*)&mask[w] & valid;
            while (free) {
                const u32 b = __ffs(free) - 1;
                const u32 old = atomicOr(&mask[w], 1u << b);
                if (!(old & (1u << b))) { __threadfence(); return w * 32 + b; }
                free &= ~old & ~(1u << b);
            }
        }
        __nanosleep(200);
    }
}
__device__ void tab_release(u32* mask, u32 id) { atomicAnd(&mask[id >> 5], ~(1u << (id & 31))); }

// PART: the launch that runs the partition engine (blocks whose quality scheme has rows of 32+ symbols); the other launch skips
// those blocks. Two kernels instead of one so that each keeps its own register allocation.
template <bool QUALITY, bool PART>
__global__ void __maxnreg__(MODEL_REGS) k_model(Workspace ws, u64 arena_stride)
{
    extern __shared__ __align__(16) u8 model_smem[];       // sizeof(ModelShared) > 48 KiB: opt-in dynamic shared memory
    ModelShared& S = *(ModelShared*)model_smem;
    TabShared& TS = S.u.tab;
    const u32 tid = threadIdx.x;
    u64* bufA = ws.elem_a + (u64)blockIdx.x * arena_stride;
    u64* bufB = ws.elem_b + (u64)blockIdx.x * arena_stride;
    long long prof_t = ws.prof ? clock64() : 0;
    const int prof_base = QUALITY ? 0 : 8;

    // blocks are handed out by a per-launch counter (the CTAs of a launch finish together whatever the blocks cost); the
    // adaptive-row table of the tile/table engine is taken from the context-wide pool on first need and returned at the end
    u32* const queue = ws.model_queue + (PART ? 2 : QUALITY ? 0 : 1);
    u8* tab = nullptr;
    auto next_block = [&]() -> u32 {
        __syncthreads();
        if (tid == 0) S.blk = atomicAdd(queue, 1u);
        __syncthreads();
        return S.blk;
    };
    for (u32 blk = next_block(); blk < ws.n_blocks; blk = next_block()) {
        const BlockDesc& d = ws.desc[blk];
        BlockState& st = ws.state[blk];
        // ---- scheme selection. Only thread 0 looks at (and may set) the block's status; everybody else learns the outcome from
        // shared memory after the barrier, so the whole CTA takes the same path through next_block()'s barriers
        if (tid == 0) {
            S.ok = st.status == ST_OK && !(st.pad[1] & (QUALITY ? 1u : 2u));      // pad[1]: the walk engines' launches (model_walk.cuh) have coded the block
            S.M = 0;
            if (!S.ok) {}
            else if (QUALITY) {
                u32 sc = quality_order_scheme(st, ws.qua_order);
                st.q_scheme = (u8)sc;
                if (sc == 255) S.ok = 0;                        // SchemeNone: just the scheme byte
                else if (!quality_cfg(ws.qua_order, sc, S.cfg)) { S.ok = 0; st.status = ST_UNSUPPORTED; }
                else {
                    // rows of 32+ symbols belong to the partition engine's launch (first), unless its per-warp 16-bit counters cannot
                    // hold the block (multi-MB blocks) or it gave the block back (st.pad[0])
                    const bool part = S.cfg.alpha > 16 && st.q_total < DSRC_WARPS * 65000u;
                    if (PART) { if (part) st.pad[0] = 0; else S.ok = 0; }
                    else if (part && !st.pad[0]) S.ok = 0;
                }
                S.M = st.q_total;
            } else {
                u32 sc = st.d_count == 0 ? 255u : (st.d_count <= 4 ? 0u : 1u);
                st.d_scheme = (u8)sc;
                if (sc == 255) S.ok = 0;
                else {
                    dna_cfg(ws.dna_order, sc, S.cfg);
                    for (u32 i = S.cfg.alpha; i < 20; ++i) if (st.dfreq[i]) { S.ok = 0; st.status = ST_UNSUPPORTED; }   // reference: out-of-row write
                }
                S.M = st.d_total;
            }
            if (S.ok && S.M > arena_stride) { S.ok = 0; st.status = ST_OVERFLOW; }
        }
        if (QUALITY) S.rank[tid] = st.qrank[tid];
        __syncthreads();
        if (!S.ok || S.M == 0) continue;
        const ModelCfg cfg = S.cfg;
        const u32 M = S.M;
        const u32 passes = (cfg.key_bits + SORT_MAX_BITS - 1) / SORT_MAX_BITS;
        const u32 pbits = (cfg.key_bits + passes - 1) / passes;

        u64* const trip = (QUALITY ? ws.trip_q : ws.trip_d) + d.sym_base;
        u8* pc = (u8*)bufB;
        u32 fixed_len = 0;
        const bool tabpath = cfg.alpha <= 16 && ws.tab != nullptr;
        const bool partpath = QUALITY && PART;
        if (QUALITY && (tabpath || partpath) && st.min_len == st.max_len && st.max_len > 0 && st.max_len <= 1024) {
            // one read length in the block: position bucket j * rescale / len (TTranslationalQualityEncoder::Encode :307) from a small table
            fixed_len = st.max_len;
            for (u32 j = tid; j < fixed_len; j += DSRC_CTA) TS.plut[j] = (u8)(j * cfg.rescale / fixed_len);
            __syncthreads();
        } else if (QUALITY) {
            // variable lengths: the bucket of every symbol is computed per record and parked in the idle sort buffer
            const RecArrays& R = ws.rec;
            for (u32 r = warp_id(); r < st.n_rec; r += DSRC_WARPS) {
                const u32 len = R.qua_len[d.rec_base + r], qo = R.qcat_off[d.rec_base + r];
                for (u32 j = lane_id(); j < len; j += 32) pc[qo + j] = (u8)(j * cfg.rescale / len);
            }
            __syncthreads();
        }
        if (partpath && PART) {
            // ---- partition engine (model_part.cuh): 32 / 64 / 128-symbol quality rows, table-free. It runs BEFORE the other launch
            // and hands a block it cannot take (a partition beyond its 16-bit offsets: one very hot context) over through st.pad[0]
            if (QUALITY) {
                FetchQ f; f.q = ws.qcat + d.sym_base; f.pctx = pc; f.rank = S.rank; f.so = cfg.sym_order; f.h = cfg.sym_order / 2; f.bits = cfg.bits; f.prev = 0; f.M = M; f.plut = TS.plut; f.fixed_len = fixed_len; f.jpos = 0;
                f.ebits = 1; while ((1u << f.ebits) < st.q_count) ++f.ebits;
                f.pbits = cfg.rescale > 8 ? cfg.bits : 3;
                if (!part_engine(S, f, M, f.pbits + cfg.sym_order * f.ebits, cfg.alpha, f.ebits, bufA, bufB, trip, ws, prof_t, 32) && tid == 0) st.pad[0] = 1;
            }
            continue;
        }
        if (PART) continue;
        if (!QUALITY && cfg.alpha == 4 && cfg.key_bits <= 12) {
            // ---- whole table in shared memory (model_dna.cuh)
            dna_direct_engine(S.u.dna, ws.dcat + d.sym_base, M, cfg.ord, trip);
            PROF_MARK(prof_base + 6);
            continue;
        }
        if (tabpath) {
            // ---- tile/table engine (model_tab.cuh)
            if (!tab) {
                if (tid == 0) S.tab_id = tab_acquire(ws.tab_mask, ws.tab_count);
                __syncthreads();
                tab = ws.tab + (u64)S.tab_id * ws.tab_stride;
            }
            if (QUALITY) {
                FetchQ f; f.q = ws.qcat + d.sym_base; f.pctx = pc; f.rank = S.rank; f.so = cfg.sym_order; f.h = cfg.sym_order / 2; f.bits = cfg.bits; f.prev = 0; f.M = M; f.plut = TS.plut; f.fixed_len = fixed_len; f.jpos = 0;
                f.ebits = 1; while ((1u << f.ebits) < st.q_count) ++f.ebits;
                f.pbits = cfg.rescale > 8 ? 4 : 3;
                tab_engine<16>(TS, S.scan, f, M, f.pbits + cfg.sym_order * f.ebits, tab, (u32*)bufA, trip, ws, prof_t, 16,
                               st.q_count <= 4 ? 1u : st.q_count <= 8 ? 2u : 4u);
            } else {
                FetchD f; f.sq = ws.dcat + d.sym_base; f.ord = cfg.ord; f.bits = cfg.bits; f.prev = 0; f.M = M;
                if (cfg.alpha == 4) tab_engine<4>(TS, S.scan, f, M, cfg.key_bits, tab, (u32*)bufA, trip, ws, prof_t, 24);
                else tab_engine<8>(TS, S.scan, f, M, cfg.key_bits, tab, (u32*)bufA, trip, ws, prof_t, 24);
            }
            continue;
        }
        // ---- sort engine: first pass straight from the symbols (contexts are a pure function of the input), then passes over elements
        if (QUALITY) {
            FetchQ f; f.q = ws.qcat + d.sym_base; f.pctx = pc; f.rank = S.rank; f.so = cfg.sym_order; f.h = cfg.sym_order / 2; f.bits = cfg.bits; f.prev = 0; f.M = M; f.plut = nullptr; f.fixed_len = 0; f.jpos = 0;
            sort_pass(S, f, bufA, M, 40, pbits);
        } else {
            FetchD f; f.sq = ws.dcat + d.sym_base; f.ord = cfg.ord; f.bits = cfg.bits; f.prev = 0; f.M = M;
            sort_pass(S, f, bufA, M, 40, pbits);
        }
        PROF_MARK(prof_base + 0);
        u64* src = bufA; u64* dst = bufB;
        for (u32 p = 1; p < passes; ++p) {
            FetchSorted f; f.src = src;
            sort_pass(S, f, dst, M, 40 + p * pbits, pbits);
            u64* t = src; src = dst; dst = t;
        }
        PROF_MARK(prof_base + 1);
        // ---- adaptive statistics per context run
        group_scan(S, src, (u32*)dst, trip, M, ws, prof_t, prof_base);
    }
    if (tab && tid == 0) { __threadfence(); tab_release(ws.tab_mask, S.tab_id); }   // all touched rows are zero again (tab_engine)
}

// ------------------------------------------------------------------------------------------------
// serial range-coder chains: one thread per (block, stream).  RangeEncoder::{Start,EncodeFrequency,End}
// (src/RangeCoder.h:51-84); stream prologues: scheme byte (DnaModelerProxy.h:50-60, QualityModelerProxy.h:48-58)
// and, for quality, the 256-bit symbol mask of TTranslationalQualityEncoder::Store (QualityEncoder.h:332-342).
// ------------------------------------------------------------------------------------------------
#define RC_CTA 64
#ifndef RC_SMEM_RING
#define RC_SMEM_RING 1                        // 1: the chains' triples travel through a cp.async ring in shared memory; 0: straight from L2 into registers
#endif
#ifndef RC_L2HINT
#define RC_L2HINT 0
#endif
#if RC_L2HINT == 128
#define RC_L2 ".L2::128B"
#elif RC_L2HINT == 256
#define RC_L2 ".L2::256B"
#else
#define RC_L2 ""
#endif
#ifndef RC_RING
#define RC_RING 6
#endif
__device__ u32 g_rcp_lut[65536];                     // floor((2^32-1) / tot), filled once per device by k_rcp_lut
__global__ void k_rcp_lut() { const u32 t = blockIdx.x * blockDim.x + threadIdx.x; if (t < 65536) g_rcp_lut[t] = t ? 0xFFFFFFFFu / t : 0u; }
// A launch codes the chains of up to RC_GROUP_MAX batches (slots) at once: the launch is latency-bound -- ~107 K dependent steps
// per chain, under one resident warp per scheduler for one 8192-block batch -- so it takes the same time for two batches as for one.
__global__ void __launch_bounds__(RC_CTA, 16) k_rc_encode(RcGroup grp, u32 do_quality, u32 do_dna)
{
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 per = do_quality + do_dna;
    u32 g = 0;
    while (g + 1 < grp.n && t >= grp.ws[g].n_blocks * per) { t -= grp.ws[g].n_blocks * per; ++g; }
    const Workspace& ws = grp.ws[g];
    const u32 n = ws.n_blocks;
    u32 blk, is_dna;
    if (do_quality && do_dna) { is_dna = t >= n; blk = is_dna ? t - n : t; }
    else { is_dna = do_dna; blk = t; }
    if (blk >= n) return;
    const BlockDesc& d = ws.desc[blk];
    BlockState& st = ws.state[blk];
    if (st.status != ST_OK) return;
    const int sidx = is_dna ? 2 : 3;
    u8* out = ws.streams + d.stream_base + stream_offset(d, sidx);       // 16-byte aligned
    const u32 cap = d.stream_cap[sidx];
    const u32 scheme = is_dna ? st.d_scheme : st.q_scheme;
    u32 pos = 0;
    out[pos++] = (u8)scheme;
    if (scheme == 255) { st.stream_size[sidx] = pos; return; }
    if (!is_dna) {
        for (u32 k = 0; k < 32; ++k) {
            u32 m = 0;
            for (u32 j = 0; j < 8; ++j) m = (m << 1) | (st.qrank[k * 8 + j] != 255);
            out[pos++] = (u8)m;
        }
    }
    const u32 M = is_dna ? st.d_total : st.q_total;
    // the arena holds 1.25 bytes per symbol (api.cu; the adaptive coder stays below log2(alphabet) bits per symbol on average) -- the
    // chain checks its position every RC_CHECK sectors (outside the inner loop: a test per sector cost 11 %) and gives up with
    // ST_OVERFLOW, which makes the host repeat the call with 3 bytes per symbol (more than a symbol can ever cost: 16 bits)
    constexpr u32 RC_CHECK = 32;                       // 32 sectors = 128 steps put out at most 512 bytes
    if ((u64)pos + 640 > cap) { st.status = ST_OVERFLOW; return; }
    const u32 pos_limit = cap - 600;
    // the chain's triples arrive 4 at a time (one 32-byte sector per load) through a per-thread ring of RC_RING sectors in shared memory (cp.async), so
    // ~RC_RING*4 symbols of DRAM latency are covered; its output bytes leave 4 at a time. `range / tot` is the one long-latency
    // instruction of the chain: the reciprocal floor((2^32-1)/tot) of each symbol's total is fetched from a 64 K-entry table
    // one sector ahead (off the dependent chain), leaving mulhi + one correction on it. q' = mulhi(range, m) is q or q-1
    // (m = 2^32/tot - e with 0 < e <= 1, so range*m/2^32 > range/tot - 1), hence a single fix-up is exact.
    const ulonglong2* trip = (const ulonglong2*)((is_dna ? ws.trip_d : ws.trip_q) + d.sym_base);
    const u32 G = M / 4;                               // full groups of 4 triples; the last M % 4 symbols are coded one by one
    u64 low = 0; u32 range = 0xFFFFFFFFu;
    // output: bytes are shifted into a 32-bit word from the top (one PRMT takes the top byte of `low` and inserts it); every
    // fourth byte the word is complete and in memory order and leaves with one predicated store
    u32 obuf = 0;
    {
        const u32 on = pos & 3u; pos &= ~3u;
        for (u32 k = 0; k < on; ++k) obuf |= (u32)out[pos + k] << (8 * (4 - on + k));
        pos += on;
    }
#define RC_PUT_TOP() do { obuf = __byte_perm(obuf, (u32)(low >> 32), 0x7321); ++pos; if ((pos & 3u) == 0) *(u32*)(out + pos - 4) = obuf; } while (0)
    // One renormalisation step of RangeEncoder::EncodeFrequency (src/RangeCoder.h:62-72), written WITHOUT a branch: the chain puts
    // a byte out iff range <= 2^24 - 1, everything is predicated on that. Divergent branches are what a lone warp of chains
    // pays most for (ncu: BSSY/BRA/BSYNC held half of the stall samples of the looped form), and at warp level some lane needs
    // a byte in 98 % of the steps anyway. The carry-less adjustment `(low ^ (low + range)) & 0xFF00..` can only be non-zero if
    // low + range carries out of the low word AND bits 32..55 of low are all ones.
#define RC_RENORM() do { \
        const bool n_ = range <= 0x00FFFFFFu; \
        const u32 lo_ = (u32)low, hi_ = (u32)(low >> 32); \
        const bool cy_ = n_ && (lo_ + range < lo_) && ((hi_ & 0x00FFFFFFu) == 0x00FFFFFFu); \
        range = cy_ ? (~lo_ & 0x00FFFFFFu) : range;                      /* (r | 0xFFFFFF) - r */ \
        obuf = __byte_perm(obuf, hi_, n_ ? 0x7321u : 0x3210u); \
        pos += n_ ? 1u : 0u; \
        if (n_ && (pos & 3u) == 0) *(u32*)(out + pos - 4) = obuf; \
        const u32 sh_ = n_ ? 8u : 0u; low <<= sh_; range <<= sh_; } while (0)
#define RC_STEP(lo_w, hi_w, m) do { \
        const u32 f_ = (lo_w) & 0xFFFFu, cum_ = (lo_w) >> 16, tot_ = (hi_w); \
        u32 q_ = __umulhi(range, (m)); q_ += (range - q_ * tot_ >= tot_) ? 1u : 0u; \
        low += (u64)(q_ * cum_); range = q_ * f_; \
        RC_RENORM();                                                     /* the first byte, if any: no branch */ \
        while (range <= 0x00FFFFFFu) RC_RENORM();                        /* a second byte in one step is rare */ \
    } while (0)
#define RC_RCP(hi_) __ldg(&g_rcp_lut[(hi_)])      /* hi word of a triple the model kernels wrote = tot < 2^16 */
#if RC_SMEM_RING
    __shared__ __align__(16) ulonglong2 ring[RC_RING][2][RC_CTA];       // [slot][half][thread]: conflict-free 16-byte cells
    constexpr u32 SLOT = 2 * RC_CTA * 16, HALF = RC_CTA * 16;          // bytes per ring slot / per half slot
    const u32 rb = (u32)__cvta_generic_to_shared(&ring[0][0][threadIdx.x]);
    const u8* const rp = (const u8*)&ring[0][0][threadIdx.x];
    // sector g of this chain -> ring slot at byte offset `so` (cp.async, L2 only); one group per call, empty past the end, so the
    // group counting of wait_group is uniform
#define RC_FETCH(g, so) do { \
        if ((g) < G) { \
            asm volatile("cp.async.cg.shared.global" RC_L2 " [%0], [%1], 16;" :: "r"(rb + (so)), "l"(trip + 2 * (g)) : "memory"); \
            asm volatile("cp.async.cg.shared.global" RC_L2 " [%0], [%1], 16;" :: "r"(rb + (so) + HALF), "l"(trip + 2 * (g) + 1) : "memory"); \
        } \
        asm volatile("cp.async.commit_group;" ::: "memory"); } while (0)
#pragma unroll
    for (u32 k = 0; k < RC_RING; ++k) RC_FETCH(k, k * SLOT);
    asm volatile("cp.async.wait_group %0;" :: "n"(RC_RING - 1) : "memory");
    u32 mn0 = 0, mn1 = 0, mn2 = 0, mn3 = 0;
    uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;                        // the next sector's 4 triples, (lo, hi) word pairs, one sector ahead in registers
    if (G) { n0 = *(const uint4*)rp; n1 = *(const uint4*)(rp + HALF); mn0 = RC_RCP(n0.y); mn1 = RC_RCP(n0.w); mn2 = RC_RCP(n1.y); mn3 = RC_RCP(n1.w); }
    u32 so = 0;                                                        // byte offset of the ring slot of sector g
    for (u32 g0 = 0; g0 < G && pos <= pos_limit; g0 += RC_CHECK) {
    const u32 g1 = min(G, g0 + RC_CHECK);
    for (u32 g = g0; g < g1; ++g) {
        const u32 sn = so + SLOT == RC_RING * SLOT ? 0u : so + SLOT;
        const uint4 c0 = n0, c1 = n1;
        const u32 m0 = mn0, m1 = mn1, m2 = mn2, m3 = mn3;
        asm volatile("cp.async.wait_group %0;" :: "n"(RC_RING - 2) : "memory");      // sector g+1 has landed
        if (g + 1 < G) {   // next sector and the reciprocals of its totals: in flight while this sector is coded
            n0 = *(const uint4*)(rp + sn); n1 = *(const uint4*)(rp + sn + HALF);
            mn0 = RC_RCP(n0.y); mn1 = RC_RCP(n0.w); mn2 = RC_RCP(n1.y); mn3 = RC_RCP(n1.w);
        }
        RC_FETCH(g + RC_RING, so);                                     // refills the slot of sector g (already in registers)
        so = sn;
        RC_STEP(c0.x, c0.y, m0); RC_STEP(c0.z, c0.w, m1); RC_STEP(c1.x, c1.y, m2); RC_STEP(c1.z, c1.w, m3);
    }
    }
    if (pos > pos_limit) { asm volatile("cp.async.wait_all;" ::: "memory"); st.status = ST_OVERFLOW; return; }
#undef RC_FETCH
#else
    // Variant without shared memory (measured, not the default): the chain's triples come straight from L2 into registers TWO sectors
    // ahead, the reciprocals of a sector's totals one sector ahead, the chain's lines are pulled from DRAM into L2 RC_AHEAD sectors
    // ahead. It was built on the theory that the 42 KiB of ring per SM keep the walk engines' 100 KiB CTAs of the next batch off the
    // SMs while the chains run. Alone it is as fast as the ring (10.2 vs 10.8 ms per launch), inside the pipeline it is slower
    // (251 vs 229 ms per 18.6 GB step), whatever the walk kernels' register count or the chains' carveout: what the chains cost the
    // other batches' kernels is DRAM time -- a launch streams 27 GB of triples in 10 ms beside the model kernels' 14 GB of triple
    // writes per batch -- not SM resources.
    constexpr u32 RC_AHEAD = 24;
    auto ld_sector = [&](u32 g, uint4& a, uint4& b) {
        if (g < G) {
            const uint4* p = (const uint4*)(trip + 2 * g);
            asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
            asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p + 1));
        }
    };
    for (u32 g = 0; g < RC_AHEAD && g < G; g += 4) asm volatile("prefetch.global.L2 [%0];" :: "l"(trip + 2 * g));
    u32 mn0 = 0, mn1 = 0, mn2 = 0, mn3 = 0;
    uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0, f0 = n0, f1 = n0;      // sectors g+1 (n) and g+2 (f) in registers
    ld_sector(0, n0, n1); ld_sector(1, f0, f1);
    if (G) { mn0 = RC_RCP(n0.y); mn1 = RC_RCP(n0.w); mn2 = RC_RCP(n1.y); mn3 = RC_RCP(n1.w); }
    for (u32 g0 = 0; g0 < G && pos <= pos_limit; g0 += RC_CHECK) {
    const u32 g1 = min(G, g0 + RC_CHECK);
    for (u32 g = g0; g < g1; ++g) {
        const uint4 c0 = n0, c1 = n1;
        const u32 m0 = mn0, m1 = mn1, m2 = mn2, m3 = mn3;
        n0 = f0; n1 = f1;                                              // sector g+1 (loaded two sectors ago) and the reciprocals of its totals
        if (g + 1 < G) { mn0 = RC_RCP(n0.y); mn1 = RC_RCP(n0.w); mn2 = RC_RCP(n1.y); mn3 = RC_RCP(n1.w); }
        ld_sector(g + 2, f0, f1);
        if ((g & 3u) == 0 && g + RC_AHEAD < G) asm volatile("prefetch.global.L2 [%0];" :: "l"(trip + 2 * (g + RC_AHEAD)));
        RC_STEP(c0.x, c0.y, m0); RC_STEP(c0.z, c0.w, m1); RC_STEP(c1.x, c1.y, m2); RC_STEP(c1.z, c1.w, m3);
    }
    }
    if (pos > pos_limit) { st.status = ST_OVERFLOW; return; }
#endif
    for (u32 i = G * 4; i < M; ++i) { const uint2 tr = ((const uint2*)trip)[i]; const u32 m = RC_RCP(tr.y); RC_STEP(tr.x, tr.y, m); }   // <= 3 steps: < 24 bytes
    for (int k = 0; k < 8; ++k) { RC_PUT_TOP(); low <<= 8; }
    {
        const u32 on = pos & 3u;
        for (u32 k = 0; k < on; ++k) out[pos - on + k] = (u8)(obuf >> (8 * (4 - on + k)));
    }
    st.stream_size[sidx] = pos;
#undef RC_RCP
#undef RC_STEP
#undef RC_RENORM
#undef RC_PUT_TOP
}

static u32 model_grid(const Workspace& ws, u32 max_ctas)
{
    u32 g = max_ctas;
    if (g > ws.n_blocks) g = ws.n_blocks;
    return g ? g : 1;
}
static void model_smem_optin()
{
    cudaFuncSetAttribute(k_model<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ModelShared));
    cudaFuncSetAttribute(k_model<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ModelShared));
    cudaFuncSetAttribute(k_model<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ModelShared));
}
void launch_model_quality(const Workspace& ws, cudaStream_t s, u32 ctas, u64 stride)
{
    model_smem_optin();
    {   // blocks with <= 5 quality values and one read length: tables in shared memory, one warp per position bucket (model_walk.cuh)
        cudaFuncSetAttribute(k_model_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WalkShared));
        const u32 g = ws.n_blocks < 148u * 2u ? ws.n_blocks : 148u * 2u;
        k_model_walk<<<g ? g : 1, WALK_CTA, sizeof(WalkShared), s>>>(ws);
    }
    k_model<true, true><<<model_grid(ws, ctas), DSRC_CTA, MODEL_SMEM_QUALITY, s>>>(ws, stride);
    k_model<true, false><<<model_grid(ws, ctas), DSRC_CTA, MODEL_SMEM_QUALITY_CLASSIC, s>>>(ws, stride);
}
void launch_model_dna(const Workspace& ws, cudaStream_t s, u32 ctas, u64 stride)
{
    model_smem_optin();
    {   // 4-symbol blocks at order <= 6: one warp per block, table in shared memory (model_walk.cuh); 6 warps per SM
        cudaMemsetAsync(ws.model_queue + 3, 0, 4, s);
        static int per_sm = 0;
        if (!per_sm) { const char* e = getenv("DSRCGPU_DWALK_PER_SM"); per_sm = e ? atoi(e) : 6; if (per_sm < 1) per_sm = 1; }
        const u32 g = ws.n_blocks < 148u * (u32)per_sm ? ws.n_blocks : 148u * (u32)per_sm;
        k_dna_walk<<<g ? g : 1, 32, sizeof(DnaWalkShared), s>>>(ws);
    }
    k_model<false, false><<<model_grid(ws, ctas), DSRC_CTA, sizeof(ModelShared), s>>>(ws, stride);
}
cudaError_t rc_init_device() { k_rcp_lut<<<65536 / 256, 256>>>(); return cudaDeviceSynchronize(); }
void launch_rc_encode(const RcGroup& grp, cudaStream_t s)
{
    const u32 dq = grp.ws[0].qua_order > 0, dd = grp.ws[0].dna_order > 0;
    if (!grp.n || (!dq && !dd)) return;
    u32 threads = 0;
    for (u32 g = 0; g < grp.n; ++g) threads += grp.ws[g].n_blocks * (dq + dd);
    // developer switch DSRCGPU_RC_CARVEOUT=1: the chains ask for the largest shared-memory carveout (measured: no gain with the ring, a loss
    // without it)
    static int carve = -1;
    if (carve < 0) { const char* e = getenv("DSRCGPU_RC_CARVEOUT"); carve = e ? atoi(e) : 0; }
    if (carve) cudaFuncSetAttribute(k_rc_encode, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    k_rc_encode<<<(threads + RC_CTA - 1) / RC_CTA, RC_CTA, 0, s>>>(grp, dq, dd);
}
