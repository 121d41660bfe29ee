// Order-k range-coder modelers for DNA and quality, re-designed for the GPU.
//
// Reference: TDnaRCOrderModeler (src/DnaModelerRCO.h:45-62,94-133), TQualityOrderModeler +
// TTranslationalQualityEncoder + TQualityModelExt (src/QualityOrderModeler.h:36-51, src/QualityEncoder.h:77-136,
// 281-342), TSymbolCoderRC (src/SymbolCoderRC.h:24-93), RangeEncoder (src/RangeCoder.h:51-84).
//
// The reference codes symbol after symbol: look the context row up in a 32 KB .. 64 MB table, sum it, emit,
// bump one counter. On a GPU that is one DRAM round trip per symbol. Two observations remove the table:
//   (1) the context of symbol i is a pure function of the INPUT (the previous <= 5 symbols and the read
//       position), never of the coder state -- so all contexts of a block are computable in parallel;
//   (2) the adaptive row a symbol sees is determined by the earlier symbols of the SAME context only:
//       freq = 1 + 2*#(same ctx, same sym before), cum/tot likewise, with the halving rescale of
//       TSymbolCoderRC::Rescale applied when tot crosses 2^16 - 2N.
// So per block (one CTA): build (ctx, sym, index) keys -> stable LSD radix sort by ctx in HBM scratch ->
// every context becomes a contiguous run in original order -> threads walk runs with N 16-bit counters in
// shared memory and write the exact (freq, cum, tot) triple each symbol would have met.  What remains serial
// is only RangeEncoder::EncodeFrequency over the triples: k_rc_encode runs one thread per (block, stream),
// thousands of independent chains in flight, each a dozen integer instructions per symbol.
#include "common.cuh"
#include "kernels.h"

#define SORT_E 8                              // elements per thread per tile
#define SORT_TILE (DSRC_CTA * SORT_E)
#define CNT_BYTES 32768                       // shared budget for the per-thread symbol counters

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

#define LONG_T 48                              // context runs longer than this are walked by a whole warp
#define LONGQ_MAX 2048

struct ModelShared {
    union {
        struct {
            u32 off[DSRC_WARPS][256];
            u16 wcnt[DSRC_WARPS][256];
            u32 base[256];
            u32 hist[2][256];
        } s;
        u16 cnt[CNT_BYTES / 2];
        struct { u32 B[DSRC_WARPS][128]; u32 P[DSRC_WARPS][128]; } l;   // per-warp row state of the long-run walker
    } u;
    u32 longq[LONGQ_MAX];
    u32 n_long;
    u8 rank[256];
    ModelCfg cfg;
    u32 M, ok;
};

#define PROF_MARK(slot) do { if (ws.prof && threadIdx.x == 0) { long long t_ = clock64(); atomicAdd((unsigned long long*)&ws.prof[slot], (unsigned long long)(t_ - prof_t)); prof_t = t_; } } while (0)

__device__ __forceinline__ void hist_add(u32* hist, u32 digit, bool active)
{
    u32 am = __ballot_sync(0xFFFFFFFFu, active);
    if (active) {
        u32 peers = __match_any_sync(am, digit);
        if ((__ffs(peers) - 1) == (int)lane_id()) atomicAdd(&hist[digit], __popc(peers));
    }
}

// one stable LSD pass (8-bit digit at `shift`) from src to dst; hist_cur = counts of this digit,
// hist_next (may be null) accumulates the counts of the next digit while scattering.
__device__ void sort_pass(ModelShared& S, const u64* src, u64* dst, u32 M, u32 shift,
                          u32* hist_cur, u32* hist_next)
{
    const u32 tid = threadIdx.x, w = warp_id(), ln = lane_id();
    // exclusive scan of the 256 digit counts
    {
        u32 v = hist_cur[tid];
        u32 inc = warp_incl_sum(v);
        __shared__ u32 wsum[DSRC_WARPS];
        if (ln == 31) wsum[w] = inc;
        __syncthreads();
        u32 b = 0;
        for (u32 k = 0; k < w; ++k) b += wsum[k];
        S.u.s.base[tid] = b + inc - v;
        for (u32 k = 0; k < DSRC_WARPS; ++k) S.u.s.wcnt[k][tid] = 0;
        if (hist_next) hist_next[tid] = 0;
        __syncthreads();
    }
    for (u32 tile = 0; tile < M; tile += SORT_TILE) {
        u64 e[SORT_E]; u32 rk[SORT_E];
        const u32 wbase = tile + w * (32 * SORT_E);
#pragma unroll
        for (int k = 0; k < SORT_E; ++k) {
            const u32 i = wbase + k * 32 + ln;
            const bool in = i < M;
            e[k] = in ? src[i] : 0;
            const u32 dg = (u32)(e[k] >> shift) & 255u;
            const u32 am = __ballot_sync(0xFFFFFFFFu, in);
            rk[k] = 0;
            if (in) {
                const u32 peers = __match_any_sync(am, dg);
                const u32 prior = S.u.s.wcnt[w][dg];
                rk[k] = prior + __popc(peers & ((1u << ln) - 1));
                __syncwarp(am);
                if ((__ffs(peers) - 1) == (int)ln) S.u.s.wcnt[w][dg] = (u16)(prior + __popc(peers));
            }
            __syncwarp();
            if (hist_next) hist_add(hist_next, (u32)(e[k] >> (shift + 8)) & 255u, in);
        }
        __syncthreads();
        {   // thread = digit: turn per-warp counts into global offsets, advance the running base
            u32 run = S.u.s.base[tid];
#pragma unroll
            for (int k = 0; k < DSRC_WARPS; ++k) { u32 c = S.u.s.wcnt[k][tid]; S.u.s.wcnt[k][tid] = 0; S.u.s.off[k][tid] = run; run += c; }
            S.u.s.base[tid] = run;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SORT_E; ++k) {
            const u32 i = wbase + k * 32 + ln;
            if (i < M) dst[S.u.s.off[w][(u32)(e[k] >> shift) & 255u] + rk[k]] = e[k];
        }
        // the next tile's wcnt writes happen after its own first barrier-free phase; off[] is only rewritten after the
        // next __syncthreads, by which time every thread has finished the scatter above
    }
    __syncthreads();
}

#define TRIP(f, cum, tot) ((u64)(f) | ((u64)(cum) << 16) | ((u64)(tot) << 32))

// A whole warp walks ONE long context run starting at sorted[a], 32 symbols per step. Within a row of 32 symbols of the
// same context the adaptive row seen by lane l is the row at the start of the row plus 2 x (the symbols of the lanes
// before l), so freq / cum / tot come from ballots and popcounts; the halving rescale of TSymbolCoderRC::Rescale
// (src/SymbolCoderRC.h:69-73) can only fire once per ~32 K symbols of a context: a row that could contain it is
// replayed symbol by symbol by lane 0.
__device__ void warp_run(ModelShared& S, const u64* sorted, u64* trip, u32 M, u32 a)
{
    const u32 N = S.cfg.alpha, limit = (1u << 16) - 2 * N, ln = lane_id(), lt = (1u << ln) - 1;
    u32* B = S.u.l.B[warp_id()]; u32* P = S.u.l.P[warp_id()];
    for (u32 s = ln; s < N; s += 32) B[s] = 1;
    u32 T = N;
    const u64 key = sorted[a] >> 40;
    __syncwarp();
    for (u32 row = a;; row += 32) {
        const u32 i = row + ln;
        const u64 e = i < M ? sorted[i] : ~0ull;
        const bool valid = (e >> 40) == key;
        const u32 vm = __ballot_sync(0xFFFFFFFFu, valid);
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
            T = __shfl_sync(0xFFFFFFFFu, T, 0);
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
                const u32 gs = __shfl_sync(0xFFFFFFFFu, sym, leader);
                const u32 gm = __shfl_sync(0xFFFFFFFFu, peers, leader);
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
// (TSymbolCoderRC<N>::EncodeSymbol / Accumulate / Rescale, src/SymbolCoderRC.h:35-48, 69-90):
// short runs one thread each (N 16-bit counters per thread in shared memory), long runs one warp each.
__device__ void group_scan(ModelShared& S, const u64* sorted, u64* trip, u32 M, const Workspace& ws, long long& prof_t, int prof_base)
{
    const u32 N = S.cfg.alpha;
    const u32 nthr = min((u32)DSRC_CTA, (u32)(CNT_BYTES / 2) / N);
    const u32 limit = (1u << 16) - 2 * N;
    const u32 tid = threadIdx.x;
    if (tid == 0) S.n_long = 0;
    __syncthreads();
    if (tid < nthr) {
        u16* cnt = S.u.cnt + tid;                        // counter of symbol s at cnt[s * nthr]
        for (u32 i = tid; i < M; i += nthr) {
            const u64 e0 = sorted[i];
            const u64 key = e0 >> 40;
            if (i > 0 && (sorted[i - 1] >> 40) == key) continue;     // not a run head
            u64 nx = (i + 1 < M) ? sorted[i + 1] : ~0ull;             // ~0 is never a key (contexts have <= 21 bits)
            if ((nx >> 40) != key) {                      // fresh row: all ones
                const u32 s = (u32)(e0 >> 32) & 255u;
                trip[(u32)e0] = TRIP(1, s, N);
                continue;
            }
            if (i + LONG_T < M && (sorted[i + LONG_T] >> 40) == key) {
                const u32 slot = atomicAdd(&S.n_long, 1u);
                if (slot < LONGQ_MAX) { S.longq[slot] = i; continue; }
            }
            for (u32 s = 0; s < N; ++s) cnt[s * nthr] = 1;
            u32 tot = N, k = i;
            u64 e = e0;
            for (;;) {
                const u32 s = (u32)(e >> 32) & 255u;
                if (tot >= limit) {                       // Rescale: stats[i] -= stats[i] >> 1
                    tot = 0;
                    for (u32 q = 0; q < N; ++q) { u32 c = cnt[q * nthr]; c -= c >> 1; cnt[q * nthr] = (u16)c; tot += c; }
                }
                const u32 f = cnt[s * nthr];
                u32 cum = 0;
                if (s * 2 <= N) { for (u32 q = 0; q < s; ++q) cum += cnt[q * nthr]; }
                else { u32 hi = 0; for (u32 q = s; q < N; ++q) hi += cnt[q * nthr]; cum = tot - hi; }
                trip[(u32)e] = TRIP(f, cum, tot);
                cnt[s * nthr] = (u16)(f + 2); tot += 2;
                e = nx; ++k;
                if ((e >> 40) != key) break;
                nx = (k + 1 < M) ? sorted[k + 1] : ~0ull;
            }
        }
    }
    __syncthreads();
    PROF_MARK(prof_base + 2);
    const u32 n_long = min(S.n_long, (u32)LONGQ_MAX);
    for (u32 r = warp_id(); r < n_long; r += DSRC_WARPS) warp_run(S, sorted, trip, M, S.longq[r]);
    __syncthreads();
    PROF_MARK(prof_base + 3);
}

template <bool QUALITY>
__global__ void __launch_bounds__(DSRC_CTA) k_model(Workspace ws, u64 arena_stride)
{
    __shared__ ModelShared S;
    const u32 tid = threadIdx.x;
    u64* bufA = ws.elem_a + (u64)blockIdx.x * arena_stride;
    u64* bufB = ws.elem_b + (u64)blockIdx.x * arena_stride;
    long long prof_t = ws.prof ? clock64() : 0;
    const int prof_base = QUALITY ? 0 : 8;

    for (u32 blk = blockIdx.x; blk < ws.n_blocks; blk += gridDim.x) {
        const BlockDesc& d = ws.desc[blk];
        BlockState& st = ws.state[blk];
        __syncthreads();
        if (st.status != ST_OK) continue;
        // ---- scheme selection
        if (tid == 0) {
            S.ok = 1;
            if (QUALITY) {
                u32 sc = quality_order_scheme(st, ws.qua_order);
                st.q_scheme = (u8)sc;
                if (sc == 255) S.ok = 0;                        // SchemeNone: just the scheme byte
                else if (!quality_cfg(ws.qua_order, sc, S.cfg)) { S.ok = 0; st.status = ST_UNSUPPORTED; }
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
            if (S.M > arena_stride) { S.ok = 0; st.status = ST_OVERFLOW; }
        }
        if (QUALITY) S.rank[tid] = st.qrank[tid];
        S.u.s.hist[0][tid] = 0;
        __syncthreads();
        if (!S.ok || S.M == 0) continue;
        const ModelCfg cfg = S.cfg;
        const u32 M = S.M;
        const u32 passes = (cfg.key_bits + 7) / 8;

        // ---- keys: (ctx << 40) | (sym << 32) | index
        if (QUALITY) {
            const u8* q = ws.qcat + d.sym_base;
            const RecArrays& R = ws.rec;
            const u32 h = cfg.sym_order / 2, so = cfg.sym_order, bits = cfg.bits;
            for (u32 r = warp_id(); r < st.n_rec; r += DSRC_WARPS) {
                const u32 len = R.qua_len[d.rec_base + r], qo = R.qcat_off[d.rec_base + r];
                for (u32 j0 = 0; j0 < len; j0 += 32) {
                    const u32 j = j0 + lane_id(); const bool in = j < len;
                    u64 el = 0; u32 ctx = 0;
                    if (in) {
                        const u32 i = qo + j;
                        u32 y[6];                                  // y[k] = symbol k steps back (0 before the block start)
#pragma unroll
                        for (int k = 0; k < 6; ++k) y[k] = (i >= (u32)k) ? S.rank[q[i - k]] : 0u;
                        // hash slots (QualityEncoder.h:77-94): raw below slot h, pairwise means from slot h on
                        u32 hash = 0;
                        if (so == 1) hash = y[1];
                        else for (u32 t = 0; t < so; ++t) {
                            u32 v = t < h ? y[t + 1] : ((y[t + 1] + y[t + 2]) >> 1);
                            hash |= v << (t * bits);
                        }
                        const u32 pctx = j * cfg.rescale / len;     // TTranslationalQualityEncoder::Encode :307
                        ctx = (hash << bits) | pctx;
                        el = ((u64)ctx << 40) | ((u64)y[0] << 32) | i;
                        bufA[i] = el;
                    }
                    hist_add(S.u.s.hist[0], ctx & 255u, in);
                }
            }
        } else {
            const u8* sq = ws.dcat + d.sym_base;
            const u32 bits = cfg.bits, ord = cfg.ord;
            for (u32 i0 = 0; i0 < M; i0 += DSRC_CTA) {
                const u32 i = i0 + tid; const bool in = i < M;
                u32 ctx = 0;
                if (in) {
                    for (u32 t = 0; t < ord; ++t) { u32 v = (i >= t + 1) ? sq[i - t - 1] : 0u; ctx |= (v & ((1u << bits) - 1)) << (t * bits); }
                    bufA[i] = ((u64)ctx << 40) | ((u64)sq[i] << 32) | i;
                }
                hist_add(S.u.s.hist[0], ctx & 255u, in);
            }
        }
        __syncthreads();
        PROF_MARK(prof_base + 0);
        // ---- stable LSD radix sort by context
        u64* src = bufA; u64* dst = bufB;
        for (u32 p = 0; p < passes; ++p) {
            sort_pass(S, src, dst, M, 40 + 8 * p, S.u.s.hist[p & 1], (p + 1 < passes) ? S.u.s.hist[(p + 1) & 1] : nullptr);
            u64* t = src; src = dst; dst = t;
        }
        PROF_MARK(prof_base + 1);
        // ---- adaptive statistics per context run
        group_scan(S, src, (QUALITY ? ws.trip_q : ws.trip_d) + d.sym_base, M, ws, prof_t, prof_base);
    }
}

// ------------------------------------------------------------------------------------------------
// serial range-coder chains: one thread per (block, stream).  RangeEncoder::{Start,EncodeFrequency,End}
// (src/RangeCoder.h:51-84); stream prologues: scheme byte (DnaModelerProxy.h:50-60, QualityModelerProxy.h:48-58)
// and, for quality, the 256-bit symbol mask of TTranslationalQualityEncoder::Store (QualityEncoder.h:332-342).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_rc_encode(Workspace ws, u32 do_quality, u32 do_dna)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 n = ws.n_blocks;
    u32 blk, is_dna;
    if (do_quality && do_dna) { is_dna = t >= n; blk = is_dna ? t - n : t; }
    else { is_dna = do_dna; blk = t; }
    if (blk >= n) return;
    const BlockDesc& d = ws.desc[blk];
    BlockState& st = ws.state[blk];
    if (st.status != ST_OK) return;
    const int sidx = is_dna ? 2 : 3;
    u8* out = ws.streams + d.stream_base + stream_offset(d, sidx);
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
    if ((u64)pos + 3ull * M + 16 > cap) { st.status = ST_OVERFLOW; return; }
    const u64* trip = (is_dna ? ws.trip_d : ws.trip_q) + d.sym_base;
    u64 low = 0; u32 range = 0xFFFFFFFFu;
    u64 nxt = M ? trip[0] : 0;
    for (u32 i = 0; i < M; ++i) {
        const u64 tr = nxt;
        if (i + 1 < M) nxt = trip[i + 1];
        const u32 f = (u32)tr & 0xFFFFu, cum = (u32)(tr >> 16) & 0xFFFFu, tot = (u32)(tr >> 32);
        range /= tot;
        low += (u64)(range * cum);
        range *= f;
        while (range <= 0x00FFFFFFu) {
            if ((low ^ (low + range)) & 0xFF00000000000000ull) { u32 r = (u32)low; range = (r | 0x00FFFFFFu) - r; }
            out[pos++] = (u8)(low >> 56);
            low <<= 8; range <<= 8;
        }
    }
    for (int k = 0; k < 8; ++k) { out[pos++] = (u8)(low >> 56); low <<= 8; }
    st.stream_size[sidx] = pos;
}

static u32 model_grid(const Workspace& ws, u32 max_ctas)
{
    u32 g = max_ctas;
    if (g > ws.n_blocks) g = ws.n_blocks;
    return g ? g : 1;
}
void launch_model_quality(const Workspace& ws, cudaStream_t s, u32 ctas, u64 stride) { k_model<true><<<model_grid(ws, ctas), DSRC_CTA, 0, s>>>(ws, stride); }
void launch_model_dna(const Workspace& ws, cudaStream_t s, u32 ctas, u64 stride) { k_model<false><<<model_grid(ws, ctas), DSRC_CTA, 0, s>>>(ws, stride); }
void launch_rc_encode(const Workspace& ws, cudaStream_t s)
{
    const u32 dq = ws.qua_order > 0, dd = ws.dna_order > 0;
    if (!dq && !dd) return;
    const u32 threads = ws.n_blocks * (dq + dd);
    k_rc_encode<<<(threads + 127) / 128, 128, 0, s>>>(ws, dq, dd);
}
