// Tile/table engine of the order-k modelers for alphabets of <= 16 symbols (DNA 4/8-symbol rows, 16-symbol quality rows).
//
// The block's symbols are taken in their original order, 2048 at a time. Inside a tile the symbols are grouped by context
// with a stable radix sort that never leaves shared memory; every context group then meets its adaptive row exactly as
// TSymbolCoderRC<N>::EncodeSymbol would (src/SymbolCoderRC.h:35-48,69-90): one thread (short groups) or one warp (long
// groups) loads the row from the per-CTA table in HBM/L2 -- the same table the reference keeps per coder
// (DnaModelerRCO.h:94-119, QualityEncoder.h:45-58), rows of 2N bytes -- walks the group's symbols in order, emits their
// (freq, cum, tot) triples into a shared-memory staging tile and writes the row back. The staged triples leave in one
// coalesced store. A row whose first counter is 0 has not been touched by this block (counters never drop below 1), so
// the table needs no per-block initialisation: touched contexts are remembered and re-zeroed when the block is done.
#pragma once

#define TT 2048                               // symbols per tile
#define TT_SHIFT 11
#ifndef TAB_C1
#define TAB_C1 6
#endif
#ifndef TAB_MATCH_ANY
#define TAB_MATCH_ANY 8                       // digits this wide rank their row with match.any (few distinct values per row), narrower ones with ballots
#endif
#ifndef TAB_LONG
#define TAB_LONG 24                           // groups longer than this take a whole warp
#endif

// partition engine (model_part.cuh): state that has to survive the engines' shared-memory union
#define PART_MAX_BITS 10
struct PartState {
    u32 pend[1 << PART_MAX_BITS];             // counting pass: start of every partition in the arena; after the scatter pass: its end
    u32 d0, d1, a, n;                         // partitions [d0, d1) of the current tile, their span [a, a + n) in the arena
};

struct TabShared {
    alignas(16) u32 el[2][TT];                // (ctx << 11) | position in tile, ping-pong of the in-tile sort (vector loads)
    u8 sym[TT];
    union { u64 trip[TT]; u16 H[DSRC_WARPS << 10]; } x;   // triple staging / sort counters (never live together)
    // group heads by size class so that the lanes of a warp walk groups of similar length:
    // class 0: 1..2 symbols, class 1: 3..TAB_C1, class 2: TAB_C1+1..TAB_LONG; longer groups take a whole warp
    u16 heads[TT];
    u16 heads1[TT / 3 + 2];
    u16 heads2[TT / (TAB_C1 + 1) + 2];
    u16 longs[TT / (TAB_LONG + 1) + 2];
    u32 B[DSRC_WARPS][16], P[DSRC_WARPS][16];
    u32 n_heads[3], n_long, n_touched;
    u32 wv[DSRC_WARPS][4], wcnt[DSRC_WARPS], wflag[DSRC_WARPS];   // scan engine: per-warp aggregates of the segmented scan
    u8 plut[1024];                            // position bucket of every read position when the block's reads have one length
    u16 pB[DSRC_WARPS][128], pP[DSRC_WARPS][128];   // partition engine: symbol counts of a warp's open run and their exclusive prefix sums
    PartState part;                           // last: beyond the part of the union the sort engine's scratch overlays (static_assert in rc_model.cu)
};

// lanes holding the same `bits`-wide digit (replaces match.any, whose cost grows with the number of distinct values)
__device__ __forceinline__ u32 match_bits(u32 d, u32 bits, u32 active)
{
    u32 peers = active;
    for (u32 k = 0; k < bits; ++k) { const u32 m = __ballot_sync(FULL, (d >> k) & 1u); peers &= ((d >> k) & 1u) ? m : ~m; }
    return peers;
}

template <int BITS>
__device__ __forceinline__ void tile_sort_pass_t(TabShared& S, u32* scan, const u32* src, u32* dst, u32 n, u32 shift)
{
    const u32 tid = threadIdx.x, w = warp_id(), ln = lane_id(), lt = (1u << ln) - 1;
    const u32 bins = 1u << BITS, dmask = bins - 1;
    const u32 wb = min(n, w * (TT / DSRC_WARPS)), we = min(n, wb + TT / DSRC_WARPS);
    u16* H = S.x.H + w * bins;
    u32* H32 = (u32*)H;
    for (u32 i = tid; i < DSRC_WARPS * bins / 2; i += DSRC_CTA) ((u32*)S.x.H)[i] = 0;
    __syncthreads();
    // the warp's 8 rows of 32 elements: element, and the lanes of its row holding the same digit -- found once, used by the
    // histogram (one shared reduction per distinct digit of a row: packed 16-bit counters, counts <= 256, no carry) and, after
    // the scan, by the scatter
    constexpr int ROWS = TT / DSRC_WARPS / 32;
    u32 e[ROWS], peers[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) { const u32 i = wb + 32 * k + ln; e[k] = i < we ? src[i] : 0u; }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
        const bool in = wb + 32 * k + ln < we;
        const u32 d = (e[k] >> shift) & dmask;
        u32 p = __ballot_sync(FULL, in);
#if TAB_MATCH_ANY
        if (BITS >= TAB_MATCH_ANY) { if (in) p = __match_any_sync(p, d); }
        else
#endif
        {
#pragma unroll
            for (int q = 0; q < BITS; ++q) { const u32 m = __ballot_sync(FULL, (d >> q) & 1u); p &= ((d >> q) & 1u) ? m : ~m; }
        }
        peers[k] = in ? p : 0u;
        if (in && (__ffs(p) - 1) == (int)ln) atomicAdd(&H32[d >> 1], (u32)__popc(p) << ((d & 1) * 16));
    }
    __syncthreads();
    {
        const u32 per = bins > DSRC_CTA ? bins / DSRC_CTA : 1u, d0 = tid * per;
        u32 sum = 0;
        if (d0 < bins) for (u32 k = 0; k < per; ++k) for (u32 ww = 0; ww < DSRC_WARPS; ++ww) sum += S.x.H[ww * bins + d0 + k];
        u32 total, run = block_excl_sum(sum, scan, &total);
        if (d0 < bins) for (u32 k = 0; k < per; ++k) for (u32 ww = 0; ww < DSRC_WARPS; ++ww) { const u32 c = S.x.H[ww * bins + d0 + k]; S.x.H[ww * bins + d0 + k] = (u16)run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
        const u32 p = peers[k], d = (e[k] >> shift) & dmask;
        const u32 pos = p ? H[d] + __popc(p & lt) : 0u;
        __syncwarp();
        if (p) {
            if ((__ffs(p) - 1) == (int)ln) H[d] += (u16)__popc(p);
            dst[pos] = e[k];
        }
        __syncwarp();
    }
    __syncthreads();
}
__device__ __forceinline__ void tile_sort_pass(TabShared& S, u32* scan, const u32* src, u32* dst, u32 n, u32 shift, u32 bits)
{
    switch (bits) {
    case 1: case 2:
    case 3: tile_sort_pass_t<3>(S, scan, src, dst, n, shift); break;
    case 4: tile_sort_pass_t<4>(S, scan, src, dst, n, shift); break;
    case 5: tile_sort_pass_t<5>(S, scan, src, dst, n, shift); break;
    case 6: tile_sort_pass_t<6>(S, scan, src, dst, n, shift); break;
    case 7: tile_sort_pass_t<7>(S, scan, src, dst, n, shift); break;
    case 8: tile_sort_pass_t<8>(S, scan, src, dst, n, shift); break;
    case 9: tile_sort_pass_t<9>(S, scan, src, dst, n, shift); break;
    default: tile_sort_pass_t<10>(S, scan, src, dst, n, shift); break;       // a wider digit than needed only costs idle bins
    }
}

// counter k of a row held as N/2 packed u16 pairs
template <int N> struct RowRegs {
    u32 c[N / 2];
    __device__ __forceinline__ void load(const u8* p)
    {
        if (N == 4) { const uint2 v = *(const uint2*)p; c[0] = v.x; c[1] = v.y; }
        else {
#pragma unroll
            for (int k = 0; k < N / 8; ++k) { const uint4 v = ((const uint4*)p)[k]; c[4 * k] = v.x; c[4 * k + 1] = v.y; c[4 * k + 2] = v.z; c[4 * k + 3] = v.w; }
        }
    }
    __device__ __forceinline__ void store(u8* p) const
    {
        if (N == 4) *(uint2*)p = make_uint2(c[0], c[1]);
        else {
#pragma unroll
            for (int k = 0; k < N / 8; ++k) ((uint4*)p)[k] = make_uint4(c[4 * k], c[4 * k + 1], c[4 * k + 2], c[4 * k + 3]);
        }
    }
    __device__ __forceinline__ void ones() {
#pragma unroll
        for (int k = 0; k < N / 2; ++k) c[k] = 0x00010001u;
    }
    __device__ __forceinline__ u32 total() const { u32 t = 0;
#pragma unroll
        for (int k = 0; k < N / 2; ++k) t += (c[k] & 0xFFFFu) + (c[k] >> 16);
        return t; }
    __device__ __forceinline__ u32 rescale() { u32 t = 0;      // stats[i] -= stats[i] >> 1
#pragma unroll
        for (int k = 0; k < N / 2; ++k) { u32 lo = c[k] & 0xFFFFu, hi = c[k] >> 16; lo -= lo >> 1; hi -= hi >> 1; c[k] = lo | (hi << 16); t += lo + hi; }
        return t; }
    __device__ __forceinline__ void get(u32 s, u32& f, u32& cum) const { f = 0; cum = 0;
#pragma unroll
        for (int k = 0; k < N; ++k) { const u32 v = (c[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu; cum += (u32)k < s ? v : 0u; f = (u32)k == s ? v : f; } }
    __device__ __forceinline__ void bump(u32 s) {
#pragma unroll
        for (int k = 0; k < N / 2; ++k) c[k] += (u32)k == (s >> 1) ? (2u << ((s & 1) * 16)) : 0u; }
};

template <int N>
__device__ void tab_short_groups(TabShared& S, const u32* sorted, u32 n, u8* tab, u32* touched)
{
    const u32 limit = (1u << 16) - 2 * N;
    for (int cls = 0; cls < 3; ++cls) {
    const u16* list = cls == 0 ? S.heads : cls == 1 ? S.heads1 : S.heads2;
    const u32 nh = S.n_heads[cls];
    for (u32 h = threadIdx.x; h < nh; h += DSRC_CTA) {
        u32 p = list[h];
        u32 e = sorted[p];
        const u32 key = e >> TT_SHIFT;
        u8* rowp = tab + (u64)key * (2 * N);
        RowRegs<N> R; R.load(rowp);
        if ((R.c[0] & 0xFFFFu) == 0) { R.ones(); touched[atomicAdd(&S.n_touched, 1u)] = key; }
        u32 tot = R.total();
        do {
            const u32 pos = e & (TT - 1), s = S.sym[pos];
            if (tot >= limit) tot = R.rescale();
            u32 f, cum; R.get(s, f, cum);
            S.x.trip[pos] = TRIP(f, cum, tot);
            R.bump(s); tot += 2;
            if (++p >= n) break;
            e = sorted[p];
        } while ((e >> TT_SHIFT) == key);
        R.store(rowp);
    }
    }
}

// one warp per long group; 32 members per step (see warp_run in rc_model.cu for the row algebra)
template <int N>
__device__ void tab_long_groups(TabShared& S, const u32* sorted, u32 n, u8* tab, u32* touched)
{
    const u32 limit = (1u << 16) - 2 * N, ln = lane_id(), lt = (1u << ln) - 1;
    u32* B = S.B[warp_id()]; u32* P = S.P[warp_id()];
    const u32 nl = S.n_long;
    for (u32 g = warp_id(); g < nl; g += DSRC_WARPS) {
        const u32 a = S.longs[g];
        const u32 key = sorted[a] >> TT_SHIFT;
        u16* rowp = (u16*)(tab + (u64)key * (2 * N));
        u32 v = ln < N ? rowp[ln] : 0u;
        const bool fresh = __shfl_sync(FULL, v, 0) == 0;
        if (fresh) { v = ln < N ? 1u : 0u; if (ln == 0) touched[atomicAdd(&S.n_touched, 1u)] = key; }
        if (ln < N) B[ln] = v;
        u32 T = warp_red_sum(v);
        __syncwarp();
        for (u32 row = a;; row += 32) {
            const u32 i = row + ln;
            const u32 e = i < n ? sorted[i] : 0xFFFFFFFFu;
            const bool valid = i < n && (e >> TT_SHIFT) == key;
            const u32 vm = __ballot_sync(FULL, valid);
            if (!vm) break;
            const u32 nv = __popc(vm);
            const u32 pos = e & (TT - 1);
            const u32 sym = valid ? S.sym[pos] : 0u;
            if (T + 2 * (nv - 1) >= limit) {
                for (u32 j = 0; j < nv; ++j) {                 // replay the row symbol by symbol (lane 0 computes)
                    const u32 sj = __shfl_sync(FULL, sym, j), pj = __shfl_sync(FULL, pos, j);
                    if (ln == 0) {
                        if (T >= limit) { T = 0; for (u32 q = 0; q < N; ++q) { u32 c = B[q]; c -= c >> 1; B[q] = c; T += c; } }
                        const u32 f = B[sj]; u32 cum = 0;
                        for (u32 q = 0; q < sj; ++q) cum += B[q];
                        S.x.trip[pj] = TRIP(f, cum, T);
                        B[sj] = f + 2; T += 2;
                    }
                }
                T = __shfl_sync(FULL, T, 0);
                __syncwarp();
            } else {
                const u32 bv = ln < N ? B[ln] : 0u;
                const u32 inc = warp_incl_sum(bv);
                if (ln < N) P[ln] = inc - bv;
                __syncwarp();
                u32 peers = 0;
                if (valid) peers = __match_any_sync(vm, sym);
                u32 c_lt = 0, rem = vm;
                while (rem) {
                    const int leader = __ffs(rem) - 1;
                    const u32 gs = __shfl_sync(FULL, sym, leader);
                    const u32 gm = __shfl_sync(FULL, peers, leader);
                    if (valid && gs < sym) c_lt += __popc(gm & lt);
                    rem &= ~gm;
                }
                if (valid) S.x.trip[pos] = TRIP(B[sym] + 2 * __popc(peers & lt), P[sym] + 2 * c_lt, T + 2 * __popc(vm & lt));
                __syncwarp();
                if (valid && (__ffs(peers) - 1) == (int)ln) B[sym] += 2 * __popc(peers);
                T += 2 * nv;
                __syncwarp();
            }
            if (nv < 32) break;
        }
        if (ln < N) rowp[ln] = (u16)B[ln];
        __syncwarp();
    }
}

// ---- scan engine (16-symbol rows): the group walk without a walk.
// In the sorted tile every context group is a contiguous run in original order. The triple a symbol meets is a function of
// the group's row at the start of the tile and of the COUNTS of the symbols before it in its run:
//   freq = row[s] + 2 #(earlier, same symbol), cum = sum_{q<s} row[q] + 2 #(earlier, smaller symbol), tot = sum row + 2 #(earlier)
// (TSymbolCoderRC::EncodeSymbol, src/SymbolCoderRC.h:35-48, as long as no rescale fires). The counts are an exclusive SEGMENTED
// prefix sum of one-hot vectors over the sorted tile: 16 byte counters in four words, every thread scans 8 consecutive elements
// in registers, thread aggregates are combined with a segmented warp scan and a pass over the 8 warp aggregates. No run heads,
// no size classes, no divergent walks: every lane does the same work whatever the run lengths are. Byte counters are exact
// for the first 255 members of a run (a count never exceeds the member's position); a run that is longer, or that reaches
// the rescale threshold of its row, is queued once for the warp-cooperative walker (tab_long_groups), which redoes it from its
// head -- a few runs per block. The last member of every other run adds the run's counts to the row.
// NW = words of byte counters = live symbols / 4: the block's symbols are dense ranks < q_count, so a block with <= 8 (<= 4)
// distinct quality values scans 2 (1) words instead of 4 and touches only the first 8 (4) counters of its rows; the dead
// counters of a row stay at their initial 1 for ever (rescaling keeps 1 at 1) and enter `tot` as a constant.
template <int NW> __device__ __forceinline__ void cnt_add(u32 (&v)[NW], u32 s)
{
    const u32 inc = 1u << ((s & 3u) * 8);
    if (NW == 1) v[0] += inc;
    else {
        const u32 k = s >> 2;
#pragma unroll
        for (int q = 0; q < NW; ++q) v[q] += k == (u32)q ? inc : 0u;
    }
}
template <int NW> __device__ __forceinline__ void cnt_zero(u32 (&v)[NW]) {
#pragma unroll
    for (int q = 0; q < NW; ++q) v[q] = 0u;
}
__device__ __forceinline__ u32 bytesum(u32 x) { return __vsadu4(x, 0u); }

template <int NW>
__device__ void tab_scan_groups16(TabShared& S, const u32* sorted, u32* scratch, u32 n, u8* tab, u32* touched)
{
    constexpr u32 NS = 4 * NW, DEAD = 16 - NS;        // live symbols; counters that never move
    const u32 tid = threadIdx.x, ln = lane_id(), w = warp_id();
    const u32 limit = (1u << 16) - 32;
    const u32 p0 = tid * 8;
    const u32 NOKEY = 0x3FFFFFu;                     // not a context (keys have <= 21 bits after the shift)
    u32 sy = 0, heads = 0, tails = 0;                // per element: symbol (4 bits), "first of its run", "last of its run"
    {
        const uint4 a = ((const uint4*)sorted)[tid * 2], b = ((const uint4*)sorted)[tid * 2 + 1];
        const u32 e[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        u32 pk = (p0 && p0 - 1 < n) ? sorted[p0 - 1] >> TT_SHIFT : NOKEY + 1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool in = p0 + j < n;
            const u32 k = in ? e[j] >> TT_SHIFT : NOKEY;
            sy |= (in ? (u32)S.sym[e[j] & (TT - 1)] : 0u) << (4 * j);
            heads |= (k != pk ? 1u : 0u) << j;
            // the rows P2 will need: on their way into L1 while the counts are scanned
            if (in && (j == 0 || k != pk)) asm volatile("prefetch.global.L1 [%0];" :: "l"(tab + (u64)k * 32));
            pk = k;
        }
        const u32 nextkey = p0 + 8 < n ? sorted[p0 + 8] >> TT_SHIFT : NOKEY + 2;
        tails = (heads >> 1) | ((pk != nextkey ? 1u : 0u) << 7);
    }
    const u32 nval = p0 < n ? min(8u, n - p0) : 0u;

    // P1: aggregate of this thread's 8 elements (counts since the last run head inside the thread, or of all 8)
    u32 v[NW], cnt = 0, flag = heads ? 1u : 0u;
    cnt_zero<NW>(v);
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
        if ((heads >> j) & 1u) { cnt_zero<NW>(v); cnt = 0; }
        cnt_add<NW>(v, (sy >> (4 * j)) & 15u); ++cnt;
    }
    // segmented inclusive scan over the warp's threads
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 av[NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) av[q] = __shfl_up_sync(FULL, v[q], o);
        const u32 ac = __shfl_up_sync(FULL, cnt, o), af = __shfl_up_sync(FULL, flag, o);
        if (ln >= (u32)o) {
            if (!flag) {
#pragma unroll
                for (int q = 0; q < NW; ++q) v[q] += av[q];
                cnt += ac;
            }
            flag |= af;
        }
    }
    if (ln == 31) {
#pragma unroll
        for (int q = 0; q < NW; ++q) S.wv[w][q] = v[q];
        S.wcnt[w] = cnt; S.wflag[w] = flag;
    }
    // exclusive: what the threads before me accumulated since the last run head
    u32 c[NW], ccnt;
    {
#pragma unroll
        for (int q = 0; q < NW; ++q) c[q] = __shfl_up_sync(FULL, v[q], 1);
        ccnt = __shfl_up_sync(FULL, cnt, 1);
        u32 cflag = __shfl_up_sync(FULL, flag, 1);
        if (ln == 0) { cnt_zero<NW>(c); ccnt = 0; cflag = 0; }
        __syncthreads();
        if (!cflag) {                                // no run head in this warp before me: the run continues from earlier warps
            for (int ww = (int)w - 1; ww >= 0; --ww) {
#pragma unroll
                for (int q = 0; q < NW; ++q) c[q] += S.wv[ww][q];
                ccnt += S.wcnt[ww];
                if (S.wflag[ww]) break;
            }
        }
    }

    // P2: every element meets its row
    u32 badmask = 0, freshmask = 0;
    {
        u32 T0 = 0; bool fresh = false;
        u32* const myscr = scratch + tid * 8;        // exclusive prefix sums of the live counters of the current row, u16 each
        const u16* P = (const u16*)myscr;
#pragma unroll
        for (int q = 0; q < NW; ++q) v[q] = c[q];
        cnt = ccnt;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
            const u32 s = (sy >> (4 * j)) & 15u;
            const bool head = (heads >> j) & 1u;
            if (head) { cnt_zero<NW>(v); cnt = 0; }
            if ((u32)j < nval) {
                const u32 e = sorted[p0 + j];
                if (j == 0 || head) {
                    const u8* rowp = tab + (u64)(e >> TT_SHIFT) * 32;
                    u32 rw[2 * NW];                  // the live counters, two per word
                    if (NW == 1) { const uint2 r = *(const uint2*)rowp; rw[0] = r.x; rw[1] = r.y; }
                    else {
#pragma unroll
                        for (int q = 0; q < NW / 2; ++q) { const uint4 r = ((const uint4*)rowp)[q]; rw[4 * q] = r.x; rw[4 * q + 1] = r.y; rw[4 * q + 2] = r.z; rw[4 * q + 3] = r.w; }
                    }
                    fresh = (rw[0] & 0xFFFFu) == 0;
                    u32 o = 0, ex[2 * NW];
#pragma unroll
                    for (int q = 0; q < 2 * NW; ++q) {
                        const u32 r = fresh ? 0x00010001u : rw[q];
                        ex[q] = o * 0x10001u + (r << 16); o += (r & 0xFFFFu) + (r >> 16);
                    }
                    if (NW == 1) *(uint2*)myscr = make_uint2(ex[0], ex[1]);
                    else {
#pragma unroll
                        for (int q = 0; q < NW / 2; ++q) ((uint4*)myscr)[q] = make_uint4(ex[4 * q], ex[4 * q + 1], ex[4 * q + 2], ex[4 * q + 3]);
                    }
                    T0 = o + DEAD;
                }
                const u32 pos = cnt;
                const bool bad = pos >= 255u || T0 + 2 * pos >= limit;
                freshmask |= (fresh ? 1u : 0u) << j;
                if (!bad) {
                    const u32 c0 = P[s], c1 = s < NS - 1 ? (u32)P[s + 1] : T0 - DEAD;
                    const u32 kk = s >> 2, sh = (s & 3u) * 8;
                    u32 vk = v[0], nc = 0;
#pragma unroll
                    for (int q = 1; q < NW; ++q) { vk = kk == (u32)q ? v[q] : vk; nc += kk >= (u32)q ? bytesum(v[q - 1]) : 0u; }
                    const u32 nf = (vk >> sh) & 255u;
                    nc += bytesum(vk & ((1u << sh) - 1u));
                    S.x.trip[e & (TT - 1)] = TRIP(c1 - c0 + 2 * nf, c0 + 2 * nc, T0 + 2 * pos);
                } else {
                    badmask |= 1u << j;
                    if (pos == 0 || !(pos - 1 >= 255u || T0 + 2 * (pos - 1) >= limit)) S.longs[atomicAdd(&S.n_long, 1u)] = (u16)(p0 + j - pos);
                }
            }
            cnt_add<NW>(v, s); ++cnt;
        }
    }
    __syncthreads();
    // P3: the last member of every run the scan covered adds the run's counts to the row -- without reading it again: a row
    // that was fresh is stored whole (ones + counts), any other receives its non-zero words as fire-and-forget reductions
#pragma unroll
    for (int q = 0; q < NW; ++q) v[q] = c[q];
    tails &= ~badmask & ((1u << nval) - 1u);
    if (tails) {
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
            if ((heads >> j) & 1u) cnt_zero<NW>(v);
            cnt_add<NW>(v, (sy >> (4 * j)) & 15u);
            if ((tails >> j) & 1u) {
                const u32 k = sorted[p0 + j] >> TT_SHIFT;
                u32* rowp = (u32*)(tab + (u64)k * 32);
                // byte counters -> 16-bit lanes, doubled; the dead counters keep their 1
                u32 d[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    d[2 * q] = q < NW ? 2 * __byte_perm(v[q < NW ? q : 0], 0u, 0x4140) : 0u;
                    d[2 * q + 1] = q < NW ? 2 * __byte_perm(v[q < NW ? q : 0], 0u, 0x4342) : 0u;
                }
                if ((freshmask >> j) & 1u) {
                    touched[atomicAdd(&S.n_touched, 1u)] = k;
                    ((uint4*)rowp)[0] = make_uint4(d[0] + 0x00010001u, d[1] + 0x00010001u, d[2] + 0x00010001u, d[3] + 0x00010001u);
                    ((uint4*)rowp)[1] = make_uint4(d[4] + 0x00010001u, d[5] + 0x00010001u, d[6] + 0x00010001u, d[7] + 0x00010001u);
                } else {
#pragma unroll
                    for (int q = 0; q < 2 * NW; ++q) if (d[q]) atomicAdd(&rowp[q], d[q]);
                }
            }
        }
    }
}

// F: FetchQ / FetchD (rc_model.cu). scratch: per-CTA global scratch for the touched-context list (M entries).
template <int N, class F>
__device__ void tab_engine(TabShared& S, u32* scan, F f, u32 M, u32 key_bits, u8* tab, u32* touched, u64* trip,
                           const Workspace& ws, long long& prof_t, int pb, u32 live_words = 4)
{
    const u32 tid = threadIdx.x, w = warp_id(), ln = lane_id(), lt = (1u << ln) - 1;
    const u32 passes = (key_bits + 9) / 10, pbits = (key_bits + passes - 1) / passes;
    if (tid == 0) S.n_touched = 0;
    for (u32 t0 = 0; t0 < M; t0 += TT) {
        const u32 n = min((u32)TT, M - t0);
        __syncthreads();
        if (tid == 0) { S.n_heads[0] = S.n_heads[1] = S.n_heads[2] = 0; S.n_long = 0; }
        f.prefetch(t0 + TT, tid);                    // next tile's input towards L2 while this one is processed
        // contexts of the tile, in original order
        if (F::TILE8) f.tile8(S, t0, n);             // 8 consecutive symbols per thread, history in registers (FetchQ, 16-symbol rows)
        else
        {
            const u32 wb = w * (TT / DSRC_WARPS);
            f.begin(t0 + wb);
            typename F::Raw raw[TT / DSRC_WARPS / 32];
#pragma unroll
            for (int r = 0; r < TT / DSRC_WARPS / 32; ++r) { const u32 p = wb + r * 32 + ln; raw[r] = f.ld(t0 + p, p < n); }
#pragma unroll
            for (int r = 0; r < TT / DSRC_WARPS / 32; ++r) {
                const u32 p = wb + r * 32 + ln; const bool in = p < n;
                const u64 e = f.mk(raw[r], t0 + p, in);
                if (in) { S.el[0][p] = ((u32)(e >> 40) << TT_SHIFT) | p; S.sym[p] = (u8)(e >> 32); }
            }
        }
        __syncthreads();
        PROF_MARK(pb + 0);
        u32 cur = 0;
        for (u32 ps = 0; ps < passes; ++ps) { tile_sort_pass(S, scan, S.el[cur], S.el[cur ^ 1], n, TT_SHIFT + ps * pbits, pbits); cur ^= 1; }
        const u32* sorted = S.el[cur];
        PROF_MARK(pb + 1);
        if (N == 16) {
            // ---- scan engine: counts by segmented prefix sums, rows met by every element in parallel
            if (live_words == 1) tab_scan_groups16<1>(S, sorted, S.el[cur ^ 1], n, tab, touched);
            else if (live_words == 2) tab_scan_groups16<2>(S, sorted, S.el[cur ^ 1], n, tab, touched);
            else tab_scan_groups16<4>(S, sorted, S.el[cur ^ 1], n, tab, touched);
            PROF_MARK(pb + 3);
            tab_long_groups<N>(S, sorted, n, tab, touched);
        } else {
            // group heads, by size class
            for (u32 p0 = 0; p0 < n; p0 += DSRC_CTA) {
                const u32 p = p0 + tid; const bool in = p < n;
                const u32 key = in ? sorted[p] >> TT_SHIFT : 0u;
                int cls = -1;
                if (in && (p == 0 || (sorted[p - 1] >> TT_SHIFT) != key)) {
                    if (!(p + 2 < n && (sorted[p + 2] >> TT_SHIFT) == key)) cls = 0;
                    else if (!(p + TAB_C1 < n && (sorted[p + TAB_C1] >> TT_SHIFT) == key)) cls = 1;
                    else if (!(p + TAB_LONG < n && (sorted[p + TAB_LONG] >> TT_SHIFT) == key)) cls = 2;
                    else S.longs[atomicAdd(&S.n_long, 1u)] = (u16)p;
                }
    #pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const u32 m = __ballot_sync(FULL, cls == c);
                    if (m) {
                        u32 base = 0;
                        const int leader = __ffs(m) - 1;
                        if ((int)ln == leader) base = atomicAdd(&S.n_heads[c], (u32)__popc(m));
                        base = __shfl_sync(FULL, base, leader);
                        u16* list = c == 0 ? S.heads : c == 1 ? S.heads1 : S.heads2;
                        if (cls == c) list[base + __popc(m & lt)] = (u16)p;
                    }
                }
            }
            __syncthreads();
            PROF_MARK(pb + 2);
            tab_short_groups<N>(S, sorted, n, tab, touched);
            if (ws.prof) { __syncthreads(); PROF_MARK(pb + 3); }
            tab_long_groups<N>(S, sorted, n, tab, touched);
        }
        __syncthreads();
        PROF_MARK(pb + 4);
        for (u32 p = tid; p < n; p += DSRC_CTA) trip[t0 + p] = S.x.trip[p];
        PROF_MARK(pb + 5);
    }
    __syncthreads();
    // leave the table as it was found: first counter 0 == untouched
    const u32 nt = S.n_touched;
    for (u32 k = tid; k < nt; k += DSRC_CTA) {       // whole rows: the next block (or the other stream) may use another row size
        u8* rowp = tab + (u64)touched[k] * (2 * N);
        if (N == 4) *(uint2*)rowp = make_uint2(0u, 0u);
        else for (int j = 0; j < N / 8; ++j) ((uint4*)rowp)[j] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    PROF_MARK(pb + 6);
}
