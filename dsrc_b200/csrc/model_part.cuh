// Partition engine of the order-k quality modelers for alphabets of 32 / 64 / 128 symbols (schemes 1-3 and 5-7 of
// QualityOrderModelerProxyLossless, src/QualityModelerProxy.h:231-254) -- table-free.
//
// TSymbolCoderRC<N>::EncodeSymbol (src/SymbolCoderRC.h:35-48) meets, for the i-th symbol of a context, the row
//     stats[q] = 1 + 2 #(earlier symbols of the context equal to q)
// as long as no rescale fired (Rescale needs >= 2^15 - N symbols in one context, src/SymbolCoderRC.h:69-90). So the triple of a symbol
//     freq = 1 + 2 #(earlier, same symbol)    cum = sym + 2 #(earlier, smaller symbol)    tot = N + 2 #(earlier)
// is a function of the symbols of its own context alone, and a 64-symbol row (128 B, in a 32 MiB table that only HBM can hold) never
// has to exist. What the engine needs is every context's symbols side by side, in original order:
//   pass 0  contexts of the block, 2048 symbols at a time (FetchQ::tile8x: hash window in registers); the row index is scrambled
//           with an odd multiplier (a bijection on key_bits-bit keys), so that its TOP bits spread sticky data evenly; count the
//           elements per partition (top `pb` bits, about 100 elements each).
//   pass 1  the same tiles again; each is sorted by partition in shared memory (one stable radix pass) and leaves for the CTA's arena
//           in HBM as (key' << 40 | symbol << 32 | index), every partition contiguous and in original order.
//   pass 2  every WARP takes a few consecutive partitions (<= 256 elements) on its own -- no CTA barrier from here on: loads them,
//           sorts them by key' in its slice of shared memory (one stable counting pass with warp-private counters: every context
//           becomes a run in original order) and walks the run a row of 32 at a time: the lanes of a row that belong to the same run
//           are a lane mask, "earlier with the same / a smaller symbol" are popcounts of two masks built from one ballot per symbol
//           bit; a run that crosses rows carries its symbol counts in a per-warp histogram. No divergent walks, no table, no
//           scattered row traffic. Triples are staged in shared memory and stored at their original index.
// A partition of 257..2048 elements is handled the same way by the whole CTA (two radix passes, runs that cross a warp's segment go
// to a second round); one of more than 2048 elements (one very hot context, or a multi-MB block) takes the sort engine's path on its
// span of the arena (sort_pass + group_scan, rc_model.cu), which also knows how to rescale.
#pragma once

#define PART_MUL 0x9E3779B1u                  // odd: key -> key * PART_MUL mod 2^key_bits is a bijection
#define PW 256                                // elements of a warp's tile
#define PW_SHIFT 8
#ifndef PART_TARGET
#define PART_TARGET 128                       // a partition holds at most about this many elements (a warp's tile takes two or three)
#endif

// One warp walks sorted[lo, hi) a row of 32 elements at a time (element = (key << SH) | position in the loaded tile).
// main mode: lo = start of the warp's segment; a run that began before the segment is left to the long walker (queued in longs).
// long mode: lo = head of ONE run that crosses a segment boundary; stops after the row in which that run ends.
template <int NB, int SH>
__device__ __noinline__ void part_rows(const u32* sorted, const u8* sym, u64* stage, u16* B, u16* P, u16* longs, u32* n_long,
                                          u32 lo, u32 hi, bool long_mode, u32 N)
{
    const u32 ln = lane_id(), lt_mask = (1u << ln) - 1, le_mask = lt_mask | (1u << ln);
    const u32 NOKEY = 0xFFFFFFFFu, PMASK = (1u << SH) - 1;
    const u32 per = N >> 5;                           // row entries per lane (N = 32, 64, 128)
    u32 openkey = NOKEY, openpos = 0;
    bool skipping = false;
    if (lo >= hi) return;
    if (!long_mode && lo > 0) {
        openkey = sorted[lo - 1] >> SH;
        if ((sorted[lo] >> SH) == openkey) {          // the segment starts inside a run
            skipping = true;
            if (ln == 0) longs[atomicAdd(n_long, 1u)] = (u16)lo;
        }
    }
    const u32 runkey = sorted[lo] >> SH;
    for (u32 row = lo; row < hi; row += 32) {
        const u32 i = row + ln; const bool valid = i < hi;
        const u32 e = valid ? sorted[i] : 0u;
        const u32 key = valid ? e >> SH : NOKEY - 1;
        u32 pk = __shfl_up_sync(FULL, key, 1); if (ln == 0) pk = openkey;
        const u32 headmask = __ballot_sync(FULL, valid && key != pk);
        const u32 validmask = __ballot_sync(FULL, valid);
        const u32 hm = headmask & le_mask;
        const int start = hm ? 31 - __clz(hm) : -1;   // lane where my run starts in this row; -1: it came in from the row before
        const u32 lowmask = start > 0 ? (lt_mask & ~((1u << start) - 1)) : lt_mask;    // earlier members of my run in this row
        const u32 s = valid ? (u32)sym[e & PMASK] : 0u;
        u32 ltm = 0, eqm = validmask;                 // lanes holding a smaller / the same symbol
#pragma unroll
        for (int k = NB - 1; k >= 0; --k) {
            const u32 b = __ballot_sync(FULL, (s >> k) & 1u);
            if ((s >> k) & 1u) { ltm |= eqm & ~b; eqm &= b; } else eqm &= ~b;
        }
        u32 nf = __popc(eqm & lowmask), nc = __popc(ltm & lowmask), pos = __popc(lowmask);
        const bool cont = start < 0;
        if (!(headmask & 1u) && !skipping) {          // lane 0 continues the open run: its earlier rows are in B
            u32 loc[4], sum = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) if ((u32)k < per) { loc[k] = sum; sum += B[ln * per + k]; }
            const u32 off = warp_incl_sum(sum) - sum;
#pragma unroll
            for (int k = 0; k < 4; ++k) if ((u32)k < per) P[ln * per + k] = (u16)(loc[k] + off);
            __syncwarp();
            if (cont && valid) { nf += B[s]; nc += P[s]; pos += openpos; }
        }
        if (valid && !(cont && skipping)) stage[e & PMASK] = TRIP(1 + 2 * nf, s + 2 * nc, N + 2 * pos);
        if (headmask) skipping = false;
        const u32 lastkey = __shfl_sync(FULL, key, 31);
        const u32 nextkey = row + 32 < hi ? sorted[row + 32] >> SH : NOKEY;
        __syncwarp();
        if (nextkey == lastkey) {                     // the row's last run goes on: leave its symbol counts for the next row
            u32 members = FULL;
            if (headmask) {
                const int L = 31 - __clz(headmask);
                members = ~((1u << L) - 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) if ((u32)k < per) B[ln * per + k] = 0;
                openpos = 0;
                __syncwarp();
            }
            if (((members >> ln) & 1u) && (eqm & members & lt_mask) == 0) B[s] += (u16)__popc(eqm & members);
            openpos += __popc(members);
            __syncwarp();
        }
        openkey = lastkey;
        if (long_mode && nextkey != runkey) break;
    }
}

// the whole CTA on one sorted tile of <= 2048 elements: every warp its 256-element segment, then the runs that cross segments
template <int NB>
__device__ void part_tile(TabShared& S, const u32* sorted, u32 n, u32 N)
{
    const u32 w = warp_id();
    part_rows<NB, TT_SHIFT>(sorted, S.sym, S.x.trip, S.pB[w], S.pP[w], S.longs, &S.n_long,
                            min(n, w * (TT / DSRC_WARPS)), min(n, (w + 1) * (TT / DSRC_WARPS)), false, N);
    __syncthreads();
    const u32 nl = S.n_long;
    for (u32 g = w; g < nl; g += DSRC_WARPS) {
        // the run holding S.longs[g] (a segment start inside it): walk back to its head; of several segment starts inside one run only
        // the first one does the work
        const u32 pos = S.longs[g];
        u32 h = pos;
        if (lane_id() == 0) { const u32 key = sorted[pos] >> TT_SHIFT; while (h > 0 && (sorted[h - 1] >> TT_SHIFT) == key) --h; }
        h = __shfl_sync(FULL, h, 0);
        if ((((h >> 8) + 1) << 8) != pos) continue;
        part_rows<NB, TT_SHIFT>(sorted, S.sym, S.x.trip, S.pB[w], S.pP[w], S.longs, &S.n_long, h, n, true, N);
    }
}

// one warp on its own tile: partitions spanning arena[a, a + n), n <= PW; keys relative to `base`, gk significant key bits (<= 10)
template <int NB>
__device__ void part_warp_tile(TabShared& S, const u64* arena, u32 a, u32 n, u32 base, u32 gk, u32 N, u64* trip)
{
    const u32 w = warp_id(), ln = lane_id(), lt = (1u << ln) - 1;
    u32* el1 = S.el[1] + w * PW; u8* sym = S.sym + w * PW;
    u64* stage = S.x.trip + w * PW;                    // 2 KiB per warp: the sort's counters, then the staged triples
    u16* H = (u16*)stage;
    constexpr int ROWS = PW / 32;
    u32 e[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
        const u32 i = k * 32 + ln;
        e[k] = 0xFFFFFFFFu;
        if (i < n) { const u64 v = arena[a + i]; e[k] = (((u32)(v >> 40) - base) << PW_SHIFT) | i; sym[i] = (u8)(v >> 32); }
    }
    const u32 bins = 1u << gk;
    for (u32 i = ln; i < (bins + 1) / 2; i += 32) ((u32*)H)[i] = 0;
    __syncwarp();
    u32 peers[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
        const bool in = (u32)k * 32 + ln < n;
        const u32 d = e[k] >> PW_SHIFT;
        u32 p = __ballot_sync(FULL, in);
        if (in) p = __match_any_sync(p, d);
        peers[k] = in ? p : 0u;
        if (in && (__ffs(p) - 1) == (int)ln) H[d] += (u16)__popc(p);
        __syncwarp();
    }
    {   // exclusive scan of the counters: every lane a contiguous run of bins
        const u32 per = bins > 32 ? bins / 32 : 1u, b0 = ln * per;
        u32 sum = 0;
        if (b0 < bins) for (u32 k = 0; k < per; ++k) sum += H[b0 + k];
        u32 run = warp_incl_sum(sum) - sum;
        if (b0 < bins) for (u32 k = 0; k < per; ++k) { const u32 c = H[b0 + k]; H[b0 + k] = (u16)run; run += c; }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
        const u32 p = peers[k], d = e[k] >> PW_SHIFT;
        const u32 pos = p ? H[d] + __popc(p & lt) : 0u;
        __syncwarp();
        if (p) {
            if ((__ffs(p) - 1) == (int)ln) H[d] += (u16)__popc(p);
            el1[pos] = e[k];
        }
        __syncwarp();
    }
    part_rows<NB, PW_SHIFT>(el1, sym, stage, S.pB[w], S.pP[w], S.longs, &S.n_long, 0, n, false, N);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
        const u32 i = k * 32 + ln;
        if (i < n) trip[(u32)arena[a + i]] = stage[i];
    }
    __syncwarp();
}

// F = FetchQ. arena / spare: the CTA's two sort buffers (M entries each); spare is only used by oversize partitions.
// returns false (nothing written) when the block has to take the sort engine instead
template <class F>
__device__ bool part_engine(ModelShared& MS, F f, u32 M, u32 key_bits, u32 N, u32 sym_bits, u64* arena, u64* spare, u64* trip,
                            const Workspace& ws, long long& prof_t, int prof_base)
{
    TabShared& S = MS.u.tab;
    PartState& PS = S.part;
    const u32 tid = threadIdx.x;
    // (a warp-private form of passes 0/1 -- every warp its own eighth of the block, per-(warp, partition) offsets, no CTA barrier --
    // was measured 20 % SLOWER on B200: the partitions are then written in eight interleaved pieces and pass 2 waits longer for them)
    // partitions of about 100 elements (a warp's tile holds two or three), at most 2^PART_MAX_BITS of them and never more than keys
    u32 pb = key_bits > 13 ? key_bits - 10 : 3;       // a warp's counting sort has 2^10 counters: at most 10 key bits below the partition
    while (pb < PART_MAX_BITS && pb < key_bits && (M >> pb) > PART_TARGET) ++pb;
    const u32 bins = 1u << pb, lowbits = key_bits - pb;
    f.kmul = PART_MUL; f.kmask = (1u << key_bits) - 1;

    // ---- pass 0: elements per partition
    for (u32 d = tid; d < bins; d += DSRC_CTA) PS.pend[d] = 0;
    __syncthreads();
    for (u32 t0 = 0; t0 < M; t0 += TT) {
        const u32 n = min((u32)TT, M - t0);
        f.prefetch(t0 + TT, tid);
        f.template tile8x<true>(S, t0, n);
        const u32 p0 = tid * 8;
        if (p0 < n) {
            const uint4 a = ((const uint4*)S.el[0])[tid * 2], b = ((const uint4*)S.el[0])[tid * 2 + 1];
            const u32 e[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            u32 run = 0, dprev = e[0] >> (TT_SHIFT + lowbits);
#pragma unroll
            for (int k = 0; k < 8; ++k) {             // neighbours may share a partition: one atomic per run
                const bool in = p0 + k < n;
                const u32 d = e[k] >> (TT_SHIFT + lowbits);
                if (in && d == dprev) ++run;
                else { if (run) atomicAdd(&PS.pend[dprev], run); run = in ? 1u : 0u; dprev = d; }
            }
            if (run) atomicAdd(&PS.pend[dprev], run);
        }
    }
    __syncthreads();
    {
        const u32 per = bins > DSRC_CTA ? bins / DSRC_CTA : 1u, d0 = tid * per;
        u32 sum = 0, c[4];
        if (d0 < bins) for (u32 k = 0; k < per; ++k) { c[k] = PS.pend[d0 + k]; sum += c[k]; }
        u32 total, run = block_excl_sum(sum, MS.scan, &total);
        if (d0 < bins) for (u32 k = 0; k < per; ++k) { PS.pend[d0 + k] = run; run += c[k]; }
    }
    __syncthreads();
    PROF_MARK(prof_base + 0);

    // ---- pass 1: every tile sorted by partition in shared memory (one stable radix pass), then appended to the partitions in the arena
    for (u32 t0 = 0; t0 < M; t0 += TT) {
        const u32 n = min((u32)TT, M - t0);
        f.prefetch(t0 + TT, tid);
        f.template tile8x<true>(S, t0, n);
        __syncthreads();
        tile_sort_pass(S, MS.scan, S.el[0], S.el[1], n, TT_SHIFT + lowbits, pb);
        const u32* sorted = S.el[1];
        const u16* HE = S.x.H + (DSRC_WARPS - 1) * bins;      // after the scatter: end of every partition's slice of the sorted tile
        for (u32 j = tid; j < n; j += DSRC_CTA) {
            const u32 e = sorted[j], d = e >> (TT_SHIFT + lowbits), p = e & (TT - 1);
            const u32 ts = d ? HE[d - 1] : 0u;
            arena[PS.pend[d] + (j - ts)] = ((u64)(e >> TT_SHIFT) << 40) | ((u64)S.sym[p] << 32) | (t0 + p);
        }
        __syncthreads();
        for (u32 d = tid; d < bins; d += DSRC_CTA) PS.pend[d] += (u32)HE[d] - (d ? (u32)HE[d - 1] : 0u);
        __syncthreads();
    }
    PROF_MARK(prof_base + 1);

    // ---- pass 2: consecutive partitions are grouped into warp tiles (<= PW elements, <= 10 significant key bits); a partition that
    // does not fit one is flagged for the whole CTA. glist[g] = first partition of group g (| 0x8000: CTA), glist[n_groups] = bins
    u16* glist = S.heads;
    if (tid == 0) {
        u32 ng = 0, d = 0;
        while (d < bins) {
            const u32 a = d ? PS.pend[d - 1] : 0u;
            if (PS.pend[d] - a > PW) { glist[ng++] = (u16)(d | 0x8000u); ++d; continue; }
            u32 d1 = d + 1;
            while (d1 < bins && PS.pend[d1] - a <= PW && lowbits + (32 - __clz(d1 - d)) <= 10) ++d1;     // bit_length(span - 1) with span = d1 + 1 - d
            glist[ng++] = (u16)d; d = d1;
        }
        glist[ng] = (u16)bins;
        PS.n = ng; PS.a = 0;
    }
    __syncthreads();
    {
        const u32 ng = PS.n;
        u32 gn = 0;                                    // the tile after this one: taken early so that its span of the arena is on its way
        if (lane_id() == 0) gn = atomicAdd(&PS.a, 1u);
        gn = __shfl_sync(FULL, gn, 0);
        for (;;) {
            const u32 g = gn;
            if (g >= ng) break;
            if (lane_id() == 0) gn = atomicAdd(&PS.a, 1u);
            gn = __shfl_sync(FULL, gn, 0);
            if (gn < ng && !(glist[gn] & 0x8000u)) {
                const u32 p0 = glist[gn], p1 = glist[gn + 1] & 0x7FFFu;
                const u32 pa = p0 ? PS.pend[p0 - 1] : 0u, pn = PS.pend[p1 - 1] - pa;
                if (lane_id() * 16 < pn) asm volatile("prefetch.global.L2 [%0];" :: "l"(arena + pa + lane_id() * 16));
            }
            const u32 d0 = glist[g];
            if (d0 & 0x8000u) continue;
            const u32 d1 = glist[g + 1] & 0x7FFFu;
            const u32 a = d0 ? PS.pend[d0 - 1] : 0u, n = PS.pend[d1 - 1] - a;
            if (n == 0) continue;
            u32 span = d1 - d0 - 1, gk = lowbits;
            while (span) { ++gk; span >>= 1; }
            if (sym_bits <= 6) part_warp_tile<6>(S, arena, a, n, d0 << lowbits, gk, N, trip);
            else part_warp_tile<7>(S, arena, a, n, d0 << lowbits, gk, N, trip);
        }
    }
    __syncthreads();
    PROF_MARK(prof_base + 2);
    // ---- the partitions that did not fit a warp's tile
    const u32 ng = PS.n;
    for (u32 g = 0; g < ng; ++g) {
        const u32 dg = glist[g];
        if (!(dg & 0x8000u)) continue;
        const u32 d0 = dg & 0x7FFFu;
        const u32 a = d0 ? PS.pend[d0 - 1] : 0u, n = PS.pend[d0] - a;
        if (n > TT) {
            // LSD passes over its span of the arena on the remaining key bits, then the sort engine's run walker. This clobbers the
            // shared-memory union (glist included): such partitions are few, the list is rebuilt by skipping to the next flagged one
            u64* src = arena + a; u64* dst = spare + a;
            if (lowbits) {
                const u32 gp = (lowbits + SORT_MAX_BITS - 1) / SORT_MAX_BITS, gb = (lowbits + gp - 1) / gp;
                for (u32 p = 0; p < gp; ++p) {
                    FetchSorted fs; fs.src = src;
                    sort_pass(MS, fs, dst, n, 40 + p * gb, gb);
                    u64* t = src; src = dst; dst = t;
                }
            }
            group_scan(MS, src, (u32*)dst, trip, n, ws, prof_t, prof_base + 8);
            __syncthreads();
            // rebuild the flags this path destroyed (only the flagged entries are read from here on)
            if (tid == 0) { for (u32 k = g + 1; k < ng; ++k) glist[k] = 0; u32 k = g + 1; for (u32 d = d0 + 1; d < bins && k < ng; ++d) if (PS.pend[d] - PS.pend[d - 1] > PW) glist[k++] = (u16)(d | 0x8000u); }
            __syncthreads();
            PROF_MARK(prof_base + 6);
            continue;
        }
        if (tid == 0) S.n_long = 0;
        for (u32 i = tid; i < n; i += DSRC_CTA) {
            const u64 e = arena[a + i];
            S.el[0][i] = (((u32)(e >> 40) - (d0 << lowbits)) << TT_SHIFT) | i;
            S.sym[i] = (u8)(e >> 32);
        }
        __syncthreads();
        u32 cur = 0;
        if (lowbits) { tile_sort_pass_t<10>(S, MS.scan, S.el[0], S.el[1], n, TT_SHIFT); cur = 1; }     // lowbits <= 10: the idle bins of a wider digit only cost their zeroing
        if (sym_bits <= 6) part_tile<6>(S, S.el[cur], n, N);
        else part_tile<7>(S, S.el[cur], n, N);
        __syncthreads();
        for (u32 i = tid; i < n; i += DSRC_CTA) trip[(u32)arena[a + i]] = S.x.trip[i];
        __syncthreads();
    }
    PROF_MARK(prof_base + 3);
    return true;
}
