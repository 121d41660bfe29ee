// Walk engine of the order-k quality modeler for blocks with at most 5 distinct quality values and one read length (binned
// Illumina qualities -- NovaSeq-style 4 levels plus the byte an 'N' of low quality is moved into, src/RecordsProcessor.cpp:228-233:
// the headline workload). The adaptive tables live in SHARED memory and the block is walked in its original order.
//
// Two facts make that possible. (1) With Q <= 5 symbols a hash slot of TQualityModelBase::UpdateHash (src/QualityEncoder.h:77-89)
// takes Q values (a slot holds a symbol rank or the mean of two), so the hash part of the row index of
// TTranslationalQualityEncoder<16, 4|3, 8|16> (src/QualityModelerProxy.h:231-254) is a number below Q^4 = 625 -- the index only has
// to be injective, the table is private to the block. (2) The other part of the index is the position bucket j * P / len of the
// symbol (src/QualityEncoder.h:307), P = 8 or 16: a bucket's symbols only ever meet the bucket's rows, so the P buckets are P
// independent coders. One CTA owns one block; warp b owns bucket b and its table (625 rows x 5 live counters = 6.1 KiB) and walks
// the bucket's symbols -- positions [ceil(b len / P), ceil((b + 1) len / P)) of every read, in read order -- 32 per step, exactly
// as TSymbolCoderRC::EncodeSymbol does (src/SymbolCoderRC.h:35-48): the lanes of a row that share a context are found with one
// match.any, the symbols of the lanes before a lane come from three ballots, so the counters a lane meets are
// `row + 2 * popcounts`; the last lane of every context writes the row back. No sort, no segmented scan, no table in HBM, no
// barrier between the warps of a walk. A row in which some lane's total reaches the rescale threshold (once per ~32 K hits of
// one context) is replayed symbol by symbol by lane 0 (Accumulate / Rescale, :69-90).
// Contexts are computed by the whole CTA, 4096 symbols at a time in original order, 8 consecutive symbols per thread with the
// hash window in registers (as FetchQ::tile8x), into a shared array the bucket warps pick their symbols from.
#pragma once

#define WALK_CTA 512
#define WALK_WARPS (WALK_CTA / 32)                     // = the largest bucket count (rescale 16)
#define WALK_ROWS 625                                  // 5^4 contexts per bucket
#define WALK_CHUNK (WALK_CTA * 8)
struct WalkShared {
    u64 tab[WALK_WARPS][WALK_ROWS];                    // counters 0..3 of every context, u16 each
    u16 tab4[WALK_WARPS][WALK_ROWS + 1];               // counter 4
    alignas(16) u16 ks[WALK_CHUNK];                    // (hash << 3) | symbol of the chunk's symbols, original order
    u8 rank[256];
    u32 blk, take, scheme;
};

// the block is the walk engine's: fixed read length, <= 5 quality symbols, 16-symbol rows (every lane computes the same answer)
__device__ __forceinline__ bool walk_takes(const BlockState& st, u32 order, ModelCfg& cfg, u32& scheme)
{
    if (st.status != ST_OK) return false;
    scheme = quality_order_scheme(st, order);
    if (scheme == 255 || !quality_cfg(order, scheme, cfg)) return false;
    return cfg.alpha == 16 && st.q_count <= 5 && st.q_total > 0 && st.q_total < (1u << 22) && st.min_len == st.max_len && st.max_len >= 32 && st.max_len <= 1024;
}

#ifdef WALK_MAXNREG
__global__ void __maxnreg__(WALK_MAXNREG) k_model_walk(Workspace ws)
#else
__global__ void __launch_bounds__(WALK_CTA, 2) k_model_walk(Workspace ws)
#endif
{
    extern __shared__ __align__(16) u8 walk_smem[];
    WalkShared& S = *(WalkShared*)walk_smem;
    const u32 tid = threadIdx.x, ln = tid & 31, w = tid >> 5, lt = (1u << ln) - 1;
    const u32 limit = (1u << 16) - 32, DEAD = 11;      // MaxAccumulatedValue of a 16-symbol coder; the 11 counters that stay 1
    u32* const queue = ws.model_queue + 3;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const u32 blk = atomicAdd(queue, 1u);
            S.blk = blk; S.take = 0;
            if (blk < ws.n_blocks) {
                ModelCfg c; u32 scheme = 255;
                BlockState& st = ws.state[blk];
                const bool take = walk_takes(st, ws.qua_order, c, scheme);
                st.pad[1] = take ? 1 : 0;                // the other launches of k_model<quality> skip the block
                if (take) st.q_scheme = (u8)scheme;
                S.take = take; S.scheme = scheme;
            }
        }
        __syncthreads();
        const u32 blk = S.blk;
        if (blk >= ws.n_blocks) break;
        if (!S.take) continue;
        const BlockDesc& d = ws.desc[blk];
        const BlockState& st = ws.state[blk];
        ModelCfg cfg; quality_cfg(ws.qua_order, S.scheme, cfg);
        const u32 M = st.q_total, L = st.max_len, so = cfg.sym_order, P = cfg.rescale;
        const u32 Q = st.q_count < 2 ? 2u : st.q_count;
        if (tid < 256) S.rank[tid] = st.qrank[tid];
        for (u32 j = tid; j < WALK_WARPS * WALK_ROWS; j += WALK_CTA) (&S.tab[0][0])[j] = 0x0001000100010001ull;
        for (u32 j = tid; j < WALK_WARPS * (WALK_ROWS + 1); j += WALK_CTA) (&S.tab4[0][0])[j] = 1;
        const u8* q = ws.qcat + d.sym_base;
        uint2* const trip = (uint2*)(ws.trip_q + d.sym_base);
        const u32 M1 = Q, M2 = Q * Q, M3 = Q * Q * Q;
        // this warp's bucket: read positions [sb, sb + nb)
        const u32 sb = w < P ? (w * L + P - 1) / P : 0u, nb = w < P ? ((w + 1) * L + P - 1) / P - sb : 0u;
        const u32 magic_nb = nb ? 0xFFFFFFFFu / nb + 1u : 0u, magic_L = 0xFFFFFFFFu / L + 1u;   // x / nb = umulhi(x, magic) for x < 2^20
        u64* const tab = S.tab[w]; u16* const tab4 = S.tab4[w];
        // the chunk's symbols are fetched one chunk ahead (the load would otherwise be exposed behind the barrier of every chunk)
        u64 ncur = 8 * tid < M ? *(const u64*)(q + 8 * tid) : 0ull, nprv = (tid && 8 * tid < M) ? *(const u64*)(q + 8 * tid - 8) : 0ull;
        for (u32 c0 = 0; c0 < M; c0 += WALK_CHUNK) {
            __syncthreads();                              // the previous chunk's picks are done (and, first time, the tables are set)
            {   // ---- contexts of 8 consecutive symbols per thread (FetchQ::tile8x with radix-Q hash slots)
                const u32 i = c0 + 8 * tid;
                const u64 cur = ncur, prv = nprv;
                if (i + WALK_CHUNK < M) { ncur = *(const u64*)(q + i + WALK_CHUNK); nprv = *(const u64*)(q + i + WALK_CHUNK - 8); }   // the arena has slack behind M
                if (i < M) {
                    u32 r[13];
#pragma unroll
                    for (int k = 0; k < 5; ++k) r[k] = i ? (u32)S.rank[(u32)(prv >> (8 * (3 + k))) & 255u] : 0u;
#pragma unroll
                    for (int k = 0; k < 8; ++k) r[5 + k] = S.rank[(u32)(cur >> (8 * k)) & 255u];
                    u32 a[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) a[k] = (r[k] + r[k + 1]) >> 1;
                    u32 el[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        u32 hash;
                        if (so == 4) hash = r[4 + k] + r[3 + k] * M1 + a[1 + k] * M2 + a[k] * M3;       // h = 2
                        else if (so == 3) hash = r[4 + k] + a[2 + k] * M1 + a[1 + k] * M2;              // h = 1
                        else if (so == 2) hash = r[4 + k] + a[2 + k] * M1;
                        else hash = r[4 + k];
                        el[k] = ((hash << 3) | (r[5 + k] & 7u)) & 0xFFFFu;
                    }
                    *(uint4*)&S.ks[8 * tid] = make_uint4(el[0] | (el[1] << 16), el[2] | (el[3] << 16), el[4] | (el[5] << 16), el[6] | (el[7] << 16));
                }
            }
            __syncthreads();
            if (!nb) continue;
            // ---- this bucket's symbols inside the chunk: elements [e_lo, e_hi) of the bucket's own numbering (read-major)
            const u32 c1 = min(c0 + (u32)WALK_CHUNK, M);
            u32 e_lo, e_hi;
            { const u32 rd = __umulhi(c0, magic_L), rem = c0 - rd * L; e_lo = rd * nb + (rem > sb ? min(rem - sb, nb) : 0u); }
            { const u32 rd = __umulhi(c1, magic_L), rem = c1 - rd * L; e_hi = rd * nb + (rem > sb ? min(rem - sb, nb) : 0u); }
            for (u32 e0 = e_lo; e0 < e_hi; e0 += 32) {
                const u32 e = e0 + ln;
                const bool valid = e < e_hi;
                const u32 rd = __umulhi(e, magic_nb);
                const u32 idx = valid ? rd * L + sb + (e - rd * nb) : c0;       // position in the block
                const u32 e16 = S.ks[idx - c0];
                const u32 key = e16 >> 3, s = e16 & 7u;
                const u32 peers = __match_any_sync(FULL, valid ? key : 0x10000u + ln);
                const u32 b1 = __ballot_sync(FULL, e16 & 1u), b2 = __ballot_sync(FULL, e16 & 2u), b4 = __ballot_sync(FULL, e16 & 4u);
                const u32 below = peers & lt, below03 = below & ~b4;
                const u32 n0 = __popc(below03 & ~b1 & ~b2), n1 = __popc(below03 & b1 & ~b2), n2 = __popc(below03 & ~b1 & b2), n3 = __popc(below03 & b1 & b2);
                const u32 n4 = __popc(below & b4);
                const u32 kk = valid ? key : 0u;
                const u64 row = tab[kk];
                const u32 c4 = tab4[kk];
                const u32 v01 = (u32)row + 2u * (n0 | (n1 << 16)), v23 = (u32)(row >> 32) + 2u * (n2 | (n3 << 16)), v4 = c4 + 2u * n4;
                const u32 p01 = v01 * 0x10001u, p23 = v23 * 0x10001u;          // low half: first counter, high half: sum of the pair
                const u32 sum03 = (p01 >> 16) + (p23 >> 16);
                const u32 tot = sum03 + v4 + DEAD;
                if (__any_sync(FULL, valid && tot >= limit)) {
                    // a rescale falls into this row: lane 0 codes it symbol by symbol (TSymbolCoderRC::Accumulate / Rescale)
                    const u32 vm = __ballot_sync(FULL, valid);
                    __syncwarp();                        // every lane has read its row
                    for (u32 x = 0; x < 32; ++x) {
                        const u32 kx = __shfl_sync(FULL, key, x), sx = __shfl_sync(FULL, s, x), ix = __shfl_sync(FULL, idx, x);
                        if (!((vm >> x) & 1u)) break;
                        if (ln == 0) {
                            const u64 rw = tab[kx];
                            u32 c[5] = {(u32)rw & 0xFFFFu, (u32)(rw >> 16) & 0xFFFFu, (u32)(rw >> 32) & 0xFFFFu, (u32)(rw >> 48), (u32)tab4[kx]};
                            u32 T = c[0] + c[1] + c[2] + c[3] + c[4] + DEAD;
                            if (T >= limit) { T = DEAD; for (int y = 0; y < 5; ++y) { c[y] -= c[y] >> 1; T += c[y]; } }
                            u32 cum = 0;
                            for (u32 y = 0; y < sx; ++y) cum += c[y];
                            trip[ix] = make_uint2(c[sx] | (cum << 16), T);
                            c[sx] += 2;
                            tab[kx] = (u64)(c[0] | (c[1] << 16)) | ((u64)(c[2] | (c[3] << 16)) << 32);
                            tab4[kx] = (u16)c[4];
                        }
                        __syncwarp();
                    }
                    continue;
                }
                // freq = counter s, cum = counters below s: 0, c0, c0 + c1, c0 + c1 + c2, c0 + c1 + c2 + c3
                const u32 pair = (s & 2u) ? v23 : v01;
                u32 f = (s & 1u) ? pair >> 16 : pair & 0xFFFFu;
                const u32 base2 = (s & 2u) ? p01 >> 16 : 0u;
                const u32 lowp = (s & 2u) ? p23 : p01;
                u32 cum = base2 + ((s & 1u) ? lowp & 0xFFFFu : 0u);
                if (s & 4u) { f = v4; cum = sum03; }
                if (valid) trip[idx] = make_uint2(f | (cum << 16), tot);
                __syncwarp();                            // every lane has its row: the rows may change
                if (valid && (peers >> ln) == 1u) {      // no peer above me: my view plus my own symbol is the row after this step
                    const u32 inc = (s & 4u) ? 0u : 2u << ((s & 1u) * 16);
                    tab[key] = (u64)(v01 + ((s & 2u) ? 0u : inc)) | ((u64)(v23 + ((s & 2u) ? inc : 0u)) << 32);
                    tab4[key] = (u16)(v4 + ((s & 4u) ? 2u : 0u));
                }
                __syncwarp();
            }
        }
    }
}

// ---- DNA: TDnaRCOrderModeler<order <= 6, 4 symbols> (src/DnaModelerRCO.h:45-62,94-119), one WARP per block. The table -- 4096
// contexts x 4 x u16 = 32 KiB -- lives in shared memory and the block is walked 32 bases per step as above (match.any on the
// context, two ballots for the 2-bit symbols, the last lane of a context writes its row back): ~2 warp instructions per base and no
// CTA barrier, against the bidding rounds of the CTA-wide direct engine (model_dna.cuh). Contexts are built 8 consecutive bases per
// lane (one 8-byte load, the 6 bases of history come from the neighbouring lane) and transposed to row order through shared memory;
// the table-independent part of the 8 rows of a step is computed before the walk so that only `load row, add, store row` is serial.
#define DWALK_STEP 256
struct DnaWalkShared {
    u64 tab[4096];
    alignas(16) u16 ks[DWALK_STEP];                    // (context << 2) | base
    u8 claim[4096];                                    // lane that wrote last, per context (collision detection inside a row)
};

__global__ void __launch_bounds__(32) k_dna_walk(Workspace ws)
{
    extern __shared__ __align__(16) u8 walk_smem[];
    DnaWalkShared& S = *(DnaWalkShared*)walk_smem;
    const u32 ln = threadIdx.x, lt = (1u << ln) - 1;
    const u32 limit = (1u << 16) - 8;                  // MaxAccumulatedValue of a 4-symbol coder
    u32* const queue = ws.model_queue + 3;             // (the quality walk launch of this batch has drained its counter: reset below)
    const u32 ord = ws.dna_order > 6 ? 6u : ws.dna_order, mask = (1u << (2 * ord)) - 1;
    for (;;) {
        u32 blk = 0;
        if (ln == 0) blk = atomicAdd(queue, 1u);
        blk = __shfl_sync(FULL, blk, 0);
        if (blk >= ws.n_blocks) break;
        const BlockDesc& d = ws.desc[blk];
        BlockState& st = ws.state[blk];
        bool take = st.status == ST_OK && ws.dna_order <= 6 && st.d_count > 0 && st.d_count <= 4 && st.d_total > 0;
        for (u32 i = 4; i < 20 && take; ++i) if (st.dfreq[i]) take = false;       // the other launch reports it
        __syncwarp();
        if (ln == 0) { st.pad[1] = (u8)((st.pad[1] & 1u) | (take ? 2u : 0u)); if (take) st.d_scheme = 0; }
        if (!take) continue;
        const u32 M = st.d_total;
        for (u32 j = ln; j <= mask; j += 32) S.tab[j] = 0x0001000100010001ull;
        __syncwarp();
        const u8* sq = ws.dcat + d.sym_base;
        uint2* const trip = (uint2*)(ws.trip_d + d.sym_base);
        u64 cur = 8 * ln < M ? *(const u64*)(sq + 8 * ln) : 0ull;         // the arena has slack behind M
        u32 carry = 0;                                                    // packed bases of the previous step's last lane
        for (u32 t0 = 0; t0 < M; t0 += DWALK_STEP) {
            const u32 i = t0 + 8 * ln;
            const u64 nxt = i + DWALK_STEP < M ? *(const u64*)(sq + i + DWALK_STEP) : 0ull;
            // 8 bases -> 16 bits, oldest in the low bits
            const u32 lo = (u32)cur, hi = (u32)(cur >> 32);
            u32 p16 = ((lo & 3u) | ((lo >> 6) & 0xCu) | ((lo >> 12) & 0x30u) | ((lo >> 18) & 0xC0u)) |
                      (((hi & 3u) | ((hi >> 6) & 0xCu) | ((hi >> 12) & 0x30u) | ((hi >> 18) & 0xC0u)) << 8);
            if (i + 8 > M) p16 &= i < M ? (1u << (2 * (M - i))) - 1 : 0u;
            u32 prev = __shfl_up_sync(FULL, p16, 1);
            if (ln == 0) prev = carry;
            carry = __shfl_sync(FULL, p16, 31);
            const u32 W = (prev >> 4) | (p16 << 12);                      // bases i-6 .. i+7
            u32 el[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) el[k] = ((((W >> (12 + 2 * k - 2 * ord)) & mask) << 2) | ((W >> (12 + 2 * k)) & 3u)) & 0xFFFFu;
            *(uint4*)&S.ks[8 * ln] = make_uint4(el[0] | (el[1] << 16), el[2] | (el[3] << 16), el[4] | (el[5] << 16), el[6] | (el[7] << 16));
            __syncwarp();
            // ---- the walk. Contexts of DNA rarely collide inside a row of 32 bases and match.any costs a step per distinct value, so
            // collisions are detected through shared memory first: every lane writes its number to its context's claim cell; if all
            // lanes read their own number back, every lane is alone with its row (the common case). Otherwise the peers come from
            // twelve ballots (row32). Two rows go through the detection together (a lane holds base l and base 32 + l of the 64; the
            // second row's claims are written after the first's): when all 64 contexts differ -- 6 times out of 10 -- the 64 bases
            // meet the table in one go, with two warp barriers less and one loop's worth of bookkeeping less than two single rows.
            auto row32 = [&](u32 j) {
                if (t0 + 32 * j >= M) return;
                const u32 e = S.ks[32 * j + ln];
                const bool valid = t0 + 32 * j + ln < M;
                const u32 key = valid ? e >> 2 : 0u, s = e & 3u;
                if (valid) S.claim[key] = (u8)ln;
                __syncwarp();
                const u32 won = S.claim[key];
                const u64 row = S.tab[key];
                u32 v01 = (u32)row, v23 = (u32)(row >> 32);
                bool last = valid;
                if (__ballot_sync(FULL, valid && won != ln)) {
                    u32 peers = __ballot_sync(FULL, valid);
#pragma unroll
                    for (int k = 0; k < 12; ++k) { const u32 m = __ballot_sync(FULL, (key >> k) & 1u); peers &= ((key >> k) & 1u) ? m : ~m; }
                    const u32 b1 = __ballot_sync(FULL, e & 1u), b2 = __ballot_sync(FULL, e & 2u);
                    const u32 below = peers & lt;
                    v01 += 2u * (__popc(below & ~b1 & ~b2) | (__popc(below & b1 & ~b2) << 16));
                    v23 += 2u * (__popc(below & ~b1 & b2) | (__popc(below & b1 & b2) << 16));
                    last = valid && (peers >> ln) == 1u;
                }
                const u32 p01 = v01 * 0x10001u, p23 = v23 * 0x10001u;    // low half: first counter, high half: sum of the pair
                const u32 tot = (p01 >> 16) + (p23 >> 16);
                if (__any_sync(FULL, valid && tot >= limit)) {
                    // a rescale falls into this row: lane 0 codes it base by base (TSymbolCoderRC::Accumulate / Rescale)
                    const u32 vm = __ballot_sync(FULL, valid);
                    __syncwarp();                        // every lane has read its row
                    for (u32 x = 0; x < 32; ++x) {
                        const u32 kx = __shfl_sync(FULL, key, x), sx = __shfl_sync(FULL, s, x);
                        if (!((vm >> x) & 1u)) break;
                        if (ln == 0) {
                            unsigned long long rw = S.tab[kx];
                            u32 f, cum, T;
                            dna_row_step(rw, sx, f, cum, T);
                            S.tab[kx] = rw;
                            trip[t0 + 32 * j + x] = make_uint2(f | (cum << 16), T);
                        }
                        __syncwarp();
                    }
                    return;
                }
                const u32 pair = (s & 2u) ? v23 : v01;
                const u32 f = (s & 1u) ? pair >> 16 : pair & 0xFFFFu;
                const u32 lowp = (s & 2u) ? p23 : p01;
                const u32 cum = ((s & 2u) ? p01 >> 16 : 0u) + ((s & 1u) ? lowp & 0xFFFFu : 0u);
                if (valid) trip[t0 + 32 * j + ln] = make_uint2(f | (cum << 16), tot);
                __syncwarp();                            // every lane has its row: the rows may change
                if (last) {                             // no peer above me: my view plus my own base is the row after this step
                    const u32 inc = 2u << ((s & 1u) * 16);
                    S.tab[key] = (u64)(v01 + ((s & 2u) ? 0u : inc)) | ((u64)(v23 + ((s & 2u) ? inc : 0u)) << 32);
                }
                __syncwarp();
            };
#pragma unroll 1
            for (u32 J = 0; J < 4; ++J) {
                if (t0 + 64 * J >= M) break;
                const u32 ea = S.ks[64 * J + ln], eb = S.ks[64 * J + 32 + ln];
                const u32 ia = t0 + 64 * J + ln, ib = ia + 32;
                const bool va = ia < M, vb = ib < M;
                const u32 ka = va ? ea >> 2 : 0u, kb = vb ? eb >> 2 : 0u, sa = ea & 3u, sb = eb & 3u;
                if (va) S.claim[ka] = (u8)ln;
                __syncwarp();
                if (vb) S.claim[kb] = (u8)(32 + ln);
                __syncwarp();
                const u32 wa = S.claim[ka], wb = S.claim[kb];
                const u64 ra = S.tab[ka], rb_ = S.tab[kb];
                const u32 a01 = (u32)ra, a23 = (u32)(ra >> 32), b01 = (u32)rb_, b23 = (u32)(rb_ >> 32);
                const u32 pa01 = a01 * 0x10001u, pa23 = a23 * 0x10001u, pb01 = b01 * 0x10001u, pb23 = b23 * 0x10001u;
                const u32 ta = (pa01 >> 16) + (pa23 >> 16), tb = (pb01 >> 16) + (pb23 >> 16);
                if (__any_sync(FULL, (va && (wa != ln || ta >= limit)) || (vb && (wb != 32 + ln || tb >= limit)))) {
                    __syncwarp();                        // (every lane has read its claim cells: row32 writes them again)
                    row32(2 * J); row32(2 * J + 1);
                    continue;
                }
                {
                    const u32 pair = (sa & 2u) ? a23 : a01, lowp = (sa & 2u) ? pa23 : pa01;
                    const u32 f = (sa & 1u) ? pair >> 16 : pair & 0xFFFFu;
                    const u32 cum = ((sa & 2u) ? pa01 >> 16 : 0u) + ((sa & 1u) ? lowp & 0xFFFFu : 0u);
                    if (va) trip[ia] = make_uint2(f | (cum << 16), ta);
                }
                {
                    const u32 pair = (sb & 2u) ? b23 : b01, lowp = (sb & 2u) ? pb23 : pb01;
                    const u32 f = (sb & 1u) ? pair >> 16 : pair & 0xFFFFu;
                    const u32 cum = ((sb & 2u) ? pb01 >> 16 : 0u) + ((sb & 1u) ? lowp & 0xFFFFu : 0u);
                    if (vb) trip[ib] = make_uint2(f | (cum << 16), tb);
                }
                __syncwarp();                            // every lane has its rows: the rows may change
                if (va) { const u32 inc = 2u << ((sa & 1u) * 16); S.tab[ka] = (u64)(a01 + ((sa & 2u) ? 0u : inc)) | ((u64)(a23 + ((sa & 2u) ? inc : 0u)) << 32); }
                if (vb) { const u32 inc = 2u << ((sb & 1u) * 16); S.tab[kb] = (u64)(b01 + ((sb & 2u) ? 0u : inc)) | ((u64)(b23 + ((sb & 2u) ? inc : 0u)) << 32); }
                __syncwarp();
            }
            cur = nxt;
        }
    }
}
