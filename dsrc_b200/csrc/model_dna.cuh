// Direct engine of the DNA order-k modeler when the whole table fits shared memory: 4-symbol alphabet, order <= 6
// (TDnaRCOrderModeler<6,4>: 4096 contexts x 4 x u16 = 32 KiB, src/DnaModelerRCO.h:94-119) -- the headline configuration.
//
// 256 consecutive bases per step, one per thread. Bases of a step that share a context must meet the row in order; contexts of
// real DNA spread over the 4096 rows, so conflicts inside 256 bases are rare: every thread bids for its context with a shared
// atomicMin of its index, the lowest bidder of each context reads the row, emits its (freq, cum, tot) triple (one coalesced
// global store for the step) and bumps the row; the few losers bid again in the next round. Degenerate stretches
// (homopolymers: hundreds of bases in one context) fall back to one thread replaying the step's leftovers in order.
#pragma once

#define DNA_DIRECT_ROUNDS 3

struct DnaDirectShared {
    unsigned long long rows[4096];                  // 4 x u16 counters per context
    u32 owner[4096];
    u16 p_ctx[DSRC_CTA];                            // leftovers of a step for the serial replay
    u8 p_sym[DSRC_CTA], p_flag[DSRC_CTA];
};

__device__ __forceinline__ void dna_row_step(unsigned long long& row, u32 s, u32& f, u32& cum, u32& tot)
{
    u32 c0 = (u32)row & 0xFFFFu, c1 = (u32)(row >> 16) & 0xFFFFu, c2 = (u32)(row >> 32) & 0xFFFFu, c3 = (u32)(row >> 48);
    tot = c0 + c1 + c2 + c3;
    if (tot >= (1u << 16) - 8) {                     // TSymbolCoderRC<4>::Rescale
        c0 -= c0 >> 1; c1 -= c1 >> 1; c2 -= c2 >> 1; c3 -= c3 >> 1;
        tot = c0 + c1 + c2 + c3;
    }
    f = s == 0 ? c0 : s == 1 ? c1 : s == 2 ? c2 : c3;
    cum = s == 0 ? 0u : s == 1 ? c0 : s == 2 ? c0 + c1 : c0 + c1 + c2;
    if (s == 0) c0 += 2; else if (s == 1) c1 += 2; else if (s == 2) c2 += 2; else c3 += 2;
    row = (unsigned long long)c0 | ((unsigned long long)c1 << 16) | ((unsigned long long)c2 << 32) | ((unsigned long long)c3 << 48);
}

__device__ void dna_direct_engine(DnaDirectShared& D, const u8* sq, u32 M, u32 ord, u64* trip)
{
    const u32 tid = threadIdx.x, ln = lane_id();
    const u32 n_ctx = 1u << (2 * ord), mask = n_ctx - 1;
    for (u32 i = tid; i < n_ctx; i += DSRC_CTA) { D.rows[i] = 0x0001000100010001ull; D.owner[i] = 0u; }
    __syncthreads();
    // bids are (epoch << 8) | (255 - tid) under atomicMax: a new round outbids every stale entry, inside a round the lowest
    // thread (= earliest base) wins -- the owner table never needs clearing
    u32 epoch = 1;
    u32 sym_n = tid < M ? (sq[tid] & 3u) : 0u;                               // bases of the next step, loaded one step ahead
    u32 prev_n = (tid >= 32 && tid - 32 < M) ? (sq[tid - 32] & 3u) : 0u;     // the base 32 positions back (previous warp row)
    for (u32 base = 0; base < M; base += DSRC_CTA) {
        const u32 i = base + tid; const bool in = i < M;
        const u32 sym = sym_n, prev = prev_n;
        {
            const u32 j = i + DSRC_CTA;
            sym_n = j < M ? (sq[j] & 3u) : 0u;
            prev_n = (j - 32 < M) ? (sq[j - 32] & 3u) : 0u;
        }
        // context = previous `ord` bases (0 before the block start), newest in the low bits
        u32 ctx = 0;
        for (u32 t = 0; t < ord; ++t) {
            const u32 k = t + 1;
            const u32 a = __shfl_up_sync(0xFFFFFFFFu, sym, k), b = __shfl_sync(0xFFFFFFFFu, prev, (ln + 32 - k) & 31);
            ctx |= (ln >= k ? a : b) << (2 * t);
        }
        ctx &= mask;
        bool pending = in;
        for (int round = 0;; ++round, ++epoch) {
            const u32 bid = (epoch << 8) | (255u - tid);
            if (pending) atomicMax(&D.owner[ctx], bid);
            __syncthreads();
            if (pending && D.owner[ctx] == bid) {
                unsigned long long row = D.rows[ctx];
                u32 f, cum, tot;
                dna_row_step(row, sym, f, cum, tot);
                D.rows[ctx] = row;
                trip[i] = TRIP(f, cum, tot);
                pending = false;
            }
            if (!__syncthreads_or(pending)) { ++epoch; break; }
            if (round + 1 >= DNA_DIRECT_ROUNDS) {
                D.p_flag[tid] = pending; D.p_ctx[tid] = (u16)ctx; D.p_sym[tid] = (u8)sym;
                __syncthreads();
                if (tid == 0) {
                    for (u32 t = 0; t < DSRC_CTA; ++t) if (D.p_flag[t]) {
                        unsigned long long row = D.rows[D.p_ctx[t]];
                        u32 f, cum, tot;
                        dna_row_step(row, D.p_sym[t], f, cum, tot);
                        D.rows[D.p_ctx[t]] = row;
                        trip[base + t] = TRIP(f, cum, tot);
                    }
                }
                __syncthreads();
                ++epoch;
                break;
            }
        }
    }
    __syncthreads();
}
