// Direct engine of the DNA order-k modeler when the whole table fits shared memory: 4-symbol alphabet, order <= 6
// (TDnaRCOrderModeler<6,4>: 4096 contexts x 4 x u16 = 32 KiB, src/DnaModelerRCO.h:94-119) -- the headline configuration.
//
// 1024 consecutive bases per step, four per thread. Bases of a step that share a context must meet the row in order; contexts of
// real DNA spread over the 4096 rows, so conflicts inside 256 bases are rare: every thread bids for its context with a shared
// atomicMin of its index, the lowest bidder of each context reads the row, emits its (freq, cum, tot) triple (one coalesced
// global store for the step) and bumps the row; the few losers bid again in the next round. Degenerate stretches
// (homopolymers: hundreds of bases in one context) fall back to one thread replaying the step's leftovers in order.
#pragma once

#define DNA_DIRECT_ROUNDS 8
#define DNA_E 4                                     // consecutive bases per thread and step
#define DNA_STEP (DSRC_CTA * DNA_E)

struct DnaDirectShared {
    unsigned long long rows[4096];                  // 4 x u16 counters per context
    u32 owner[4096];
    u16 p_ctx[DNA_STEP];                            // leftovers of a step for the serial replay
    u8 p_sym[DNA_STEP], p_flag[DNA_STEP];
};

__device__ __forceinline__ void dna_row_step(unsigned long long& row, u32 s, u32& f, u32& cum, u32& tot)
{
    tot = (u32)(row & 0xFFFFu) + (u32)((row >> 16) & 0xFFFFu) + (u32)((row >> 32) & 0xFFFFu) + (u32)(row >> 48);
    if (tot >= (1u << 16) - 8) {                     // TSymbolCoderRC<4>::Rescale: c -= c >> 1, per counter
        u32 c0 = (u32)row & 0xFFFFu, c1 = (u32)(row >> 16) & 0xFFFFu, c2 = (u32)(row >> 32) & 0xFFFFu, c3 = (u32)(row >> 48);
        c0 -= c0 >> 1; c1 -= c1 >> 1; c2 -= c2 >> 1; c3 -= c3 >> 1;
        tot = c0 + c1 + c2 + c3;
        row = (unsigned long long)c0 | ((unsigned long long)c1 << 16) | ((unsigned long long)c2 << 32) | ((unsigned long long)c3 << 48);
    }
    const u32 sh = 16 * s;
    f = (u32)(row >> sh) & 0xFFFFu;
    const unsigned long long low = row & ((1ull << sh) - 1);          // the counters below s
    cum = (u32)(low & 0xFFFFu) + (u32)((low >> 16) & 0xFFFFu) + (u32)(low >> 32);
    row += 2ull << sh;                                // no carry: every counter stays below 2^16 - 8
}

__device__ __forceinline__ u32 dna_load4(const u8* sq, u32 i, u32 M)      // 4 bases at i..i+3 (i multiple of 4), 2 bits each, 0 outside [0, M)
{
    if (i >= M) return 0u;                            // also covers the wrapped negative positions before the block start
    const u32 w = *(const u32*)(sq + i);
    u32 p = (w & 3u) | ((w >> 6) & 0xCu) | ((w >> 12) & 0x30u) | ((w >> 18) & 0xC0u);
    if (i + 4 > M) p &= (1u << (2 * (M - i))) - 1;
    return p;
}

__device__ void dna_direct_engine(DnaDirectShared& D, const u8* sq, u32 M, u32 ord, u64* trip)
{
    const u32 tid = threadIdx.x;
    const u32 n_ctx = 1u << (2 * ord), mask = n_ctx - 1;
    for (u32 i = tid; i < n_ctx; i += DSRC_CTA) { D.rows[i] = 0x0001000100010001ull; D.owner[i] = 0u; }
    __syncthreads();
    // bids are (epoch << 11) | (1023 - position in step) under a shared atomicMax: a new round outbids every stale entry, inside
    // a round the earliest base of a context wins -- the owner table never needs clearing
    u32 epoch = 1;
    u32 hist_n = dna_load4(sq, 4 * tid - 8, M) | (dna_load4(sq, 4 * tid - 4, M) << 8) | (dna_load4(sq, 4 * tid, M) << 16);
    for (u32 base = 0; base < M; base += DNA_STEP) {
        const u32 i0 = base + 4 * tid;
        const u32 hist = hist_n;                      // bases i0-8 .. i0+3, oldest in the low bits
        {
            const u32 j = i0 + DNA_STEP;              // next step, loaded one step ahead
            hist_n = dna_load4(sq, j - 8, M) | (dna_load4(sq, j - 4, M) << 8) | (dna_load4(sq, j, M) << 16);
        }
        u32 ctx[DNA_E], sym[DNA_E], pend = 0;
#pragma unroll
        for (int k = 0; k < DNA_E; ++k) {
            sym[k] = (hist >> (16 + 2 * k)) & 3u;
            // the table is private to this kernel: any one-to-one numbering of the contexts will do, so the window of the
            // previous `ord` bases is used as it lies in `hist` (oldest base in the low bits)
            ctx[k] = (hist >> (16 + 2 * k - 2 * ord)) & mask;
            if (i0 + k < M) pend |= 1u << k;
        }
        for (int round = 0;; ++round, ++epoch) {
#pragma unroll
            for (int k = 0; k < DNA_E; ++k) if ((pend >> k) & 1u) atomicMax(&D.owner[ctx[k]], (epoch << 11) | (1023u - (4 * tid + k)));
            __syncthreads();
#pragma unroll
            for (int k = 0; k < DNA_E; ++k) {
                if (((pend >> k) & 1u) && D.owner[ctx[k]] == ((epoch << 11) | (1023u - (4 * tid + k)))) {
                    unsigned long long row = D.rows[ctx[k]];
                    u32 f, cum, tot;
                    dna_row_step(row, sym[k], f, cum, tot);
                    D.rows[ctx[k]] = row;
                    trip[i0 + k] = TRIP(f, cum, tot);
                    pend &= ~(1u << k);
                }
            }
            if (!__syncthreads_or(pend != 0)) { ++epoch; break; }
            if (round + 1 >= DNA_DIRECT_ROUNDS) {
#pragma unroll
                for (int k = 0; k < DNA_E; ++k) { D.p_flag[4 * tid + k] = (pend >> k) & 1u; D.p_ctx[4 * tid + k] = (u16)ctx[k]; D.p_sym[4 * tid + k] = (u8)sym[k]; }
                __syncthreads();
                if (tid == 0) {
                    for (u32 t = 0; t < DNA_STEP; ++t) if (D.p_flag[t]) {
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
