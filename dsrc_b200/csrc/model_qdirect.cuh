// Direct engine of the 16-symbol quality model (QualityOrderModeler schemes 0/4: <16,3,*> for -q1, <16,4,*> for -q2).
//
// Same idea as model_dna.cuh, with the 2 MiB .. 32 MiB table in HBM (its hot rows live in L2): the block is taken 1024 symbols at
// a time, four consecutive symbols per thread. Consecutive symbols of a thread that share a context form one segment (inside a
// run of equal qualities the context repeats); every pending segment bids for its context with a shared atomicMax of its
// position, the earliest segment of each context wins the round, loads the row, walks its symbols (freq/cum/tot of
// TSymbolCoderRC<16>::EncodeSymbol, src/SymbolCoderRC.h:35-48, with the halving rescale :69-73), stores the row, and the losers bid
// again. Bids of different contexts that hash to the same slot just cost the later one a round. A row whose first counter is 0 has
// not been touched by this block; touched contexts are listed and re-zeroed when the block is done (as in model_tab.cuh).
#pragma once

#define QD_E 4
#define QD_STEP (DSRC_CTA * QD_E)
#define QD_ROUNDS 48                               // then one thread replays the step's leftovers in order
#define QD_SLOTS 4096

struct QDirectShared {
    u32 owner[QD_SLOTS];
    u32 s_ctx[QD_STEP + 1];                         // context of every symbol of the step (+ sentinel)
    u8 s_sym[QD_STEP], s_flag[QD_STEP];             // symbol rank; bit 0: the symbol starts a run of one context, bit 1: that run is pending
    u8 plut[1024];
    u16 win[QD_STEP];                               // run heads that won the current round, compacted so that every lane walks a run
    u32 n_win, n_touched;
};

__device__ __forceinline__ u32 qd_slot(u32 ctx) { return (ctx * 0x9E3779B1u) >> 20; }

// walks one run (consecutive symbols of one context, starting at step position p) against its row
__device__ __forceinline__ void qd_run(QDirectShared& D, u32 p, u32 n_step, u8* tab, u32* touched, u64* trip_step)
{
    const u32 ctx = D.s_ctx[p];
    u8* rowp = tab + (u64)ctx * 32;
    RowRegs<16> R; R.load(rowp);
    if ((R.c[0] & 0xFFFFu) == 0) { R.ones(); touched[atomicAdd(&D.n_touched, 1u)] = ctx; }
    u32 tot = R.total();
    u32 last = 0xFFFFFFFFu, f = 0, cum = 0;
    do {
        const u32 s = D.s_sym[p];
        if (tot >= (1u << 16) - 32) { tot = R.rescale(); last = 0xFFFFFFFFu; }
        if (s != last) { R.get(s, f, cum); last = s; }          // inside a run of equal symbols only freq and tot move
        trip_step[p] = TRIP(f, cum, tot);
        R.bump(s); f += 2; tot += 2;
        ++p;
    } while (p < n_step && !(D.s_flag[p] & 1u));
    R.store(rowp);
}

// q: processed quality bytes of the block; pc: per-symbol position buckets (variable read lengths) or null with fixed_len != 0
__device__ void quality_direct_engine(QDirectShared& D, const u8* rank, const u8* q, const u8* pc, u32 fixed_len, u32 rescale, u32 M,
                                      u32 so, u32 bits, u8* tab, u32* touched, u64* trip, u64* prof)
{
    const u32 tid = threadIdx.x;
    const u32 h = so / 2;
    for (u32 i = tid; i < QD_SLOTS; i += DSRC_CTA) D.owner[i] = 0u;
    if (fixed_len) for (u32 j = tid; j < fixed_len; j += DSRC_CTA) D.plut[j] = (u8)(j * rescale / fixed_len);
    if (tid == 0) D.n_touched = 0;
    __syncthreads();
    u32 epoch = 1;
    u32 jpos = fixed_len ? (4 * tid) % fixed_len : 0u;               // read position of this thread's first symbol
    const u32 jstep = fixed_len ? QD_STEP % fixed_len : 0u;
    for (u32 base = 0; base < M; base += QD_STEP) {
        const u32 i0 = base + 4 * tid;
        const u32 n_step = min((u32)QD_STEP, M - base);
        // ranks of the symbols i0-8 .. i0+3 (0 before the block start)
        u32 r[12];
        {
            const u32 w0 = (i0 >= 8 && i0 - 8 < M) ? *(const u32*)(q + i0 - 8) : 0u;
            const u32 w1 = (i0 >= 4 && i0 - 4 < M) ? *(const u32*)(q + i0 - 4) : 0u;
            const u32 w2 = i0 < M ? *(const u32*)(q + i0) : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                r[k] = (i0 >= 8 && i0 - 8 + k < M) ? rank[(w0 >> (8 * k)) & 255u] : 0u;
                r[4 + k] = (i0 >= 4 && i0 - 4 + k < M) ? rank[(w1 >> (8 * k)) & 255u] : 0u;
                r[8 + k] = (i0 + k < M) ? rank[(w2 >> (8 * k)) & 255u] : 0u;
            }
        }
        u32 pcw = 0;
        if (!fixed_len && i0 < M) pcw = *(const u32*)(pc + i0);
#pragma unroll
        for (int k = 0; k < QD_E; ++k) {
            // TQualityModelBase::UpdateHash / GetHash (QualityEncoder.h:77-94): raw symbols below slot h, pairwise means from slot h on
            u32 hash = 0;
            if (so == 1) hash = r[8 + k - 1];
            else {
#pragma unroll
                for (int t = 0; t < 4; ++t) if ((u32)t < so) {
                    const u32 v = (u32)t < h ? r[8 + k - 1 - t] : ((r[8 + k - 1 - t] + r[8 + k - 2 - t]) >> 1);
                    hash |= v << (t * bits);
                }
            }
            u32 pctx;
            if (fixed_len) { u32 j = jpos + k; while (j >= fixed_len) j -= fixed_len; pctx = D.plut[j]; }
            else pctx = (pcw >> (8 * k)) & 255u;
            D.s_ctx[4 * tid + k] = (hash << bits) | pctx;
            D.s_sym[4 * tid + k] = (u8)r[8 + k];
        }
        if (fixed_len) { jpos += jstep; if (jpos >= fixed_len) jpos -= fixed_len; }
        __syncthreads();
        // run heads: a symbol whose context differs from its predecessor's (the step's first symbol always starts a run)
        u32 pend = 0;                                   // bit k: symbol 4*tid+k heads a run that still has to meet its row
#pragma unroll
        for (int k = 0; k < QD_E; ++k) {
            const u32 p = 4 * tid + k;
            const bool head = p < n_step && (p == 0 || D.s_ctx[p] != D.s_ctx[p - 1]);
            D.s_flag[p] = head ? 1u : 0u;
            if (head) pend |= 1u << k;
        }
        __syncthreads();
        for (int round = 0;; ++round, ++epoch) {
            if (tid == 0) D.n_win = 0;
#pragma unroll
            for (int k = 0; k < QD_E; ++k)
                if ((pend >> k) & 1u) atomicMax(&D.owner[qd_slot(D.s_ctx[4 * tid + k])], (epoch << 10) | (1023u - (4 * tid + k)));
            __syncthreads();
            u32 wins = 0;
#pragma unroll
            for (int k = 0; k < QD_E; ++k)
                if (((pend >> k) & 1u) && D.owner[qd_slot(D.s_ctx[4 * tid + k])] == ((epoch << 10) | (1023u - (4 * tid + k)))) wins |= 1u << k;
            if (wins) {
                u32 at = atomicAdd(&D.n_win, (u32)__popc(wins));
#pragma unroll
                for (int k = 0; k < QD_E; ++k) if ((wins >> k) & 1u) D.win[at++] = (u16)(4 * tid + k);
                pend &= ~wins;
            }
            __syncthreads();
            const u32 n_win = D.n_win;
            for (u32 w = tid; w < n_win; w += DSRC_CTA) qd_run(D, D.win[w], n_step, tab, touched, trip + base);
            if (!__syncthreads_or(pend != 0)) { ++epoch; if (prof && tid == 0) { atomicAdd((unsigned long long*)&prof[40], (unsigned long long)(round + 1)); atomicAdd((unsigned long long*)&prof[41], 1ull); } break; }
            if (round + 1 >= QD_ROUNDS) {
                if (prof && tid == 0) atomicAdd((unsigned long long*)&prof[42], 1ull);
#pragma unroll
                for (int k = 0; k < QD_E; ++k) if ((pend >> k) & 1u) D.s_flag[4 * tid + k] |= 2u;
                __syncthreads();
                if (tid == 0) for (u32 p = 0; p < n_step; ++p) if (D.s_flag[p] & 2u) qd_run(D, p, n_step, tab, touched, trip + base);
                __syncthreads();
                ++epoch;
                break;
            }
        }
    }
    __syncthreads();
    const u32 nt = D.n_touched;
    for (u32 k = tid; k < nt; k += DSRC_CTA) { uint4* rp = (uint4*)(tab + (u64)touched[k] * 32); rp[0] = make_uint4(0u, 0u, 0u, 0u); rp[1] = make_uint4(0u, 0u, 0u, 0u); }
    __syncthreads();
}
