// Direct engine of the 16-symbol quality model (QualityOrderModeler schemes 0/4: <16,3,*> for -q1, <16,4,*> for -q2).
//
// The 2 MiB .. 32 MiB table of adaptive rows stays in HBM (its hot rows live in L2); the block is taken 1024 symbols at a time,
// four consecutive symbols per thread. Consecutive symbols that share a context form a run (inside a stretch of equal qualities the
// context repeats). Every run head pushes itself onto the chain of its context's hash slot (shared atomicExch, no ordering
// needed); the earliest run of a context then owns the context for this step: its thread loads the row once, walks ALL runs of
// that context in position order -- freq / cum / tot of TSymbolCoderRC<16>::EncodeSymbol (src/SymbolCoderRC.h:35-48) with the
// halving rescale (:69-73) -- and stores the row once. One pass per step, three barriers, no retries. A row whose first counter is
// 0 has not been touched by this block; touched contexts are listed and re-zeroed when the block is done (as in model_tab.cuh).
#pragma once

#define QD_E 4
#define QD_STEP (DSRC_CTA * QD_E)
#define QD_SLOTS 4096

struct QDirectShared {
    u32 head[QD_SLOTS];                             // (epoch << 10) | step position of the newest run pushed onto the slot's chain
    u32 s_ctx[QD_STEP + 1];                         // context of every symbol of the step
    u16 next[QD_STEP];                              // chain link (step position) or 0xFFFF
    u8 s_sym[QD_STEP], s_flag[QD_STEP];             // symbol rank; 1 = the symbol starts a run
    u8 plut[1024];
    u16 win[QD_STEP];                               // owners of the step, compacted so that every lane walks a chain
    u32 n_win, n_touched;
};

__device__ __forceinline__ u32 qd_slot(u32 ctx) { return (ctx * 0x9E3779B1u) >> 20; }

// q: processed quality bytes of the block; pc: per-symbol position buckets (variable read lengths) or null with fixed_len != 0
__device__ void quality_direct_engine(QDirectShared& D, const u8* rank, const u8* q, const u8* pc, u32 fixed_len, u32 rescale, u32 M,
                                      u32 so, u32 bits, u8* tab, u32* touched, u64* trip, u64* prof)
{
    const u32 tid = threadIdx.x;
    const u32 h = so / 2;
    for (u32 i = tid; i < QD_SLOTS; i += DSRC_CTA) D.head[i] = 0u;
    if (fixed_len) for (u32 j = tid; j < fixed_len; j += DSRC_CTA) D.plut[j] = (u8)(j * rescale / fixed_len);
    if (tid == 0) D.n_touched = 0;
    __syncthreads();
    u32 epoch = 1;                                  // stale chain heads of earlier steps are recognised by their epoch
    u32 jpos = fixed_len ? (4 * tid) % fixed_len : 0u;               // read position of this thread's first symbol
    const u32 jstep = fixed_len ? QD_STEP % fixed_len : 0u;
    for (u32 base = 0; base < M; base += QD_STEP, ++epoch) {
        const u32 i0 = base + 4 * tid;
        const u32 n_step = min((u32)QD_STEP, M - base);
        u64* trip_step = trip + base;
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
        // run heads push themselves onto the chain of their context's slot
        u32 heads = 0;
#pragma unroll
        for (int k = 0; k < QD_E; ++k) {
            const u32 p = 4 * tid + k;
            const bool head = p < n_step && (p == 0 || D.s_ctx[p] != D.s_ctx[p - 1]);
            D.s_flag[p] = head ? 1u : 0u;
            if (head) {
                heads |= 1u << k;
                const u32 old = atomicExch(&D.head[qd_slot(D.s_ctx[p])], (epoch << 10) | p);
                D.next[p] = (old >> 10) == epoch ? (u16)(old & 1023u) : (u16)0xFFFFu;
            }
        }
        __syncthreads();
        // the earliest run of a context owns it for this step; owners are compacted so that every lane walks a chain
        if (tid == 0) D.n_win = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < QD_E; ++k) {
            if (!((heads >> k) & 1u)) continue;
            const u32 p0 = 4 * tid + k, ctx = D.s_ctx[p0];
            bool owner = true;
            for (u32 e = D.head[qd_slot(ctx)] & 1023u; e != 0xFFFFu; e = D.next[e]) if (e < p0 && D.s_ctx[e] == ctx) { owner = false; break; }
            if (owner) D.win[atomicAdd(&D.n_win, 1u)] = (u16)p0;
        }
        __syncthreads();
        const u32 n_win = D.n_win;
        // walk every run of the context, in position order, with the row in registers
        for (u32 w = tid; w < n_win; w += DSRC_CTA) {
            const u32 p0 = D.win[w], ctx = D.s_ctx[p0];
            const u32 first = D.head[qd_slot(ctx)] & 1023u;
            u8* rowp = tab + (u64)ctx * 32;
            RowRegs<16> R; R.load(rowp);
            if ((R.c[0] & 0xFFFFu) == 0) { R.ones(); touched[atomicAdd(&D.n_touched, 1u)] = ctx; }
            u32 tot = R.total();
            u32 last = 0xFFFFFFFFu, f = 0, cum = 0;
            for (u32 cur = p0;;) {
                u32 p = cur;
                do {
                    const u32 s = D.s_sym[p];
                    if (tot >= (1u << 16) - 32) { tot = R.rescale(); last = 0xFFFFFFFFu; }
                    if (s != last) { R.get(s, f, cum); last = s; }          // inside a run of equal symbols only freq and tot move
                    trip_step[p] = TRIP(f, cum, tot);
                    R.bump(s); f += 2; tot += 2;
                    ++p;
                } while (p < n_step && !D.s_flag[p]);
                u32 nxt = 0xFFFFu;                       // the next run of this context
                for (u32 e = first; e != 0xFFFFu; e = D.next[e]) if (e > cur && e < nxt && D.s_ctx[e] == ctx) nxt = e;
                if (nxt == 0xFFFFu) break;
                cur = nxt;
            }
            R.store(rowp);
        }
        __syncthreads();
    }
    const u32 nt = D.n_touched;
    for (u32 k = tid; k < nt; k += DSRC_CTA) { uint4* rp = (uint4*)(tab + (u64)touched[k] * 32); rp[0] = make_uint4(0u, 0u, 0u, 0u); rp[1] = make_uint4(0u, 0u, 0u, 0u); }
    __syncthreads();
}
