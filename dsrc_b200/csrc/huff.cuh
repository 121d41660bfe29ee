// Warp-cooperative Huffman construction + tree serialisation, bit-exact with HuffmanEncoder
// (src/huffman.cpp:94-221, src/huffman.h:67-70,109-122).
//
// The reference extracts minima from a binary heap ordered by the strict total order
// (frequency asc, symbol asc), so the merge sequence is independent of the heap layout. Here the
// 32 lanes of a warp scan the active list in stride and an argmin shuffle-reduction picks each minimum.
#pragma once
#include "common.cuh"

#define HUF_NMAX 512

struct HufWork {                 // per-warp scratch (global memory)
    u32 wfreq[HUF_NMAX + 2];
    u16 wsym[HUF_NMAX + 2];
    u16 left[HUF_NMAX], right[HUF_NMAX];   // children of internal node n+i
    u32 icode[HUF_NMAX];
    u8 ilen[HUF_NMAX];
    u16 stack[2 * HUF_NMAX];
};

#ifdef __CUDACC__
__device__ __forceinline__ u64 huf_warp_argmin(const HufWork* w, u32 cnt)
{
    u64 best = ~0ull;
    for (u32 k = lane_id(); k < cnt; k += 32) {
        u64 key = ((u64)w->wfreq[k] << 20) | ((u64)w->wsym[k] << 10) | k;
        best = key < best ? key : best;
    }
    for (int o = 16; o; o >>= 1) { u64 t = __shfl_xor_sync(0xFFFFFFFFu, best, o); best = t < best ? t : best; }
    return best;
}

// Builds codes for `n_in` symbols with frequencies freq[]. All lanes of the calling warp must participate.
// code[]/len[] receive the n leaf codes; ser receives the StoreTree serialisation; returns its size in bytes
// (0xFFFFFFFF on overflow of ser_cap).
static __device__ u32 huf_build_warp(const u32* freq, u32 n_in, HufWork* w, u32* code, u8* len, u8* ser, u32 ser_cap)
{
    const u32 ln = lane_id();
    u32 n = n_in < 2 ? 2 : n_in;                      // huffman.cpp:101
    for (u32 k = ln; k < n; k += 32) {
        w->wsym[k] = (u16)(k < n_in ? k : 0);         // an absent 2nd slot is a default Frequency{0,0}
        w->wfreq[k] = k < n_in ? freq[k] : 0;
        code[k] = 0; len[k] = 0;
    }
    __syncwarp();
    u32 cnt = n;
    {
        u64 m = huf_warp_argmin(w, cnt);
        if (cnt == 2 && (m >> 20) == 0) {             // :128-133: the reference forces both frequencies to >= 1 in place, without
            // re-heapifying: the former minimum stays heap[0] and becomes the LEFT child even if the other symbol is smaller and
            // also has frequency 1. Leaving the minimum's key at 0 keeps that order (the merged frequency of the only merge is unused).
            if (ln == 0) { u32 i = (u32)(m & 1023); if (w->wfreq[1 - i] == 0) w->wfreq[1 - i] = 1; }
            __syncwarp();
        } else {
            // :134-138 -- drop zero-frequency symbols (ascending symbol order) while more than two remain.
            // Equivalent bulk form: with z zero symbols, remove min(z, cnt-2) of them, smallest symbols first.
            u32 z = 0;
            for (u32 k = ln; k < cnt; k += 32) z += (w->wfreq[k] == 0);
            z = warp_red_sum(z);
            u32 drop = min(z, cnt - 2);
            if (drop) {
                // stable compaction by one lane keeps the code simple; n <= 512 and this runs once per tree
                if (ln == 0) {
                    u32 o = 0, dropped = 0;
                    for (u32 k = 0; k < cnt; ++k) {
                        if (w->wfreq[k] == 0 && dropped < drop) { ++dropped; continue; }   // wsym ascending == k ascending here
                        w->wfreq[o] = w->wfreq[k]; w->wsym[o] = w->wsym[k]; ++o;
                    }
                }
                cnt -= drop;
                __syncwarp();
            }
        }
    }
    const u32 present = cnt;
    for (u32 i = 0; i + 1 < present; ++i) {           // :146-158
        u64 m = huf_warp_argmin(w, cnt);
        u32 li = (u32)(m & 1023), ls = (u32)((m >> 10) & 1023); u32 lf = (u32)(m >> 20);
        __syncwarp();
        if (ln == 0) { w->wfreq[li] = w->wfreq[cnt - 1]; w->wsym[li] = w->wsym[cnt - 1]; }
        --cnt;
        __syncwarp();
        m = huf_warp_argmin(w, cnt);
        u32 ri = (u32)(m & 1023), rs = (u32)((m >> 10) & 1023); u32 rf = (u32)(m >> 20);
        __syncwarp();
        if (ln == 0) {
            w->wfreq[ri] = w->wfreq[cnt - 1]; w->wsym[ri] = w->wsym[cnt - 1];
            w->wsym[cnt - 1] = (u16)(n + i); w->wfreq[cnt - 1] = lf + rf;
            w->left[i] = (u16)ls; w->right[i] = (u16)rs;
        }
        __syncwarp();
    }
    // codes, root downwards (:161-168); serial, short
    u32 ser_size = 0;
    if (ln == 0) {
        const u32 root = n + present - 2;
        for (u32 i = 0; i + 1 < present; ++i) { w->icode[i] = 0; w->ilen[i] = 0; }
        for (i32 k = (i32)present - 2; k >= 0; --k) {
            u32 c = w->icode[k]; u32 l = w->ilen[k];
            u32 lc = w->left[k], rc = w->right[k];
            if (lc >= n) { w->icode[lc - n] = c << 1; w->ilen[lc - n] = (u8)(l + 1); } else { code[lc] = c << 1; len[lc] = (u8)(l + 1); }
            if (rc >= n) { w->icode[rc - n] = (c << 1) | 1; w->ilen[rc - n] = (u8)(l + 1); } else { code[rc] = (c << 1) | 1; len[rc] = (u8)(l + 1); }
        }
        // StoreTree (:177-221)
        BitW bw; bw.init(ser, ser_cap);
        u32 bpi = dsrc_ilog2(n) + ((n & (n - 1)) ? 1 : 0);
        u32 min_len = n;
        for (u32 k = 0; k < n; ++k) if (len[k] < min_len && len[k] > 0) min_len = len[k];
        bw.be32(0); bw.be32(root); bw.be32(n); bw.byte((u8)min_len);
        u32 sp = 0; w->stack[sp++] = (u16)root;
        while (sp) {                                  // pre-order: node, left subtree, right subtree
            u32 id = w->stack[--sp];
            if (id < n) { bw.bit(1); bw.bits(id, bpi); }
            else { bw.bit(0); w->stack[sp++] = w->right[id - n]; w->stack[sp++] = w->left[id - n]; }
        }
        bw.flush();
        if (bw.ovf) ser_size = 0xFFFFFFFFu;
        else { ser_size = bw.pos; ser[0] = (u8)(ser_size >> 24); ser[1] = (u8)(ser_size >> 16); ser[2] = (u8)(ser_size >> 8); ser[3] = (u8)ser_size; }
    }
    ser_size = __shfl_sync(0xFFFFFFFFu, ser_size, 0);
    __syncwarp();
    return ser_size;
}
#endif
