// -c: per-block CRC-32 of the raw titles / sequences / qualities (FastqChecksumHasher, src/RecordsProcessor.h:28-68; core::Crc32Hasher,
// src/Crc32.h:24-92: reflected polynomial 0xEDB88320, seed ~0, final xor ~0), stored in the block header by StoreMetaData
// (src/BlockCompressor.cpp:424-440) and checked against the decoded records by VerifyChecksum (:576-594).
//
// The reference runs one CRC over the concatenation of a field over all records. CRCs of adjacent pieces combine as
// crc(A||B) = x^(8|B|) * crc(A) + crc(B) over GF(2)[x] mod P, so every thread hashes a contiguous run of records and the CTA
// folds the 256 partial (crc, length) pairs with an ordered tree reduction.
#include "common.cuh"
#include "kernels.h"

#define CRC_POLY 0xEDB88320u

// a(x) * b(x) mod P, reflected bit order (bit 31 = x^0)
__device__ __forceinline__ u32 crc_mulmod(u32 a, u32 b)
{
    u32 m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}
// x^(8n) mod P from the table of x^(2^k)
__device__ __forceinline__ u32 crc_x8n(const u32* x2n, u32 n)
{
    u32 p = 1u << 31, k = 3;
    while (n) { if (n & 1u) p = crc_mulmod(x2n[k & 31], p); n >>= 1; ++k; }
    return p;
}
__device__ __forceinline__ u32 crc_combine(const u32* x2n, u32 crc1, u32 crc2, u32 len2) { return crc_mulmod(crc_x8n(x2n, len2), crc1) ^ crc2; }

// decode_mode = 0: fields of the raw input block (record SoA from k_parse) -> st.crc
// decode_mode = 1: fields of the decoded FASTQ (layout of k_dec_assemble) -> compared with st.crc_expected
__global__ void __launch_bounds__(DSRC_CTA) k_crc(Workspace ws, u32 decode_mode)
{
    const u32 blk = blockIdx.x;
    const BlockDesc& d = ws.desc[blk];
    BlockState& st = ws.state[blk];
    if (st.status != ST_OK) return;
    __shared__ u32 s_tab[256], s_x2n[32];
    __shared__ u32 s_crc[3][DSRC_CTA], s_len[3][DSRC_CTA];
    const u32 tid = threadIdx.x;
    {
        u32 h = tid;
        for (int j = 0; j < 8; ++j) h = (h & 1u) ? CRC_POLY ^ (h >> 1) : h >> 1;
        s_tab[tid] = h;
    }
    if (tid == 0) {
        u32 p = 1u << 30;                                   // x^1
        s_x2n[0] = p;
        for (int k = 1; k < 32; ++k) { p = crc_mulmod(p, p); s_x2n[k] = p; }
    }
    __syncthreads();
    const RecArrays& R = ws.rec;
    const u32 n = st.n_rec, rb = d.rec_base;
    const u8* base = decode_mode ? ws.out + d.out_off : ws.in + d.in_off;
    const u32 per = (n + DSRC_CTA - 1) / DSRC_CTA;
    const u32 r0 = min(n, tid * per), r1 = min(n, r0 + per);
    u32 c[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, len[3] = {0, 0, 0};
    for (u32 r = r0; r < r1; ++r) {
        const u32 g = rb + r;
        const u32 tl = R.title_len[g], ql = R.qua_len[g];
        const u8 *t, *s, *q;
        if (decode_mode) { t = base + R.seq_off[g]; s = t + tl + 1; q = s + ql + 1 + (ws.plus_rep ? tl : 1) + 1; }
        else { t = base + R.title_off[g]; s = base + R.seq_off[g]; q = base + R.qua_off[g]; }
        for (u32 i = 0; i < tl; ++i) c[0] = (c[0] >> 8) ^ s_tab[(t[i] ^ c[0]) & 0xFF];
        for (u32 i = 0; i < ql; ++i) c[1] = (c[1] >> 8) ^ s_tab[(s[i] ^ c[1]) & 0xFF];
        for (u32 i = 0; i < ql; ++i) c[2] = (c[2] >> 8) ^ s_tab[(q[i] ^ c[2]) & 0xFF];
        len[0] += tl; len[1] += ql; len[2] += ql;
    }
    for (int k = 0; k < 3; ++k) { s_crc[k][tid] = ~c[k]; s_len[k][tid] = len[k]; }      // an empty piece has crc 0 (~~0)
    __syncthreads();
    for (u32 off = 1; off < DSRC_CTA; off <<= 1) {
        if ((tid & (2 * off - 1)) == 0) {
            for (int k = 0; k < 3; ++k) {
                const u32 l2 = s_len[k][tid + off];
                if (l2) { s_crc[k][tid] = s_len[k][tid] ? crc_combine(s_x2n, s_crc[k][tid], s_crc[k][tid + off], l2) : s_crc[k][tid + off]; s_len[k][tid] += l2; }
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (!decode_mode) { for (int k = 0; k < 3; ++k) st.crc[k] = s_crc[k][0]; }
        else if (s_crc[0][0] != st.crc_expected[0] || s_crc[1][0] != st.crc_expected[1] || s_crc[2][0] != st.crc_expected[2]) {
            st.status = ST_CRC; ws.result[blk].status = ST_CRC; ws.result[blk].total_size = 0;
        }
    }
}

void launch_crc(const Workspace& ws, cudaStream_t s, u32 decode_mode) { k_crc<<<ws.n_blocks, DSRC_CTA, 0, s>>>(ws, decode_mode); }
