// -q0 and -d0 paths: Huffman / 2-bit modelers (BASELINE config[0], "fast mode").
//   quality: QualityNormalModelerProxy (src/QualityModelerProxy.h:113-122) selecting
//            QualityPositionModelerPlain / Truncated (src/QualityPositionModeler.cpp:57-72,107-187,240-287) or
//            QualityRLEModeler (src/QualityRLEModeler.cpp:121-373)
//   DNA    : DnaNormalModelerProxy (src/DnaModelerProxy.h:102-122) selecting DnaModelerBasicB2
//            (src/DnaModelerBasicB2.h:34-46) or DnaModelerHuffman (src/DnaModelerHuffman.cpp:21-73)
// All of these are "histogram -> Huffman trees -> variable-length codes": the reference emits codes one PutBits at a
// time; here code lengths are prefix-summed (CTA scans with a 64-bit carry) and every thread ORs its codes into the
// zero-initialised big-endian stream at its own bit offset. Trees are built one per warp (huff.cuh).
#include "common.cuh"
#include "huff.cuh"
#include "kernels.h"

#define Q0_ENTRIES 262144u       // max (positions x symbols) histogram cells per block
#define Q0_MAXTREES 8192u
#define Q0_SER_BYTES (640u << 10)

struct Q0Arena {                 // per persistent CTA; followed by a u64[max_syms + 2] scratch array
    HufWork work[DSRC_WARPS];
    u32 hist[Q0_ENTRIES];
    u32 code[Q0_ENTRIES];
    u8 len[Q0_ENTRIES];
    u32 ser_size[Q0_MAXTREES];
    u32 ser_off[Q0_MAXTREES];
    u8 ser[Q0_SER_BYTES];
};
u64 q0_arena_bytes(u64 max_block_bytes) { return ((sizeof(Q0Arena) + 255) & ~(u64)255) + (max_block_bytes / 2 + 64) * 8; }

__device__ __forceinline__ u32 ser_stride_for(u32 n_sym)
{
    const u32 n = n_sym < 2 ? 2 : n_sym;
    const u32 bpi = dsrc_ilog2(n) + ((n & (n - 1)) ? 1 : 0);
    return (13 + ((2 * n - 1) + n * bpi + 7) / 8 + 8 + 3) & ~3u;
}

struct Q0Shared {
    u32 scan[DSRC_WARPS + 1];
    u32 qf[256], lf[256];
    u8 qrank[256], lrank[256];
    u32 status, scheme, nq, nl, runs, hdr_bytes, carry_u32, carry_max;
    unsigned long long carry, total_bits;
};

// copies the serialised trees [0, n_trees) behind each other at out + *pos (parallel over trees), returns via shared hdr_bytes
__device__ void copy_trees(Q0Shared& S, Q0Arena* A, u32 n_trees, u32 ser_stride, u8* out, u32 out_cap)
{
    // exclusive scan of the tree sizes
    if (threadIdx.x == 0) S.carry_u32 = S.hdr_bytes;
    __syncthreads();
    for (u32 base = 0; base < n_trees; base += DSRC_CTA) {
        const u32 j = base + threadIdx.x;
        u32 sz = j < n_trees ? A->ser_size[j] : 0;
        if (sz == 0xFFFFFFFFu) { S.status = ST_OVERFLOW; sz = 0; }
        u32 total, ex = block_excl_sum(sz, S.scan, &total);
        const u32 carry = S.carry_u32;
        if (j < n_trees) A->ser_off[j] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) S.carry_u32 = carry + total;
        __syncthreads();
    }
    const u32 end = S.carry_u32;
    if (end + 16 > out_cap) { if (threadIdx.x == 0) S.status = ST_OVERFLOW; __syncthreads(); return; }
    for (u32 j = warp_id(); j < n_trees; j += DSRC_WARPS) {
        const u8* src = A->ser + (u64)j * ser_stride; u8* dst = out + A->ser_off[j];
        for (u32 i = lane_id(); i < A->ser_size[j]; i += 32) dst[i] = src[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) S.hdr_bytes = end;
    __syncthreads();
}

__global__ void __launch_bounds__(DSRC_CTA) k_q0_quality(Workspace ws, u8* arena, u64 arena_stride)
{
    __shared__ Q0Shared S;
    Q0Arena* A = (Q0Arena*)(arena + (u64)blockIdx.x * arena_stride);
    unsigned long long* big = (unsigned long long*)((u8*)A + ((sizeof(Q0Arena) + 255) & ~(u64)255));
    const u32 tid = threadIdx.x, w = warp_id(), ln = lane_id();

    for (u32 blk = blockIdx.x; blk < ws.n_blocks; blk += gridDim.x) {
        const BlockDesc& d = ws.desc[blk];
        BlockState& st = ws.state[blk];
        __syncthreads();
        if (st.status != ST_OK) continue;
        const RecArrays& R = ws.rec;
        const u32 n_rec = st.n_rec, rb = d.rec_base, M = st.q_total;
        const u8* q = ws.qcat + d.sym_base;
        u8* out = ws.streams + d.stream_base + stream_offset(d, 3);
        const u32 out_cap = d.stream_cap[3];
        if (tid == 0) {
            // QualityNormalModelerProxy::SelectSchemeId (QualityModelerProxy.h:113-122)
            u32 sc = 0;
            if (__fdiv_rn((float)st.th_len, (float)st.rle_len) > 1.25f) sc = 2;
            else if (__fdiv_rn((float)st.raw_len, (float)st.th_len) > 1.10f) sc = 1;
            S.scheme = sc; st.q_scheme = (u8)sc; S.status = ST_OK;
            out[0] = (u8)sc; S.hdr_bytes = 1;
        }
        S.qrank[tid] = st.qrank[tid];
        __syncthreads();
        const u32 scheme = S.scheme;

        if (scheme < 2) {
            // ============ positional Huffman (Plain / Truncated) ============
            const bool trunc = scheme == 1;
            const u32 L = st.max_len, Sy = st.q_count;
            if (Sy < 2 || L > Q0_MAXTREES || (u64)L * Sy > Q0_ENTRIES || (u64)L * ser_stride_for(Sy) > Q0_SER_BYTES) {
                if (tid == 0) st.status = ST_UNSUPPORTED;     // Sy == 1 is undefined behaviour upstream
                continue;
            }
            const u32 sstr = ser_stride_for(Sy);
            for (u32 i = tid; i < L * Sy; i += DSRC_CTA) A->hist[i] = 0;
            __syncthreads();
            for (u32 r = w; r < n_rec; r += DSRC_WARPS) {      // CalculatePositionStats (:143-157 / :226-238)
                const u32 lim = trunc ? R.trunc_len[rb + r] : R.qua_len[rb + r];
                const u8* qr = q + R.qcat_off[rb + r];
                for (u32 j = ln; j < lim; j += 32) atomicAdd(&A->hist[j * Sy + S.qrank[qr[j]]], 1u);
            }
            __syncthreads();
            for (u32 j = w; j < L; j += DSRC_WARPS) {          // ComputeHuffmanContext (:107-138)
                u32 sz = huf_build_warp(A->hist + j * Sy, Sy, &A->work[w], A->code + j * Sy, A->len + j * Sy, A->ser + (u64)j * sstr, sstr);
                if (ln == 0) A->ser_size[j] = sz;
            }
            __syncthreads();
            if (tid == 0) {                                    // Encode (:57-72): flush, maxLength, symbol mask
                BitW hw; hw.init(out, out_cap); hw.pos = 1;
                hw.be32(L);
                for (u32 i = 0; i < 256; ++i) hw.bit(S.qrank[i] != 255);
                hw.flush();
                if (hw.ovf) S.status = ST_OVERFLOW;
                S.hdr_bytes = hw.pos;
            }
            __syncthreads();
            copy_trees(S, A, L, sstr, out, out_cap);
            if (S.status != ST_OK) { if (tid == 0) st.status = S.status; continue; }
            // per-record bit counts -> big[r]
            const bool variable = st.min_len != st.max_len;
            const u32 max_bits = dsrc_bit_length((u64)L);
            for (u32 r = w; r < n_rec; r += DSRC_WARPS) {
                const u32 len = R.qua_len[rb + r], tl = R.trunc_len[rb + r];
                const u32 lim = trunc ? tl : len;
                const u8* qr = q + R.qcat_off[rb + r];
                u32 nb = 0;
                for (u32 j = ln; j < lim; j += 32) nb += A->len[j * Sy + S.qrank[qr[j]]];
                nb = warp_red_sum(nb);
                if (trunc) nb += 1 + (len != tl ? (variable ? dsrc_bit_length((u64)len) : max_bits) : 0);
                if (ln == 0) big[r] = nb;
            }
            __syncthreads();
            if (tid == 0) S.carry = trunc ? 1 : 0;               // Truncated: leading variableLength bit (:256)
            __syncthreads();
            for (u32 base = 0; base < n_rec; base += DSRC_CTA) {
                const u32 r = base + tid;
                u32 nb = r < n_rec ? (u32)big[r] : 0;
                u32 total, ex = block_excl_sum(nb, S.scan, &total);
                const unsigned long long carry = S.carry;
                if (r < n_rec) big[r] = carry + ex;
                __syncthreads();
                if (tid == 0) S.carry = carry + total;
                __syncthreads();
            }
            const u64 total_bits = S.carry, nbytes = (total_bits + 7) / 8;
            if ((u64)S.hdr_bytes + nbytes + 8 > out_cap) { if (tid == 0) st.status = ST_OVERFLOW; continue; }
            for (u64 i = tid; i < nbytes + 8; i += DSRC_CTA) out[S.hdr_bytes + i] = 0;
            __syncthreads();
            u32* words = (u32*)out;
            const u64 bit0 = (u64)S.hdr_bytes * 8;
            if (trunc && tid == 0) bits_or(words, bit0, variable ? 1u : 0u, 1);
            for (u32 r = w; r < n_rec; r += DSRC_WARPS) {      // EncodeRecords (:160-187 / :240-287)
                const u32 len = R.qua_len[rb + r], tl = R.trunc_len[rb + r];
                const u32 lim = trunc ? tl : len;
                const u8* qr = q + R.qcat_off[rb + r];
                u64 pos = bit0 + big[r];
                if (trunc) {
                    if (ln == 0) {
                        bits_or(words, pos, len != tl ? 1u : 0u, 1);
                        if (len != tl) bits_or(words, pos + 1, tl, variable ? dsrc_bit_length((u64)len) : max_bits);
                    }
                    pos += 1 + (len != tl ? (variable ? dsrc_bit_length((u64)len) : max_bits) : 0);
                }
                for (u32 j0 = 0; j0 < lim; j0 += 32) {
                    const u32 j = j0 + ln; const bool in = j < lim;
                    u32 c = 0, l = 0;
                    if (in) { const u32 e = j * Sy + S.qrank[qr[j]]; c = A->code[e]; l = A->len[e]; }
                    const u32 inc = warp_incl_sum(l);
                    if (in) bits_or(words, pos + inc - l, c, l);
                    pos += __shfl_sync(0xFFFFFFFFu, inc, 31);
                }
            }
            __syncthreads();
            if (tid == 0) st.stream_size[3] = S.hdr_bytes + (u32)nbytes;
            continue;
        }

        // ============ run-length + context Huffman (QualityRLEModeler) ============
        u32* runpos = (u32*)big;
        if (tid == 0) { S.carry_u32 = 0; S.carry_max = 0; }
        S.qf[tid] = 0; S.lf[tid] = 0;
        __syncthreads();
        // EncodeRecords (:142-205): runs over the whole block, at most 255 symbols each
        for (u32 base = 0; base < M; base += DSRC_CTA) {
            const u32 i = base + tid; const bool in = i < M;
            const bool head = in && (i == 0 || q[i] != q[i - 1]);
            // natural-run head index through a max-scan (two barriers), then sub-run starts every 255 symbols
            u32 v = head ? i : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, v, o); if (ln >= (u32)o) v = max(v, t); }
            __syncthreads();
            if (ln == 31) S.scan[w] = v;
            __syncthreads();
            u32 m = S.carry_max;
            for (u32 k = 0; k < w; ++k) m = max(m, S.scan[k]);
            const u32 h = max(v, m);
            const bool start = in && ((i - h) % 255u == 0);
            u32 total, ex = block_excl_sum(start ? 1u : 0u, S.scan, &total);
            const u32 carry = S.carry_u32;
            if (start) runpos[carry + ex] = i;
            __syncthreads();
            if (tid == DSRC_CTA - 1) S.carry_max = h;
            if (tid == 0) S.carry_u32 = carry + total;
            __syncthreads();
        }
        const u32 runs = S.carry_u32;
        if (tid == 0) runpos[runs] = M;
        __syncthreads();
        for (u32 k = tid; k < runs; k += DSRC_CTA) {
            atomicAdd(&S.qf[q[runpos[k]]], 1u);
            atomicAdd(&S.lf[runpos[k + 1] - runpos[k] - 1], 1u);
        }
        __syncthreads();
        if (tid == 0) {                                        // CalculateSymbolIndices (:207-231)
            u32 nq = 0, nl = 0;
            for (u32 i = 0; i < 256; ++i) { S.qrank[i] = S.qf[i] ? (u8)nq++ : (u8)255; S.lrank[i] = S.lf[i] ? (u8)nl++ : (u8)255; }
            S.nq = nq; S.nl = nl;
            if (nq > 1 && nl < 2) S.status = ST_UNSUPPORTED;   // single-symbol length tree: undefined behaviour upstream (Q5)
            if ((u64)nq * nq + (u64)nq * nl > Q0_ENTRIES) S.status = ST_UNSUPPORTED;
            BitW hw; hw.init(out, out_cap); hw.pos = 1;
            hw.be32(runs);
            for (u32 i = 0; i < 256; ++i) hw.bit(S.qrank[i] != 255);
            for (u32 i = 0; i < 256; ++i) hw.bit(S.lrank[i] != 255);
            hw.flush();
            S.hdr_bytes = hw.pos;
        }
        __syncthreads();
        if (S.status != ST_OK) { if (tid == 0) st.status = S.status; continue; }
        const u32 nq = S.nq, nl = S.nl;
        if (nq == 1) {                                         // EncodeRuns degenerate branch (:360-372)
            if (tid == 0) {
                u32 p = S.hdr_bytes;
                if (nl > 1) out[p++] = S.lrank[runpos[1] - runpos[0] - 1];
                st.stream_size[3] = p;
            }
            continue;
        }
        u32* QF = A->hist; u32* LF = A->hist + nq * nq;
        for (u32 i = tid; i < nq * nq + nq * nl; i += DSRC_CTA) A->hist[i] = 0;
        __syncthreads();
        for (u32 k = tid; k < runs; k += DSRC_CTA) {           // ComputeHuffmanContext (:233-310)
            const u32 qs = S.qrank[q[runpos[k]]], ls = S.lrank[runpos[k + 1] - runpos[k] - 1];
            const u32 prev = k ? S.qrank[q[runpos[k - 1]]] : 0u;
            atomicAdd(&QF[prev * nq + qs], 1u); atomicAdd(&LF[qs * nl + ls], 1u);
        }
        __syncthreads();
        const u32 sstr = ser_stride_for(max(nq, nl));
        u32* qcode = A->code; u32* lcode = A->code + nq * nq; u8* qlen = A->len; u8* llen = A->len + nq * nq;
        for (u32 j = w; j < 2 * nq; j += DSRC_WARPS) {          // tree order in the stream: q0, l0, q1, l1, ...  (:312-322)
            const u32 c = j >> 1;
            u32 sz;
            if ((j & 1) == 0) sz = huf_build_warp(QF + c * nq, nq, &A->work[w], qcode + c * nq, qlen + c * nq, A->ser + (u64)j * sstr, sstr);
            else sz = huf_build_warp(LF + c * nl, nl, &A->work[w], lcode + c * nl, llen + c * nl, A->ser + (u64)j * sstr, sstr);
            if (ln == 0) A->ser_size[j] = sz;
        }
        __syncthreads();
        copy_trees(S, A, 2 * nq, sstr, out, out_cap);
        if (S.status != ST_OK) { if (tid == 0) st.status = S.status; continue; }
        // EncodeRuns (:324-358): two passes (count, write) over the runs
        bool failed = false;
        for (int pass = 0; pass < 2 && !failed; ++pass) {
            if (pass == 1) {
                const u64 nbytes = (S.total_bits + 7) / 8;
                if ((u64)S.hdr_bytes + nbytes + 8 > out_cap) { if (tid == 0) st.status = ST_OVERFLOW; failed = true; break; }
                for (u64 i = tid; i < nbytes + 8; i += DSRC_CTA) out[S.hdr_bytes + i] = 0;
                __syncthreads();
            }
            unsigned long long run = 0;
            for (u32 base = 0; base < runs; base += DSRC_CTA) {
                const u32 k = base + tid; const bool in = k < runs;
                u32 qs = 0, ls = 0, prev = 0, nb = 0;
                if (in) {
                    qs = S.qrank[q[runpos[k]]]; ls = S.lrank[runpos[k + 1] - runpos[k] - 1];
                    prev = k ? S.qrank[q[runpos[k - 1]]] : 0u;
                    nb = qlen[prev * nq + qs] + llen[qs * nl + ls];
                }
                u32 total, ex = block_excl_sum(nb, S.scan, &total);
                if (pass == 1 && in) {
                    const u64 pos = (u64)S.hdr_bytes * 8 + run + ex;
                    bits_or((u32*)out, pos, qcode[prev * nq + qs], qlen[prev * nq + qs]);
                    bits_or((u32*)out, pos + qlen[prev * nq + qs], lcode[qs * nl + ls], llen[qs * nl + ls]);
                }
                run += total;
                __syncthreads();
            }
            if (pass == 0) { if (tid == 0) S.total_bits = run; __syncthreads(); }
        }
        __syncthreads();
        if (tid == 0 && !failed) st.stream_size[3] = S.hdr_bytes + (u32)((S.total_bits + 7) / 8);
    }
}

__global__ void __launch_bounds__(DSRC_CTA) k_d0_dna(Workspace ws, u8* arena, u64 arena_stride)
{
    __shared__ Q0Shared S;
    Q0Arena* A = (Q0Arena*)(arena + (u64)blockIdx.x * arena_stride);
    const u32 tid = threadIdx.x;
    for (u32 blk = blockIdx.x; blk < ws.n_blocks; blk += gridDim.x) {
        const BlockDesc& d = ws.desc[blk];
        BlockState& st = ws.state[blk];
        __syncthreads();
        if (st.status != ST_OK) continue;
        const u32 M = st.d_total;
        const u8* sq = ws.dcat + d.sym_base;
        u8* out = ws.streams + d.stream_base + stream_offset(d, 2);
        const u32 out_cap = d.stream_cap[2];
        const u32 scheme = st.d_count == 0 ? 255u : (st.d_count <= 4 ? 0u : 1u);   // DnaNormalModelerProxy::SelectSchemeId
        if (tid == 0) { st.d_scheme = (u8)scheme; out[0] = (u8)scheme; S.status = ST_OK; }
        if (scheme == 255) { if (tid == 0) st.stream_size[2] = 1; continue; }
        if (scheme == 0) {                                     // DnaModelerBasicB2::Encode: Put2Bits(sym & 3)
            const u32 nbytes = (M + 3) / 4;
            if (1 + nbytes > out_cap) { if (tid == 0) st.status = ST_OVERFLOW; continue; }
            for (u32 k = tid; k < nbytes; k += DSRC_CTA) {
                u32 v = 0;
#pragma unroll
                for (u32 t = 0; t < 4; ++t) { const u32 i = 4 * k + t; v = (v << 2) | (i < M ? (sq[i] & 3u) : 0u); }
                out[1 + k] = (u8)v;
            }
            if (tid == 0) st.stream_size[2] = 1 + nbytes;
            continue;
        }
        // DnaModelerHuffman (ProcessStats :21-39, Encode :41-73)
        const u32 Sy = st.d_count;
        if (tid < 32) {
            if (tid < 20) A->hist[tid] = (tid < Sy && st.drank[tid] < 20) ? st.dfreq[st.drank[tid]] : 0u;   // symbolFreqs[symbols[i]] (:36), see SURVEY a11
            __syncwarp();
            u32 sz = huf_build_warp(A->hist, Sy, &A->work[0], A->code, A->len, A->ser, 256);
            if (tid == 0) {
                BitW hw; hw.init(out, out_cap); hw.pos = 1;
                for (u32 i = 0; i < 20; ++i) hw.bit(st.drank[i] != 255);
                hw.flush();
                if (sz == 0xFFFFFFFFu) S.status = ST_OVERFLOW; else for (u32 i = 0; i < sz; ++i) hw.byte(A->ser[i]);
                if (hw.ovf) S.status = ST_OVERFLOW;
                S.hdr_bytes = hw.pos;
                u64 tb = 0;
                for (u32 s = 0; s < 20; ++s) if (st.drank[s] != 255) tb += (u64)st.dfreq[s] * A->len[st.drank[s]];
                S.total_bits = tb;
            }
        }
        if (tid < 20) S.qrank[tid] = st.drank[tid];
        __syncthreads();
        if (S.status != ST_OK) { if (tid == 0) st.status = S.status; continue; }
        const u64 nbytes = (S.total_bits + 7) / 8;
        if ((u64)S.hdr_bytes + nbytes + 8 > out_cap) { if (tid == 0) st.status = ST_OVERFLOW; continue; }
        for (u64 i = tid; i < nbytes + 8; i += DSRC_CTA) out[S.hdr_bytes + i] = 0;
        __syncthreads();
        unsigned long long run = (u64)S.hdr_bytes * 8;
        for (u32 base = 0; base < M; base += DSRC_CTA) {
            const u32 i = base + tid; const bool in = i < M;
            u32 c = 0, l = 0;
            if (in) { const u32 r = S.qrank[sq[i] < 20 ? sq[i] : 0]; c = A->code[r]; l = A->len[r]; }
            u32 total, ex = block_excl_sum(l, S.scan, &total);
            if (in) bits_or((u32*)out, run + ex, c, l);
            run += total;
        }
        __syncthreads();
        if (tid == 0) st.stream_size[2] = S.hdr_bytes + (u32)nbytes;
    }
}

void launch_q0_quality(const Workspace& ws, cudaStream_t s, u8* arena, u64 stride, u32 ctas)
{
    u32 g = ctas < ws.n_blocks ? ctas : ws.n_blocks;
    k_q0_quality<<<g ? g : 1, DSRC_CTA, 0, s>>>(ws, arena, stride);
}
void launch_d0_dna(const Workspace& ws, cudaStream_t s, u8* arena, u64 stride, u32 ctas)
{
    u32 g = ctas < ws.n_blocks ? ctas : ws.n_blocks;
    k_d0_dna<<<g ? g : 1, DSRC_CTA, 0, s>>>(ws, arena, stride);
}
