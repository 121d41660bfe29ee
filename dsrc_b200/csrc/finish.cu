// Block assembly: StoreMetaData (src/BlockCompressor.cpp:403-443), the meta|tags|quality|dna layout of
// StoreRecords (:223-259), and the dense packing of the batch's blocks (the offsets the reference's writer
// derives from DsrcFileWriter::WriteNextChunk, src/DsrcFile.cpp:59-73).
#include "common.cuh"
#include "kernels.h"

// one CTA: meta streams + per-block totals + exclusive scan of the totals -> out_off
__global__ void __launch_bounds__(DSRC_CTA) k_meta_sizes(Workspace ws, u64 out_base, unsigned long long* cursor)
{
    __shared__ u32 sm[DSRC_WARPS + 1];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = cursor ? *cursor : out_base;      // cursor: running end of the dense output across batches (device-resident output)
    __syncthreads();
    for (u32 base = 0; base < ws.n_blocks; base += DSRC_CTA) {
        const u32 blk = base + threadIdx.x;
        u32 total = 0;
        if (blk < ws.n_blocks) {
            const BlockDesc& d = ws.desc[blk];
            BlockState& st = ws.state[blk];
            if (st.status == ST_OK) {
                u8* m = ws.streams + d.stream_base;       // META is the first sub-arena
                BitW w; w.init(m, d.stream_cap[0]);
                w.be32(st.n_rec); w.be32(st.max_len); w.be32(st.flags); w.be32(st.chunk_size);
                if (st.flags & 2u) w.be32(st.min_len);
                if (ws.calc_crc) { w.be32(st.crc[0]); w.be32(st.crc[1]); w.be32(st.crc[2]); }   // BlockCompressor.cpp:424-440
                st.stream_size[0] = w.pos;
                total = st.stream_size[0] + st.stream_size[1] + st.stream_size[2] + st.stream_size[3];
                st.total_size = total;
            }
        }
        // totals are < 2^31, a batch may exceed 2^32: scan in 32 bits inside the tile, carry in 64
        u32 tile_total, ex = block_excl_sum(total, sm, &tile_total);
        const unsigned long long carry = s_carry;
        if (blk < ws.n_blocks) {
            const BlockState& st = ws.state[blk];
            BlockResult& r = ws.result[blk];
            r.status = st.status; r.total_size = total;
            for (int k = 0; k < 4; ++k) { r.stream_size[k] = st.status == ST_OK ? st.stream_size[k] : 0; r.raw[k] = st.status == ST_OK ? st.raw[k] : 0; }
            r.out_off = carry + ex;
            if (st.status == ST_OK && carry + ex + total > ws.out_cap) r.status = 0x100 | ST_OVERFLOW;   // caller's buffer too small
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tile_total;
        __syncthreads();
    }
    if (cursor && threadIdx.x == 0) *cursor = s_carry;
}

__global__ void __launch_bounds__(DSRC_CTA) k_gather(Workspace ws)
{
    const u32 blk = blockIdx.x;
    const BlockDesc& d = ws.desc[blk];
    const BlockResult& r = ws.result[blk];
    if (r.status != ST_OK) return;
    u8* dst = ws.out + r.out_off;
    const int order[4] = {0, 1, 3, 2};                    // meta | tags | quality | dna
    for (int k = 0; k < 4; ++k) {
        const u8* src = ws.streams + d.stream_base + stream_offset(d, order[k]);
        const u32 n = r.stream_size[order[k]];
        // 16-byte chunks where source and destination are co-aligned, bytes otherwise
        if ((((uintptr_t)src ^ (uintptr_t)dst) & 15) == 0) {
            u32 head = (u32)((16 - ((uintptr_t)dst & 15)) & 15); if (head > n) head = n;
            for (u32 i = threadIdx.x; i < head; i += DSRC_CTA) dst[i] = src[i];
            const u32 nv = (n - head) / 16;
            const uint4* s4 = (const uint4*)(src + head); uint4* d4 = (uint4*)(dst + head);
            for (u32 i = threadIdx.x; i < nv; i += DSRC_CTA) d4[i] = s4[i];
            for (u32 i = head + nv * 16 + threadIdx.x; i < n; i += DSRC_CTA) dst[i] = src[i];
        } else {
            for (u32 i = threadIdx.x; i < n; i += DSRC_CTA) dst[i] = src[i];
        }
        dst += n;
    }
}

void launch_meta_and_sizes(const Workspace& ws, cudaStream_t s, u64 out_base, u64* cursor) { k_meta_sizes<<<1, DSRC_CTA, 0, s>>>(ws, out_base, (unsigned long long*)cursor); }
void launch_gather(const Workspace& ws, cudaStream_t s) { k_gather<<<ws.n_blocks, DSRC_CTA, 0, s>>>(ws); }

// Small host <-> device transfers of the scheduler (block descriptors up, layout probes and block results down) as a kernel over
// MAPPED pinned host memory instead of a cudaMemcpyAsync: a DMA copy of a few hundred KB would queue behind the 2 GiB payload copies
// of the other batches on the same copy engine, and the host waits for the probe.
__global__ void k_copy_words(u32* dst, const u32* src, u32 n)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
void launch_copy_words(void* dst, const void* src, size_t bytes, cudaStream_t s)
{
    const u32 n = (u32)((bytes + 3) / 4);
    if (!n) return;
    const u32 g = (n + 1023) / 1024;
    k_copy_words<<<g < 296u ? g : 296u, 256, 0, s>>>((u32*)dst, (const u32*)src, n);
}
