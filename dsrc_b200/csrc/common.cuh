// dsrc_b200 -- shared device/host declarations for the sm_100a block codec.
// Reference citations are file:line relative to /root/reference/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint64_t u64;
typedef int64_t i64;

#define DSRC_CTA 256            // threads per CTA for the per-block kernels
#define DSRC_WARPS (DSRC_CTA / 32)

// per-block status codes written by kernels (mapped to DSRCGPU_E_* by the host)
enum : u32 {
    ST_OK = 0,
    ST_MALFORMED = 1,           // invalid / truncated record, line > 65535
    ST_UNSUPPORTED = 2,         // outside the supported envelope (too many fields, symbol >= alphabet, ...)
    ST_OVERFLOW = 3,            // a scratch stream buffer was too small
    ST_CRC = 5,                 // -c: the decoded block does not match its stored CRC-32 words (4 = decode retry, decode.cu)
};

// limits of the tag tokenizer (DESIGN.md "limits")
#define TAG_MAX_FIELDS 64
#define TAG_STAT_LEN 128        // Field::MAX_FIELD_STAT_LEN (src/TagModeler.h:27)
#define TAG_NUM_HUF 512         // Field::MAX_NUM_VAL_HUF
#define TAG_TEXT_SLOTS 640      // Huffman slots (256 symbols) per block for text-field positions (five 128-position fields)
#define TAG_NUM_SLOTS 8         // Huffman slots (512 symbols) per block for ValueVar/DeltaVar fields

// host-filled description of one block of a batch
struct BlockDesc {
    u64 in_off;                 // byte offset of the block in the device input buffer
    u32 in_len;
    u32 tag_cap;                // capacity of TagStats::fields before this block; 0xFFFFFFFF = warm
    u32 line_base, line_cap;    // slice of lines[]
    u32 rec_base, rec_cap;      // slice of the record arrays
    u64 sym_base;               // slice of qcat/dcat (bytes) and of elem/triple arrays (entries)
    u32 sym_cap;
    u32 n_fields;               // field count of the block's first title (host-computed)
    u64 ftab_base;              // slice of the field table (entries = n_fields * rec_cap)
    u64 stream_base;            // byte offset of this block's stream arena
    u32 stream_cap[4];          // META, TAG, DNA, QUALITY capacities (bytes); offsets are cumulative in this order
    u64 out_off;                // filled by the size scan: offset of the block in the dense output
};

// device-written per-block state
struct BlockState {
    u32 status;
    u32 n_lines, n_rec, chunk_size;
    u64 raw[4];                 // StreamsInfo raw sizes (src/FastqParser.cpp:152-157)
    u32 stream_size[4];         // compressed sizes, StreamsInfo order
    u32 flags;                  // BlockCompressor.h:76-81 : 2 = variable length, 4 = mixed tag formatting
    u32 min_len, max_len, raw_len, th_len, rle_len;   // QualityStats (src/Stats.h:69-101)
    u32 q_count, d_count;       // distinct symbols
    u32 q_total, d_total;       // symbols in the quality / DNA streams
    u32 qfreq[256];
    u32 dfreq[20];
    u8 qrank[256];
    u8 drank[20];
    u8 q_scheme, d_scheme;
    u8 pad[2];
    u32 total_size;
    u32 crc[3], crc_expected[3];   // -c: CRC-32 of titles / sequences / qualities (computed; read from the block header)
    u32 pre_flat;                  // the block was preprocessed by k_preprocess_flat (parse.cu); k_preprocess skips it
};

// compact per-block result copied back to the host
struct BlockResult {
    u32 status, total_size;
    u32 stream_size[4];
    u64 raw[4];
    u64 out_off;
};
// written by k_count_lines so the host can lay the batch out
struct BlockProbe { u32 n_lines, n_fields; };

// per-record arrays (structure of arrays, indexed rec_base + r)
struct RecArrays {
    u32* title_off; u32* seq_off; u32* qua_off;       // offsets relative to the block start
    u16* title_len; u16* qua_len; u16* dna_len; u16* trunc_len;
    u32* qcat_off; u32* dcat_off;                      // exclusive prefix sums of qua_len / dna_len inside the block
};

struct Workspace {
    const u8* in;               // device input
    const BlockDesc* desc;
    BlockState* state;
    u32* lines;
    RecArrays rec;
    u8* qcat; u8* dcat;         // processed quality bytes / compacted DNA indices, concatenated per block
    u64* elem_a; u64* elem_b;   // sort ping-pong (entries at sym_base)
    u64* trip_q; u64* trip_d;   // (freq | cum<<16 | tot<<32) per symbol
    u64* ftab;                  // tag field table
    u8* streams;                // stream arenas
    u8* tagpool;                // Huffman slots
    u64 tagpool_stride;
    u8* out;                    // dense output
    BlockResult* result;
    BlockProbe* probe;
    u64 out_cap;
    u32 n_blocks;
    u32 qoff;                   // quality offset
    u32 plus_rep;
    u32 dna_order, qua_order;
    u32 calc_crc;               // CompressionSettings::calculateCrc32
    u8* tab; u64 tab_stride;    // pool of adaptive-row tables of the tile/table model engine (a table is all-zero between blocks)
    u32* tab_mask; u32 tab_count;   // pool bitmap (bit set = in use) and size
    u32* model_queue;           // [3] next block of the quality / DNA / quality-partition-engine model launch of this batch (zeroed per batch)
    u64* prof;                  // optional: 64 phase cycle counters (clock64 deltas of thread 0 of every CTA), or null
};

__host__ __device__ inline u32 stream_offset(const BlockDesc& d, int s)
{
    u32 o = 0;
    for (int i = 0; i < s; ++i) o += d.stream_cap[i];
    return o;
}

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ u32 warp_incl_sum(u32 v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane_id() >= (u32)o) v += t; }
    return v;
}

// exclusive prefix sum over the CTA (DSRC_CTA threads); returns the exclusive value, *total = CTA sum.
// `sm` must hold DSRC_WARPS+1 words; contains two barriers.
__device__ __forceinline__ u32 block_excl_sum(u32 v, u32* sm, u32* total)
{
    u32 inc = warp_incl_sum(v);
    __syncthreads();
    if (lane_id() == 31) sm[warp_id()] = inc;
    __syncthreads();
    u32 base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < DSRC_WARPS; ++w) { u32 s = sm[w]; if ((u32)w < warp_id()) base += s; tot += s; }
    *total = tot;
    return base + inc - v;
}

__device__ __forceinline__ u32 warp_red_min(u32 v) { for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, o)); return v; }
__device__ __forceinline__ u32 warp_red_max(u32 v) { for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o)); return v; }
__device__ __forceinline__ u32 warp_red_sum(u32 v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o); return v; }
__device__ __forceinline__ i32 warp_red_imin(i32 v) { for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, o)); return v; }
__device__ __forceinline__ i32 warp_red_imax(i32 v) { for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o)); return v; }

// utils.h:177 bit_length: i for x < 2^i (i < 32), else 64
__host__ __device__ inline u32 dsrc_bit_length(u64 x)
{
    for (u32 i = 0; i < 32; ++i)
        if (x < (1ull << i)) return i;
    return 64;
}
__host__ __device__ inline u32 dsrc_ilog2(u32 x) { u32 r = 0; while (x > 1) { x >>= 1; ++r; } return r; }

// Serial MSB-first bit writer into global/shared bytes (BitMemoryWriter semantics: a flush pads to a byte,
// byte-level puts are only issued when aligned -- src/BitMemory.h:274-409).
struct BitW {
    u8* p; u32 pos, cap; u64 acc; u32 nacc; bool ovf;
    __device__ void init(u8* p_, u32 cap_) { p = p_; cap = cap_; pos = 0; acc = 0; nacc = 0; ovf = false; }
    __device__ void raw(u8 b) { if (pos < cap) p[pos] = b; else ovf = true; ++pos; }
    __device__ void bits(u32 v, u32 n)
    {
        if (n == 0) return;
        if (n < 32) v &= (1u << n) - 1;
        acc = (acc << n) | v; nacc += n;
        while (nacc >= 8) { raw((u8)(acc >> (nacc - 8))); nacc -= 8; }
    }
    __device__ void bit(u32 b) { bits(b & 1, 1); }
    __device__ void flush() { if (nacc) bits(0, 8 - nacc); }
    __device__ void byte(u8 b) { raw(b); }
    __device__ void be32(u32 v) { raw((u8)(v >> 24)); raw((u8)(v >> 16)); raw((u8)(v >> 8)); raw((u8)v); }
};

// OR `n` (<= 32) bits of v, MSB first, into a zero-initialised big-endian bit stream at bit position bitpos.
// Neighbouring writers may share words, hence the atomics.
__device__ __forceinline__ void bits_or(u32* words, u64 bitpos, u32 v, u32 n)
{
    if (n == 0) return;
    if (n < 32) v &= (1u << n) - 1;
    u64 w = bitpos >> 5; u32 sh = (u32)(bitpos & 31);
    u64 x = ((u64)v << (64 - n)) >> sh;           // bits placed in a 64-bit window starting at word w
    u32 hi = (u32)(x >> 32), lo = (u32)x;
    if (hi) atomicOr(&words[w], __byte_perm(hi, 0, 0x0123));
    if (lo) atomicOr(&words[w + 1], __byte_perm(lo, 0, 0x0123));
}

#endif  // __CUDACC__
