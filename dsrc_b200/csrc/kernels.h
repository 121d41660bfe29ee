// host-side launchers of the per-block kernels (each .cu defines its own)
#pragma once
#include "common.cuh"

void launch_count_lines(const Workspace& ws, cudaStream_t s);
void launch_parse(const Workspace& ws, cudaStream_t s);
void launch_preprocess(const Workspace& ws, cudaStream_t s);
void launch_tags(const Workspace& ws, cudaStream_t s, u32 ctas);   // ctas = number of TagPool arenas
// order-k contexts -> (freq,cum,tot) triples; ctas = persistent CTAs owning a sort arena of `stride` entries each
void launch_model_quality(const Workspace& ws, cudaStream_t s, u32 ctas, u64 stride);
void launch_model_dna(const Workspace& ws, cudaStream_t s, u32 ctas, u64 stride);
#define RC_GROUP_MAX 4
struct RcGroup { Workspace ws[RC_GROUP_MAX]; u32 n; };            // the batches (slots) whose chains one launch codes
void launch_rc_encode(const RcGroup& grp, cudaStream_t s);        // serial range-coder chains, one thread per (block, stream)
cudaError_t rc_init_device();                                      // fills the reciprocal table of the chains on the current device (once per context)
// -q0: positional / truncated / RLE Huffman; -d0: 2-bit pack / Huffman. arena: ctas x stride bytes of per-CTA scratch
void launch_q0_quality(const Workspace& ws, cudaStream_t s, u8* arena, u64 stride, u32 ctas);
void launch_d0_dna(const Workspace& ws, cudaStream_t s, u8* arena, u64 stride, u32 ctas);
u64 q0_arena_bytes(u64 max_block_bytes);
void launch_meta_and_sizes(const Workspace& ws, cudaStream_t s, u64 out_base, u64* cursor);  // StoreMetaData + dense output offsets (first block at out_base, or at *cursor which is then advanced)
void launch_gather(const Workspace& ws, cudaStream_t s);          // meta|tags|quality|dna -> dense output
void launch_copy_words(void* dst, const void* src, size_t bytes, cudaStream_t s);   // small transfers over mapped pinned memory (bytes rounded up to 4)

u64 tagpool_bytes_per_block();

// decode (decode.cu): one thread per block for the bit-serial parts, one CTA per block for the FASTQ assembly
void launch_dec_probe(const Workspace& ws, cudaStream_t s);
void launch_dec_tags(const Workspace& ws, cudaStream_t s, void* pool, u32 pool_stride);
void launch_dec_quality(const Workspace& ws, cudaStream_t s, void* pool, u32 pool_stride, u8* arena, u64 arena_stride, u64 arena_bytes);
void launch_dec_dna(const Workspace& ws, cudaStream_t s, void* pool, u32 pool_stride, u8* arena, u64 arena_stride, u64 arena_bytes);
void launch_dec_assemble(const Workspace& ws, cudaStream_t s);
void launch_crc(const Workspace& ws, cudaStream_t s, u32 decode_mode);   // -c: CRC-32 words of the raw (0) / decoded (1) records
