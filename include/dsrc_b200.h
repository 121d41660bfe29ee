/*
 * dsrc_b200 -- C ABI of the B200-native DSRC block codec (the drop-in boundary).
 *
 * The reference has no FFI; its seam for the hot path is the C++ class comp::BlockCompressor
 * (/root/reference/src/BlockCompressor.h:66-73):
 *     void Store(core::BitMemoryWriter&, fq::StreamsInfo& raw, fq::StreamsInfo& comp, const fq::FastqDataChunk&);
 *     void Read (core::BitMemoryReader&, fq::FastqDataChunk&);
 * called once per block from DsrcCompressor::Process / DsrcDecompressor::Process
 * (src/DsrcWorker.cpp:48,94) and DsrcCompressorST/DsrcDecompressorST::Process (src/DsrcOperator.cpp:107,205).
 * A GPU wants thousands of blocks per call, so the entry points below take a BATCH of blocks; a
 * BlockCompressor shim calls them with n = 1, the operators with whole block queues
 * (INTEGRATION.md shows both bindings).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 or a negative DSRCGPU_E_* code
 * and never throws; input buffers are read-only (Store's in-place preprocessing,
 * src/RecordsProcessor.cpp:209-267, happens on the device copy); a context is single-caller
 * (one per GPU), internally multi-stream. There is NO CPU fallback: without a CUDA device
 * dsrcgpu_create fails with DSRCGPU_E_CUDA.
 */
#ifndef DSRC_B200_H
#define DSRC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsrcgpu_ctx dsrcgpu_ctx;

/* fq::FastqDatasetType (src/Common.h:46-69) */
typedef struct {
    uint32_t quality_offset;   /* 33 / 59 / 64; must be resolved (no auto = 0) */
    uint8_t plus_repetition;   /* '+' line repeats the title */
    uint8_t color_space;       /* must be 0: SOLiD colour space is out of scope */
} dsrcgpu_dataset_t;

/* comp::CompressionSettings (src/Common.h:115-147) */
typedef struct {
    uint32_t dna_order;        /* 0, 3, 6, 9  (= 3 * CLI -d) */
    uint32_t quality_order;    /* 0, 1, 2     (= CLI -q, lossless) */
    uint64_t tag_preserve_flags; /* must be 0 (-f field filtering out of scope) */
    uint8_t lossy;             /* must be 0 */
    uint8_t calc_crc32;        /* -c: CRC-32 words of titles / sequences / qualities in every block header (encode), verified on decode */
} dsrcgpu_settings_t;

enum {
    DSRCGPU_OK = 0,
    DSRCGPU_E_CUDA = -1,        /* no device / CUDA runtime failure (dsrcgpu_last_error has the text) */
    DSRCGPU_E_ARG = -2,         /* bad argument or unsupported setting */
    DSRCGPU_E_CAPACITY = -3,    /* caller's output buffer too small */
    DSRCGPU_E_MALFORMED = -4,   /* a block is not well-formed FASTQ / a compressed block is corrupt */
    DSRCGPU_E_UNSUPPORTED = -5, /* input outside the supported envelope (see DESIGN.md "limits") */
    DSRCGPU_E_NOMEM = -6
};

/* StreamsInfo order (src/Common.h:75-82) */
enum { DSRCGPU_STREAM_META = 0, DSRCGPU_STREAM_TAG = 1, DSRCGPU_STREAM_DNA = 2, DSRCGPU_STREAM_QUALITY = 3 };

/* == BlockCompressor::BlockCompressor(datasetType, settings)  (src/BlockCompressor.cpp:53-94).
 * max_block_bytes: largest FASTQ block that will be submitted (CLI: -b MB << 20);
 * max_inflight_blocks: blocks processed per internal batch (0 = about 2 GiB of FASTQ, at most 8192 blocks; the workspace of a batch is
 * about 13x its FASTQ bytes, and 3 batches are in flight). */
int dsrcgpu_create(dsrcgpu_ctx** ctx, int device, const dsrcgpu_dataset_t* dataset,
                   const dsrcgpu_settings_t* settings, uint32_t max_block_bytes, uint32_t max_inflight_blocks);
void dsrcgpu_destroy(dsrcgpu_ctx* ctx);
const char* dsrcgpu_last_error(dsrcgpu_ctx* ctx);

/* == BlockCompressor::Store for n blocks (src/BlockCompressor.cpp:208-259).
 *  fastq        host memory holding the blocks; block i = fastq[blk_off[i] .. +blk_len[i]) exactly as
 *               IFastqStreamReader::ReadNextChunk cuts them (no trailing '\n', src/FastqStream.cpp:18-72)
 *  blk_tagcap   per block: capacity of the reference's TagStats::fields vector BEFORE the block (the one
 *               cross-block state that changes bytes; track it with dsrcgpu_tag_capacity_after).
 *               NULL = every block "warm" (capacity >= field count).
 *  out          host buffer; compressed blocks are written back to back in order; out_sizes[i] = size of block i
 *  raw_stream_sizes / comp_stream_sizes   n x 4 StreamsInfo (may be NULL)                                  */
int dsrcgpu_encode_blocks(dsrcgpu_ctx* ctx, const uint8_t* fastq, const uint64_t* blk_off, const uint32_t* blk_len,
                          const uint32_t* blk_tagcap, uint32_t n, uint8_t* out, uint64_t out_cap,
                          uint32_t* out_sizes, uint64_t* raw_stream_sizes, uint64_t* comp_stream_sizes);

/* Same contract with `fastq` and `out` in DEVICE memory of the context's GPU (inputs already resident in
 * HBM): no host<->device payload copies, only sizes come back. Used for the kernel-throughput measurement. */
int dsrcgpu_encode_blocks_device(dsrcgpu_ctx* ctx, const uint8_t* d_fastq, const uint64_t* blk_off, const uint32_t* blk_len,
                                 const uint32_t* blk_tagcap, uint32_t n, uint8_t* d_out, uint64_t out_cap,
                                 uint32_t* out_sizes, uint64_t* raw_stream_sizes, uint64_t* comp_stream_sizes);

/* == BlockCompressor::Read for n blocks (src/BlockCompressor.cpp:262-297). dsrc = host memory with the compressed
 * blocks; decoded FASTQ chunks (chunkSize + 1 bytes each, final '\n' included) are written back to back. */
int dsrcgpu_decode_blocks(dsrcgpu_ctx* ctx, const uint8_t* dsrc, const uint64_t* blk_off, const uint32_t* blk_len,
                          uint32_t n, uint8_t* fastq_out, uint64_t out_cap, uint64_t* out_sizes);
int dsrcgpu_decode_blocks_device(dsrcgpu_ctx* ctx, const uint8_t* d_dsrc, const uint64_t* blk_off, const uint32_t* blk_len,
                                 uint32_t n, uint8_t* d_fastq_out, uint64_t out_cap, uint64_t* out_sizes);

/* Host helpers for the Q1 state (pure functions, no device work):
 * number of fields TagAnalyzer::InitializeFieldsStats (src/TagModeler.cpp:159-224) creates for a title, and the
 * std::vector<Field> capacity after a block with that many fields (libstdc++ doubling growth). */
uint32_t dsrcgpu_tag_field_count(const uint8_t* title, uint32_t title_len);
uint32_t dsrcgpu_tag_capacity_after(uint32_t capacity_before, uint32_t n_fields);

/* == IFastqStreamReader::ReadNextChunk + GetNextRecordPos (src/FastqStream.cpp:18-98) over an in-memory FASTQ file:
 * fills off[]/len[] with the chunks the reference's reader hands to Store for a chunk buffer of cbuf bytes
 * (CLI: -b MB << 20). Pure host function. Returns the number of blocks (only the first max_blocks are stored). */
uint64_t dsrcgpu_cut_blocks(const uint8_t* data, uint64_t size, uint64_t cbuf, uint64_t* off, uint32_t* len, uint64_t max_blocks);

/* The same for a host that STREAMS the file through a bounded window (the reference's reader keeps one chunk buffer, src/FastqStream.h:74-89):
 * data[0..size) is a window of the file that starts at a block boundary; reader_state carries the one piece of reader state that outlives
 * a chunk (usesCrlf) from window to window -- 0 before the first one. The last block of a window that is not the end of the file was cut
 * at the window's end, not at a record: drop it and start the next window at its offset. */
uint64_t dsrcgpu_cut_blocks_window(const uint8_t* data, uint64_t size, uint64_t cbuf, uint64_t* off, uint32_t* len, uint64_t max_blocks,
                                   uint32_t* reader_state);

/* == FastqParser::Analyze (src/FastqParser.cpp:27-138) on the first chunk: fills plus_repetition / color_space and, when
 * ds->quality_offset is 0 on entry, the auto-detected quality offset. DSRCGPU_E_MALFORMED == "Error analyzing FASTQ dataset". Pure host. */
int dsrcgpu_analyze_first_chunk(const uint8_t* chunk, uint64_t size, dsrcgpu_dataset_t* ds);

/* == DsrcFileWriter::WriteFileHeader / WriteFileFooter (src/DsrcFile.cpp:112-170) and DsrcFileReader::ReadFileHeader / ReadFileFooter
 * (:264-314): archive = 40-byte header | blocks back to back | footer. Pure host. */
uint64_t dsrcgpu_archive_footer_size(uint64_t n_blocks);
int dsrcgpu_write_archive_header(uint8_t* out40, uint64_t n_blocks, uint64_t blocks_total_bytes);
int dsrcgpu_write_archive_footer(uint8_t* out, uint64_t out_cap, const uint32_t* block_sizes, uint64_t n_blocks,
                                 const dsrcgpu_dataset_t* ds, const dsrcgpu_settings_t* cs);
int dsrcgpu_read_archive_index(const uint8_t* arc, uint64_t size, uint64_t* n_blocks, uint64_t* blk_off, uint32_t* blk_len,
                               uint64_t max_blocks, dsrcgpu_dataset_t* ds, dsrcgpu_settings_t* cs);

/* header / footer for a host that streams the archive: footer position from the 40-byte header, then the index from the footer alone
 * (block i starts at 40 + the sizes before it). footer == NULL: validate the header and return the block count only. */
int dsrcgpu_archive_footer_span(const uint8_t* header40, uint64_t* footer_offset, uint64_t* footer_bytes);
int dsrcgpu_read_archive_footer(const uint8_t* header40, const uint8_t* footer, uint64_t footer_bytes, uint64_t file_size,
                                uint64_t* n_blocks, uint32_t* blk_len, uint64_t max_blocks, dsrcgpu_dataset_t* ds, dsrcgpu_settings_t* cs);

/* device-timed duration (ms, CUDA events on the context's stream) of the last encode/decode call */
float dsrcgpu_last_call_ms(dsrcgpu_ctx* ctx);

/* Measurement support: device time (ms, CUDA events on the context's streams) spent in each kernel family during
 * the last encode/decode call, and launch counts. names[i] are static strings. Returns the number of entries. */
int dsrcgpu_last_kernel_times(dsrcgpu_ctx* ctx, const char** names, float* ms, uint32_t* launches, int max_entries);
/* bit 0: per-kernel event timing; bit 1: in-kernel phase cycle counters (clock64 of thread 0 of every CTA, summed
 * over CTAs; slot map in DESIGN.md "instrumentation"). */
void dsrcgpu_set_profiling(dsrcgpu_ctx* ctx, int on);
/* copies the 64 phase counters accumulated since the last reset */
int dsrcgpu_phase_cycles(dsrcgpu_ctx* ctx, uint64_t* out64, int reset);

/* frees the internal device workspaces (they are re-created on demand by the next encode/decode call) */
int dsrcgpu_release_workspace(dsrcgpu_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
