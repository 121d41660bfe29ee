/*
 * TEST INFRASTRUCTURE ONLY -- the CPU oracle for the DSRC 2.02 per-block codec.
 *
 * A plain-C restatement of the reference algorithm on the hot path named by BASELINE.json
 * (BlockCompressor::Store / ::Read and everything below it). Nothing in the product
 * (dsrc_b200/, include/) may include, link or call this file; only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() use it, as the checker.
 *
 * Parity status: PINNED BY EXECUTION. The reference ships no golden vectors;
 * this restatement is checked byte-for-byte against the unmodified reference compiled into
 * oracle/_ref/libdsrcref.so (tests/test_oracle_vs_ref.py) and against fixtures generated from it
 * (tests/golden/, made by tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line (relative to /root/reference/) it follows.
 */
#ifndef DSRC_ORACLE_H
#define DSRC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsrc_oracle dsrc_oracle_t;

/* One instance == one comp::BlockCompressor (src/BlockCompressor.h:66). It carries the one piece of
 * cross-block state that influences the bitstream: the capacity of TagStats::fields. */
dsrc_oracle_t* dsrc_oracle_create(uint32_t quality_offset, int plus_repetition,
                                  uint32_t dna_order, uint32_t quality_order);
void dsrc_oracle_destroy(dsrc_oracle_t* o);

/* BlockCompressor::Store (src/BlockCompressor.cpp:208). fastq = one chunk WITHOUT its final '\n'.
 * Returns the compressed size or a negative error. raw4/comp4 = StreamsInfo in enum order
 * Meta, Tag, Dna, Quality (src/Common.h:75-82); either may be NULL. */
int64_t dsrc_oracle_store(dsrc_oracle_t* o, const uint8_t* fastq, uint64_t size,
                          uint8_t* out, uint64_t cap, uint64_t* raw4, uint64_t* comp4);

/* BlockCompressor::Read (src/BlockCompressor.cpp:262). Returns chunkSize+1 bytes of FASTQ. */
int64_t dsrc_oracle_read(dsrc_oracle_t* o, const uint8_t* blk, uint64_t size,
                         uint8_t* out, uint64_t cap);

/* CompressionSettings::calculateCrc32 (CLI -c): Store adds the three CRC-32 words (titles, sequences, qualities of the
 * raw records, src/RecordsProcessor.cpp:135-152) to the block header (BlockCompressor.cpp:424-440); Read parses them and
 * dsrc_oracle_last_crc_ok tells whether the decoded block matched (VerifyChecksum, :576-594). */
void dsrc_oracle_set_crc(dsrc_oracle_t* o, int on);
int dsrc_oracle_last_crc_ok(const dsrc_oracle_t* o);

/* Current emulated capacity of TagStats::fields (for tests of Q1). */
uint32_t dsrc_oracle_tag_capacity(const dsrc_oracle_t* o);

/* IFastqStreamReader::ReadNextChunk + GetNextRecordPos (src/FastqStream.cpp:18-98) over an in-memory
 * file: fills off[]/len[] with the chunks the reference would hand to Store. Returns the block count
 * (may exceed max_blocks; only the first max_blocks are written). */
uint64_t dsrc_oracle_cut_blocks(const uint8_t* file, uint64_t size, uint64_t cbuf,
                                uint64_t* off, uint64_t* len, uint64_t max_blocks);

/* FastqParser::Analyze (src/FastqParser.cpp:27-138) on the first chunk. *qoff in: 0 = auto-detect.
 * Returns 1 on success, 0 on failure (reference: "Error analyzing FASTQ dataset"). */
int dsrc_oracle_analyze(const uint8_t* chunk, uint64_t size, uint32_t* qoff, int* plus_rep, int* color_space);

/* DsrcFileWriter header/footer (src/DsrcFile.cpp:112-170): assembles a whole .dsrc archive from
 * already-compressed blocks laid back to back in blocks[]. Returns archive size or negative. */
int64_t dsrc_oracle_write_archive(const uint8_t* blocks, const uint32_t* block_sizes, uint64_t n_blocks,
                                  uint32_t quality_offset, int plus_rep, int color_space,
                                  uint32_t dna_order, uint32_t quality_order,
                                  uint8_t* out, uint64_t cap);

/* Whole-file single-thread compress == DsrcCompressorST::Process (src/DsrcOperator.cpp:47) on memory.
 * dna_level/quality_level as on the CLI (-d/-q); buf_bytes = chunk buffer size (CLI: MB<<20). */
int64_t dsrc_oracle_compress_mem(const uint8_t* file, uint64_t size, uint32_t dna_level, uint32_t quality_level,
                                 uint64_t buf_bytes, uint32_t quality_offset, uint8_t* out, uint64_t cap);

int64_t dsrc_oracle_compress_mem_crc(const uint8_t* file, uint64_t size, uint32_t dna_level, uint32_t quality_level,
                                     uint64_t buf_bytes, uint32_t quality_offset, int crc, uint8_t* out, uint64_t cap);

/* Whole-archive decompress == DsrcDecompressorST::Process (src/DsrcOperator.cpp:165) on memory. */
int64_t dsrc_oracle_decompress_mem(const uint8_t* arc, uint64_t size, uint8_t* out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
