// BlockCompressorGpu -- the binding a DSRC maintainer adds to call dsrc_b200 through the reference's own seam.
//
// Same interface as comp::BlockCompressor (reference src/BlockCompressor.h:66-73: constructor, Store, Read, Reset), written
// against the reference's own types (core::BitMemoryWriter/Reader, fq::StreamsInfo, fq::FastqDataChunk) and, below them,
// against include/dsrc_b200.h only. Dropping it in: DsrcCompressor::Process / DsrcDecompressor::Process
// (src/DsrcWorker.cpp:37,83) and DsrcCompressorST/DsrcDecompressorST::Process (src/DsrcOperator.cpp:95,193) instantiate
// BlockCompressorGpu instead of BlockCompressor; nothing else changes. It needs the reference's headers on the include path,
// so it is compiled only where they are (oracle/Makefile target `shim` builds it into the test harness; tests/test_gpu_shim.py
// checks it byte for byte against the unmodified BlockCompressor). One block per call is the parity form; for throughput the
// operators hand whole block queues to dsrcgpu_encode_blocks (binding 2).
#pragma once
#include "BlockCompressor.h"      // reference: comp::CompressionSettings, fq::*, core::BitMemory*
#include "dsrc_b200.h"
#include <vector>

namespace dsrc { namespace comp {

class BlockCompressorGpu
{
public:
	BlockCompressorGpu(const fq::FastqDatasetType& type_, const CompressionSettings& settings_, uint32 maxBlockBytes_ = 8u << 20, int device_ = 0)
		:	ctx(NULL)
		,	tagCapacity(0)
	{
		dsrcgpu_dataset_t ds = { type_.qualityOffset, (uint8_t)type_.plusRepetition, (uint8_t)type_.colorSpace };
		dsrcgpu_settings_t cs = { settings_.dnaOrder, settings_.qualityOrder, settings_.tagPreserveFlags,
								  (uint8_t)settings_.lossy, (uint8_t)settings_.calculateCrc32 };
		if (dsrcgpu_create(&ctx, device_, &ds, &cs, maxBlockBytes_, 1) != DSRCGPU_OK)
			throw DsrcException("dsrc_b200: no CUDA device or unsupported settings");
	}

	~BlockCompressorGpu()
	{
		dsrcgpu_destroy(ctx);
	}

	// BlockCompressor::Store (src/BlockCompressor.cpp:208-259). The reference destroys the chunk in place; here it is read-only.
	void Store(core::BitMemoryWriter& memory_, fq::StreamsInfo& rawStreamInfo_, fq::StreamsInfo& compStreamInfo_,
			   const fq::FastqDataChunk& chunk_)
	{
		uint64_t off = 0, raw[4], cmp[4];
		uint32_t len = (uint32_t)chunk_.size, size = 0, cap = tagCapacity;
		const uint8_t* p = chunk_.data.Pointer();
		// the capacity of TagStats::fields this compressor instance has reached so far decides which fields lose a count
		uint32_t titleLen = 0;
		while (titleLen < len && p[titleLen] != '\n' && p[titleLen] != '\r')
			++titleLen;
		tagCapacity = dsrcgpu_tag_capacity_after(tagCapacity, dsrcgpu_tag_field_count(p, titleLen));
		if (out.size() < (size_t)len + len / 2 + 4096)
			out.resize((size_t)len + len / 2 + 4096);
		int rc = DSRCGPU_E_CAPACITY;
		for (int attempt = 0; attempt < 4 && rc == DSRCGPU_E_CAPACITY; ++attempt)		// a tiny block with many text fields can exceed 1.5x its input
		{
			if (attempt > 0)
				out.resize(out.size() * 4);
			rc = dsrcgpu_encode_blocks(ctx, p, &off, &len, &cap, 1, out.data(), out.size(), &size, raw, cmp);
		}
		if (rc != DSRCGPU_OK)
			throw DsrcException(dsrcgpu_last_error(ctx));
		for (int i = 0; i < 4; ++i)
		{
			rawStreamInfo_.sizes[i] = raw[i];
			compStreamInfo_.sizes[i] = cmp[i];
		}
		memory_.PutBytes(out.data(), size);						// BitMemory.h:356: byte-aligned append
	}

	// BlockCompressor::Read (src/BlockCompressor.cpp:262-297)
	void Read(core::BitMemoryReader& memory_, fq::FastqDataChunk& chunk_)
	{
		uint64_t off = 0, got = 0;
		uint32_t len = (uint32_t)(memory_.Size() - memory_.Position());
		const uint8_t* blk = memory_.Pointer() + memory_.Position();
		// chunkSize is the fourth big-endian word of the block header (ReadMetaData, :302-308); the chunk holds chunkSize + 1 bytes (:279)
		uint64 need = (((uint64)blk[12] << 24) | ((uint64)blk[13] << 16) | ((uint64)blk[14] << 8) | blk[15]) + 1;
		if (chunk_.data.Size() < need)
			chunk_.data.Extend(need);							// :282-285
		int rc = dsrcgpu_decode_blocks(ctx, blk, &off, &len, 1, chunk_.data.Pointer(), chunk_.data.Size(), &got);
		if (rc != DSRCGPU_OK)
			throw DsrcException(dsrcgpu_last_error(ctx));		// with -c: "CRC32 checksums mismatch." (src/DsrcWorker.cpp:60)
		chunk_.size = got;
		memory_.SetPosition(memory_.Size());
	}

	void Reset()												// BlockCompressor::Reset: statistics only, the field vector keeps its capacity
	{}

private:
	dsrcgpu_ctx* ctx;
	uint32_t tagCapacity;
	std::vector<uint8_t> out;
};

} }
