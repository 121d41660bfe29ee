// DsrcCompressorGpu / DsrcDecompressorGpu -- IDsrcOperator implementations (reference src/DsrcOperator.h:28-119) over dsrc_b200.
//
// Where the reference's DsrcCompressorMT::Process (src/DsrcOperator.cpp:230-394) starts a reader thread, `threadNum` workers each
// owning a BlockCompressor, and a writer, these operators hand the whole block queue to the library in one call
// (INTEGRATION.md binding 2): the GPU's stream scheduler is the worker pool. Same interface (`bool Process(const InputParameters&)`,
// IsError/GetError), same InputParameters, same .dsrc bytes as `dsrc c -t1` -- main.cpp:61-74 only has to pick these classes.
// Needs the reference's headers (IDsrcOperator, InputParameters, CompressionSettings); built by oracle/Makefile target `shim`,
// checked against the reference's own operators by tests/test_gpu_shim.py. File I/O is whole-file (the reference streams; a
// maintainer would keep FastqFileReader/DsrcFileWriter and feed batches -- the codec calls are the same).
#pragma once
#include "DsrcOperator.h"
#include "dsrc_b200.h"
#include <cstdio>
#include <string>
#include <vector>

namespace dsrc { namespace comp {

namespace gpuop {

inline bool ReadWholeFile(const std::string& name_, std::vector<uint8_t>& data_)
{
	FILE* f = std::fopen(name_.c_str(), "rb");
	if (f == NULL)
		return false;
	std::fseek(f, 0, SEEK_END);
	long long n = std::ftell(f);
	std::fseek(f, 0, SEEK_SET);
	data_.resize(n > 0 ? (size_t)n : 0);
	bool ok = n <= 0 || std::fread(data_.data(), 1, (size_t)n, f) == (size_t)n;
	std::fclose(f);
	return ok;
}

struct Context													// dsrcgpu_ctx with scope
{
	dsrcgpu_ctx* ctx;
	Context() : ctx(NULL) {}
	~Context() { if (ctx != NULL) dsrcgpu_destroy(ctx); }
};

} // namespace gpuop


class DsrcCompressorGpu : public IDsrcOperator
{
public:
	bool Process(const InputParameters& args_)
	{
		ClearError();
		if (args_.useFastqStdIo || args_.lossyCompression || args_.tagPreserveFlags != 0)
		{
			AddError("stdin/stdout, lossy mode and field filtering are outside the GPU operator's envelope");
			return false;
		}
		std::vector<uint8_t> fastq;
		if (!gpuop::ReadWholeFile(args_.inputFilename, fastq))
		{
			AddError("Cannot open file to read:" + args_.inputFilename);						// FileStream.cpp:78
			return false;
		}

		// the block queue IFastqStreamReader::ReadNextChunk would produce (src/FastqStream.cpp:18-98)
		const uint64_t cbuf = (uint64_t)args_.fastqBufferSizeMB << 20;
		const uint64_t n = dsrcgpu_cut_blocks(fastq.data(), fastq.size(), cbuf, NULL, NULL, 0);
		std::vector<uint64_t> off(n + 1);
		std::vector<uint32_t> len(n + 1), cap(n + 1), size(n + 1);
		dsrcgpu_cut_blocks(fastq.data(), fastq.size(), cbuf, off.data(), len.data(), n);

		// FastqFileReader::AnalyzeFirstChunk (src/FastqIo.cpp:26-44): quality offset, '+' repetition, colour space
		dsrcgpu_dataset_t ds = { args_.qualityOffset, 0, 0 };
		if (n == 0 || dsrcgpu_analyze_first_chunk(fastq.data() + off[0], len[0], &ds) != DSRCGPU_OK || ds.color_space)
		{
			AddError("Error analyzing FASTQ dataset");											// DsrcOperator.cpp:284
			return false;
		}
		const CompressionSettings settings = GetCompressionSettings(args_);
		dsrcgpu_settings_t cs = { settings.dnaOrder, settings.qualityOrder, 0, 0, (uint8_t)settings.calculateCrc32 };

		// one compressor in file order == `dsrc c -t1`: capacity of TagStats::fields before every block
		uint32_t capacity = 0, maxLen = 0;
		uint64_t bound = 0;
		for (uint64_t i = 0; i < n; ++i)
		{
			const uint8_t* p = fastq.data() + off[i];
			uint32_t tl = 0;
			while (tl < len[i] && p[tl] != '\n' && p[tl] != '\r')
				++tl;
			cap[i] = capacity;
			capacity = dsrcgpu_tag_capacity_after(capacity, dsrcgpu_tag_field_count(p, tl));
			maxLen = len[i] > maxLen ? len[i] : maxLen;
			bound += (uint64_t)len[i] + len[i] / 2 + 4096;
		}

		gpuop::Context c;
		if (dsrcgpu_create(&c.ctx, 0, &ds, &cs, maxLen + 64, 0) != DSRCGPU_OK)
		{
			AddError("dsrc_b200: no CUDA device or unsupported settings");
			return false;
		}
		std::vector<uint8_t> blocks(bound);
		if (dsrcgpu_encode_blocks(c.ctx, fastq.data(), off.data(), len.data(), cap.data(), (uint32_t)n, blocks.data(), blocks.size(),
								  size.data(), NULL, NULL) != DSRCGPU_OK)
		{
			AddError(dsrcgpu_last_error(c.ctx));
			return false;
		}
		uint64_t total = 0;
		for (uint64_t i = 0; i < n; ++i)
			total += size[i];

		if (args_.calculateCrc32)									// DsrcOperator.cpp:113-123: verify by decoding what was written
		{
			std::vector<uint64_t> boff(n), got(n);
			uint64_t p = 0, outBytes = 0;
			for (uint64_t i = 0; i < n; ++i) { boff[i] = p; p += size[i]; outBytes += (uint64_t)len[i] + 2; }
			std::vector<uint8_t> back(outBytes + 64);
			if (dsrcgpu_decode_blocks(c.ctx, blocks.data(), boff.data(), size.data(), (uint32_t)n, back.data(), back.size(), got.data()) != DSRCGPU_OK)
			{
				AddError("CRC32 checksums mismatch.");
				return false;
			}
		}

		// DsrcFileWriter: 40-byte header | blocks | footer (src/DsrcFile.cpp:112-170)
		uint8_t header[40];
		std::vector<uint8_t> footer(dsrcgpu_archive_footer_size(n));
		dsrcgpu_write_archive_header(header, n, total);
		dsrcgpu_write_archive_footer(footer.data(), footer.size(), size.data(), n, &ds, &cs);
		FILE* f = std::fopen(args_.outputFilename.c_str(), "wb");
		if (f == NULL)
		{
			AddError("Cannot open file to write:" + args_.outputFilename);						// FileStream.cpp:142
			return false;
		}
		bool ok = std::fwrite(header, 1, 40, f) == 40 && std::fwrite(blocks.data(), 1, total, f) == total
			&& std::fwrite(footer.data(), 1, footer.size(), f) == footer.size();
		std::fclose(f);
		if (!ok)
			AddError("Error writing " + args_.outputFilename);
		return !IsError();
	}
};


class DsrcDecompressorGpu : public IDsrcOperator
{
public:
	bool Process(const InputParameters& args_)
	{
		ClearError();
		std::vector<uint8_t> arc;
		if (!gpuop::ReadWholeFile(args_.inputFilename, arc))
		{
			AddError("Cannot open file to read:" + args_.inputFilename);
			return false;
		}
		if (arc.empty())
		{
			AddError("Empty file.");																// DsrcFile.cpp:193
			return false;
		}
		// DsrcFileReader::ReadFileHeader / ReadFileFooter (src/DsrcFile.cpp:186-314)
		uint64_t n = 0;
		dsrcgpu_dataset_t ds;
		dsrcgpu_settings_t cs;
		if (dsrcgpu_read_archive_index(arc.data(), arc.size(), &n, NULL, NULL, 0, &ds, &cs) != DSRCGPU_OK)
		{
			AddError("Invalid archive or old unsupported version");								// DsrcFile.cpp:205
			return false;
		}
		std::vector<uint64_t> off(n), got(n);
		std::vector<uint32_t> len(n);
		dsrcgpu_read_archive_index(arc.data(), arc.size(), &n, off.data(), len.data(), n, &ds, &cs);
		uint64_t total = 0;
		uint32_t maxChunk = 0;
		for (uint64_t i = 0; i < n; ++i)							// chunkSize + 1 of every block (BlockCompressor.cpp:279,302-308)
		{
			const uint8_t* b = arc.data() + off[i];
			const uint32_t chunk = (((uint32_t)b[12] << 24) | ((uint32_t)b[13] << 16) | ((uint32_t)b[14] << 8) | b[15]) + 1;
			total += chunk;
			maxChunk = chunk > maxChunk ? chunk : maxChunk;
		}
		gpuop::Context c;
		if (dsrcgpu_create(&c.ctx, 0, &ds, &cs, maxChunk + 64, 0) != DSRCGPU_OK)
		{
			AddError("dsrc_b200: no CUDA device or unsupported archive settings");
			return false;
		}
		std::vector<uint8_t> fastq(total + 64);
		if (dsrcgpu_decode_blocks(c.ctx, arc.data(), off.data(), len.data(), (uint32_t)n, fastq.data(), fastq.size(), got.data()) != DSRCGPU_OK)
		{
			AddError(dsrcgpu_last_error(c.ctx));					// with -c archives: "CRC32 checksums mismatch." (src/DsrcWorker.cpp:60)
			return false;
		}
		uint64_t bytes = 0;
		for (uint64_t i = 0; i < n; ++i)
			bytes += got[i];
		FILE* f = std::fopen(args_.outputFilename.c_str(), "wb");
		if (f == NULL)
		{
			AddError("Cannot open file to write:" + args_.outputFilename);
			return false;
		}
		bool ok = std::fwrite(fastq.data(), 1, bytes, f) == bytes;
		std::fclose(f);
		if (!ok)
			AddError("Error writing " + args_.outputFilename);
		return !IsError();
	}
};

} }
