// DsrcCompressorGpu / DsrcDecompressorGpu -- IDsrcOperator implementations (reference src/DsrcOperator.h:28-119) over dsrc_b200.
//
// Where the reference's DsrcCompressorMT::Process (src/DsrcOperator.cpp:230-394) starts a reader thread, `threadNum` workers each
// owning a BlockCompressor, and a writer, these operators stream the file through a bounded window and hand every window's block
// queue to the library in one call (INTEGRATION.md binding 2): the GPU's stream scheduler is the worker pool. Same interface
// (`bool Process(const InputParameters&)`, IsError/GetError), same InputParameters, same .dsrc bytes as `dsrc c -t1` --
// main.cpp:61-74 only has to pick these classes. Needs the reference's headers (IDsrcOperator, InputParameters,
// CompressionSettings); built by oracle/Makefile target `shim`, checked against the reference's own operators by
// tests/test_gpu_shim.py.
//
// Host memory is bounded by the window (default 1 GiB of FASTQ, at least 4 chunk buffers; DSRCGPU_WINDOW_MB overrides) whatever the
// file size: the reference streams `-b`-sized chunks through its pools (src/DsrcOperator.cpp:261-338), this streams windows of
// them. Block counts are 64-bit; the device is DSRCGPU_DEVICE (default 0) or the constructor's argument.
#pragma once
#include "DsrcOperator.h"
#include "dsrc_b200.h"
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace dsrc { namespace comp {

namespace gpuop {

struct Context													// dsrcgpu_ctx with scope
{
	dsrcgpu_ctx* ctx;
	Context() : ctx(NULL) {}
	~Context() { if (ctx != NULL) dsrcgpu_destroy(ctx); }
};

struct File														// FILE* with scope and 64-bit offsets
{
	FILE* f;
	File() : f(NULL) {}
	~File() { if (f != NULL) std::fclose(f); }
	bool Open(const std::string& name_, const char* mode_) { f = std::fopen(name_.c_str(), mode_); return f != NULL; }
	uint64_t Size() { fseeko(f, 0, SEEK_END); const uint64_t n = (uint64_t)ftello(f); fseeko(f, 0, SEEK_SET); return n; }
	bool ReadAt(uint64_t pos_, void* p_, uint64_t n_) { return fseeko(f, (off_t)pos_, SEEK_SET) == 0 && std::fread(p_, 1, n_, f) == n_; }
	bool Write(const void* p_, uint64_t n_) { return std::fwrite(p_, 1, n_, f) == n_; }
};

inline int DeviceFromEnv(int device_)
{
	if (device_ >= 0)
		return device_;
	const char* e = std::getenv("DSRCGPU_DEVICE");
	return e != NULL ? std::atoi(e) : 0;
}

inline uint64_t WindowBytes(uint64_t cbuf_)
{
	uint64_t w = 1ull << 30;
	if (const char* e = std::getenv("DSRCGPU_WINDOW_MB"))
		w = (uint64_t)std::atoll(e) << 20;
	return w < 4 * cbuf_ ? 4 * cbuf_ : w;
}

// dsrcgpu_encode_blocks with the output buffer grown on DSRCGPU_E_CAPACITY: tiny blocks with many text fields can come out larger
// than 1.5x their input (hundreds of Huffman trees in the tag header)
inline int EncodeGrowing(dsrcgpu_ctx* ctx_, const uint8_t* fastq_, const uint64_t* off_, const uint32_t* len_, const uint32_t* cap_,
						 uint32_t n_, std::vector<uint8_t>& out_, uint32_t* size_, uint64_t* raw_, uint64_t* cmp_)
{
	int rc = DSRCGPU_E_CAPACITY;
	for (int attempt = 0; attempt < 4 && rc == DSRCGPU_E_CAPACITY; ++attempt)
	{
		if (attempt > 0)
			out_.resize(out_.size() * 4);
		rc = dsrcgpu_encode_blocks(ctx_, fastq_, off_, len_, cap_, n_, out_.data(), out_.size(), size_, raw_, cmp_);
	}
	return rc;
}

} // namespace gpuop


class DsrcCompressorGpu : public IDsrcOperator
{
public:
	explicit DsrcCompressorGpu(int device_ = -1) : device(gpuop::DeviceFromEnv(device_)) {}

	bool Process(const InputParameters& args_)
	{
		ClearError();
		if (args_.useFastqStdIo || args_.lossyCompression || args_.tagPreserveFlags != 0)
		{
			AddError("stdin/stdout, lossy mode and field filtering are outside the GPU operator's envelope");
			return false;
		}
		gpuop::File in, out;
		if (!in.Open(args_.inputFilename, "rb"))
		{
			AddError("Cannot open file to read:" + args_.inputFilename);						// FileStream.cpp:78
			return false;
		}
		const uint64_t fileSize = in.Size();
		const uint64_t cbuf = (uint64_t)args_.fastqBufferSizeMB << 20;
		const uint64_t window = gpuop::WindowBytes(cbuf);
		const CompressionSettings settings = GetCompressionSettings(args_);
		dsrcgpu_dataset_t ds = { args_.qualityOffset, 0, 0 };
		dsrcgpu_settings_t cs = { settings.dnaOrder, settings.qualityOrder, 0, 0, (uint8_t)settings.calculateCrc32 };

		gpuop::Context c;
		std::vector<uint8_t> fastq, blocks, back;
		std::vector<uint64_t> off, boff, got;
		std::vector<uint32_t> len, cap, size;
		std::vector<uint32_t> allSizes;							// one u32 per block of the file: the footer (src/DsrcFile.cpp:142)
		uint32_t capacity = 0, readerState = 0;					// TagStats::fields capacity, usesCrlf: the state that crosses blocks
		uint64_t pos = 0, total = 0;
		bool headerWritten = false;
		uint8_t header[40] = { 0 };

		while (pos < fileSize || (fileSize == 0 && !headerWritten))
		{
			const uint64_t w = fileSize - pos < window ? fileSize - pos : window;
			const bool last = pos + w == fileSize;
			fastq.resize(w);
			if (w > 0 && !in.ReadAt(pos, fastq.data(), w))
			{
				AddError("Error reading " + args_.inputFilename);
				return false;
			}
			// the block queue IFastqStreamReader::ReadNextChunk would produce from here (src/FastqStream.cpp:18-98)
			uint32_t st = readerState;
			const uint64_t k = dsrcgpu_cut_blocks_window(fastq.data(), w, cbuf, NULL, NULL, 0, &st);
			off.resize(k + 1); len.resize(k + 1);
			dsrcgpu_cut_blocks_window(fastq.data(), w, cbuf, off.data(), len.data(), k, &readerState);
			const uint64_t n = last ? k : k - 1;				// a window's last block was cut at the window's end: it opens the next window
			if (!headerWritten)
			{
				// FastqFileReader::AnalyzeFirstChunk (src/FastqIo.cpp:26-44): quality offset, '+' repetition, colour space
				if (k == 0 || dsrcgpu_analyze_first_chunk(fastq.data() + off[0], len[0], &ds) != DSRCGPU_OK || ds.color_space)
				{
					AddError("Error analyzing FASTQ dataset");										// DsrcOperator.cpp:284
					return false;
				}
				if (dsrcgpu_create(&c.ctx, device, &ds, &cs, (uint32_t)cbuf + 64, 0) != DSRCGPU_OK)
				{
					AddError("dsrc_b200: no CUDA device or unsupported settings");
					return false;
				}
				if (!out.Open(args_.outputFilename, "wb"))
				{
					AddError("Cannot open file to write:" + args_.outputFilename);				// FileStream.cpp:142
					return false;
				}
				if (!out.Write(header, 40))						// placeholder: block count and footer position are known at the end
				{
					AddError("Error writing " + args_.outputFilename);
					return false;
				}
				headerWritten = true;
			}
			if (n == 0)
			{
				AddError("Error analyzing FASTQ dataset");		// a record longer than the window: not FASTQ the reference could read either
				return false;
			}
			// one compressor in file order == `dsrc c -t1`: capacity of TagStats::fields before every block
			cap.resize(n); size.resize(n);
			uint64_t bound = 0;
			for (uint64_t i = 0; i < n; ++i)
			{
				const uint8_t* p = fastq.data() + off[i];
				uint32_t tl = 0;
				while (tl < len[i] && p[tl] != '\n' && p[tl] != '\r')
					++tl;
				cap[i] = capacity;
				capacity = dsrcgpu_tag_capacity_after(capacity, dsrcgpu_tag_field_count(p, tl));
				bound += (uint64_t)len[i] + len[i] / 2 + 4096;
			}
			if (blocks.size() < bound)
				blocks.resize(bound);
			if (gpuop::EncodeGrowing(c.ctx, fastq.data(), off.data(), len.data(), cap.data(), (uint32_t)n, blocks, size.data(), NULL, NULL) != DSRCGPU_OK)
			{
				AddError(dsrcgpu_last_error(c.ctx));
				return false;
			}
			uint64_t bytes = 0;
			for (uint64_t i = 0; i < n; ++i)
				bytes += size[i];
			if (args_.calculateCrc32)								// DsrcOperator.cpp:113-123: verify by decoding what was written
			{
				boff.resize(n); got.resize(n);
				uint64_t p = 0, outBytes = 0;
				for (uint64_t i = 0; i < n; ++i) { boff[i] = p; p += size[i]; outBytes += (uint64_t)len[i] + 2; }
				back.resize(outBytes + 64);
				if (dsrcgpu_decode_blocks(c.ctx, blocks.data(), boff.data(), size.data(), (uint32_t)n, back.data(), back.size(), got.data()) != DSRCGPU_OK)
				{
					AddError("CRC32 checksums mismatch.");
					return false;
				}
			}
			if (!out.Write(blocks.data(), bytes))				// DsrcFileWriter::WriteNextChunk: blocks back to back from offset 40
			{
				AddError("Error writing " + args_.outputFilename);
				return false;
			}
			allSizes.insert(allSizes.end(), size.begin(), size.begin() + n);
			total += bytes;
			pos = last ? fileSize : pos + off[n];
			if (fileSize == 0)
				break;
		}
		if (!headerWritten)
		{
			AddError("Error analyzing FASTQ dataset");
			return false;
		}
		// DsrcFileWriter::WriteFileFooter, then the header over its placeholder (src/DsrcFile.cpp:112-170)
		const uint64_t nBlocks = allSizes.size();
		std::vector<uint8_t> footer(dsrcgpu_archive_footer_size(nBlocks));
		dsrcgpu_write_archive_header(header, nBlocks, total);
		dsrcgpu_write_archive_footer(footer.data(), footer.size(), allSizes.data(), nBlocks, &ds, &cs);
		bool ok = out.Write(footer.data(), footer.size()) && fseeko(out.f, 0, SEEK_SET) == 0 && out.Write(header, 40);
		if (!ok)
			AddError("Error writing " + args_.outputFilename);
		return !IsError();
	}

private:
	int device;
};


class DsrcDecompressorGpu : public IDsrcOperator
{
public:
	explicit DsrcDecompressorGpu(int device_ = -1) : device(gpuop::DeviceFromEnv(device_)) {}

	bool Process(const InputParameters& args_)
	{
		ClearError();
		gpuop::File in, out;
		if (!in.Open(args_.inputFilename, "rb"))
		{
			AddError("Cannot open file to read:" + args_.inputFilename);
			return false;
		}
		const uint64_t fileSize = in.Size();
		if (fileSize == 0)
		{
			AddError("Empty file.");																// DsrcFile.cpp:193
			return false;
		}
		// DsrcFileReader::ReadFileHeader / ReadFileFooter (src/DsrcFile.cpp:186-314): 40-byte header, then the footer it points at
		uint8_t header[40];
		uint64_t n = 0, footOff = 0, footBytes = 0;
		dsrcgpu_dataset_t ds;
		dsrcgpu_settings_t cs;
		if (fileSize < 40 || !in.ReadAt(0, header, 40) || dsrcgpu_archive_footer_span(header, &footOff, &footBytes) != DSRCGPU_OK
			|| dsrcgpu_read_archive_footer(header, NULL, 0, fileSize, &n, NULL, 0, NULL, NULL) != DSRCGPU_OK)
		{
			AddError("Invalid archive or old unsupported version");								// DsrcFile.cpp:205
			return false;
		}
		std::vector<uint8_t> footer(footBytes);
		std::vector<uint32_t> len(n);
		if (!in.ReadAt(footOff, footer.data(), footBytes)
			|| dsrcgpu_read_archive_footer(header, footer.data(), footBytes, fileSize, &n, len.data(), n, &ds, &cs) != DSRCGPU_OK)
		{
			AddError("Invalid archive or old unsupported version");
			return false;
		}
		if (!out.Open(args_.outputFilename, "wb"))
		{
			AddError("Cannot open file to write:" + args_.outputFilename);
			return false;
		}
		gpuop::Context c;
		const uint64_t window = gpuop::WindowBytes(1u << 20) / 4;		// compressed bytes per window (the FASTQ is ~4-6x that)
		std::vector<uint8_t> arc, fastq;
		std::vector<uint64_t> off, got;
		uint64_t pos = 40;
		for (uint64_t b0 = 0; b0 < n;)
		{
			uint64_t b1 = b0, bytes = 0;
			while (b1 < n && (b1 == b0 || bytes + len[b1] <= window) && b1 - b0 < 0x7FFFFFFFu) { bytes += len[b1]; ++b1; }
			arc.resize(bytes);
			if (!in.ReadAt(pos, arc.data(), bytes))
			{
				AddError("Error reading " + args_.inputFilename);
				return false;
			}
			const uint64_t k = b1 - b0;
			off.resize(k); got.resize(k);
			uint64_t p = 0, outBytes = 0;
			uint32_t maxChunk = 0;
			for (uint64_t i = 0; i < k; ++i)						// chunkSize + 1 of every block (BlockCompressor.cpp:279,302-308)
			{
				off[i] = p;
				const uint8_t* b = arc.data() + p;
				const uint32_t chunk = len[b0 + i] >= 16 ? (((uint32_t)b[12] << 24) | ((uint32_t)b[13] << 16) | ((uint32_t)b[14] << 8) | b[15]) + 1 : 1;
				outBytes += chunk;
				maxChunk = chunk > maxChunk ? chunk : maxChunk;
				p += len[b0 + i];
			}
			if (c.ctx == NULL && dsrcgpu_create(&c.ctx, device, &ds, &cs, (maxChunk > (1u << 20) ? maxChunk : (1u << 20)) + 64, 0) != DSRCGPU_OK)
			{
				AddError("dsrc_b200: no CUDA device or unsupported archive settings");
				return false;
			}
			fastq.resize(outBytes + 64);
			if (dsrcgpu_decode_blocks(c.ctx, arc.data(), off.data(), len.data() + b0, (uint32_t)k, fastq.data(), fastq.size(), got.data()) != DSRCGPU_OK)
			{
				AddError(dsrcgpu_last_error(c.ctx));				// with -c archives: "CRC32 checksums mismatch." (src/DsrcWorker.cpp:60)
				return false;
			}
			uint64_t w = 0;
			for (uint64_t i = 0; i < k; ++i)
				w += got[i];
			if (!out.Write(fastq.data(), w))
			{
				AddError("Error writing " + args_.outputFilename);
				return false;
			}
			pos += bytes;
			b0 = b1;
		}
		return !IsError();
	}

private:
	int device;
};

} }
