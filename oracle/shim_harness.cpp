// TEST INFRASTRUCTURE ONLY. Compiles host/BlockCompressorGpu.h -- the C++ binding of INTEGRATION.md -- against the UNMODIFIED
// reference headers (where they lie under /root/reference/src) and exposes it with the same extern "C" signatures as
// ref_harness.cpp exposes comp::BlockCompressor, so tests/test_gpu_shim.py can drive both through identical calls and compare
// bytes. Built into oracle/_ref/libdsrcshim.so (links ../../dsrc_b200/libdsrc_b200.so; git-ignored, travels to the GPU box).
#include "../host/BlockCompressorGpu.h"
#include "../host/DsrcOperatorGpu.h"
#include "../host/DsrcModuleGpu.h"
#include "Buffer.h"
#include <cstring>

using namespace dsrc;

struct ShimBc
{
	comp::BlockCompressorGpu* bc;
	core::Buffer* outBuf;
};

extern "C" {

void* shim_bc_create(uint32 qoff, int plusRep, int colorSpace, uint32 dnaOrder, uint32 quaOrder, int lossy, int crc, uint32 maxBlock)
{
	fq::FastqDatasetType ds;
	ds.qualityOffset = qoff;
	ds.plusRepetition = plusRep != 0;
	ds.colorSpace = colorSpace != 0;
	comp::CompressionSettings cs;
	cs.dnaOrder = dnaOrder;
	cs.qualityOrder = quaOrder;
	cs.lossy = lossy != 0;
	cs.calculateCrc32 = crc != 0;
	try
	{
		ShimBc* h = new ShimBc;
		h->bc = new comp::BlockCompressorGpu(ds, cs, maxBlock);
		h->outBuf = new core::Buffer(1 << 20);
		return h;
	}
	catch (const DsrcException&)
	{
		return NULL;
	}
}

void shim_bc_destroy(void* h_)
{
	ShimBc* h = (ShimBc*)h_;
	delete h->bc;
	delete h->outBuf;
	delete h;
}

long long shim_bc_store(void* h_, const unsigned char* fastq, unsigned long long size,
						unsigned char* out_, unsigned long long cap, unsigned long long* raw4, unsigned long long* comp4)
{
	ShimBc* h = (ShimBc*)h_;
	fq::FastqDataChunk chunk(size + 16);
	std::memcpy(chunk.data.Pointer(), fastq, size);
	chunk.data.Pointer()[size] = '\n';
	chunk.size = size;
	core::BitMemoryWriter writer(*h->outBuf);
	fq::StreamsInfo raw, cmp;
	try { h->bc->Store(writer, raw, cmp, chunk); } catch (const DsrcException&) { return -2; }
	writer.Flush();
	unsigned long long n = writer.Position();
	for (int i = 0; i < 4; ++i)
	{
		if (raw4) raw4[i] = raw.sizes[i];
		if (comp4) comp4[i] = cmp.sizes[i];
	}
	if (n > cap)
		return -1;
	std::memcpy(out_, writer.Pointer(), n);
	return (long long)n;
}

long long shim_bc_read(void* h_, const unsigned char* in_, unsigned long long size, unsigned char* out_, unsigned long long cap)
{
	ShimBc* h = (ShimBc*)h_;
	core::Buffer inBuf(size + 16);
	std::memcpy(inBuf.Pointer(), in_, size);
	core::BitMemoryReader reader(inBuf.Pointer(), size);
	fq::FastqDataChunk chunk(1 << 20);
	try { h->bc->Read(reader, chunk); } catch (const DsrcException&) { return -2; }
	if (chunk.size > cap)
		return -1;
	std::memcpy(out_, chunk.data.Pointer(), chunk.size);
	return (long long)chunk.size;
}

// whole-file operators, same signatures as ref_compress_file_crc / ref_decompress_file of ref_harness.cpp
int shim_compress_file(const char* in_, const char* out_, int dnaLevel, int quaLevel, int bufMB, unsigned qoff, int crc, char* err, int errCap)
{
	comp::InputParameters p;
	p.inputFilename = in_;
	p.outputFilename = out_;
	p.dnaCompressionLevel = dnaLevel;
	p.qualityCompressionLevel = quaLevel;
	p.fastqBufferSizeMB = bufMB;
	p.qualityOffset = qoff;
	p.calculateCrc32 = crc != 0;
	comp::DsrcCompressorGpu op;
	bool ok = op.Process(p);
	if (err && errCap > 0) { std::strncpy(err, op.GetError().c_str(), errCap - 1); err[errCap - 1] = 0; }
	return ok ? 0 : -1;
}

int shim_decompress_file(const char* in_, const char* out_, char* err, int errCap)
{
	comp::InputParameters p;
	p.inputFilename = in_;
	p.outputFilename = out_;
	comp::DsrcDecompressorGpu op;
	bool ok = op.Process(p);
	if (err && errCap > 0) { std::strncpy(err, op.GetError().c_str(), errCap - 1); err[errCap - 1] = 0; }
	return ok ? 0 : -1;
}

// wrap::DsrcModule surface: Configurable setters, then Compress / Decompress; returns 0 or -1 with the exception text
int shim_module_roundtrip(const char* fastq_, const char* archive_, const char* back_, int dnaLevel, int quaLevel, int bufMB, int crc,
						  char* err, int errCap)
{
	try
	{
		wrap::DsrcModuleGpu m;
		m.SetDnaCompressionLevel(dnaLevel);
		m.SetQualityCompressionLevel(quaLevel);
		m.SetFastqBufferSizeMB(bufMB);
		m.SetCrc32Checking(crc != 0);
		m.Compress(fastq_, archive_);
		m.Decompress(archive_, back_);
		return 0;
	}
	catch (const DsrcException& e)
	{
		if (err && errCap > 0) { std::strncpy(err, e.what(), errCap - 1); err[errCap - 1] = 0; }
		return -1;
	}
}

} // extern "C"
