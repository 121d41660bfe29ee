// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// Thin extern "C" harness around the UNMODIFIED reference (DSRC 2.02) compiled from
// /root/reference/src where it lies (see oracle/Makefile). Exposes the block-level seam
// comp::BlockCompressor::{Store,Read} (src/BlockCompressor.h:66-73) with arbitrary chunk sizes
// (the reference CLI cannot go below 1 MB blocks, src/main.cpp:300) and the single-/multi-thread
// file operators (src/DsrcOperator.h:93-119). Built into oracle/_ref/libdsrcref.so (git-ignored).
#include "BlockCompressor.h"
#include "DsrcOperator.h"
#include "BitMemory.h"
#include "Buffer.h"
#include <cstring>
#include <algorithm>

using namespace dsrc;

struct RefBc
{
	comp::BlockCompressor* bc;
	core::Buffer* outBuf;
};

extern "C" {

void* ref_bc_create(uint32 qoff, int plusRep, int colorSpace, uint32 dnaOrder, uint32 quaOrder, int lossy, int crc)
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
	RefBc* h = new RefBc;
	h->bc = new comp::BlockCompressor(ds, cs);
	h->outBuf = new core::Buffer(1 << 20);
	return h;
}

void ref_bc_destroy(void* h_)
{
	RefBc* h = (RefBc*)h_;
	delete h->bc;
	delete h->outBuf;
	delete h;
}

// Store one FASTQ chunk (no trailing '\n', as cut by IFastqStreamReader::ReadNextChunk).
// Returns compressed size, or -1 if out_ is too small. raw4/comp4: StreamsInfo sizes
// in enum order Meta, Tag, Dna, Quality (src/Common.h:75-82).
long long ref_bc_store(void* h_, const unsigned char* fastq, unsigned long long size,
					   unsigned char* out_, unsigned long long cap, unsigned long long* raw4, unsigned long long* comp4)
{
	RefBc* h = (RefBc*)h_;
	fq::FastqDataChunk chunk(size + 16);
	std::memcpy(chunk.data.Pointer(), fastq, size);
	// the reference reads title[titleLen] of the last record (one byte past the chunk)
	chunk.data.Pointer()[size] = '\n';
	chunk.size = size;
	core::BitMemoryWriter writer(*h->outBuf);
	fq::StreamsInfo raw, cmp;
	h->bc->Store(writer, raw, cmp, chunk);
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

long long ref_bc_read(void* h_, const unsigned char* in_, unsigned long long size,
					  unsigned char* out_, unsigned long long cap)
{
	RefBc* h = (RefBc*)h_;
	core::Buffer inBuf(size + 16);
	std::memcpy(inBuf.Pointer(), in_, size);
	core::BitMemoryReader reader(inBuf.Pointer(), size);
	fq::FastqDataChunk chunk(1 << 20);
	h->bc->Read(reader, chunk);
	if (chunk.size > cap)
		return -1;
	std::memcpy(out_, chunk.data.Pointer(), chunk.size);
	return (long long)chunk.size;
}

// whole-file operators; threads==1 selects the ST operators exactly as src/main.cpp:61-74 does
int ref_compress_file(const char* in_, const char* out_, int dnaLevel, int quaLevel, int bufMB, int threads, unsigned qoff)
{
	comp::InputParameters p;
	p.inputFilename = in_;
	p.outputFilename = out_;
	p.dnaCompressionLevel = dnaLevel;
	p.qualityCompressionLevel = quaLevel;
	p.fastqBufferSizeMB = bufMB;
	p.threadNum = threads;
	p.qualityOffset = qoff;
	comp::IDsrcOperator* op = (threads == 1) ? (comp::IDsrcOperator*)new comp::DsrcCompressorST()
											 : (comp::IDsrcOperator*)new comp::DsrcCompressorMT();
	bool ok = op->Process(p);
	delete op;
	return ok ? 0 : -1;
}

int ref_compress_file_crc(const char* in_, const char* out_, int dnaLevel, int quaLevel, int bufMB, int threads, unsigned qoff, int crc)
{
	comp::InputParameters p;
	p.inputFilename = in_;
	p.outputFilename = out_;
	p.dnaCompressionLevel = dnaLevel;
	p.qualityCompressionLevel = quaLevel;
	p.fastqBufferSizeMB = bufMB;
	p.threadNum = threads;
	p.qualityOffset = qoff;
	p.calculateCrc32 = crc != 0;
	comp::IDsrcOperator* op = (threads == 1) ? (comp::IDsrcOperator*)new comp::DsrcCompressorST()
											 : (comp::IDsrcOperator*)new comp::DsrcCompressorMT();
	bool ok = op->Process(p);
	delete op;
	return ok ? 0 : -1;
}

int ref_decompress_file(const char* in_, const char* out_, int threads)
{
	comp::InputParameters p;
	p.inputFilename = in_;
	p.outputFilename = out_;
	p.threadNum = threads;
	comp::IDsrcOperator* op = (threads == 1) ? (comp::IDsrcOperator*)new comp::DsrcDecompressorST()
											 : (comp::IDsrcOperator*)new comp::DsrcDecompressorMT();
	bool ok = op->Process(p);
	delete op;
	return ok ? 0 : -1;
}

} // extern "C"
