// Seeded synthetic FASTQ generator on the device (SURVEY.md 8d shapes) -- measurement support, not part of the codec.
// Every record is a pure function of (seed, read index), so any range of reads can be generated anywhere
// (each rank of a multi-GPU run generates its own shard) and the CPU twin in synth_host() gives the same bytes.
//
// Illumina shape, fixed 372-byte records:
//   @SIM.<9 digits> A00123:45:HXXXXXXX:<lane>:<tile 4d>:<x 5d>:<y 5d> 1:N:0:ACGTACGT \n <150 bases> \n + \n <150 quals> \n
// profile 0: 4-level binned qualities {2,12,23,37} (NovaSeq-like), profile 1: 41 levels (HiSeq-like), profile 2: 454 / Ion shape (below).
#include "../../include/dsrc_b200_bench.h"
#include "common.cuh"
#include <cstdlib>

#include "synth_records.h"

__global__ void k_synth454_sizes(u32* sizes, u64 seed, u64 first, u64 n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sizes[i] = syn454_size(seed, first + i);
}
// offsets: per-chunk exclusive sums on the device (one thread per 1024 records), chunk bases on the host -- a generator, not a hot path
__global__ void k_synth454_chunks(const u32* sizes, u64* chunk_sum, u64 n)
{
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c * 1024 >= n) return;
    u64 s = 0;
    for (u64 i = c * 1024; i < n && i < (c + 1) * 1024; ++i) s += sizes[i];
    chunk_sum[c] = s;
}
__global__ void k_synth454_write(u8* out, const u32* sizes, const u64* chunk_base, u64 seed, u64 first, u64 n)
{
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c * 1024 >= n) return;
    u64 off = chunk_base[c];
    for (u64 i = c * 1024; i < n && i < (c + 1) * 1024; ++i) { syn454_record(out + off, seed, first + i); off += sizes[i]; }
}

__global__ void k_synth(u8* out, u32 profile, u64 seed, u64 first, u64 n)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u8 rec[SYN_REC];
    syn_record(rec, profile, seed, first + i);
    u8* o = out + i * SYN_REC;                           // 372 = 4 * 93: records are 4-byte aligned when out is
    if (((uintptr_t)o & 3) == 0) { const u32* s = (const u32*)rec; u32* d = (u32*)o; for (int k = 0; k < SYN_REC / 4; ++k) d[k] = s[k]; }
    else for (int k = 0; k < SYN_REC; ++k) o[k] = rec[k];
}

extern "C" int dsrcgpu_synth_fastq_device(dsrcgpu_ctx* ctx, uint32_t profile, uint64_t seed, uint64_t first_read, uint64_t n_reads,
                                          uint8_t* d_out, uint64_t out_cap, uint64_t* bytes)
{
    (void)ctx;
    if (profile > 2) return DSRCGPU_E_UNSUPPORTED;
    if (profile == 2) {
        if (n_reads == 0) { if (bytes) *bytes = 0; return DSRCGPU_OK; }
        const u64 nch = (n_reads + 1023) / 1024;
        u32* d_sizes = nullptr; u64* d_chunk = nullptr;
        if (cudaMalloc(&d_sizes, n_reads * 4) != cudaSuccess || cudaMalloc(&d_chunk, nch * 8) != cudaSuccess) { cudaFree(d_sizes); return DSRCGPU_E_NOMEM; }
        k_synth454_sizes<<<(unsigned)((n_reads + 255) / 256), 256>>>(d_sizes, seed, first_read, n_reads);
        k_synth454_chunks<<<(unsigned)((nch + 127) / 128), 128>>>(d_sizes, d_chunk, n_reads);
        u64* h_chunk = (u64*)malloc(nch * 8);
        int rc = DSRCGPU_OK;
        if (cudaMemcpy(h_chunk, d_chunk, nch * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = DSRCGPU_E_CUDA;
        u64 total = 0;
        for (u64 c = 0; c < nch && rc == DSRCGPU_OK; ++c) { const u64 sz = h_chunk[c]; h_chunk[c] = total; total += sz; }
        if (rc == DSRCGPU_OK && total > out_cap) rc = DSRCGPU_E_CAPACITY;
        if (rc == DSRCGPU_OK && cudaMemcpy(d_chunk, h_chunk, nch * 8, cudaMemcpyHostToDevice) != cudaSuccess) rc = DSRCGPU_E_CUDA;
        if (rc == DSRCGPU_OK) {
            k_synth454_write<<<(unsigned)((nch + 31) / 32), 32>>>(d_out, d_sizes, d_chunk, seed, first_read, n_reads);
            if (cudaDeviceSynchronize() != cudaSuccess) rc = DSRCGPU_E_CUDA;
        }
        free(h_chunk); cudaFree(d_sizes); cudaFree(d_chunk);
        if (bytes) *bytes = total;
        return rc;
    }
    if (n_reads * SYN_REC > out_cap) return DSRCGPU_E_CAPACITY;
    if (n_reads) k_synth<<<(unsigned)((n_reads + 127) / 128), 128>>>(d_out, profile, seed, first_read, n_reads);
    if (cudaDeviceSynchronize() != cudaSuccess) return DSRCGPU_E_CUDA;
    if (bytes) *bytes = n_reads * SYN_REC;
    return DSRCGPU_OK;
}

// CPU twin (same bytes), for tests and for hosts that want the data without a device round trip
extern "C" int dsrcgpu_synth_fastq_host(uint32_t profile, uint64_t seed, uint64_t first_read, uint64_t n_reads, uint8_t* out, uint64_t out_cap, uint64_t* bytes)
{
    if (profile > 2) return DSRCGPU_E_UNSUPPORTED;
    if (profile == 2) {
        u64 p = 0;
        for (u64 i = 0; i < n_reads; ++i) {
            const u32 sz = syn454_size(seed, first_read + i);
            if (p + sz > out_cap) return DSRCGPU_E_CAPACITY;
            syn454_record(out + p, seed, first_read + i);
            p += sz;
        }
        if (bytes) *bytes = p;
        return DSRCGPU_OK;
    }
    if (n_reads * SYN_REC > out_cap) return DSRCGPU_E_CAPACITY;
    for (u64 i = 0; i < n_reads; ++i) syn_record(out + i * SYN_REC, profile, seed, first_read + i);
    if (bytes) *bytes = n_reads * SYN_REC;
    return DSRCGPU_OK;
}
