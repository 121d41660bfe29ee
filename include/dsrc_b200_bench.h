/*
 * dsrc_b200 -- measurement / staging helpers exported by libdsrc_b200.so beside the codec ABI of dsrc_b200.h.
 * None of these replaces a reference interface: device and pinned-host allocation for hosts without a CUDA binding, and the
 * seeded synthetic FASTQ generator of the bench workload (SURVEY.md 8d shapes).
 */
#ifndef DSRC_B200_BENCH_H
#define DSRC_B200_BENCH_H

#include "dsrc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Device allocation helpers so a host language without a CUDA binding can stage resident inputs. */
int dsrcgpu_device_alloc(dsrcgpu_ctx* ctx, uint64_t bytes, void** d_ptr);
int dsrcgpu_device_free(dsrcgpu_ctx* ctx, void* d_ptr);
int dsrcgpu_memcpy_h2d(dsrcgpu_ctx* ctx, void* d_dst, const void* h_src, uint64_t bytes);
int dsrcgpu_memcpy_d2h(dsrcgpu_ctx* ctx, void* h_dst, const void* d_src, uint64_t bytes);
/* pinned host buffers (cudaHostAlloc) for full-speed PCIe transfers */
int dsrcgpu_host_alloc(uint64_t bytes, void** h_ptr);
int dsrcgpu_host_free(void* h_ptr);

/* Seeded synthetic FASTQ generator (SURVEY.md 8d shapes) running on the device; fills d_out with whole records,
 * returns bytes written in *bytes. profile: 0 Illumina 4-level binned, 1 Illumina 41-level, 2 454/Ion variable. */
int dsrcgpu_synth_fastq_device(dsrcgpu_ctx* ctx, uint32_t profile, uint64_t seed, uint64_t first_read, uint64_t n_reads,
                               uint8_t* d_out, uint64_t out_cap, uint64_t* bytes);
/* CPU twin producing the same bytes into host memory (tests; hosts that want the data without a device round trip) */
int dsrcgpu_synth_fastq_host(uint32_t profile, uint64_t seed, uint64_t first_read, uint64_t n_reads,
                             uint8_t* out, uint64_t out_cap, uint64_t* bytes);

#ifdef __cplusplus
}
#endif
#endif
