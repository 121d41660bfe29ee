// TEST / BENCH INFRASTRUCTURE. Host-only build of the synthetic FASTQ generator (the record functions of
// dsrc_b200/csrc/synth_records.h compiled with g++): bench.py's reference arm generates its input with this library so that the
// process timing the reference never loads the product library. Same bytes as dsrcgpu_synth_fastq_host / _device.
#include "../dsrc_b200/csrc/synth_records.h"

extern "C" int bench_synth_fastq_host(uint32_t profile, uint64_t seed, uint64_t first_read, uint64_t n_reads, uint8_t* out, uint64_t out_cap, uint64_t* bytes)
{
    if (profile > 2) return -5;
    uint64_t p = 0;
    for (uint64_t i = 0; i < n_reads; ++i) {
        const uint32_t sz = profile == 2 ? syn454_size(seed, first_read + i) : (uint32_t)SYN_REC;
        if (p + sz > out_cap) return -3;
        if (profile == 2) syn454_record(out + p, seed, first_read + i); else syn_record(out + p, profile, seed, first_read + i);
        p += sz;
    }
    if (bytes) *bytes = p;
    return 0;
}
