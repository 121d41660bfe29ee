// Seeded synthetic FASTQ generator on the device (SURVEY.md 8d shapes) -- measurement support, not part of the codec.
// Every record is a pure function of (seed, read index), so any range of reads can be generated anywhere
// (each rank of a multi-GPU run generates its own shard) and the CPU twin in synth_host() gives the same bytes.
//
// Illumina shape, fixed 372-byte records:
//   @SIM.<9 digits> A00123:45:HXXXXXXX:<lane>:<tile 4d>:<x 5d>:<y 5d> 1:N:0:ACGTACGT \n <150 bases> \n + \n <150 quals> \n
// profile 0: 4-level binned qualities {2,12,23,37} (NovaSeq-like), profile 1: 41 levels (HiSeq-like).
#include "../../include/dsrc_b200.h"
#include "common.cuh"

#define SYN_READ_LEN 150
#define SYN_TITLE_LEN 67
#define SYN_REC (SYN_TITLE_LEN + 1 + SYN_READ_LEN + 1 + 2 + SYN_READ_LEN + 1)   // 372

__host__ __device__ inline u64 syn_mix(u64 x)
{
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
struct SynRng { u64 s; __host__ __device__ u64 next() { s += 0x9E3779B97F4A7C15ull; return syn_mix(s); } };

__host__ __device__ inline void syn_put_num(u8* p, u32 v, int digits) { for (int k = digits - 1; k >= 0; --k) { p[k] = (u8)('0' + v % 10); v /= 10; } }

__host__ __device__ inline void syn_record(u8* o, u32 profile, u64 seed, u64 idx)
{
    SynRng rng; rng.s = syn_mix(seed ^ (idx * 0xD1342543DE82EF95ull));
    const u64 h = rng.next();
    // title
    const char* head = "@SIM.";
    for (int k = 0; k < 5; ++k) o[k] = (u8)head[k];
    syn_put_num(o + 5, (u32)(100000000ull + idx % 800000000ull), 9);
    const char* mid = " A00123:45:HXXXXXXX:";
    for (int k = 0; k < 20; ++k) o[14 + k] = (u8)mid[k];
    o[34] = (u8)('1' + ((idx >> 22) & 3)); o[35] = ':';
    syn_put_num(o + 36, 1101 + (u32)((idx >> 14) % 1000), 4); o[40] = ':';
    syn_put_num(o + 41, 10000 + (u32)(h % 22000), 5); o[46] = ':';
    syn_put_num(o + 47, 10000 + (u32)(idx & 16383) * 5 + (u32)((h >> 32) % 5), 5);
    const char* tail = " 1:N:0:ACGTACGT\n";
    for (int k = 0; k < 16; ++k) o[52 + k] = (u8)tail[k];
    u8* seq = o + SYN_TITLE_LEN + 1;
    u8* qua = seq + SYN_READ_LEN + 1 + 2;
    // bases
    for (int j = 0; j < SYN_READ_LEN; j += 32) {
        u64 r = rng.next();
        for (int k = 0; k < 32 && j + k < SYN_READ_LEN; ++k) { seq[j + k] = (u8)"ACGT"[r & 3]; r >>= 2; }
    }
    seq[SYN_READ_LEN] = '\n'; seq[SYN_READ_LEN + 1] = '+'; seq[SYN_READ_LEN + 2] = '\n';
    // qualities: sticky Markov chain
    const u32 nlev = profile == 0 ? 4u : 41u;
    u64 r = rng.next();
    u32 state = (u32)(r % nlev); int have = 0;
    for (int j = 0; j < SYN_READ_LEN; ++j) {
        if (have == 0) { r = rng.next(); have = 8; }
        const u32 b = (u32)(r & 255); r >>= 8; --have;
        if (b < 38) {                                   // ~15 %: move by -2..+2
            int step = (int)(b % 5) - 2; int ns = (int)state + step;
            state = (u32)(ns < 0 ? 0 : (ns >= (int)nlev ? (int)nlev - 1 : ns));
        }
        const u32 q = profile == 0 ? (state == 0 ? 2u : state == 1 ? 12u : state == 2 ? 23u : 37u) : state;
        qua[j] = (u8)(33 + q);
    }
    qua[SYN_READ_LEN] = '\n';
    // '#' tail on ~10 % of the reads, 1..40 long
    const u64 t = rng.next();
    if (t % 10 == 0) { u32 tl = 1 + (u32)((t >> 8) % 40); for (u32 k = 0; k < tl; ++k) qua[SYN_READ_LEN - 1 - k] = '#'; }
    // N at ~2e-3 per base, carrying quality '#'
    const u64 nn = rng.next();
    u32 n_n = (nn & 1023) < 266 ? 1u : 0u; if ((nn & 1023) < 40) n_n = 2;
    for (u32 k = 0; k < n_n; ++k) { u32 p = (u32)((nn >> (16 + 16 * k)) % SYN_READ_LEN); seq[p] = 'N'; qua[p] = '#'; }
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
    if (profile > 1) return DSRCGPU_E_UNSUPPORTED;      // the variable-length 454 shape is generated on the host (tests/synth.py)
    if (n_reads * SYN_REC > out_cap) return DSRCGPU_E_CAPACITY;
    if (n_reads) k_synth<<<(unsigned)((n_reads + 127) / 128), 128>>>(d_out, profile, seed, first_read, n_reads);
    if (cudaDeviceSynchronize() != cudaSuccess) return DSRCGPU_E_CUDA;
    if (bytes) *bytes = n_reads * SYN_REC;
    return DSRCGPU_OK;
}

// CPU twin (same bytes), for tests and for hosts that want the data without a device round trip
extern "C" int dsrcgpu_synth_fastq_host(uint32_t profile, uint64_t seed, uint64_t first_read, uint64_t n_reads, uint8_t* out, uint64_t out_cap, uint64_t* bytes)
{
    if (profile > 1) return DSRCGPU_E_UNSUPPORTED;
    if (n_reads * SYN_REC > out_cap) return DSRCGPU_E_CAPACITY;
    for (u64 i = 0; i < n_reads; ++i) syn_record(out + i * SYN_REC, profile, seed, first_read + i);
    if (bytes) *bytes = n_reads * SYN_REC;
    return DSRCGPU_OK;
}
