// Record functions of the seeded synthetic FASTQ generator (SURVEY.md 8d shapes) -- measurement support, not part of the codec.
// Plain C++ so that the same bytes can be produced by the device kernels (synth.cu, nvcc) and by a host-only build
// (oracle/bench_synth.cpp, g++: the reference arm of bench.py must not load the product library).
// Every record is a pure function of (seed, read index).
//
// Illumina shape, fixed 372-byte records:
//   @SIM.<9 digits> A00123:45:HXXXXXXX:<lane>:<tile 4d>:<x 5d>:<y 5d> 1:N:0:ACGTACGT \n <150 bases> \n + \n <150 quals> \n
// profile 0: 4-level binned qualities {2,12,23,37} (NovaSeq-like), profile 1: 41 levels (HiSeq-like), profile 2: 454 / Ion shape (below).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define SYN_HD __host__ __device__
#else
#define SYN_HD
#endif
typedef uint8_t u8; typedef uint32_t u32; typedef uint64_t u64;

#define SYN_READ_LEN 150
#define SYN_TITLE_LEN 67
#define SYN_REC (SYN_TITLE_LEN + 1 + SYN_READ_LEN + 1 + 2 + SYN_READ_LEN + 1)   // 372

SYN_HD inline u64 syn_mix(u64 x)
{
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
struct SynRng { u64 s; SYN_HD u64 next() { s += 0x9E3779B97F4A7C15ull; return syn_mix(s); } };

SYN_HD inline void syn_put_num(u8* p, u32 v, int digits) { for (int k = digits - 1; k >= 0; --k) { p[k] = (u8)('0' + v % 10); v /= 10; } }

SYN_HD inline void syn_record(u8* o, u32 profile, u64 seed, u64 idx)
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

// ---- 454 / Ion-Torrent shape (profile 2, SURVEY.md 8d, BASELINE configs[2]): variable-length reads, lengths ~N(350, 90) clipped to
// [40, 600], homopolymer-biased bases, qualities falling from ~40 to ~8 along the read with noise (about 45 distinct values), ambiguity
// codes N / R / W / S at 4e-3 per base -- N mostly with q < 7 (it moves into the quality byte, src/RecordsProcessor.cpp:230-233), the
// others with q >= 7 (they stay in the DNA stream: the 8-symbol order-7 model) -- titles
//   @<14 alnum> rank=<7 digits> x=<1..4095> y=<1..4095> length=<L>
// A record is a pure function of (seed, read index); its size is too, so a first pass sizes the records and a scan places them.
#define SYN454_MAX_REC 1300
SYN_HD inline u32 syn_digits(u32 v) { u32 d = 1; while (v >= 10) { v /= 10; ++d; } return d; }
struct Syn454Head { u32 len, x, y; u64 name; };
SYN_HD inline Syn454Head syn454_head(SynRng& rng)
{
    Syn454Head h;
    const u64 a = rng.next();
    // sum of four 16-bit uniforms: mean 2 * 65535, sigma 65535 / sqrt(3) -> N(350, 90)
    const int sum = (int)(a & 0xFFFF) + (int)((a >> 16) & 0xFFFF) + (int)((a >> 32) & 0xFFFF) + (int)(a >> 48);
    int L = 350 + (int)(((long long)(sum - 131070) * 90) / 37837);
    h.len = (u32)(L < 40 ? 40 : (L > 600 ? 600 : L));
    const u64 b = rng.next();
    h.x = 1 + (u32)(b % 4095); h.y = 1 + (u32)((b >> 32) % 4095);
    h.name = rng.next();
    return h;
}
SYN_HD inline u32 syn454_size(u64 seed, u64 idx)
{
    SynRng rng; rng.s = syn_mix(seed ^ (idx * 0xD1342543DE82EF95ull) ^ 0x454ull);
    const Syn454Head h = syn454_head(rng);
    return 1 + 14 + 6 + 7 + 3 + syn_digits(h.x) + 3 + syn_digits(h.y) + 8 + syn_digits(h.len) + 1 + h.len + 3 + h.len + 1;
}
SYN_HD inline u32 syn454_record(u8* o, u64 seed, u64 idx)
{
    SynRng rng; rng.s = syn_mix(seed ^ (idx * 0xD1342543DE82EF95ull) ^ 0x454ull);
    const Syn454Head h = syn454_head(rng);
    u32 p = 0;
    o[p++] = '@';
    { u64 nm = h.name; for (int k = 0; k < 14; ++k) { const u32 c = (u32)(nm % 36); nm /= 36; o[p++] = (u8)(c < 26 ? 'A' + c : '0' + (c - 26)); } }
    { const char* t = " rank="; for (int k = 0; k < 6; ++k) o[p++] = (u8)t[k]; }
    syn_put_num(o + p, (u32)((idx + 1) % 10000000ull), 7); p += 7;
    { const char* t = " x="; for (int k = 0; k < 3; ++k) o[p++] = (u8)t[k]; }
    { const u32 d = syn_digits(h.x); syn_put_num(o + p, h.x, (int)d); p += d; }
    { const char* t = " y="; for (int k = 0; k < 3; ++k) o[p++] = (u8)t[k]; }
    { const u32 d = syn_digits(h.y); syn_put_num(o + p, h.y, (int)d); p += d; }
    { const char* t = " length="; for (int k = 0; k < 8; ++k) o[p++] = (u8)t[k]; }
    { const u32 d = syn_digits(h.len); syn_put_num(o + p, h.len, (int)d); p += d; }
    o[p++] = '\n';
    const u32 L = h.len;
    u8* seq = o + p; u8* qua = seq + L + 3;
    u32 prev = 0;
    for (u32 j = 0; j < L; ++j) {
        const u64 r = rng.next();
        // base: 30 % repeat the previous one (homopolymers)
        u32 b = (u32)(r & 3);
        if (j && ((r >> 2) & 1023) < 307) b = prev;
        prev = b;
        u8 c = (u8)"ACGT"[b];
        // quality: 40 -> 8 along the read + noise (sum of four 4-bit uniforms - 30, sigma ~ 4), clipped to [0, 44]
        int q = 40 - (int)((32 * j) / (L > 1 ? L - 1 : 1));
        q += ((int)((r >> 12) & 15) + (int)((r >> 16) & 15) + (int)((r >> 20) & 15) + (int)((r >> 24) & 15) - 30) * 7 / 16;
        q = q < 0 ? 0 : (q > 44 ? 44 : q);
        // ambiguity codes at ~4e-3 per base
        if (((r >> 28) & 4095) < 16) {
            const u32 k = (u32)((r >> 40) % 6);
            c = (u8)"NNNRWS"[k];
            if (k < 3 && ((r >> 44) & 15) < 13) q = (int)((r >> 48) % 7);      // N, mostly below the transfer threshold
            else if (q < 7) q = 7;
        }
        seq[j] = c; qua[j] = (u8)(33 + q);
    }
    seq[L] = '\n'; seq[L + 1] = '+'; seq[L + 2] = '\n';
    qua[L] = '\n';
    return p + L + 3 + L + 1;
}
