// C ABI of dsrc_b200 (include/dsrc_b200.h): context, HBM layout, the CUDA-stream block scheduler.
//
// Replaces DsrcCompressor::Process / DsrcDecompressor::Process (src/DsrcWorker.cpp:30-108): instead of N CPU
// threads each owning a BlockCompressor, the block queue is cut into batches; each batch is laid out in HBM
// (dense per-batch arrays + persistent per-CTA arenas) and pushed through the per-block kernels on a stream.
#include "../../include/dsrc_b200.h"
#include "../../include/dsrc_b200_bench.h"
#include "common.cuh"
#include "kernels.h"
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <functional>
#include <time.h>

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return DSRCGPU_E_CUDA; } } while (0)

static unsigned long long g_devbuf_allocs = 0;       // device allocations made so far (a call that allocated says nothing about steady-state timing)
static thread_local unsigned t_alloc_scale = 1;      // a host-buffer call with half-size batches allocates for full-size ones: a later switch does not reallocate
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        ++g_devbuf_allocs;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes * t_alloc_scale + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess && t_alloc_scale > 1) {           // the head-room is a convenience, not a requirement
            cudaGetLastError();
            want = bytes + bytes / 8 + 256;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

enum { K_COUNT, K_PARSE, K_PREP, K_TAGS, K_MODEL_Q, K_MODEL_D, K_RC, K_Q0, K_D0, K_SIZES, K_GATHER, K_DECODE, K_DEC_TAGS, K_DEC_Q, K_DEC_D, K_DEC_ASM, K_CRC, K_H2D, K_D2H, K_NUM };
static const char* K_NAMES[K_NUM] = {"count_lines", "parse", "preprocess", "tags", "model_quality", "model_dna", "rc_encode",
                                     "q0_quality", "d0_dna", "meta_sizes", "gather", "decode_probe", "decode_tags", "decode_quality", "decode_dna", "decode_assemble", "crc32",
                                     "copy_h2d", "copy_d2h"};      // the last two: payload copies of the host-buffer calls (not kernels)
#define MAX_SLOTS 8

// One in-flight batch of blocks: its own stream, workspace and pinned staging. The scheduler keeps several slots busy so
// that the copies of one batch, the parallel kernels of the next and the serial range-coder chains of a third overlap.
struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t stream_r = nullptr;                         // serial stage of the batch (range-coder chains, sizes, gather, output copy): high priority
    cudaEvent_t ev_pdone = nullptr;                          // the batch's parallel stage (everything up to the model kernels) has finished
    DevBuf in, desc, state, result, probe, lines, qcat, dcat, trip_q, trip_d, ftab, streams, out;
    DevBuf r_title_off, r_seq_off, r_qua_off, r_title_len, r_qua_len, r_dna_len, r_trunc_len, r_qcat_off, r_dcat_off;
    DevBuf elem_a, elem_b, tagpool, q0_arena, queue;        // per-CTA arenas of the persistent kernels, block queue counters
    BlockDesc* h_desc = nullptr; BlockResult* h_result = nullptr; BlockProbe* h_probe = nullptr; u32 h_cap = 0;
    cudaEvent_t ev_results = nullptr, ev_sizes = nullptr, ev_probe = nullptr;
    bool busy = false; u32 first = 0, cnt = 0, batch = 0;
    bool r_enq = false;                                      // the batch's serial stage (chains, sizes, gather, result copy) has been enqueued
    u64 r_out_base = 0; u64* r_cursor = nullptr;             // arguments of the size scan, kept until then
    Workspace ws{};                                          // of the batch in flight (re-used if its output staging has to grow)
    std::vector<u64> offs;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_used;
    void release()
    {
        DevBuf* bufs[] = {&in, &desc, &state, &result, &probe, &lines, &qcat, &dcat, &trip_q, &trip_d, &ftab, &streams, &out, &r_title_off, &r_seq_off,
                          &r_qua_off, &r_title_len, &r_qua_len, &r_dna_len, &r_trunc_len, &r_qcat_off, &r_dcat_off, &elem_a, &elem_b, &tagpool, &q0_arena, &queue};
        for (DevBuf* b : bufs) b->release();
        if (h_desc) cudaFreeHost(h_desc);
        if (h_result) cudaFreeHost(h_result);
        if (h_probe) cudaFreeHost(h_probe);
        h_desc = nullptr; h_result = nullptr; h_probe = nullptr; h_cap = 0;
        if (ev_results) cudaEventDestroy(ev_results);
        if (ev_sizes) cudaEventDestroy(ev_sizes);
        if (ev_probe) cudaEventDestroy(ev_probe);
        if (ev_pdone) cudaEventDestroy(ev_pdone);
        if (stream) cudaStreamDestroy(stream);
        if (stream_r) cudaStreamDestroy(stream_r);
        ev_results = ev_sizes = ev_probe = ev_pdone = nullptr; stream = nullptr; stream_r = nullptr;
    }
};

struct dsrcgpu_ctx {
    int device = 0, sms = 148;
    dsrcgpu_dataset_t ds{};
    dsrcgpu_settings_t cs{};
    u32 max_block = 0, max_inflight = 0;
    std::string err;
    Slot slots[MAX_SLOTS]; int n_slots = 1;
    DevBuf prof, cursor, dec_arena;
    // host-buffer calls: the next batch's input travels on a stream of its own into a spare staging buffer while the host still waits
    // for that batch's slot, so the input link never idles behind a slot; the buffers are swapped when the slot is free
    DevBuf spare_in[2]; cudaStream_t copy_stream = nullptr; cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    DevBuf dec_out2;                                 // decode with host buffers: second output staging buffer (a batch's output copy runs beside the next batch's chains)
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> copy_ev;     // timers of those copies (collected at the end of the call)
    DevBuf tab, tab_mask; u32 tab_count = 0;         // pool of adaptive-row tables shared by all slots (rc_model.cu: tab_acquire)
    u64 tab_stride = 0;
    u32 model_ctas = 0; u64 model_stride = 0; u32 tag_ctas = 0; u32 q0_ctas = 0; u64 q0_stride = 0;
    // per-kernel timing
    std::vector<cudaEvent_t> ev_pool;
    float k_ms[K_NUM]; u32 k_launches[K_NUM];
    cudaEvent_t call_a = nullptr, call_b = nullptr; float call_ms = 0;
    bool profiling = true, phase_prof = false;
    u32 narrow_div = 0;                              // test switch DSRCGPU_NARROW_STREAMS=<d>: narrow arenas of 1/d byte per symbol (forces the retry)
    bool host_full = false;                          // host-buffer calls use full-size batches (the last one was bound by the coding, not by the link)
    bool host_call = false;                          // the call in progress takes host buffers: small transfers go by kernel over mapped memory
    bool last_overflow = false;                      // the last call failed with ST_OVERFLOW
    bool wide_streams = false;                       // range-coder stream arenas at 3 bytes per symbol (set after a chain ran out of its 1.25)
    int rc_group_host = 1;                           // the same for host-buffer calls (DSRCGPU_RC_GROUP_HOST)
    int rc_group = 2;                                // batches whose range-coder chains share one launch (<= n_slots, RC_GROUP_MAX)
    int p_serial = 2;                                // where a batch waits for the previous batch's parallel stage: 0 nowhere, 1 before parse, 2 before the model kernels
};
static inline cudaStream_t rstream(const Slot& sl) { return sl.stream_r ? sl.stream_r : sl.stream; }

static cudaEvent_t get_event(dsrcgpu_ctx* ctx)
{
    if (!ctx->ev_pool.empty()) { cudaEvent_t e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct KTimer {
    dsrcgpu_ctx* ctx; Slot* sl; int k; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    KTimer(dsrcgpu_ctx* c, Slot* s, int kid, cudaStream_t on = nullptr, u32 kernels = 1) : ctx(c), sl(s), k(kid), st(on ? on : s->stream)
    {
        ctx->k_launches[k] += kernels;                 // kernel launches the timer brackets
        if (ctx->profiling) { a = get_event(ctx); b = get_event(ctx); cudaEventRecord(a, st); }
    }
    ~KTimer() { if (ctx->profiling) { cudaEventRecord(b, st); sl->ev_used.push_back({k, {a, b}}); } }
};
static void collect_times(dsrcgpu_ctx* ctx, Slot* sl)
{
    static FILE* tl = nullptr;                        // developer aid: DSRCGPU_TIMELINE=<file> dumps every kernel's start/end on its stream
    static bool tl_checked = false;
    if (!tl_checked) { tl_checked = true; if (const char* e = getenv("DSRCGPU_TIMELINE")) tl = fopen(e, "w"); }
    for (auto& u : sl->ev_used) {
        float ms = 0; cudaEventElapsedTime(&ms, u.second.first, u.second.second);
        if (tl && ctx->call_a) {
            float t0 = 0; cudaEventElapsedTime(&t0, ctx->call_a, u.second.first);
            fprintf(tl, "%d,%s,%.3f,%.3f\n", (int)(sl - ctx->slots), K_NAMES[u.first], t0, t0 + ms); fflush(tl);
        }
        ctx->k_ms[u.first] += ms;
        ctx->ev_pool.push_back(u.second.first); ctx->ev_pool.push_back(u.second.second);
    }
    sl->ev_used.clear();
}

extern "C" const char* dsrcgpu_last_error(dsrcgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

extern "C" int dsrcgpu_create(dsrcgpu_ctx** out, int device, const dsrcgpu_dataset_t* dataset, const dsrcgpu_settings_t* settings,
                              uint32_t max_block_bytes, uint32_t max_inflight_blocks)
{
    if (!out || !dataset || !settings) return DSRCGPU_E_ARG;
    *out = nullptr;
    if (dataset->color_space || settings->lossy || settings->tag_preserve_flags) return DSRCGPU_E_ARG;
    if (dataset->quality_offset < 33 || dataset->quality_offset > 64) return DSRCGPU_E_ARG;
    if (settings->dna_order > 9 || settings->quality_order > 2) return DSRCGPU_E_ARG;
    if (max_block_bytes < 64 || max_block_bytes > (1u << 30)) return DSRCGPU_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device >= ndev) return DSRCGPU_E_CUDA;   // no CPU fallback, by design
    dsrcgpu_ctx* ctx = new dsrcgpu_ctx();
    ctx->device = device; ctx->ds = *dataset; ctx->cs = *settings; ctx->max_block = max_block_bytes;
    memset(ctx->k_ms, 0, sizeof(ctx->k_ms)); memset(ctx->k_launches, 0, sizeof(ctx->k_launches));
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return DSRCGPU_E_CUDA; }
    cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, device);
    if (settings->dna_order || settings->quality_order) { if (rc_init_device() != cudaSuccess) { delete ctx; return DSRCGPU_E_CUDA; } }
    if (max_inflight_blocks == 0) {
        u64 n = (2048ull << 20) / max_block_bytes;         // ~2 GiB of FASTQ per batch: the range-coder chains of a batch are latency-bound,
        max_inflight_blocks = (u32)std::min<u64>(std::max<u64>(n, 1), 8192);   // their launch takes the same time for 1 or 16 K blocks
    }
    ctx->max_inflight = max_inflight_blocks;
    // batches in flight (streams): 3 by default -- copies, parallel kernels and the range-coder chains of different batches overlap
    int ns = 3;
    if (const char* e = getenv("DSRCGPU_SLOTS")) ns = atoi(e);
    ctx->n_slots = std::max(1, std::min(ns, (int)MAX_SLOTS));
    if (const char* e = getenv("DSRCGPU_PSERIAL")) ctx->p_serial = atoi(e);
    if (const char* e = getenv("DSRCGPU_NARROW_STREAMS")) ctx->narrow_div = (u32)std::max(0, atoi(e));
    if (const char* e = getenv("DSRCGPU_RC_GROUP")) ctx->rc_group = atoi(e);
    ctx->rc_group = std::max(1, std::min(ctx->rc_group, std::min((int)RC_GROUP_MAX, std::max(1, ctx->n_slots - 1))));
    if (const char* e = getenv("DSRCGPU_RC_GROUP_HOST")) ctx->rc_group_host = atoi(e);
    ctx->rc_group_host = std::max(1, std::min(ctx->rc_group_host, std::min((int)RC_GROUP_MAX, std::max(1, ctx->n_slots - 1))));
    // The model launches of successive batches run one after another (DSRCGPU_PSERIAL=2: two model launches side by side only stretch
    // each other), the serial stage of a batch -- latency-bound range-coder chains at ~5 % occupancy -- runs beside the next batch's
    // parallel stage. DSRCGPU_RSTREAM=1 moves the serial stage to a high-priority stream of its own (developer switch).
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    bool rstream_on = false;                         // measured on B200: no gain over the slot's own stream (the chains are placed early enough either way)
    if (const char* e = getenv("DSRCGPU_RSTREAM")) rstream_on = atoi(e) != 0;
    for (int i = 0; i < ctx->n_slots; ++i) {
        Slot& sl = ctx->slots[i];
        if (cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking) != cudaSuccess ||
            (rstream_on && cudaStreamCreateWithPriority(&sl.stream_r, cudaStreamNonBlocking, prio_hi) != cudaSuccess) ||
            cudaEventCreateWithFlags(&sl.ev_pdone, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&sl.ev_results, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&sl.ev_sizes, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&sl.ev_probe, cudaEventDisableTiming) != cudaSuccess) {
            for (int k = 0; k <= i; ++k) ctx->slots[k].release();
            delete ctx; return DSRCGPU_E_CUDA;
        }
    }
    // persistent model CTAs: 4 per SM (what the kernels' registers allow; the range-coder chains of other batches fit beside
    // them), fewer if the sort arenas (2 x 8 B x block/2 entries each) would exceed ~8 GiB per slot
    ctx->model_stride = (u64)max_block_bytes / 2 + 64;
    u64 per_cta = ctx->model_stride * 8 * 2;
    u64 want = (u64)ctx->sms * 4;
    if (const char* e = getenv("DSRCGPU_MODEL_CTAS")) want = (u64)std::max(1, atoi(e));
    u64 ctas = std::min<u64>(want, std::max<u64>(1, (8ull << 30) / per_cta));
    ctas = std::min<u64>(ctas, max_inflight_blocks);
    ctx->model_ctas = (u32)ctas;
    {   // adaptive-row tables of the tile/table model engine: rows of 2N bytes, one table per persistent CTA
        u64 tq = 0, td = 0;
        if (settings->quality_order == 1) tq = (1ull << 16) * 32; else if (settings->quality_order == 2) tq = (1ull << 20) * 32;   // <16,3,*> / <16,4,*>
        if (settings->dna_order) { const u32 o = settings->dna_order, o8 = o > 7 ? 7 : o; td = std::max<u64>((1ull << (2 * o)) * 8, (1ull << (3 * o8)) * 16); }
        ctx->tab_stride = std::max(tq, td);
    }
    ctx->tag_ctas = (u32)std::min<u64>((u64)ctx->sms * 4, max_inflight_blocks);
    ctx->q0_stride = q0_arena_bytes(max_block_bytes);
    ctx->q0_ctas = (u32)std::min<u64>(std::min<u64>((u64)ctx->sms * 2, max_inflight_blocks), std::max<u64>(1, (8ull << 30) / ctx->q0_stride));
    *out = ctx;
    return DSRCGPU_OK;
}

extern "C" void dsrcgpu_destroy(dsrcgpu_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < ctx->n_slots; ++i) { if (ctx->slots[i].stream) cudaStreamSynchronize(ctx->slots[i].stream); if (ctx->slots[i].stream_r) cudaStreamSynchronize(ctx->slots[i].stream_r); ctx->slots[i].release(); }
    ctx->prof.release(); ctx->cursor.release(); ctx->dec_arena.release(); ctx->tab.release(); ctx->tab_mask.release(); ctx->spare_in[0].release(); ctx->spare_in[1].release(); ctx->dec_out2.release();
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    for (cudaEvent_t e : ctx->ev_copy) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->call_a) cudaEventDestroy(ctx->call_a);
    if (ctx->call_b) cudaEventDestroy(ctx->call_b);
    delete ctx;
}

extern "C" void dsrcgpu_set_profiling(dsrcgpu_ctx* ctx, int on) { if (ctx) { ctx->profiling = (on & 1) != 0; ctx->phase_prof = (on & 2) != 0; } }
extern "C" int dsrcgpu_phase_cycles(dsrcgpu_ctx* ctx, uint64_t* out64, int reset)
{
    if (!ctx || !out64) return DSRCGPU_E_ARG;
    memset(out64, 0, 64 * 8);
    if (!ctx->prof.p) return DSRCGPU_OK;
    cudaSetDevice(ctx->device);
    CK(cudaMemcpy(out64, ctx->prof.p, 64 * 8, cudaMemcpyDeviceToHost));
    if (reset) CK(cudaMemset(ctx->prof.p, 0, 64 * 8));
    return DSRCGPU_OK;
}
extern "C" int dsrcgpu_last_kernel_times(dsrcgpu_ctx* ctx, const char** names, float* ms, uint32_t* launches, int max_entries)
{
    int n = 0;
    for (int k = 0; k < K_NUM && n < max_entries; ++k) {
        if (!ctx->k_launches[k]) continue;
        if (names) names[n] = K_NAMES[k];
        if (ms) ms[n] = ctx->k_ms[k];
        if (launches) launches[n] = ctx->k_launches[k];
        ++n;
    }
    return n;
}

static int ensure_host(dsrcgpu_ctx* ctx, Slot& sl, u32 n)
{
    if (n <= sl.h_cap) return DSRCGPU_OK;
    if (sl.h_desc) cudaFreeHost(sl.h_desc);
    if (sl.h_result) cudaFreeHost(sl.h_result);
    if (sl.h_probe) cudaFreeHost(sl.h_probe);
    sl.h_desc = nullptr; sl.h_result = nullptr; sl.h_probe = nullptr; sl.h_cap = 0;
    // mapped: the encode scheduler moves these with a kernel (launch_copy_words), not with the copy engines
    CK(cudaHostAlloc((void**)&sl.h_desc, sizeof(BlockDesc) * n, cudaHostAllocMapped | cudaHostAllocPortable));
    CK(cudaHostAlloc((void**)&sl.h_result, sizeof(BlockResult) * n, cudaHostAllocMapped | cudaHostAllocPortable));
    CK(cudaHostAlloc((void**)&sl.h_probe, sizeof(BlockProbe) * n, cudaHostAllocMapped | cudaHostAllocPortable));
    sl.h_cap = n;
    return DSRCGPU_OK;
}

static inline u64 align_up(u64 v, u64 a) { return (v + a - 1) / a * a; }

static int status_to_error(dsrcgpu_ctx* ctx, u32 status, u32 blk)
{
    char msg[160];
    int code = DSRCGPU_E_MALFORMED;
    const char* what = "malformed FASTQ block";
    if (status & 0x100) { code = DSRCGPU_E_CAPACITY; what = "output buffer too small"; }
    else if (status == ST_UNSUPPORTED) { code = DSRCGPU_E_UNSUPPORTED; what = "block outside the supported envelope (see DESIGN.md limits)"; }
    else if (status == ST_OVERFLOW) { code = DSRCGPU_E_UNSUPPORTED; what = "internal stream arena too small for this block"; ctx->last_overflow = true; }
    else if (status == ST_CRC) { code = DSRCGPU_E_MALFORMED; what = "CRC32 checksums mismatch."; }      // message of src/DsrcWorker.cpp:60
    snprintf(msg, sizeof(msg), "block %u: %s (status %u)", blk, what, status);
    ctx->err = msg;
    return code;
}

// Enqueues one batch of blocks on its slot's stream: layout probe (one short host sync), then every per-block kernel and the
// result read-back, all asynchronous. d_in / d_out are device pointers; with `cursor` the dense output continues where the
// previous batch ended (device-resident output), else the batch starts at out_base of its own staging buffer.
// `idle` is called over and over while the host waits for the layout probe (i.e. for this batch's input copy): the scheduler uses it
// to retire batches that finish meanwhile, so that their output copies start at once and run beside this input copy.
static int enqueue_batch(dsrcgpu_ctx* ctx, Slot& sl, const u8* d_in, const u32* blk_len, const u32* blk_tagcap, u32 n,
                         u8* d_out, u64 out_base, u64 out_cap, u64* cursor, cudaEvent_t wait_p, const std::function<int()>& idle)
{
    int rc = ensure_host(ctx, sl, n);
    if (rc) return rc;
    cudaStream_t s = sl.stream;
    CK(sl.desc.ensure(sizeof(BlockDesc) * n));
    CK(sl.state.ensure(sizeof(BlockState) * n));
    CK(sl.result.ensure(sizeof(BlockResult) * n));
    CK(sl.probe.ensure(sizeof(BlockProbe) * n));
    BlockDesc* hd = sl.h_desc;
    for (u32 i = 0; i < n; ++i) {
        memset(&hd[i], 0, sizeof(BlockDesc));
        hd[i].in_off = sl.offs[i]; hd[i].in_len = blk_len[i];
        hd[i].tag_cap = blk_tagcap ? blk_tagcap[i] : 0xFFFFFFFFu;
        if (blk_len[i] == 0 || blk_len[i] > ctx->max_block) { ctx->err = "block length 0 or above max_block_bytes"; return DSRCGPU_E_ARG; }
    }
    Workspace ws{};
    ws.in = d_in; ws.desc = (const BlockDesc*)sl.desc.p; ws.state = (BlockState*)sl.state.p;
    ws.result = (BlockResult*)sl.result.p; ws.probe = (BlockProbe*)sl.probe.p;
    ws.n_blocks = n; ws.qoff = ctx->ds.quality_offset; ws.plus_rep = ctx->ds.plus_repetition;
    ws.dna_order = ctx->cs.dna_order; ws.qua_order = ctx->cs.quality_order; ws.calc_crc = ctx->cs.calc_crc32 ? 1u : 0u;
    ws.out = d_out; ws.out_cap = out_cap;
    if (ctx->phase_prof) { if (!ctx->prof.p) { CK(ctx->prof.ensure(64 * 8)); CK(cudaMemsetAsync(ctx->prof.p, 0, 64 * 8, s)); } ws.prof = (u64*)ctx->prof.p; }

    // pass 1: count lines / fields so the batch can be laid out exactly
    void *m_desc = nullptr, *m_probe = nullptr;      // device addresses of the mapped host arrays
    CK(cudaHostGetDevicePointer(&m_desc, hd, 0)); CK(cudaHostGetDevicePointer(&m_probe, sl.h_probe, 0));
    // host-buffer calls move the small arrays with a kernel over mapped memory (a DMA copy would queue behind the 2 GiB payload copies
    // of the other batches); device-resident calls keep the copy engines, which are idle there, while a kernel would wait for an SM
    // among the persistent model CTAs (measured: -6 %)
    const bool by_kernel = ctx->host_call;
    if (by_kernel) launch_copy_words(sl.desc.p, m_desc, sizeof(BlockDesc) * n, s); else CK(cudaMemcpyAsync(sl.desc.p, hd, sizeof(BlockDesc) * n, cudaMemcpyHostToDevice, s));
    { KTimer t(ctx, &sl, K_COUNT); launch_count_lines(ws, s); }
    if (by_kernel) launch_copy_words(m_probe, sl.probe.p, sizeof(BlockProbe) * n, s); else CK(cudaMemcpyAsync(sl.h_probe, sl.probe.p, sizeof(BlockProbe) * n, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(sl.ev_probe, s));
    for (;;) {
        const cudaError_t q = cudaEventQuery(sl.ev_probe);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) { ctx->err = std::string("layout probe: ") + cudaGetErrorString(q); return DSRCGPU_E_CUDA; }
        if (idle) { const int r = idle(); if (r) return r; }
        const struct timespec ts = {0, 20000};
        nanosleep(&ts, nullptr);
    }

    u64 lines = 0, recs = 0, syms = 0, ftab = 0, streams = 0;
    for (u32 i = 0; i < n; ++i) {
        BlockDesc& d = hd[i];
        const u32 nl = sl.h_probe[i].n_lines;
        d.line_base = (u32)lines; d.line_cap = nl; lines += nl;
        d.rec_base = (u32)recs; d.rec_cap = nl / 4 + 1; recs += d.rec_cap;
        d.sym_base = syms; d.sym_cap = d.in_len / 2 + 16; syms += align_up(d.sym_cap, 16);
        d.n_fields = sl.h_probe[i].n_fields;
        d.ftab_base = ftab; ftab += (u64)d.n_fields * d.rec_cap;
        d.stream_base = streams;
        d.stream_cap[0] = 64;
        d.stream_cap[1] = (u32)align_up((u64)d.in_len / 2 + (128u << 10), 16);
        // range-coded streams: 1.25 bytes per symbol (the adaptive coder averages below log2(alphabet) bits per symbol; a chain that
        // runs out reports ST_OVERFLOW and the call is repeated with `wide_streams`: 3 bytes per symbol, more than a symbol can
        // cost). The -d0 / -q0 coders (bit packing, Huffman tables in the stream) keep the generous bound.
        const u64 half = (u64)d.in_len / 2, narrow = ctx->narrow_div ? half / ctx->narrow_div : half + half / 4;
        d.stream_cap[2] = (u32)align_up((ctx->cs.dna_order && !ctx->wide_streams) ? narrow + 1024 : half * 3 + 256, 16);
        d.stream_cap[3] = (u32)align_up((ctx->cs.quality_order && !ctx->wide_streams) ? narrow + 2048 : half * 3 + (256u << 10), 16);
        streams += (u64)d.stream_cap[0] + d.stream_cap[1] + d.stream_cap[2] + d.stream_cap[3];
        if (lines >= (1ull << 32) || recs >= (1ull << 32)) { ctx->err = "batch too large"; return DSRCGPU_E_ARG; }
    }
    CK(sl.lines.ensure(lines * 4));
    CK(sl.r_title_off.ensure(recs * 4)); CK(sl.r_seq_off.ensure(recs * 4)); CK(sl.r_qua_off.ensure(recs * 4));
    CK(sl.r_qcat_off.ensure(recs * 4)); CK(sl.r_dcat_off.ensure(recs * 4));
    CK(sl.r_title_len.ensure(recs * 2)); CK(sl.r_qua_len.ensure(recs * 2)); CK(sl.r_dna_len.ensure(recs * 2)); CK(sl.r_trunc_len.ensure(recs * 2));
    CK(sl.qcat.ensure(syms)); CK(sl.dcat.ensure(syms));
    const bool rc_q = ctx->cs.quality_order > 0, rc_d = ctx->cs.dna_order > 0;
    if (rc_q) CK(sl.trip_q.ensure(syms * 8));
    if (rc_d) CK(sl.trip_d.ensure(syms * 8));
    CK(sl.ftab.ensure(ftab * 8));
    CK(sl.streams.ensure(streams));
    if (rc_q || rc_d) { CK(sl.elem_a.ensure(ctx->model_stride * 8 * ctx->model_ctas)); CK(sl.elem_b.ensure(ctx->model_stride * 8 * ctx->model_ctas)); }
    if ((rc_q || rc_d) && !ctx->tab.p && ctx->tab_stride) {
        // one table per model CTA that can be resident at a time (at most 5 per SM), whatever the slot; small contexts need fewer
        ctx->tab_count = (u32)std::max<u64>(1, std::min<u64>(std::min<u64>((u64)ctx->sms * 5, (u64)ctx->model_ctas * ctx->n_slots), (24ull << 30) / ctx->tab_stride));
        CK(ctx->tab.ensure(ctx->tab_stride * ctx->tab_count));
        CK(ctx->tab_mask.ensure(((ctx->tab_count + 31) / 32) * 4));
        CK(cudaMemsetAsync(ctx->tab.p, 0, ctx->tab.cap, s));    // invariant between blocks: first counter of every row is 0
        CK(cudaMemsetAsync(ctx->tab_mask.p, 0, ctx->tab_mask.cap, s));
        CK(cudaStreamSynchronize(s));                          // once per context: the other slots' streams use the pool too
    }
    CK(sl.queue.ensure(16));
    CK(cudaMemsetAsync(sl.queue.p, 0, 16, s));
    CK(sl.tagpool.ensure(tagpool_bytes_per_block() * ctx->tag_ctas));
    if (!rc_q || !rc_d) CK(sl.q0_arena.ensure(ctx->q0_stride * ctx->q0_ctas));

    ws.lines = (u32*)sl.lines.p;
    ws.rec.title_off = (u32*)sl.r_title_off.p; ws.rec.seq_off = (u32*)sl.r_seq_off.p; ws.rec.qua_off = (u32*)sl.r_qua_off.p;
    ws.rec.title_len = (u16*)sl.r_title_len.p; ws.rec.qua_len = (u16*)sl.r_qua_len.p; ws.rec.dna_len = (u16*)sl.r_dna_len.p;
    ws.rec.trunc_len = (u16*)sl.r_trunc_len.p; ws.rec.qcat_off = (u32*)sl.r_qcat_off.p; ws.rec.dcat_off = (u32*)sl.r_dcat_off.p;
    ws.qcat = (u8*)sl.qcat.p; ws.dcat = (u8*)sl.dcat.p;
    ws.trip_q = (u64*)sl.trip_q.p; ws.trip_d = (u64*)sl.trip_d.p;
    ws.elem_a = (u64*)sl.elem_a.p; ws.elem_b = (u64*)sl.elem_b.p;
    ws.ftab = (u64*)sl.ftab.p; ws.streams = (u8*)sl.streams.p;
    ws.tab = (u8*)ctx->tab.p; ws.tab_stride = ctx->tab_stride; ws.tab_mask = (u32*)ctx->tab_mask.p; ws.tab_count = ctx->tab_count;
    ws.model_queue = (u32*)sl.queue.p;
    ws.tagpool = (u8*)sl.tagpool.p; ws.tagpool_stride = tagpool_bytes_per_block();

    if (by_kernel) launch_copy_words(sl.desc.p, m_desc, sizeof(BlockDesc) * n, s); else CK(cudaMemcpyAsync(sl.desc.p, hd, sizeof(BlockDesc) * n, cudaMemcpyHostToDevice, s));
    if (wait_p && ctx->p_serial == 1) CK(cudaStreamWaitEvent(s, wait_p, 0));
    { KTimer t(ctx, &sl, K_PARSE); launch_parse(ws, s); }
    if (ws.calc_crc) { KTimer t(ctx, &sl, K_CRC); launch_crc(ws, s, 0); }
    { KTimer t(ctx, &sl, K_PREP); launch_preprocess(ws, s); }
    { KTimer t(ctx, &sl, K_TAGS); launch_tags(ws, s, ctx->tag_ctas); }
    if (wait_p && ctx->p_serial == 2) CK(cudaStreamWaitEvent(s, wait_p, 0));
    if (rc_q) { KTimer t(ctx, &sl, K_MODEL_Q, nullptr, 3); launch_model_quality(ws, s, ctx->model_ctas, ctx->model_stride); }
    else { KTimer t(ctx, &sl, K_Q0); launch_q0_quality(ws, s, (u8*)sl.q0_arena.p, ctx->q0_stride, ctx->q0_ctas); }
    if (rc_d) { KTimer t(ctx, &sl, K_MODEL_D, nullptr, 2); launch_model_dna(ws, s, ctx->model_ctas, ctx->model_stride); }
    else { KTimer t(ctx, &sl, K_D0); launch_d0_dna(ws, s, (u8*)sl.q0_arena.p, ctx->q0_stride, ctx->q0_ctas); }
    CK(cudaEventRecord(sl.ev_pdone, s));
    CK(cudaGetLastError());
    sl.ws = ws; sl.r_enq = false; sl.r_out_base = out_base; sl.r_cursor = cursor;
    return DSRCGPU_OK;
}

// Serial stage of a group of batches (consecutive batches, in order): ONE launch of the range-coder chains for all of them -- the
// launch is latency-bound, its duration does not depend on the number of chains at these sizes -- then per batch the size scan
// (chained through the device cursor for device-resident output), the gather and the result read-back.
static int finish_group(dsrcgpu_ctx* ctx, Slot** members, u32 n, cudaEvent_t wait_sizes)
{
    if (!n) return DSRCGPU_OK;
    const bool rc_q = ctx->cs.quality_order > 0, rc_d = ctx->cs.dna_order > 0;
    Slot& last = *members[n - 1];
    const cudaStream_t r = rstream(last);
    if (rc_q || rc_d) {
        RcGroup grp{}; grp.n = n;
        for (u32 k = 0; k < n; ++k) { grp.ws[k] = members[k]->ws; if (members[k] != &last || r != last.stream) CK(cudaStreamWaitEvent(r, members[k]->ev_pdone, 0)); }
        static const bool skip_rc = getenv("DSRCGPU_DEV_SKIP_RC") != nullptr;      // developer timing experiment: output is wrong without the chains
        if (!skip_rc) { KTimer t(ctx, &last, K_RC, r); launch_rc_encode(grp, r); }
    }
    cudaEvent_t ev_rc = get_event(ctx);
    CK(cudaEventRecord(ev_rc, r));
    for (u32 k = 0; k < n; ++k) {
        Slot& sl = *members[k];
        const cudaStream_t rs = rstream(sl);
        if (rs != r) CK(cudaStreamWaitEvent(rs, ev_rc, 0));
        if (rs != sl.stream) CK(cudaStreamWaitEvent(rs, sl.ev_pdone, 0));
        if (wait_sizes) CK(cudaStreamWaitEvent(rs, wait_sizes, 0));       // the previous batch has to publish where its output ends
        { KTimer t(ctx, &sl, K_SIZES, rs); launch_meta_and_sizes(sl.ws, rs, sl.r_out_base, sl.r_cursor); }
        CK(cudaEventRecord(sl.ev_sizes, rs));
        { KTimer t(ctx, &sl, K_GATHER, rs); launch_gather(sl.ws, rs); }
        if (ctx->host_call) { void* m_res = nullptr; CK(cudaHostGetDevicePointer(&m_res, sl.h_result, 0)); launch_copy_words(m_res, sl.result.p, sizeof(BlockResult) * sl.cnt, rs); }
        else CK(cudaMemcpyAsync(sl.h_result, sl.result.p, sizeof(BlockResult) * sl.cnt, cudaMemcpyDeviceToHost, rs));
        CK(cudaEventRecord(sl.ev_results, rs));
        sl.r_enq = true;
        wait_sizes = sl.r_cursor ? sl.ev_sizes : nullptr;
    }
    ctx->ev_pool.push_back(ev_rc);
    CK(cudaGetLastError());
    return DSRCGPU_OK;
}

// timers of the input copies that travelled on the copy stream (call this when they are complete: after the streams that waited for
// them have been joined); they are listed under slot 0 in the timeline
static void collect_copy_times(dsrcgpu_ctx* ctx)
{
    Slot& s0 = ctx->slots[0];
    s0.ev_used.insert(s0.ev_used.end(), ctx->copy_ev.begin(), ctx->copy_ev.end());
    ctx->copy_ev.clear();
    collect_times(ctx, &s0);
}

static void abort_all(dsrcgpu_ctx* ctx)
{
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < ctx->n_slots; ++i) { cudaStreamSynchronize(ctx->slots[i].stream); cudaStreamSynchronize(rstream(ctx->slots[i])); collect_times(ctx, &ctx->slots[i]); ctx->slots[i].busy = false; }
    collect_copy_times(ctx);
}

// The CUDA-stream block scheduler (replaces the worker pool + queues of DsrcCompressorMT, src/DsrcOperator.cpp:230-394):
// the block queue is cut into batches, batch b runs on slot b % n_slots, results are retired strictly in block order.
static int encode_impl(dsrcgpu_ctx* ctx, const u8* fastq, bool on_device, const u64* blk_off, const u32* blk_len, const u32* blk_tagcap, u32 n,
                       u8* out, u64 out_cap, u32* out_sizes, u64* raw_sizes, u64* comp_sizes)
{
    if (!ctx) return DSRCGPU_E_ARG;
    if (!fastq || !blk_off || !blk_len || !out || !out_sizes) { ctx->err = "null argument"; return DSRCGPU_E_ARG; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return DSRCGPU_E_CUDA; }
    memset(ctx->k_ms, 0, sizeof(ctx->k_ms)); memset(ctx->k_launches, 0, sizeof(ctx->k_launches));
    if (!ctx->call_a) { cudaEventCreate(&ctx->call_a); cudaEventCreate(&ctx->call_b); }
    const int S = ctx->n_slots;
    const unsigned long long allocs_before = g_devbuf_allocs;
    ctx->host_call = !on_device;
    cudaEventRecord(ctx->call_a, ctx->slots[0].stream);
    if (on_device) {
        CK(ctx->cursor.ensure(8));
        CK(cudaMemsetAsync(ctx->cursor.p, 0, 8, ctx->slots[0].stream));
    }
    // batch schedule: uniform batches. (Measured on B200 for host buffers, where the input link is the bound and the call ends one
    // batch's coding after the last byte has arrived: cutting the last batch into 2-4 parts (DSRCGPU_TAIL_SPLIT) does not end the call
    // sooner -- the range-coder chains take ~10 ms whatever the batch size and the parts wait for slots.)
    std::vector<u32> bfirst;
    // host buffers: half-size batches while the input link is the bound (measured 47.6 -> 49.7 GB/s end to end: with 1 GiB batches the
    // call ends sooner after the last byte has arrived; resident input prefers the full size, 91 vs 77 GB/s). A context whose last
    // host-buffer call was bound by the coding instead (41-level qualities: 36 GB/s of coding against 52 GB/s of link) goes back to
    // full-size batches, which cost fewer range-coder launches (host_full, decided at the end of every call from the copy timers).
    const u32 per_batch = (on_device || ctx->host_full) ? ctx->max_inflight : std::max(1u, ctx->max_inflight / 2);
    struct ScaleGuard { unsigned old; ScaleGuard(unsigned v) : old(t_alloc_scale) { t_alloc_scale = v; } ~ScaleGuard() { t_alloc_scale = old; } }
        scale_guard((!on_device && !ctx->host_full && n > per_batch) ? 2u : 1u);
    for (u32 pos = 0; pos < n; pos += per_batch) bfirst.push_back(pos);
    if (!on_device && bfirst.size() >= 2) {
        const u32 a = bfirst.back(), len = n - a;
        u32 parts = 1;
        if (const char* e = getenv("DSRCGPU_TAIL_SPLIT")) parts = (u32)std::max(1, atoi(e));
        if (len >= 512 * parts) for (u32 k = 1; k < parts; ++k) bfirst.push_back(a + (u32)((u64)len * k / parts));
    }
    bfirst.push_back(n);
    const u32 nb = (u32)bfirst.size() - 1;
    u64 out_pos = 0;
    u32 retired = 0;
    int rc = DSRCGPU_OK;

    // batches whose parallel stage is enqueued and whose serial stage waits for the group to fill
    Slot* group[RC_GROUP_MAX]; u32 n_group = 0;
    cudaEvent_t prev_sizes = nullptr;
    auto close_group = [&]() -> int {
        const int r = finish_group(ctx, group, n_group, prev_sizes);
        if (n_group && on_device) prev_sizes = group[n_group - 1]->ev_sizes;
        n_group = 0;
        return r;
    };

    auto retire = [&](u32 r) -> int {
        Slot& t = ctx->slots[r % S];
        if (!t.r_enq) { const int g = close_group(); if (g) return g; }      // (the schedule closes groups before their batches are due)
        CK(cudaEventSynchronize(t.ev_results));
        collect_times(ctx, &t);
        u64 end = on_device ? out_pos : 0;
        if (!on_device) {
            // tiny blocks with long read-ID fields can come out LARGER than the staging estimate (hundreds of Huffman trees in the
            // tag header): the streams are still in the slot, so grow the staging buffer and repeat only the size scan + gather
            bool grow = false; u64 need = 0;
            for (u32 i = 0; i < t.cnt; ++i) { if (t.h_result[i].status & 0x100) grow = true; need += t.h_result[i].total_size; }
            if (grow) {
                CK(t.out.ensure(need + 64));
                t.ws.out = (u8*)t.out.p; t.ws.out_cap = need + 64;
                launch_meta_and_sizes(t.ws, rstream(t), 0, nullptr);
                launch_gather(t.ws, rstream(t));
                CK(cudaMemcpyAsync(t.h_result, t.result.p, sizeof(BlockResult) * t.cnt, cudaMemcpyDeviceToHost, rstream(t)));
                CK(cudaStreamSynchronize(rstream(t)));
            }
        }
        for (u32 i = 0; i < t.cnt; ++i) {
            const BlockResult& br = t.h_result[i];
            if (br.status != ST_OK) return status_to_error(ctx, br.status, t.first + i);
            end = br.out_off + br.total_size;
            out_sizes[t.first + i] = br.total_size;
            if (raw_sizes) for (int k = 0; k < 4; ++k) raw_sizes[(u64)(t.first + i) * 4 + k] = br.raw[k];
            if (comp_sizes) for (int k = 0; k < 4; ++k) comp_sizes[(u64)(t.first + i) * 4 + k] = br.stream_size[k];
        }
        if (on_device) out_pos = end;
        else {
            if (out_pos + end > out_cap) { ctx->err = "output buffer too small"; return DSRCGPU_E_CAPACITY; }
            { KTimer tcopy(ctx, &t, K_D2H, rstream(t)); CK(cudaMemcpyAsync(out + out_pos, t.out.p, end, cudaMemcpyDeviceToHost, rstream(t))); }
            out_pos += end;
        }
        return DSRCGPU_OK;
    };

    // batches that have finished are retired as soon as the host gets here, not when their slot is needed again: in host mode that is
    // what starts their output copy, and it should run beside the NEXT batch's input copy (the two directions of the link), not
    // between two input copies
    auto retire_ready = [&](u32 enqueued) {
        while (rc == DSRCGPU_OK && retired < enqueued && ctx->slots[retired % S].r_enq && cudaEventQuery(ctx->slots[retired % S].ev_results) == cudaSuccess) rc = retire(retired++);
    };
    // host buffers: the input of the next two batches is sent ahead on the copy stream into two spare staging buffers -- one copy per
    // batch when its blocks are a (near-)contiguous ascending span, packed copies otherwise -- so the input link keeps running while
    // the host waits for a slot; a batch takes its buffer over (swap with the slot's) when its turn comes
    struct Staged { u32 batch; int buf; std::vector<u64> offs; };
    std::vector<Staged> staged;                            // in batch order, at most 2
    bool spare_free[2] = {true, true};
    u32 next_stage = 0;
    auto stage_more = [&]() -> int {
        while (!on_device && next_stage < nb && (spare_free[0] || spare_free[1])) {
            if (!ctx->copy_stream) {
                CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
                for (cudaEvent_t& e : ctx->ev_copy) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            }
            const int k = spare_free[0] ? 0 : 1;
            DevBuf& buf = ctx->spare_in[k];
            const u32 first = bfirst[next_stage], cnt = bfirst[next_stage + 1] - first;
            Staged st; st.batch = next_stage; st.buf = k; st.offs.resize(cnt);
            u64 lo = blk_off[first], hi = 0, sum = 0; bool asc = true;
            for (u32 i = 0; i < cnt; ++i) {
                u64 o = blk_off[first + i]; u32 l = blk_len[first + i];
                if (i && o < blk_off[first + i - 1] + blk_len[first + i - 1]) asc = false;
                lo = std::min(lo, o); hi = std::max(hi, o + l); sum += l;
            }
            cudaEvent_t ta = nullptr, tb = nullptr;
            ctx->k_launches[K_H2D]++;
            if (ctx->profiling) { ta = get_event(ctx); tb = get_event(ctx); cudaEventRecord(ta, ctx->copy_stream); }
            if (asc && hi - lo <= sum + (u64)cnt * 64) {
                CK(buf.ensure(hi - lo + 16));
                CK(cudaMemcpyAsync(buf.p, fastq + lo, hi - lo, cudaMemcpyHostToDevice, ctx->copy_stream));
                for (u32 i = 0; i < cnt; ++i) st.offs[i] = blk_off[first + i] - lo;
            } else {
                CK(buf.ensure(sum + (u64)cnt * 16 + 16));
                u64 p = 0;
                for (u32 i = 0; i < cnt; ++i) {
                    CK(cudaMemcpyAsync((u8*)buf.p + p, fastq + blk_off[first + i], blk_len[first + i], cudaMemcpyHostToDevice, ctx->copy_stream));
                    st.offs[i] = p; p += align_up(blk_len[first + i], 16);
                }
            }
            if (ctx->profiling) { cudaEventRecord(tb, ctx->copy_stream); ctx->copy_ev.push_back({K_H2D, {ta, tb}}); }
            CK(cudaEventRecord(ctx->ev_copy[k], ctx->copy_stream));
            spare_free[k] = false;
            staged.push_back(std::move(st));
            ++next_stage;
        }
        return DSRCGPU_OK;
    };
    rc = stage_more();                                     // the first inputs leave before anything else
    for (u32 b = 0; b < nb && rc == DSRCGPU_OK; ++b) {
        Slot& sl = ctx->slots[b % S];
        retire_ready(b);
        if (rc) break;
        if (sl.busy) {
            while (rc == DSRCGPU_OK && retired <= sl.batch) rc = retire(retired++);
            if (rc) break;
            cudaError_t e = cudaStreamSynchronize(rstream(sl));           // its output copy has left the staging buffer
            if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = DSRCGPU_E_CUDA; break; }
            sl.busy = false;
        }
        const u32 first = bfirst[b], cnt = bfirst[b + 1] - first;
        sl.first = first; sl.cnt = cnt; sl.batch = b;
        sl.offs.resize(cnt);
        const u8* d_in; u8* d_out; u64 batch_out_base = 0, batch_out_cap;
        if (on_device) {
            d_in = fastq;
            for (u32 i = 0; i < cnt; ++i) sl.offs[i] = blk_off[first + i];
            d_out = out; batch_out_cap = out_cap;
        } else {
            // the batch's input was sent ahead (stage_more): take its buffer over and send the next batch's into the one given back
            if (staged.empty() || staged.front().batch != b) { ctx->err = "internal: input staging out of step"; rc = DSRCGPU_E_ARG; break; }
            const int k = staged.front().buf;
            std::swap(sl.in, ctx->spare_in[k]);
            sl.offs.swap(staged.front().offs);
            staged.erase(staged.begin());
            cudaError_t e = cudaStreamWaitEvent(sl.stream, ctx->ev_copy[k], 0);
            spare_free[k] = true;
            if (e == cudaSuccess) { rc = stage_more(); if (rc) break; }
            u64 bound = 0;
            for (u32 i = 0; i < cnt; ++i) bound += (u64)blk_len[first + i] + (blk_len[first + i] >> 1) + 4096;   // generous: DSRC never expands by 1.5x
            if (e == cudaSuccess) e = sl.out.ensure(bound);
            if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = DSRCGPU_E_CUDA; break; }
            d_in = (const u8*)sl.in.p;
            d_out = (u8*)sl.out.p; batch_out_cap = bound;
        }
        sl.busy = true;
        rc = enqueue_batch(ctx, sl, d_in, blk_len + first, blk_tagcap ? blk_tagcap + first : nullptr, cnt, d_out, batch_out_base, batch_out_cap,
                           on_device ? (u64*)ctx->cursor.p : nullptr,
                           (b > 0 && S > 1) ? ctx->slots[(b - 1) % S].ev_pdone : nullptr,
                           [&]() -> int { retire_ready(b); return rc; });
        if (rc) break;
        group[n_group++] = &sl;
        // (host buffers: the input link is the bound, not the GPU -- every batch finishes on its own, which frees its slot sooner)
        if ((int)n_group >= (on_device ? ctx->rc_group : ctx->rc_group_host) || b + 1 == nb) { rc = close_group(); if (rc) break; }
        // keep S-1 batches queued behind the one the host waits for
        while (rc == DSRCGPU_OK && retired + (u32)(S - 1) <= b && S > 1 && retired < b) rc = retire(retired++);
    }
    while (rc == DSRCGPU_OK && retired < nb) rc = retire(retired++);
    if (rc) { abort_all(ctx); return rc; }
    // join every stream into slot 0's for the call timing, then wait
    for (int i = 0; i < S; ++i) {
        for (cudaStream_t st : {ctx->slots[i].stream, ctx->slots[i].stream_r}) {
            if (!st || st == ctx->slots[0].stream) continue;
            cudaEvent_t e = get_event(ctx);
            cudaEventRecord(e, st); cudaStreamWaitEvent(ctx->slots[0].stream, e, 0);
            ctx->ev_pool.push_back(e);
        }
    }
    cudaEventRecord(ctx->call_b, ctx->slots[0].stream);
    for (int i = 0; i < S; ++i) { CK(cudaStreamSynchronize(ctx->slots[i].stream)); CK(cudaStreamSynchronize(rstream(ctx->slots[i]))); ctx->slots[i].busy = false; }
    CK(cudaEventSynchronize(ctx->call_b));
    cudaEventElapsedTime(&ctx->call_ms, ctx->call_a, ctx->call_b);
    collect_copy_times(ctx);
    if (!on_device && ctx->profiling && nb >= 4 && ctx->k_ms[K_H2D] > 0 && g_devbuf_allocs == allocs_before)      // (a warm call)
        ctx->host_full = ctx->call_ms > 1.25f * ctx->k_ms[K_H2D];
    return DSRCGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// decode: BlockCompressor::Read for a queue of blocks
// ------------------------------------------------------------------------------------------------
#define DEC_ST_RETRY 4u

// one (sub-)batch: blocks `idx[0..n)` of the call, already staged at offs[] in d_in; out_offs[] are absolute offsets in d_out
static int decode_batch(dsrcgpu_ctx* ctx, Slot& sl, const u8* d_in, const u32* idx, const u64* offs, const u32* blk_len, const u64* out_offs,
                        u32 n, u8* d_out, u64 out_cap, u64 arena_bytes, u32 pool_nodes, u32* status_out, u32* size_out)
{
    int rc = ensure_host(ctx, sl, n);
    if (rc) return rc;
    cudaStream_t s = sl.stream;
    CK(sl.desc.ensure(sizeof(BlockDesc) * n)); CK(sl.state.ensure(sizeof(BlockState) * n));
    CK(sl.result.ensure(sizeof(BlockResult) * n)); CK(sl.probe.ensure(sizeof(BlockProbe) * n));
    BlockDesc* hd = sl.h_desc;
    for (u32 i = 0; i < n; ++i) { memset(&hd[i], 0, sizeof(BlockDesc)); hd[i].in_off = offs[i]; hd[i].in_len = blk_len[idx[i]]; }
    Workspace ws{};
    ws.in = d_in; ws.desc = (const BlockDesc*)sl.desc.p; ws.state = (BlockState*)sl.state.p;
    ws.result = (BlockResult*)sl.result.p; ws.probe = (BlockProbe*)sl.probe.p;
    ws.n_blocks = n; ws.qoff = ctx->ds.quality_offset; ws.plus_rep = ctx->ds.plus_repetition;
    ws.dna_order = ctx->cs.dna_order; ws.qua_order = ctx->cs.quality_order; ws.calc_crc = ctx->cs.calc_crc32 ? 1u : 0u;
    ws.out = d_out; ws.out_cap = out_cap;
    CK(cudaMemcpyAsync(sl.desc.p, hd, sizeof(BlockDesc) * n, cudaMemcpyHostToDevice, s));
    { KTimer t(ctx, &sl, K_DECODE); launch_dec_probe(ws, s); }
    CK(cudaMemcpyAsync(sl.h_probe, sl.probe.p, sizeof(BlockProbe) * n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    u64 recs = 0, syms = 0, titles = 0;
    for (u32 i = 0; i < n; ++i) {
        BlockDesc& d = hd[i];
        const u32 nr = sl.h_probe[i].n_lines, chunk = sl.h_probe[i].n_fields;      // records, chunkSize + 1
        d.rec_base = (u32)recs; d.rec_cap = nr; recs += nr;
        d.sym_base = syms; d.sym_cap = chunk / 2 + 16; syms += align_up(d.sym_cap, 16);
        d.stream_base = titles; d.stream_cap[1] = (u32)align_up((u64)chunk + 16, 16); titles += d.stream_cap[1];
        d.out_off = out_offs[i];
        if (chunk && out_offs[i] + chunk > out_cap) { ctx->err = "output buffer too small"; return DSRCGPU_E_CAPACITY; }
        if (recs >= (1ull << 32)) { ctx->err = "batch too large"; return DSRCGPU_E_ARG; }
    }
    CK(sl.r_title_off.ensure(recs * 4 + 16)); CK(sl.r_seq_off.ensure(recs * 4 + 16)); CK(sl.r_qcat_off.ensure(recs * 4 + 16)); CK(sl.r_dcat_off.ensure(recs * 4 + 16));
    CK(sl.r_title_len.ensure(recs * 2 + 16)); CK(sl.r_qua_len.ensure(recs * 2 + 16)); CK(sl.r_dna_len.ensure(recs * 2 + 16));
    CK(sl.qcat.ensure(syms)); CK(sl.dcat.ensure(syms)); CK(sl.streams.ensure(titles + 16));
    CK(sl.ftab.ensure((u64)pool_nodes * 8 * n));
    const bool rcq = ctx->cs.quality_order > 0, rcd = ctx->cs.dna_order > 0;
    if (rcq || rcd) CK(ctx->dec_arena.ensure(arena_bytes * n));
    ws.rec.title_off = (u32*)sl.r_title_off.p; ws.rec.seq_off = (u32*)sl.r_seq_off.p; ws.rec.qcat_off = (u32*)sl.r_qcat_off.p; ws.rec.dcat_off = (u32*)sl.r_dcat_off.p;
    ws.rec.title_len = (u16*)sl.r_title_len.p; ws.rec.qua_len = (u16*)sl.r_qua_len.p; ws.rec.dna_len = (u16*)sl.r_dna_len.p;
    ws.qcat = (u8*)sl.qcat.p; ws.dcat = (u8*)sl.dcat.p; ws.streams = (u8*)sl.streams.p;
    CK(cudaMemcpyAsync(sl.desc.p, hd, sizeof(BlockDesc) * n, cudaMemcpyHostToDevice, s));
    { KTimer t(ctx, &sl, K_DEC_TAGS); launch_dec_tags(ws, s, sl.ftab.p, pool_nodes); }
    if (rcq) CK(cudaMemsetAsync(ctx->dec_arena.p, 0, arena_bytes * n, s));
    { KTimer t(ctx, &sl, K_DEC_Q); launch_dec_quality(ws, s, sl.ftab.p, pool_nodes, (u8*)ctx->dec_arena.p, arena_bytes, arena_bytes); }
    if (rcd) CK(cudaMemsetAsync(ctx->dec_arena.p, 0, arena_bytes * n, s));
    { KTimer t(ctx, &sl, K_DEC_D); launch_dec_dna(ws, s, sl.ftab.p, pool_nodes, (u8*)ctx->dec_arena.p, arena_bytes, arena_bytes); }
    { KTimer t(ctx, &sl, K_DEC_ASM); launch_dec_assemble(ws, s); }
    if (ws.calc_crc) { KTimer t(ctx, &sl, K_CRC); launch_crc(ws, s, 1); }
    CK(cudaMemcpyAsync(sl.h_result, sl.result.p, sizeof(BlockResult) * n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    collect_times(ctx, &sl);
    for (u32 i = 0; i < n; ++i) { status_out[i] = sl.h_result[i].status; size_out[i] = sl.h_probe[i].n_fields; }
    return DSRCGPU_OK;
}

static int decode_run(dsrcgpu_ctx* ctx, const u8* dsrc, bool on_device, const u64* blk_off, const u32* blk_len, u32 n,
                      u8* out, u64 out_cap, u64* out_sizes);
static int decode_impl(dsrcgpu_ctx* ctx, const u8* dsrc, bool on_device, const u64* blk_off, const u32* blk_len, u32 n,
                       u8* out, u64 out_cap, u64* out_sizes)
{
    const int rc = decode_run(ctx, dsrc, on_device, blk_off, blk_len, n, out, out_cap, out_sizes);
    if (rc && ctx && ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);     // no copy into the caller's buffer outlives a failed call
    return rc;
}
static int decode_run(dsrcgpu_ctx* ctx, const u8* dsrc, bool on_device, const u64* blk_off, const u32* blk_len, u32 n,
                      u8* out, u64 out_cap, u64* out_sizes)
{
    if (!ctx) return DSRCGPU_E_ARG;
    if (!dsrc || !blk_off || !blk_len || !out || !out_sizes) { ctx->err = "null argument"; return DSRCGPU_E_ARG; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return DSRCGPU_E_CUDA; }
    memset(ctx->k_ms, 0, sizeof(ctx->k_ms)); memset(ctx->k_launches, 0, sizeof(ctx->k_launches));
    if (!ctx->call_a) { cudaEventCreate(&ctx->call_a); cudaEventCreate(&ctx->call_b); }
    Slot& sl = ctx->slots[0];
    cudaEventRecord(ctx->call_a, sl.stream);
    // full-size tables of the chain rows for this configuration (second-chance decode of blocks whose hash filled up)
    u64 big = 4ull << 20;
    if (ctx->cs.quality_order == 2) big = 64ull << 20;
    if (ctx->cs.dna_order) { const u32 o = ctx->cs.dna_order, o8 = o > 7 ? 7 : o; big = std::max<u64>(big, std::max<u64>((1ull << (2 * o)) * 8, (1ull << (3 * o8)) * 16)); }
    // chain arenas come in tiers: every block is first decoded with a small context hash (many chains in flight -- the chains
    // are latency bound, so throughput is the number of chains); a chain that fills its hash retries in the next tier
    u64 tier0 = 512ull << 10;
    if (const char* e = getenv("DSRCGPU_DEC_ARENA_KB")) tier0 = std::max<u64>(64, (u64)atoll(e)) << 10;
    const int NT = 4;
    const u64 tiers[NT] = {tier0, std::max<u64>(tier0, 3ull << 20), std::max<u64>(tier0, 16ull << 20), std::max<u64>(big, 16ull << 20)};
    const u32 pools[NT] = {8192u, 32768u, 131072u, 262144u};        // Huffman nodes (8 B) per block
    u64 budget = 32ull << 30;                             // HBM given to the chains' arenas: bounds the chains in flight
    if (const char* e = getenv("DSRCGPU_DEC_BUDGET_GB")) budget = (u64)std::max(1, atoi(e)) << 30;
    int start_tier = 0;                                   // raised when most blocks of a batch had to retry (large-alphabet data)
    u32 dec_batch = 65536;
    if (const char* e = getenv("DSRCGPU_DEC_BATCH")) dec_batch = (u32)std::max(1, atoi(e));
    u32 per_batch0 = (u32)std::max<u64>(1, std::min<u64>(dec_batch, budget / tiers[0]));
    per_batch0 = (n + (n + per_batch0 - 1) / per_batch0 - 1) / ((n + per_batch0 - 1) / per_batch0);   // equal batches: the chains of a small last batch would be latency-bound
    std::vector<u32> idx, status, sizes, retry; std::vector<u64> offs, ooffs;
    u64 out_pos = 0;
    u32 dec_batches = 0;
    for (u32 first = 0; first < n;) {
        const u32 per_batch = start_tier == 0 ? per_batch0 : (u32)std::max<u64>(1, std::min<u64>(dec_batch, budget / (tiers[start_tier] + (u64)pools[start_tier] * 8)));
        const u32 cnt = std::min(per_batch, n - first);
        idx.resize(cnt); offs.resize(cnt); ooffs.resize(cnt); status.resize(cnt); sizes.resize(cnt);
        for (u32 i = 0; i < cnt; ++i) idx[i] = first + i;
        const u8* d_in; u8* d_out; u64 cap;
        if (on_device) { d_in = dsrc; for (u32 i = 0; i < cnt; ++i) offs[i] = blk_off[first + i]; d_out = out; cap = out_cap; }
        else {
            u64 p = 0;
            for (u32 i = 0; i < cnt; ++i) p += align_up(blk_len[first + i], 16);
            CK(sl.in.ensure(p + 16));
            p = 0;
            // compressed blocks of a batch are normally adjacent in the archive: coalesce adjacent ones into one copy
            for (u32 i = 0; i < cnt;) {
                u32 j = i; u64 span = blk_len[first + i];
                while (j + 1 < cnt && blk_off[first + j + 1] == blk_off[first + j] + blk_len[first + j]) { ++j; span += blk_len[first + j]; }
                CK(cudaMemcpyAsync((u8*)sl.in.p + p, dsrc + blk_off[first + i], span, cudaMemcpyHostToDevice, sl.stream));
                u64 q = p;
                for (u32 k = i; k <= j; ++k) { offs[k] = q; q += blk_len[first + k]; }
                p = align_up(q, 16); i = j + 1;
            }
            d_in = (const u8*)sl.in.p; d_out = nullptr; cap = 0;
        }
        // sizes first: every block stores chunkSize (BlockCompressor.cpp:302-308) -- probe pass inside decode_batch gives them, but the
        // output offsets are needed before; read the 4 bytes at offset 12 of each block on the host when the input is host memory
        u64 batch_bytes = 0;
        if (!on_device) {
            for (u32 i = 0; i < cnt; ++i) {
                const u8* b = dsrc + blk_off[first + i];
                const u32 chunk = blk_len[first + i] >= 16 ? (((u32)b[12] << 24) | ((u32)b[13] << 16) | ((u32)b[14] << 8) | b[15]) : 0;
                ooffs[i] = batch_bytes; batch_bytes += (u64)chunk + 1;
            }
            // two staging buffers in turn: the copy of a batch's FASTQ to the host runs on the copy stream while the next batch is decoded
            DevBuf& ob = (dec_batches & 1u) ? ctx->dec_out2 : sl.out;
            if (!ctx->copy_stream) {
                CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
                for (cudaEvent_t& e : ctx->ev_copy) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            }
            if (dec_batches >= 2) CK(cudaEventSynchronize(ctx->ev_copy[dec_batches & 1u]));      // its previous contents have left
            CK(ob.ensure(batch_bytes + 16));
            d_out = (u8*)ob.p; cap = batch_bytes;
            if (out_pos + batch_bytes > out_cap) { ctx->err = "output buffer too small"; return DSRCGPU_E_CAPACITY; }
        } else {
            // device-resident input: one probe launch reads every block header (ReadMetaData) for the output offsets
            int prc = ensure_host(ctx, sl, cnt);
            if (prc) return prc;
            CK(sl.desc.ensure(sizeof(BlockDesc) * cnt)); CK(sl.state.ensure(sizeof(BlockState) * cnt)); CK(sl.probe.ensure(sizeof(BlockProbe) * cnt));
            for (u32 i = 0; i < cnt; ++i) { memset(&sl.h_desc[i], 0, sizeof(BlockDesc)); sl.h_desc[i].in_off = offs[i]; sl.h_desc[i].in_len = blk_len[first + i]; }
            Workspace pw{};
            pw.in = d_in; pw.desc = (const BlockDesc*)sl.desc.p; pw.state = (BlockState*)sl.state.p; pw.probe = (BlockProbe*)sl.probe.p; pw.n_blocks = cnt; pw.calc_crc = ctx->cs.calc_crc32 ? 1u : 0u;
            CK(cudaMemcpyAsync(sl.desc.p, sl.h_desc, sizeof(BlockDesc) * cnt, cudaMemcpyHostToDevice, sl.stream));
            launch_dec_probe(pw, sl.stream);
            CK(cudaMemcpyAsync(sl.h_probe, sl.probe.p, sizeof(BlockProbe) * cnt, cudaMemcpyDeviceToHost, sl.stream));
            CK(cudaStreamSynchronize(sl.stream));
            for (u32 i = 0; i < cnt; ++i) { ooffs[i] = out_pos + batch_bytes; batch_bytes += sl.h_probe[i].n_fields ? sl.h_probe[i].n_fields : 1; }
        }
        int rc = decode_batch(ctx, sl, d_in, idx.data(), offs.data(), blk_len, ooffs.data(), cnt, d_out, cap, tiers[start_tier], pools[start_tier], status.data(), sizes.data());
        if (rc) return rc;
        retry.clear();
        for (u32 i = 0; i < cnt; ++i) {
            if (status[i] == DEC_ST_RETRY) { retry.push_back(i); continue; }
            if (status[i] != ST_OK) { int e = status_to_error(ctx, status[i], first + i); if (e == DSRCGPU_E_MALFORMED) ctx->err += " (corrupt compressed block)"; return e; }
        }
        const size_t n_retry0 = retry.size();
        for (int tier = start_tier + 1; tier < NT && !retry.empty(); ++tier) {
            const u32 group = (u32)std::max<u64>(1, budget / (tiers[tier] + (u64)pools[tier] * 8));
            std::vector<u32> again;
            for (size_t g0 = 0; g0 < retry.size(); g0 += group) {
                const u32 gn = (u32)std::min<size_t>(group, retry.size() - g0);
                std::vector<u32> gi(gn), gs(gn), gz(gn); std::vector<u64> go(gn), goo(gn);
                for (u32 k = 0; k < gn; ++k) { const u32 i = retry[g0 + k]; gi[k] = first + i; go[k] = offs[i]; goo[k] = ooffs[i]; }
                rc = decode_batch(ctx, sl, d_in, gi.data(), go.data(), blk_len, goo.data(), gn, d_out, cap, tiers[tier], pools[tier], gs.data(), gz.data());
                if (rc) return rc;
                for (u32 k = 0; k < gn; ++k) {
                    if (gs[k] == DEC_ST_RETRY && tier < NT - 1) again.push_back(retry[g0 + k]);
                    else if (gs[k] != ST_OK) return status_to_error(ctx, gs[k], gi[k]);
                }
            }
            retry.swap(again);
        }
        for (u32 i = 0; i < cnt; ++i) out_sizes[first + i] = sizes[i];
        if (!on_device) {
            // (decode_batch has waited for the batch's kernels)
            CK(cudaMemcpyAsync(out + out_pos, d_out, batch_bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
            CK(cudaEventRecord(ctx->ev_copy[dec_batches & 1u], ctx->copy_stream));
        }
        ++dec_batches;
        out_pos += batch_bytes;
        if (start_tier < 1 && n_retry0 * 2 > cnt) start_tier = 1;
        first += cnt;
    }
    if (!on_device && ctx->copy_stream) {                 // the last output copies
        CK(cudaStreamSynchronize(ctx->copy_stream));
    }
    cudaEventRecord(ctx->call_b, sl.stream);
    CK(cudaEventSynchronize(ctx->call_b));
    cudaEventElapsedTime(&ctx->call_ms, ctx->call_a, ctx->call_b);
    return DSRCGPU_OK;
}

extern "C" int dsrcgpu_decode_blocks(dsrcgpu_ctx* ctx, const uint8_t* dsrc, const uint64_t* blk_off, const uint32_t* blk_len,
                                     uint32_t n, uint8_t* fastq_out, uint64_t out_cap, uint64_t* out_sizes)
{
    return decode_impl(ctx, dsrc, false, blk_off, blk_len, n, fastq_out, out_cap, out_sizes);
}
extern "C" int dsrcgpu_decode_blocks_device(dsrcgpu_ctx* ctx, const uint8_t* d_dsrc, const uint64_t* blk_off, const uint32_t* blk_len,
                                            uint32_t n, uint8_t* d_fastq_out, uint64_t out_cap, uint64_t* out_sizes)
{
    return decode_impl(ctx, d_dsrc, true, blk_off, blk_len, n, d_fastq_out, out_cap, out_sizes);
}

extern "C" int dsrcgpu_release_workspace(dsrcgpu_ctx* ctx)
{
    if (!ctx) return DSRCGPU_E_ARG;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < ctx->n_slots; ++i) {
        Slot& sl = ctx->slots[i];
        CK(cudaStreamSynchronize(sl.stream));
        DevBuf* bufs[] = {&sl.in, &sl.desc, &sl.state, &sl.result, &sl.probe, &sl.lines, &sl.qcat, &sl.dcat, &sl.trip_q, &sl.trip_d, &sl.ftab, &sl.streams, &sl.out,
                          &sl.r_title_off, &sl.r_seq_off, &sl.r_qua_off, &sl.r_title_len, &sl.r_qua_len, &sl.r_dna_len, &sl.r_trunc_len, &sl.r_qcat_off, &sl.r_dcat_off,
                          &sl.elem_a, &sl.elem_b, &sl.tagpool, &sl.q0_arena, &sl.queue};
        for (DevBuf* b : bufs) b->release();
    }
    ctx->dec_arena.release(); ctx->tab.release(); ctx->tab_mask.release();
    return DSRCGPU_OK;
}

extern "C" float dsrcgpu_last_call_ms(dsrcgpu_ctx* ctx) { return ctx ? ctx->call_ms : 0.f; }

// IFastqStreamReader::ReadNextChunk + GetNextRecordPos (src/FastqStream.cpp:18-98) over an in-memory file: the block
// queue the reference's reader thread would produce for a chunk buffer of `cbuf` bytes (CLI: -b MB << 20).
static void cut_skip_eol(const u8* d, u64& p, u64 size, bool& crlf)
{
    while (p < size && d[p] != '\n' && d[p] != '\r') ++p;
    if (p + 1 < size && d[p] == '\r' && d[p + 1] == '\n') { crlf = true; ++p; }
}
extern "C" uint64_t dsrcgpu_cut_blocks(const uint8_t* data, uint64_t size, uint64_t cbuf, uint64_t* off, uint32_t* len, uint64_t max_blocks)
{
    uint32_t state = 0;
    return dsrcgpu_cut_blocks_window(data, size, cbuf, off, len, max_blocks, &state);
}
extern "C" uint64_t dsrcgpu_cut_blocks_window(const uint8_t* data, uint64_t size, uint64_t cbuf, uint64_t* off, uint32_t* len, uint64_t max_blocks,
                                              uint32_t* reader_state)
{
    u64 start = 0, nb = 0; bool crlf = reader_state && (*reader_state & 1u);
    bool tail_only = false;     // the previous full window ended exactly at EOF: the next Read() returns 0 and the chunk is the
                                // carry-over as it stands -- no "- 1" for the final newline, no CRLF adjustment (FastqStream.cpp:41-69)
    if (cbuf <= 8192) return 0;
    while (start < size) {
        const u64 avail = size - start;
        u64 blk, adv;
        if (!tail_only && avail >= cbuf) {
            // the window is full: resume 8 KiB before its end, cut at the next line that starts a record
            const u8* w = data + start; u64 p = cbuf - 8192;
            cut_skip_eol(w, p, cbuf, crlf); ++p;
            while (p < cbuf && w[p] != '@') { cut_skip_eol(w, p, cbuf, crlf); ++p; }
            u64 cand = p;
            cut_skip_eol(w, p, cbuf, crlf); ++p;
            if (p < cbuf && w[p] == '@') cand = p;      // the first '@' line was a quality string
            adv = cand; blk = cand - 1 - (crlf ? 1 : 0);
            tail_only = avail == cbuf;
        } else {
            const u64 drop = tail_only ? 0 : 1 + (crlf ? 1 : 0);
            adv = avail; blk = avail > drop ? avail - drop : 0;
        }
        if (nb < max_blocks && off && len) { off[nb] = start; len[nb] = (u32)blk; }
        ++nb; start += adv;
    }
    if (reader_state && off && len) *reader_state = crlf ? 1u : 0u;      // only the pass that stores blocks advances the reader
    return nb;
}

// a range-coder chain that ran out of its 1.25 bytes per symbol (enqueue_batch): the call is repeated once with the wide arenas, which
// no chain can outgrow, and the context keeps them
static int encode_retry(dsrcgpu_ctx* ctx, const u8* fastq, bool on_device, const u64* blk_off, const u32* blk_len, const u32* blk_tagcap, u32 n,
                        u8* out, u64 out_cap, u32* out_sizes, u64* raw_sizes, u64* comp_sizes)
{
    if (ctx) ctx->last_overflow = false;
    int rc = encode_impl(ctx, fastq, on_device, blk_off, blk_len, blk_tagcap, n, out, out_cap, out_sizes, raw_sizes, comp_sizes);
    if (rc == DSRCGPU_E_UNSUPPORTED && ctx->last_overflow && !ctx->wide_streams && (ctx->cs.dna_order || ctx->cs.quality_order)) {
        ctx->wide_streams = true;
        rc = encode_impl(ctx, fastq, on_device, blk_off, blk_len, blk_tagcap, n, out, out_cap, out_sizes, raw_sizes, comp_sizes);
    }
    return rc;
}

extern "C" int dsrcgpu_encode_blocks(dsrcgpu_ctx* ctx, const uint8_t* fastq, const uint64_t* blk_off, const uint32_t* blk_len,
                                     const uint32_t* blk_tagcap, uint32_t n, uint8_t* out, uint64_t out_cap,
                                     uint32_t* out_sizes, uint64_t* raw_stream_sizes, uint64_t* comp_stream_sizes)
{
    return encode_retry(ctx, fastq, false, blk_off, blk_len, blk_tagcap, n, out, out_cap, out_sizes, raw_stream_sizes, comp_stream_sizes);
}
extern "C" int dsrcgpu_encode_blocks_device(dsrcgpu_ctx* ctx, const uint8_t* d_fastq, const uint64_t* blk_off, const uint32_t* blk_len,
                                            const uint32_t* blk_tagcap, uint32_t n, uint8_t* d_out, uint64_t out_cap,
                                            uint32_t* out_sizes, uint64_t* raw_stream_sizes, uint64_t* comp_stream_sizes)
{
    return encode_retry(ctx, d_fastq, true, blk_off, blk_len, blk_tagcap, n, d_out, out_cap, out_sizes, raw_stream_sizes, comp_stream_sizes);
}

// ---- Q1 helpers (pure host) ----
extern "C" uint32_t dsrcgpu_tag_field_count(const uint8_t* t, uint32_t len)
{
    u32 nf = 1;
    for (u32 i = 0; i < len; ++i) {
        u8 c = t[i];
        nf += (c == ' ' || c == '.' || c == '_' || c == ',' || c == '=' || c == ':' || c == '/' || c == '-' || c == '#' || c == 0);
    }
    return nf;
}
extern "C" uint32_t dsrcgpu_tag_capacity_after(uint32_t cap, uint32_t n_fields)
{
    for (u32 k = 0; k < n_fields; ++k) if (k == cap) cap = cap ? cap * 2 : 1;
    return cap;
}

// ---- memory helpers ----
extern "C" int dsrcgpu_device_alloc(dsrcgpu_ctx* ctx, uint64_t bytes, void** p) { cudaSetDevice(ctx->device); CK(cudaMalloc(p, bytes)); return DSRCGPU_OK; }
extern "C" int dsrcgpu_device_free(dsrcgpu_ctx* ctx, void* p) { cudaSetDevice(ctx->device); CK(cudaFree(p)); return DSRCGPU_OK; }
extern "C" int dsrcgpu_memcpy_h2d(dsrcgpu_ctx* ctx, void* d, const void* h, uint64_t bytes)
{
    cudaSetDevice(ctx->device);
    CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->slots[0].stream)); CK(cudaStreamSynchronize(ctx->slots[0].stream)); return DSRCGPU_OK;
}
extern "C" int dsrcgpu_memcpy_d2h(dsrcgpu_ctx* ctx, void* h, const void* d, uint64_t bytes)
{
    cudaSetDevice(ctx->device);
    CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->slots[0].stream)); CK(cudaStreamSynchronize(ctx->slots[0].stream)); return DSRCGPU_OK;
}
extern "C" int dsrcgpu_host_alloc(uint64_t bytes, void** p) { return cudaHostAlloc(p, bytes, cudaHostAllocDefault) == cudaSuccess ? DSRCGPU_OK : DSRCGPU_E_NOMEM; }
extern "C" int dsrcgpu_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? DSRCGPU_OK : DSRCGPU_E_CUDA; }
