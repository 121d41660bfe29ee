#!/usr/bin/env python
"""bench.py -- FASTQ compress throughput of the B200 block codec (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W          # ours
    python bench.py --impl reference ...                    # the reference's CPU path (oracle/_ref), same JSON line

A "step" is one pass of the hot path (BlockCompressor::Store for every block) over the whole synthetic workload:
configs[1] of BASELINE.json -- 50 M synthetic 150 bp Illumina-shape reads per GPU, -d2 -q2, 256 KB blocks.
`value`  : inputs resident in HBM when the timed region starts (dsrcgpu_encode_blocks_device)
`e2e`    : same metric through the host-buffer C-ABI call (dsrcgpu_encode_blocks): pinned host FASTQ in, compressed
           blocks out in host memory, H2D/D2H inside the timed region
`roofline`: dominant kernel, CUDA-event time measured inside the timed steps, against MEASURED_PEAKS.json
`cpu_baseline`: the unmodified reference's BlockCompressor (oracle/_ref) on the host cores, bounded sample
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REC_BYTES = 372
BLOCK_BYTES = 256 << 10
DNA_ORDER, QUA_ORDER = 6, 2          # -d2 -q2
METRIC = "fastq_compress_MBps"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.samples.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max(int(s[1]) for s in self.samples if s[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 2 + k and s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.samples)}


def reference_cpu_rate(host_ptr, offs, lens, n_sample, threads):
    """the unmodified reference's BlockCompressor::Store (oracle/_ref/libdsrcref.so), one instance per thread like
    DsrcCompressorMT's workers (src/DsrcWorker.cpp:30-72); returns (MB/s, seconds, bytes)."""
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libdsrcref.so"))
    lib.ref_bc_create.restype = C.c_void_p
    lib.ref_bc_create.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
    lib.ref_bc_destroy.argtypes = [C.c_void_p]
    lib.ref_bc_store.restype = C.c_longlong
    lib.ref_bc_store.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_void_p]
    n_sample = min(n_sample, len(offs))
    total = int(sum(int(lens[i]) for i in range(n_sample)))
    start_evt = threading.Event()

    def work(tid):
        h = lib.ref_bc_create(33, 0, 0, DNA_ORDER, QUA_ORDER, 0, 0)
        out = (C.c_uint8 * (BLOCK_BYTES * 2))()
        start_evt.wait()
        for i in range(tid, n_sample, threads):
            lib.ref_bc_store(h, C.c_void_p(host_ptr + int(offs[i])), int(lens[i]), out, len(out), None, None)
        lib.ref_bc_destroy(h)

    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for t in ts:
        t.start()
    time.sleep(0.05)
    t0 = time.perf_counter()
    start_evt.set()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    return total / dt / 1e6, dt, total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--reads", type=int, default=50_000_000, help="reads per GPU (BASELINE configs[1]: 50 M)")
    ap.add_argument("--profile", type=int, default=0)
    ap.add_argument("--inflight", type=int, default=8192, help="blocks per batch (one batch per stream slot); the range-coder chains of a batch take the same time for 1 or 16 K blocks")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-serial", action="store_true", help="skip the extra single-slot pass that isolates per-kernel durations")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, world)
    config = {"workload": "configs[1]: %d synthetic 150bp Illumina-shape reads per GPU (profile %d, seed 99), -d2 -q2, 256KB blocks"
                          % (args.reads, args.profile),
              "reads_per_gpu": args.reads, "block_bytes": BLOCK_BYTES, "dna_order": DNA_ORDER, "quality_order": QUA_ORDER,
              "l2": "inputs (%.1f GB per GPU) exceed the 126 MB L2; no explicit flush" % (args.reads * REC_BYTES / 1e9),
              "parallelism": "blocks sharded over %d GPU(s), no data-path collective; NCCL all_gather of block sizes for the footer" % n_gpus}

    from dsrc_b200 import _lib
    L = _lib.lib()

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdsrcref.so")):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libdsrcref.so not built"}))
            return
        threads = os.cpu_count() or 1
        n_reads = min(args.reads, threads * 400 * 704)
        buf = np.empty(n_reads * REC_BYTES, dtype=np.uint8)
        nb = C.c_uint64()
        L.dsrcgpu_synth_fastq_host(args.profile, 99, 0, n_reads, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(nb))
        n = L.dsrcgpu_cut_blocks(buf.ctypes.data_as(C.c_void_p), buf.size, BLOCK_BYTES, None, None, 0)
        offs = np.zeros(n, dtype=np.uint64)
        lens = np.zeros(n, dtype=np.uint32)
        L.dsrcgpu_cut_blocks(buf.ctypes.data_as(C.c_void_p), buf.size, BLOCK_BYTES, offs.ctypes.data_as(_lib.u64p), lens.ctypes.data_as(_lib.u32p), n)
        rates = []
        secs = []
        for it in range(args.warmup + args.steps):
            r, dt, tot = reference_cpu_rate(buf.ctypes.data, offs, lens, n, threads)
            if it >= args.warmup:
                rates.append(r)
                secs.append(dt)
        val = float(np.mean(rates))
        sample = "%d blocks (%.1f MB) of the same synthetic stream per step, reference BlockCompressor::Store on %d threads" % (n, buf.size / 1e6, threads)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "MB/s", "n_gpus": n_gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": "MB/s", "cores": threads, "kind": "reference", "sample": sample},
                          "e2e": {"value": val, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep NCCL's own log lines (e.g. its version banner) off stdout: one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ds = _lib.Dataset(33, 0, 0)
    cs = _lib.Settings(DNA_ORDER, QUA_ORDER, 0, 0, 0)
    ctx = C.c_void_p()
    rc = L.dsrcgpu_create(C.byref(ctx), local_rank, C.byref(ds), C.byref(cs), BLOCK_BYTES, args.inflight)
    if rc:
        raise SystemExit("dsrcgpu_create failed: %d" % rc)

    def check(rc_, what):
        if rc_:
            raise SystemExit("%s failed: %d %s" % (what, rc_, L.dsrcgpu_last_error(ctx).decode()))

    # synthetic shard of this rank, generated on the device; host copy in pinned memory for the cutter and the e2e leg
    n_reads = args.reads
    in_bytes = n_reads * REC_BYTES
    d_in = torch.empty(in_bytes, dtype=torch.uint8, device="cuda")
    nb = C.c_uint64()
    check(L.dsrcgpu_synth_fastq_device(ctx, args.profile, 99, rank * n_reads, n_reads, C.c_void_p(d_in.data_ptr()), in_bytes, C.byref(nb)), "synth")
    # block cut (a host function: IFastqStreamReader::ReadNextChunk) over a sliding pinned window of the device-resident shard --
    # a full host copy per rank would not fit the box's RAM at 8 ranks
    WIN = min(in_bytes, 512 << 20)
    win = torch.empty(WIN, dtype=torch.uint8, pin_memory=True)
    cap_blocks = WIN // (BLOCK_BYTES - 8192) + 8
    woff = np.zeros(cap_blocks, dtype=np.uint64)
    wlen = np.zeros(cap_blocks, dtype=np.uint32)
    offs_l, lens_l = [], []
    start = 0
    while start < in_bytes:
        w = min(WIN, in_bytes - start)
        win[:w].copy_(d_in[start:start + w])
        torch.cuda.synchronize()
        k = int(L.dsrcgpu_cut_blocks(C.c_void_p(win.data_ptr()), w, BLOCK_BYTES, woff.ctypes.data_as(_lib.u64p), wlen.ctypes.data_as(_lib.u32p), cap_blocks))
        last = start + w == in_bytes
        take = k if last else k - 1              # the window's final block was cut at the window end, not at a record boundary rule
        if take <= 0:
            raise SystemExit("bench.py: cut window too small")
        offs_l.append(woff[:take] + np.uint64(start))
        lens_l.append(wlen[:take].copy())
        start = in_bytes if last else start + int(woff[take])
    del win
    offs = np.ascontiguousarray(np.concatenate(offs_l), dtype=np.uint64)
    lens = np.ascontiguousarray(np.concatenate(lens_l), dtype=np.uint32)
    n = len(offs)
    payload = int(lens.astype(np.uint64).sum())
    out_cap = in_bytes // 2 + (1 << 20)
    d_out = torch.empty(out_cap, dtype=torch.uint8, device="cuda")
    sizes = np.zeros(n, dtype=np.uint32)
    comp = np.zeros((n, 4), dtype=np.uint64)

    def step_device():
        check(L.dsrcgpu_encode_blocks_device(ctx, C.c_void_p(d_in.data_ptr()), offs.ctypes.data_as(_lib.u64p), lens.ctypes.data_as(_lib.u32p),
                                             None, n, C.c_void_p(d_out.data_ptr()), out_cap, sizes.ctypes.data_as(_lib.u32p), None,
                                             comp.ctypes.data_as(_lib.u64p)), "encode_device")
        if world > 1:   # the one real exchange of the path: block sizes for the single archive footer (src/DsrcFile.cpp:142)
            t = torch.from_numpy(sizes.astype(np.int32)).cuda()
            pad = torch.zeros(n + 64, dtype=torch.int32, device="cuda")
            pad[:n] = t
            outl = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(outl, pad)
        return float(L.dsrcgpu_last_call_ms(ctx))

    def kernel_times():
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        ln = (C.c_uint32 * 16)()
        k = L.dsrcgpu_last_kernel_times(ctx, names, ms, ln, 16)
        return {names[i].decode(): (float(ms[i]), int(ln[i])) for i in range(k)}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # parity gate on a few blocks before anything is timed (oracle = checker only)
    parity_blocks = 0
    if rank == 0:
        import refbind
        ora = refbind.Oracle(33, 0, DNA_ORDER, QUA_ORDER)
        k = min(n, 6)
        so = np.zeros(k, dtype=np.uint32)
        check(L.dsrcgpu_encode_blocks_device(ctx, C.c_void_p(d_in.data_ptr()), offs.ctypes.data_as(_lib.u64p), lens.ctypes.data_as(_lib.u32p),
                                             None, k, C.c_void_p(d_out.data_ptr()), out_cap, so.ctypes.data_as(_lib.u32p), None, None), "parity encode")
        got = d_out[:int(so.sum())].cpu().numpy().tobytes()
        p = 0
        hv = d_in[:int(offs[k - 1]) + int(lens[k - 1])].cpu().numpy()
        ora.store(hv[int(offs[0]):int(offs[0]) + int(lens[0])].tobytes())       # warm the oracle's field vector like blk_tagcap=NULL does
        for i in range(k):
            exp, _, _ = ora.store(hv[int(offs[i]):int(offs[i]) + int(lens[i])].tobytes())
            if got[p:p + int(so[i])] != exp:
                raise SystemExit("bench.py: parity check failed on block %d" % i)
            p += int(so[i])
            parity_blocks += 1

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    ktot = {}
    launches = 0
    for _ in range(args.steps):
        dev_ms += step_device()
        for k, (ms, ln) in kernel_times().items():
            a = ktot.get(k, (0.0, 0))
            ktot[k] = (a[0] + ms, a[1] + ln)
            launches += ln
    barrier()
    wall = time.perf_counter() - t0
    comp_bytes = int(sizes.astype(np.uint64).sum())
    tvals = torch.tensor([wall, dev_ms / 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tvals, op=dist.ReduceOp.MAX)
    wall_max, dev_max = float(tvals[0]), float(tvals[1])
    step_s = wall_max / args.steps
    value = payload * world / step_s / 1e6

    # e2e through the host-buffer call
    e2e = None
    if not args.no_e2e:
        # pinned host staging is bounded by the box's RAM shared between the ranks: the e2e leg runs on the first ne blocks
        try:
            import psutil
            host_total = psutil.virtual_memory().total
        except Exception:
            host_total = 64 << 30
        budget = int(host_total * 0.5 / max(1, world))
        ne = n
        if in_bytes + out_cap > budget:
            ends = offs + lens.astype(np.uint64)
            ne = max(1, int(np.searchsorted(ends, np.uint64(budget * 2 // 3), side="right")))
        e_in = int(offs[ne - 1]) + int(lens[ne - 1])
        e_payload = int(lens[:ne].astype(np.uint64).sum())
        e_cap = e_in // 2 + (1 << 20)
        h_in = torch.empty(e_in, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(d_in[:e_in])
        torch.cuda.synchronize()
        hp = h_in.data_ptr()
        h_out = torch.empty(e_cap, dtype=torch.uint8, pin_memory=True)
        esz = np.zeros(ne, dtype=np.uint32)

        def step_host():
            check(L.dsrcgpu_encode_blocks(ctx, C.c_void_p(hp), offs.ctypes.data_as(_lib.u64p), lens.ctypes.data_as(_lib.u32p), None, ne,
                                          C.c_void_p(h_out.data_ptr()), e_cap, esz.ctypes.data_as(_lib.u32p), None, None), "encode_host")
        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        e_wall = time.perf_counter() - t0
        tv = torch.tensor([e_wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        e2e = {"value": e_payload * world / (float(tv[0]) / args.steps) / 1e6, "unit": "MB/s",
               "h2d_bytes_per_step": int(e_in), "d2h_bytes_per_step": int(esz.astype(np.uint64).sum()),
               "blocks_per_gpu": int(ne), "note": "whole workload" if ne == n else "first %d of %d blocks per GPU (pinned staging bounded by host RAM / ranks)" % (ne, n)}
        del h_in, h_out
    # decode leg (BASELINE configs[4], bounded sample): BlockCompressor::Read of the first blocks of the archive just written,
    # device-resident, verified byte for byte against the input
    dec = None
    if not args.no_decode:
        nd = min(n, 65536)
        step_device()
        L.dsrcgpu_release_workspace(ctx)      # the encode slots' workspaces make room for the decode chains' arenas
        coffs = np.concatenate([[0], np.cumsum(sizes.astype(np.uint64))[:-1]]).astype(np.uint64)
        dbytes = int(lens[:nd].astype(np.uint64).sum()) + nd
        d_dec = torch.empty(dbytes + 64, dtype=torch.uint8, device="cuda")
        osz = np.zeros(nd, dtype=np.uint64)

        def step_decode():
            check(L.dsrcgpu_decode_blocks_device(ctx, C.c_void_p(d_out.data_ptr()), coffs.ctypes.data_as(_lib.u64p), sizes.ctypes.data_as(_lib.u32p), nd,
                                                 C.c_void_p(d_dec.data_ptr()), dbytes + 64, osz.ctypes.data_as(_lib.u64p)), "decode_device")
            return float(L.dsrcgpu_last_call_ms(ctx))
        step_decode()
        dms = min(step_decode() for _ in range(2))
        ok = bool(int(osz.sum()) == dbytes and torch.equal(d_dec[:dbytes], d_in[int(offs[0]):int(offs[0]) + dbytes]))
        dec = {"value": dbytes / dms / 1e3, "unit": "MB/s", "blocks": int(nd), "bytes": dbytes, "verified_identical": ok,
               "note": "device-resident decode of the first blocks of this step's archive, one stream, max over 2 runs"}
        del d_dec
    sampler.stop = True
    sampler.join(timeout=2)

    # one more pass of the same step with a single batch in flight: with several slots the CUDA-event pairs of a kernel include the
    # time it shares the SMs with (or waits for) the other slots' kernels, so the roofline below uses these isolated durations
    kser = None
    if not args.no_serial:
        L.dsrcgpu_release_workspace(ctx)
        old = os.environ.get("DSRCGPU_SLOTS")
        os.environ["DSRCGPU_SLOTS"] = "1"
        ctx1 = C.c_void_p()
        rc = L.dsrcgpu_create(C.byref(ctx1), local_rank, C.byref(ds), C.byref(cs), BLOCK_BYTES, args.inflight)
        if old is None:
            del os.environ["DSRCGPU_SLOTS"]
        else:
            os.environ["DSRCGPU_SLOTS"] = old
        if rc == 0:
            ssz = np.zeros(n, dtype=np.uint32)
            for _ in range(2):
                rc1 = L.dsrcgpu_encode_blocks_device(ctx1, C.c_void_p(d_in.data_ptr()), offs.ctypes.data_as(_lib.u64p), lens.ctypes.data_as(_lib.u32p),
                                                     None, n, C.c_void_p(d_out.data_ptr()), out_cap, ssz.ctypes.data_as(_lib.u32p), None, None)
            if rc1 == 0 and bool((ssz == sizes).all()):
                names = (C.c_char_p * 16)()
                ms = (C.c_float * 16)()
                ln = (C.c_uint32 * 16)()
                k = L.dsrcgpu_last_kernel_times(ctx1, names, ms, ln, 16)
                kser = {names[i].decode(): (float(ms[i]), int(ln[i])) for i in range(k)}
                kser["_step_ms"] = (float(L.dsrcgpu_last_call_ms(ctx1)), 1)
            L.dsrcgpu_destroy(ctx1)

    # roofline of the dominant kernel (CUDA-event time of its launches; isolated durations when the serialised pass ran)
    peak, peak_kind = load_peaks()
    ser_step_ms = kser.pop("_step_ms")[0] if kser else None
    kbase = {k: (v[0] * args.steps, v[1] * args.steps) for k, v in kser.items()} if kser else ktot
    dom = max(kbase.items(), key=lambda kv: kv[1][0]) if kbase else ("none", (0.0, 0))
    syms = n_reads * 150
    alg = {   # algorithmic bytes per step of each kernel family (DESIGN.md "kernels"): what the kernel must read + write
        "count_lines": payload, "parse": payload, "preprocess": payload + 2 * syms,
        "tags": n_reads * 68 + int(comp[:, 1].sum()),
        "model_quality": syms * 9, "model_dna": syms * 9,
        "rc_encode": 2 * syms * 8 + int(comp[:, 2].sum() + comp[:, 3].sum()),
        "meta_sizes": n * 64, "gather": 2 * comp_bytes,
    }
    # measured DRAM traffic per launch of the dominant kernel, from the committed `ncu --set full` capture of this command
    traffic = None
    try:
        import csv
        kmap = {"model_quality": "k_model<1>", "model_dna": "k_model<0>"}
        rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r01_ncu_bench_model_full.csv"))))
        h = rows[0]
        for r in rows[2:]:
            if kmap.get(dom[0], "?") in r[0]:
                traffic = (float(r[h.index("dram__bytes_read.sum")]) + float(r[h.index("dram__bytes_write.sum")])) * 1e9
    except Exception:
        traffic = None
    dom_ms_per_step = dom[1][0] / args.steps
    ach = alg.get(dom[0], 0) / (dom_ms_per_step / 1e3) / 1e9 if dom_ms_per_step > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": alg.get(dom[0], 0) / max(1, dom[1][1] // max(1, args.steps)),
                "note": "achieved = algorithmic bytes / CUDA-event time of the kernel's launches in one extra pass of the same step with a single batch "
                        "in flight (kernel_ms_serialized); kernel_ms_per_step are the event pairs inside the timed steps, where 3 batches are in flight "
                        "and a kernel's pair includes time shared with the other streams' kernels" if kser else
                        "kernel times are CUDA-event pairs on the launching stream; with 3 batches in flight they include time shared with other streams' kernels",
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in ktot.items()},
                "kernel_ms_serialized": {k: v[0] for k, v in kser.items()} if kser else None,
                "serialized_step_ms": ser_step_ms,
                "launches_per_kernel_per_step": max(1, dom[1][1] // max(1, args.steps)),
                "block_path_compulsory_frac": (payload + comp_bytes) / step_s / 1e9 / peak,
                "block_path_survey_A_frac": (payload + comp_bytes + syms * 64) / step_s / 1e9 / peak}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdsrcref.so")):
        threads = os.cpu_count() or 1
        ns = min(n, threads * 1500)      # ~10-20 s of CPU work
        h_cpu = d_in[:int(offs[ns - 1]) + int(lens[ns - 1])].cpu().numpy()
        r, dt, tot = reference_cpu_rate(h_cpu.ctypes.data, offs, lens, ns, threads)
        cpu = {"value": r, "unit": "MB/s", "cores": threads, "kind": "reference",
               "sample": "first %d blocks (%.1f MB) of the workload, reference BlockCompressor::Store, one instance per thread, %.1f s" % (ns, tot / 1e6, dt)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config, "clocks": sampler.summary(), "gpu_launches": launches,
                "device_ms_per_step": dev_max * 1e3 / args.steps, "ratio": payload / max(1, comp_bytes),
                "blocks_per_gpu": int(n), "parity_checked_blocks": parity_blocks,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "decode": dec}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    L.dsrcgpu_destroy(ctx)


if __name__ == "__main__":
    main()
