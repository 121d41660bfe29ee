#!/usr/bin/env python
"""bench.py -- FASTQ compress throughput of the B200 block codec (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W          # ours
    python bench.py --impl reference ...                    # the reference's CPU path (oracle/_ref), same JSON line

A "step" is one pass of the hot path (BlockCompressor::Store for every block) over the whole synthetic workload:
configs[1] of BASELINE.json -- 50 M synthetic 150 bp Illumina-shape reads per GPU, -d2 -q2, 256 KB blocks.
`value`   : inputs resident in HBM when the timed region starts (dsrcgpu_encode_blocks_device)
`e2e`     : same metric through the host-buffer C-ABI call (dsrcgpu_encode_blocks): pinned host FASTQ in, compressed
            blocks out in host memory, H2D/D2H inside the timed region; its bytes are compared with the resident leg's
`roofline`: dominant kernel, CUDA-event time measured live, against MEASURED_PEAKS.json
`cpu_baseline`: the unmodified reference's BlockCompressor (oracle/_ref) on the host cores, bounded sample
`parity`  : before anything is timed EVERY rank encodes 64 blocks spread over its shard (first / middle / last), each with a cold
            and a warm field vector, and compares them byte for byte with the oracle (checker only)
`extras`  : (N = 1) the other single-GPU workloads of BASELINE.json at reduced size, each with its own parity gate, value, e2e,
            roofline and CPU reference: 41-level Illumina data (64-symbol quality models), configs[2] = 454 / Ion shape at
            -d3 -q2, and the -q0 RLE quality path
`decode`  : BASELINE configs[4]: BlockCompressor::Read of this step's archive on every rank (device-resident and through host
            buffers), verified against the input (torch.equal over everything + SHA-256 of a sample)
`archive` : (N > 1) the multi-rank archive: ranks encode contiguous block ranges of one file, NCCL all-gather of the block sizes,
            rank 0 writes header + footer, every rank pwrites its slice; SHA-256 == the single-rank archive of the same file
"""
import argparse
import ctypes as C
import hashlib
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
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libdsrcref.so")


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


# ------------------------------------------------------------------------------------------------ the reference on the host cores
def _ref_lib():
    lib = C.CDLL(REF_LIB)
    lib.ref_bc_create.restype = C.c_void_p
    lib.ref_bc_create.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
    lib.ref_bc_destroy.argtypes = [C.c_void_p]
    lib.ref_bc_store.restype = C.c_longlong
    lib.ref_bc_store.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_void_p]
    lib.ref_bc_read.restype = C.c_longlong
    lib.ref_bc_read.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_ulonglong]
    lib.ref_compress_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint]
    lib.ref_decompress_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    return lib


def reference_cpu_rate(host_ptr, offs, lens, n_sample, threads, d_order=DNA_ORDER, q_order=QUA_ORDER, keep=None):
    """the unmodified reference's BlockCompressor::Store (oracle/_ref/libdsrcref.so), one instance per thread like
    DsrcCompressorMT's workers (src/DsrcWorker.cpp:30-72); returns (MB/s, seconds, bytes). keep: list receiving (i, block bytes)."""
    lib = _ref_lib()
    n_sample = min(n_sample, len(offs))
    total = int(sum(int(lens[i]) for i in range(n_sample)))
    start_evt = threading.Event()

    def work(tid):
        h = lib.ref_bc_create(33, 0, 0, d_order, q_order, 0, 0)
        out = (C.c_uint8 * (int(max(lens[:n_sample])) * 2 + 65536))()
        start_evt.wait()
        for i in range(tid, n_sample, threads):
            k = lib.ref_bc_store(h, C.c_void_p(host_ptr + int(offs[i])), int(lens[i]), out, len(out), None, None)
            if keep is not None:
                keep.append((i, bytes(out[:k])))
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


def reference_cpu_decode_rate(blocks, out_bytes, threads, d_order=DNA_ORDER, q_order=QUA_ORDER):
    """BlockCompressor::Read of the given compressed blocks, one instance per thread (DsrcDecompressor::Process, src/DsrcWorker.cpp:74-108)"""
    lib = _ref_lib()
    start_evt = threading.Event()

    def work(tid):
        h = lib.ref_bc_create(33, 0, 0, d_order, q_order, 0, 0)
        out = (C.c_uint8 * (2 << 20))()
        start_evt.wait()
        for i in range(tid, len(blocks), threads):
            lib.ref_bc_read(h, blocks[i], len(blocks[i]), out, len(out))
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
    return out_bytes / dt / 1e6, dt


def reference_native_rate(host_arr, d_level, q_level, threads, buf_mb=8):
    """the reference as it ships: DsrcCompressorMT::Process (src/DsrcOperator.cpp:230-394), file to file on tmpfs, its own reader /
    worker pool / ordered writer, at its default 8 MB chunk buffer (the CLI cannot go below 1 MB, src/main.cpp:300)."""
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    src = os.path.join(tmp, "dsrc_b200_bench_%d.fq" % os.getpid())
    dst = src + ".dsrc"
    try:
        host_arr.tofile(src)
        lib = _ref_lib()
        t0 = time.perf_counter()
        rc = lib.ref_compress_file(src.encode(), dst.encode(), d_level, q_level, buf_mb, min(threads, 64), 0)
        dt = time.perf_counter() - t0
        if rc != 0:
            return None
        return {"value": host_arr.size / dt / 1e6, "unit": "MB/s", "cores": min(threads, 64), "seconds": dt, "bytes": int(host_arr.size),
                "how": "DsrcCompressorMT::Process, -d%d -q%d -b%d -t%d, file to file on %s" % (d_level, q_level, buf_mb, min(threads, 64), tmp)}
    finally:
        for f in (src, dst):
            if os.path.exists(f):
                os.remove(f)


# ------------------------------------------------------------------------------------------------ the reference arm
def reference_arm(args, n_gpus, config):
    """Times the reference's own CPU implementation of the path. This process never loads the product library: the input comes
    from oracle/libbenchsynth.so (the generator's record functions compiled for the host), the block cut from the oracle."""
    if not os.path.exists(REF_LIB):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libdsrcref.so not built"}))
        return
    syn = C.CDLL(os.path.join(ROOT, "oracle", "libbenchsynth.so"))
    syn.bench_synth_fastq_host.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    ora = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    ora.dsrc_oracle_cut_blocks.restype = C.c_uint64
    ora.dsrc_oracle_cut_blocks.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint64]
    threads = os.cpu_count() or 1
    n_reads = min(args.reads, threads * 400 * 704)
    buf = np.empty(n_reads * REC_BYTES, dtype=np.uint8)
    nb = C.c_uint64()
    syn.bench_synth_fastq_host(args.profile, 99, 0, n_reads, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(nb))
    n = int(ora.dsrc_oracle_cut_blocks(buf.ctypes.data_as(C.c_void_p), buf.size, BLOCK_BYTES, None, None, 0))
    offs = np.zeros(n + 1, dtype=np.uint64)
    lens = np.zeros(n + 1, dtype=np.uint64)
    ora.dsrc_oracle_cut_blocks(buf.ctypes.data_as(C.c_void_p), buf.size, BLOCK_BYTES, offs.ctypes.data_as(C.POINTER(C.c_uint64)),
                               lens.ctypes.data_as(C.POINTER(C.c_uint64)), n)
    rates, secs = [], []
    for it in range(args.warmup + args.steps):
        r, dt, tot = reference_cpu_rate(buf.ctypes.data, offs, lens, n, threads)
        if it >= args.warmup:
            rates.append(r)
            secs.append(dt)
    val = float(np.mean(rates))
    kb = max(0, int(np.searchsorted(offs[:n] + lens[:n], np.uint64(threads * (96 << 20)), side="right")) - 1)
    native = reference_native_rate(buf[:int(offs[kb] + lens[kb]) + 1], DNA_ORDER // 3, QUA_ORDER, threads)     # whole records: up to a block's end
    sample = "%d blocks (%.1f MB) of the same synthetic stream per step, reference BlockCompressor::Store on %d threads" % (n, buf.size / 1e6, threads)
    config = dict(config)
    config["reference_note"] = ("256 KB blocks are the reference's worst operating point (a 32 MB model Clear() per block, src/QualityEncoder.h:45-51; "
                                "its CLI cannot go below 1 MB): `native_b8` is the reference as it ships (DsrcCompressorMT, 8 MB blocks) on the same data")
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "MB/s", "n_gpus": n_gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                      "cpu_baseline": {"value": val, "unit": "MB/s", "cores": threads, "kind": "reference", "sample": sample},
                      "native_b8": native,
                      "e2e": {"value": val, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ our arm: helpers
class Shard:
    """one rank's synthetic FASTQ, generated on the device, plus the block queue the reference's reader would cut from it"""

    def __init__(self, L, _lib, ctx, torch, profile, first_read, n_reads, check):
        rec_cap = REC_BYTES if profile < 2 else 800
        cap = n_reads * rec_cap + (1 << 20)
        self.d_in = torch.empty(cap, dtype=torch.uint8, device="cuda")
        nb = C.c_uint64()
        check(L.dsrcgpu_synth_fastq_device(ctx, profile, 99, first_read, n_reads, C.c_void_p(self.d_in.data_ptr()), cap, C.byref(nb)), "synth")
        self.in_bytes = in_bytes = int(nb.value)
        self.n_reads = n_reads
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
            win[:w].copy_(self.d_in[start:start + w])
            torch.cuda.synchronize()
            k = int(L.dsrcgpu_cut_blocks(C.c_void_p(win.data_ptr()), w, BLOCK_BYTES, woff.ctypes.data_as(_lib.u64p), wlen.ctypes.data_as(_lib.u32p), cap_blocks))
            last = start + w == in_bytes
            take = k if last else k - 1              # the window's final block was cut at the window end, not at a record boundary
            if take <= 0:
                raise SystemExit("bench.py: cut window too small")
            offs_l.append(woff[:take] + np.uint64(start))
            lens_l.append(wlen[:take].copy())
            start = in_bytes if last else start + int(woff[take])
        del win
        self.offs = np.ascontiguousarray(np.concatenate(offs_l), dtype=np.uint64)
        self.lens = np.ascontiguousarray(np.concatenate(lens_l), dtype=np.uint32)
        self.n = len(self.offs)
        self.payload = int(self.lens.astype(np.uint64).sum())
        self.symbols = None

    def block_bytes(self, i):
        o, l = int(self.offs[i]), int(self.lens[i])
        return self.d_in[o:o + l].cpu().numpy().tobytes()


def parity_gate(L, _lib, ctx, torch, sh, d_order, q_order, d_out, out_cap, check, want=64):
    """encodes `want` blocks spread over the shard (first quarter of them from the start, half from the middle, the rest from the
    end), every one with a cold field vector (tag capacity 0) and with a warm one, and compares with the oracle byte for byte"""
    import refbind
    n = sh.n
    k = min(want, n)
    a = k // 4
    idx = sorted(set(list(range(a)) + [int(x) for x in np.linspace(a, max(a, n - 1 - a), k - 2 * a)] + list(range(n - a, n))))
    idx = np.array([i for i in idx if 0 <= i < n], dtype=np.int64)
    k = len(idx)
    so = np.ascontiguousarray(sh.offs[idx])
    sl = np.ascontiguousarray(sh.lens[idx])
    inputs = [sh.block_bytes(int(i)) for i in idx]
    checked = 0
    for cold in (True, False):
        caps = np.zeros(k, dtype=np.uint32)
        ssz = np.zeros(k, dtype=np.uint32)
        check(L.dsrcgpu_encode_blocks_device(ctx, C.c_void_p(sh.d_in.data_ptr()), so.ctypes.data_as(_lib.u64p), sl.ctypes.data_as(_lib.u32p),
                                             caps.ctypes.data_as(_lib.u32p) if cold else None, k, C.c_void_p(d_out.data_ptr()), out_cap,
                                             ssz.ctypes.data_as(_lib.u32p), None, None), "parity encode")
        got = d_out[:int(ssz.astype(np.uint64).sum())].cpu().numpy().tobytes()
        p = 0
        for j in range(k):
            ora = refbind.Oracle(33, 0, d_order, q_order)
            exp, _, _ = ora.store(inputs[j])
            if not cold:
                exp, _, _ = ora.store(inputs[j])          # second call of one compressor: warm field vector
            if got[p:p + int(ssz[j])] != exp:
                raise SystemExit("bench.py: parity check failed on block %d (%s)" % (int(idx[j]), "cold" if cold else "warm"))
            p += int(ssz[j])
            checked += 1
    return checked, [int(i) for i in idx[:3]] + ["..."] + [int(i) for i in idx[-2:]]


def kernel_times(L, ctx):
    names = (C.c_char_p * 24)()
    ms = (C.c_float * 24)()
    ln = (C.c_uint32 * 24)()
    k = L.dsrcgpu_last_kernel_times(ctx, names, ms, ln, 24)
    return {names[i].decode(): (float(ms[i]), int(ln[i])) for i in range(k)}


def serialized_pass(L, _lib, local_rank, ds, cs, inflight, sh, d_out, out_cap, sizes):
    """one more pass of the same step with a single batch in flight: with several slots the CUDA-event pairs of a kernel include the
    time it shares the SMs with (or waits for) the other slots' kernels, so the roofline uses these isolated durations"""
    old = os.environ.get("DSRCGPU_SLOTS")
    os.environ["DSRCGPU_SLOTS"] = "1"
    ctx1 = C.c_void_p()
    rc = L.dsrcgpu_create(C.byref(ctx1), local_rank, C.byref(ds), C.byref(cs), BLOCK_BYTES, inflight)
    if old is None:
        del os.environ["DSRCGPU_SLOTS"]
    else:
        os.environ["DSRCGPU_SLOTS"] = old
    if rc:
        return None
    kser = None
    ssz = np.zeros(sh.n, dtype=np.uint32)
    for _ in range(2):
        rc1 = L.dsrcgpu_encode_blocks_device(ctx1, C.c_void_p(sh.d_in.data_ptr()), sh.offs.ctypes.data_as(_lib.u64p), sh.lens.ctypes.data_as(_lib.u32p),
                                             None, sh.n, C.c_void_p(d_out.data_ptr()), out_cap, ssz.ctypes.data_as(_lib.u32p), None, None)
    if rc1 == 0 and bool((ssz == sizes).all()):
        kser = kernel_times(L, ctx1)
        kser["_step_ms"] = (float(L.dsrcgpu_last_call_ms(ctx1)), 1)
    else:
        sys.stderr.write("bench.py: serialized pass skipped (rc %d: %s; sizes equal: %s)\n" % (rc1, L.dsrcgpu_last_error(ctx1).decode(), bool((ssz == sizes).all())))
    L.dsrcgpu_destroy(ctx1)
    return kser


def algorithmic_bytes(sh, symbols, comp, comp_bytes):
    """what each kernel family must read + write per step (DESIGN.md "kernels")"""
    return {"count_lines": sh.payload, "parse": sh.payload, "preprocess": sh.payload + 2 * symbols,
            "tags": sh.n_reads * 68 + int(comp[:, 1].sum()),
            "model_quality": symbols * 9, "model_dna": symbols * 9,
            "rc_encode": 2 * symbols * 8 + int(comp[:, 2].sum() + comp[:, 3].sum()),
            "q0_quality": 2 * symbols + int(comp[:, 3].sum()), "d0_dna": 2 * symbols + int(comp[:, 2].sum()),
            "meta_sizes": sh.n * 64, "gather": 2 * comp_bytes}


def make_roofline(kbase, ktot, kser_step_ms, steps, alg, peak, peak_kind, payload, comp_bytes, symbols, step_s, traffic_of):
    dom = max(kbase.items(), key=lambda kv: kv[1][0]) if kbase else ("none", (0.0, 0))
    dom_ms_per_step = dom[1][0] / steps
    ach = alg.get(dom[0], 0) / (dom_ms_per_step / 1e3) / 1e9 if dom_ms_per_step > 0 else 0.0
    # (the model timers bracket several kernels per batch: walk / partition / tile engine launches, each taking its own blocks)
    launches = max(1, dom[1][1] // max(1, steps) // {"model_quality": 3, "model_dna": 2}.get(dom[0], 1))
    return {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
            "frac": ach / peak, "traffic": traffic_of(dom[0]),
            "algorithmic_bytes_per_launch": alg.get(dom[0], 0) / launches,
            "note": "achieved = algorithmic bytes (1 B symbol in + 8 B triple out per range-coded symbol for the model kernels) / CUDA-event time of the "
                    "kernel's launches in one extra pass of the same step with a single batch in flight (kernel_ms_serialized); kernel_ms_per_step are the "
                    "event pairs inside the timed steps, where 3 batches are in flight and a pair includes time shared with the other streams' kernels",
            "kernel_ms_per_step": {k: v[0] / steps for k, v in ktot.items()},
            "kernel_ms_serialized": {k: v[0] / steps for k, v in kbase.items()},
            "serialized_step_ms": kser_step_ms,
            "launches_per_kernel_per_step": launches,
            "block_path_compulsory_frac": (payload + comp_bytes) / step_s / 1e9 / peak}


def ncu_traffic(kernel):
    """measured DRAM bytes per launch of a kernel family from the committed `ncu --set full` capture (profiles/), or None"""
    import csv
    pats = {"model_quality": ("r02_ncu_bench_model_full.csv", "k_model_walk"), "model_dna": ("r02_ncu_bench_model_full.csv", "k_dna_walk"),
            "rc_encode": ("r02_ncu_bench_model_full.csv", "k_rc_encode"), "preprocess": ("r02_ncu_bench_model_full.csv", "k_preprocess"),
            "tags": ("r02_ncu_bench_model_full.csv", "k_tags")}
    if kernel not in pats:
        return None
    try:
        rows = list(csv.reader(open(os.path.join(ROOT, "profiles", pats[kernel][0]))))
        h = rows[0]
        for r in rows[2:]:
            if pats[kernel][1] in r[h.index("Kernel Name")]:
                return (float(r[h.index("dram__bytes_read.sum")]) + float(r[h.index("dram__bytes_write.sum")])) * 1e9
    except Exception:
        return None
    return None


def copy_ceiling(torch, nbytes_in, nbytes_out):
    """pinned host -> device and device -> pinned host copies of the e2e leg's sizes, alone and running together: the wall no
    host-buffer number can pass on this box"""
    n_in, n_out = min(nbytes_in, 2 << 30), max(1 << 20, min(nbytes_out, 512 << 20))
    h_a = torch.empty(n_in, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(n_in, dtype=torch.uint8, device="cuda")
    h_b = torch.empty(n_out, dtype=torch.uint8, pin_memory=True)
    d_b = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)
    h2d()
    d2h()
    t_h = min(timed(h2d) for _ in range(2))
    t_d = min(timed(d2h) for _ in range(2))
    # together, in the proportion of the e2e leg: out bytes per in byte
    reps = max(1, int(round((n_in * (nbytes_out / max(1, nbytes_in))) / n_out)))

    def both():
        h2d()
        for _ in range(reps):
            d2h()
    t_b = min(timed(both) for _ in range(2))
    return {"h2d_GBps": n_in / t_h / 1e9, "d2h_GBps": n_out / t_d / 1e9, "concurrent_h2d_GBps": n_in / t_b / 1e9,
            "note": "pinned-host copies alone and running together in the e2e leg's in:out proportion (%.0f MiB in, %d x %.0f MiB out)" % (n_in / 2**20, reps, n_out / 2**20)}


def run_workload(L, _lib, torch, dist, args, rank, local_rank, world, name, profile, d_order, q_order, n_reads, steps, warmup,
                 want_parity, do_e2e, do_cpu, do_serial, symbols_per_read=None):
    """parity gate + resident steps + e2e + isolated kernel times + CPU reference for one workload on this rank; returns a dict and
    leaves nothing allocated"""
    ds = _lib.Dataset(33, 0, 0)
    cs = _lib.Settings(d_order, q_order, 0, 0, 0)
    ctx = C.c_void_p()
    if L.dsrcgpu_create(C.byref(ctx), local_rank, C.byref(ds), C.byref(cs), BLOCK_BYTES, args.inflight):
        raise SystemExit("dsrcgpu_create failed")

    def check(rc_, what):
        if rc_:
            raise SystemExit("%s failed: %d %s" % (what, rc_, L.dsrcgpu_last_error(ctx).decode()))
    sh = Shard(L, _lib, ctx, torch, profile, rank * n_reads, n_reads, check)
    out_cap = sh.in_bytes // 2 + (1 << 20)
    d_out = torch.empty(out_cap, dtype=torch.uint8, device="cuda")
    sizes = np.zeros(sh.n, dtype=np.uint32)
    comp = np.zeros((sh.n, 4), dtype=np.uint64)
    raw = np.zeros((sh.n, 4), dtype=np.uint64)
    checked, where = parity_gate(L, _lib, ctx, torch, sh, d_order, q_order, d_out, out_cap, check, want_parity)

    def step_device():
        check(L.dsrcgpu_encode_blocks_device(ctx, C.c_void_p(sh.d_in.data_ptr()), sh.offs.ctypes.data_as(_lib.u64p), sh.lens.ctypes.data_as(_lib.u32p),
                                             None, sh.n, C.c_void_p(d_out.data_ptr()), out_cap, sizes.ctypes.data_as(_lib.u32p),
                                             raw.ctypes.data_as(_lib.u64p), comp.ctypes.data_as(_lib.u64p)), "encode_device")
    for _ in range(warmup):
        step_device()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ktot = {}
    for _ in range(steps):
        step_device()
        for k, (ms, ln) in kernel_times(L, ctx).items():
            a = ktot.get(k, (0.0, 0))
            ktot[k] = (a[0] + ms, a[1] + ln)
    torch.cuda.synchronize()
    step_s = (time.perf_counter() - t0) / steps
    comp_bytes = int(sizes.astype(np.uint64).sum())
    symbols = int(raw[:, 3].sum())                     # quality symbols == bases (raw quality stream size, src/FastqParser.cpp:152-157)
    res = {"workload": name, "value": sh.payload / step_s / 1e6, "unit": "MB/s", "ms_per_step": step_s * 1e3, "steps": steps, "warmup": warmup,
           "reads": n_reads, "bytes": sh.payload, "blocks": sh.n, "ratio": sh.payload / max(1, comp_bytes),
           "parity_checked_blocks": checked, "parity_blocks": where, "dna_order": d_order, "quality_order": q_order}
    if do_e2e:
        h_in = torch.empty(sh.in_bytes, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(sh.d_in[:sh.in_bytes])
        h_out = torch.empty(out_cap, dtype=torch.uint8, pin_memory=True)
        esz = np.zeros(sh.n, dtype=np.uint32)
        torch.cuda.synchronize()

        def step_host():
            check(L.dsrcgpu_encode_blocks(ctx, C.c_void_p(h_in.data_ptr()), sh.offs.ctypes.data_as(_lib.u64p), sh.lens.ctypes.data_as(_lib.u32p), None, sh.n,
                                          C.c_void_p(h_out.data_ptr()), out_cap, esz.ctypes.data_as(_lib.u32p), None, None), "encode_host")
        step_host()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_host()
        e_s = (time.perf_counter() - t0) / steps
        same = bool((esz == sizes).all()) and bool(torch.equal(h_out[:comp_bytes].cuda(), d_out[:comp_bytes]))
        if not same:
            raise SystemExit("bench.py: %s: the host-buffer leg's bytes differ from the resident leg's" % name)
        res["e2e"] = {"value": sh.payload / e_s / 1e6, "unit": "MB/s", "h2d_bytes_per_step": int(sh.in_bytes), "d2h_bytes_per_step": comp_bytes,
                      "output_identical_to_resident_leg": same}
        del h_in, h_out
    if do_serial:
        L.dsrcgpu_release_workspace(ctx)      # room for the single-slot context of the serialized pass (each context has its own table pool)
        kser = serialized_pass(L, _lib, local_rank, ds, cs, args.inflight, sh, d_out, out_cap, sizes)
        if kser:
            peak, peak_kind = load_peaks()
            ser_ms = kser.pop("_step_ms")[0]
            alg = algorithmic_bytes(sh, symbols, comp, comp_bytes)
            res["roofline"] = make_roofline(kser, {k: (v[0] / steps, v[1]) for k, v in ktot.items()}, ser_ms, 1, alg, peak, peak_kind, sh.payload,
                                            comp_bytes, symbols, step_s,
                                            lambda k: ncu_traffic("model_quality_part" if (k == "model_quality" and profile != 0) else k))
    if do_cpu and os.path.exists(REF_LIB):
        threads = os.cpu_count() or 1
        ns = min(sh.n, threads * 300)
        h_cpu = sh.d_in[:int(sh.offs[ns - 1]) + int(sh.lens[ns - 1])].cpu().numpy()
        r, dt, tot = reference_cpu_rate(h_cpu.ctypes.data, sh.offs, sh.lens, ns, threads, d_order, q_order)
        res["cpu_baseline"] = {"value": r, "unit": "MB/s", "cores": threads, "kind": "reference",
                               "sample": "first %d blocks (%.1f MB), reference BlockCompressor::Store, one instance per thread, %.1f s" % (ns, tot / 1e6, dt)}
    del d_out, sh
    L.dsrcgpu_destroy(ctx)
    torch.cuda.empty_cache()
    return res


def archive_leg(L, _lib, torch, dist, rank, local_rank, world, n_reads=3_000_000):
    """BASELINE configs[3] as the product does it: ONE file (every rank generates the same 1.1 GB sample), contiguous block ranges per
    rank, NCCL all-gather of the block sizes, rank 0 writes header + footer, every rank pwrites its slice. Rank 0 also writes the
    single-rank archive of the same file; the two SHA-256 must agree. Then every rank decodes its share of the N-rank archive and
    pwrites it into one FASTQ whose SHA-256 must equal the input's (configs[4])."""
    from dsrc_b200 import operators as op
    nbytes = n_reads * REC_BYTES
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    ctx = C.c_void_p()
    ds = _lib.Dataset(33, 0, 0)
    cs = _lib.Settings(DNA_ORDER, QUA_ORDER, 0, 0, 0)
    L.dsrcgpu_create(C.byref(ctx), local_rank, C.byref(ds), C.byref(cs), BLOCK_BYTES, 0)
    nb = C.c_uint64()
    L.dsrcgpu_synth_fastq_device(ctx, 0, 4242, 0, n_reads, C.c_void_p(d.data_ptr()), nbytes, C.byref(nb))
    L.dsrcgpu_destroy(ctx)
    fastq = d.cpu().numpy()
    del d
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(tmp, "dsrc_b200_bench_archive.dsrc")
    back = os.path.join(tmp, "dsrc_b200_bench_archive.fq")
    args = op.InputParameters(DNA_ORDER // 3, QUA_ORDER, 1, 0, block_bytes=BLOCK_BYTES)
    gdev = torch.device("cuda", local_rank)
    dist.barrier()
    t0 = time.perf_counter()
    arc, sl = op.DsrcCompressorMT(device=local_rank, gather=lambda s: op.dist_gather_sizes(s, gdev), rank=rank, world=world).process(args, fastq)
    op.write_archive_sharded(path, arc, sl, rank, barrier=dist.barrier)
    t_enc = time.perf_counter() - t0
    archive = np.fromfile(path, dtype=np.uint8)
    if rank == 0 and os.path.exists(back):
        os.remove(back)
    dist.barrier()
    t0 = time.perf_counter()
    off, part = op.DsrcDecompressorMT(device=local_rank, rank=rank, world=world).process(archive.tobytes())
    fd = os.open(back, os.O_RDWR | os.O_CREAT, 0o644)
    os.pwrite(fd, part, off)
    os.close(fd)
    dist.barrier()
    t_dec = time.perf_counter() - t0
    out = None
    if rank == 0:
        single, _ = op.DsrcCompressorMT(device=local_rank).process(args, fastq)
        sha_n = hashlib.sha256(archive.tobytes()).hexdigest()
        sha_1 = hashlib.sha256(single).hexdigest()
        sha_in = hashlib.sha256(fastq.tobytes()).hexdigest()
        sha_back = hashlib.sha256(np.fromfile(back, dtype=np.uint8).tobytes()).hexdigest()
        out = {"fastq_bytes": int(nbytes), "archive_bytes": int(archive.size), "ranks": world,
               "sha256_n_rank_archive": sha_n, "sha256_single_rank_archive": sha_1, "archives_identical": sha_n == sha_1,
               "sha256_input": sha_in, "sha256_decoded_by_n_ranks": sha_back, "roundtrip_identical": sha_in == sha_back,
               "encode_wall_s": t_enc, "decode_wall_s": t_dec,
               "how": "operators.DsrcCompressorMT / DsrcDecompressorMT (host buffers, file on %s), block sizes over NCCL all_gather, every rank pwrites its slice" % tmp}
        if not (out["archives_identical"] and out["roundtrip_identical"]):
            raise SystemExit("bench.py: multi-rank archive check failed: %s" % json.dumps(out))
    dist.barrier()
    if rank == 0:
        for f in (path, back):
            if os.path.exists(f):
                os.remove(f)
    return out


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
    ap.add_argument("--no-extras", action="store_true", help="skip the other single-GPU workloads (41-level Illumina, configs[2], -q0)")
    ap.add_argument("--no-archive", action="store_true", help="skip the multi-rank archive check (N > 1)")
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

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, n_gpus, config)
        return

    # ------------------------------------------------------------------ our arm
    from dsrc_b200 import _lib
    L = _lib.lib()
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

    # synthetic shard of this rank, generated on the device
    sh = Shard(L, _lib, ctx, torch, args.profile, rank * args.reads, args.reads, check)
    n, offs, lens, d_in, in_bytes, payload = sh.n, sh.offs, sh.lens, sh.d_in, sh.in_bytes, sh.payload
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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # parity gate on EVERY rank before anything is timed (oracle = checker only): 64 blocks spread over the shard, cold and warm
    parity_blocks, parity_where = parity_gate(L, _lib, ctx, torch, sh, DNA_ORDER, QUA_ORDER, d_out, out_cap, check, 64)
    pv = torch.tensor([parity_blocks], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(pv, op=dist.ReduceOp.MIN)
    parity_min_per_rank = int(pv[0])

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
        for k, (ms, ln) in kernel_times(L, ctx).items():
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

    # e2e through the host-buffer call; its output must equal the resident leg's
    e2e = None
    ceiling = None
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
        e_out = int(esz.astype(np.uint64).sum())
        # the e2e leg's bytes against the resident leg's (same blocks, same warm field vectors): sizes and every byte, on every rank
        same = bool((esz == sizes[:ne]).all())
        if same:
            chunk = 1 << 30
            for p in range(0, e_out, chunk):
                q = min(e_out, p + chunk)
                if not torch.equal(h_out[p:q].cuda(), d_out[p:q]):
                    same = False
                    break
        sv = torch.tensor([1 if same else 0], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(sv, op=dist.ReduceOp.MIN)
        if int(sv[0]) != 1:
            raise SystemExit("bench.py: the host-buffer leg's output differs from the resident leg's")
        e_val = e_payload * world / (float(tv[0]) / args.steps) / 1e6
        del h_in, h_out
        torch.cuda.empty_cache()
        barrier()
        ceiling = copy_ceiling(torch, e_in, e_out)
        cv = torch.tensor([ceiling["concurrent_h2d_GBps"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(cv, op=dist.ReduceOp.SUM)        # all ranks copy at the same time: the box's aggregate
        ceiling["aggregate_concurrent_h2d_GBps"] = float(cv[0])
        ceiling["ranks_copying_together"] = world
        e2e = {"value": e_val, "unit": "MB/s",
               "h2d_bytes_per_step": int(e_in), "d2h_bytes_per_step": e_out,
               "blocks_per_gpu": int(ne), "output_identical_to_resident_leg": True,
               "copy_ceiling": ceiling, "frac_of_copy_ceiling": (e_val / 1e3) * (e_in / max(1, e_payload)) / max(1e-9, float(cv[0])),
               "note": "whole workload" if ne == n else "first %d of %d blocks per GPU (pinned staging bounded by host RAM / ranks)" % (ne, n)}

    # decode leg (BASELINE configs[4]): BlockCompressor::Read of this step's archive, on every rank -- device-resident (whole shard) and
    # through host buffers (bounded sample) -- verified against the input
    dec = None
    if not args.no_decode:
        step_device()
        L.dsrcgpu_release_workspace(ctx)      # the encode slots' workspaces make room for the decode chains' arenas
        coffs = np.concatenate([[0], np.cumsum(sizes.astype(np.uint64))[:-1]]).astype(np.uint64)
        nd = n
        dbytes = int(lens[:nd].astype(np.uint64).sum()) + nd
        d_dec = torch.empty(dbytes + 64, dtype=torch.uint8, device="cuda")
        osz = np.zeros(nd, dtype=np.uint64)

        def step_decode():
            check(L.dsrcgpu_decode_blocks_device(ctx, C.c_void_p(d_out.data_ptr()), coffs.ctypes.data_as(_lib.u64p), sizes.ctypes.data_as(_lib.u32p), nd,
                                                 C.c_void_p(d_dec.data_ptr()), dbytes + 64, osz.ctypes.data_as(_lib.u64p)), "decode_device")
        step_decode()
        barrier()
        t0 = time.perf_counter()
        step_decode()
        barrier()
        d_wall = time.perf_counter() - t0
        dv = torch.tensor([d_wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dv, op=dist.ReduceOp.MAX)
        ok = bool(int(osz.sum()) == dbytes and torch.equal(d_dec[:dbytes], d_in[int(offs[0]):int(offs[0]) + dbytes]))
        # SHA-256 of a bounded sample of the decoded bytes against the same bytes of the input
        sn = min(dbytes, 256 << 20)
        sha_ok = hashlib.sha256(d_dec[:sn].cpu().numpy().tobytes()).hexdigest() == hashlib.sha256(d_in[:sn].cpu().numpy().tobytes()).hexdigest()
        okv = torch.tensor([1 if (ok and sha_ok) else 0], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(okv, op=dist.ReduceOp.MIN)
        if int(okv[0]) != 1:
            raise SystemExit("bench.py: decode leg: the decoded FASTQ differs from the input")
        del d_dec
        torch.cuda.empty_cache()
        # through host buffers: compressed blocks in pinned host memory, FASTQ out in pinned host memory
        # the whole shard when the box's RAM allows (its share per rank, as in the e2e leg): a small sample's chains are latency-bound
        try:
            import psutil
            h_budget = int(psutil.virtual_memory().total * 0.5 / max(1, world))
        except Exception:
            h_budget = 32 << 30
        nh = n
        if dbytes + comp_bytes > h_budget:
            ends_ = np.cumsum(lens.astype(np.uint64) + np.uint64(1))
            nh = max(1, int(np.searchsorted(ends_, np.uint64(h_budget * 3 // 4), side="right")))
        cbytes = int(sizes[:nh].astype(np.uint64).sum())
        hbytes = int(lens[:nh].astype(np.uint64).sum()) + nh
        h_c = torch.empty(cbytes, dtype=torch.uint8, pin_memory=True)
        h_c.copy_(d_out[:cbytes])
        h_f = torch.empty(hbytes + 64, dtype=torch.uint8, pin_memory=True)
        hsz = np.zeros(nh, dtype=np.uint64)
        torch.cuda.synchronize()

        def step_decode_host():
            check(L.dsrcgpu_decode_blocks(ctx, C.c_void_p(h_c.data_ptr()), coffs.ctypes.data_as(_lib.u64p), sizes.ctypes.data_as(_lib.u32p), nh,
                                          C.c_void_p(h_f.data_ptr()), hbytes + 64, hsz.ctypes.data_as(_lib.u64p)), "decode_host")
        step_decode_host()
        barrier()
        t0 = time.perf_counter()
        step_decode_host()
        barrier()
        h_wall = time.perf_counter() - t0
        hv = torch.tensor([h_wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(hv, op=dist.ReduceOp.MAX)
        h_ok = bool(torch.equal(h_f[:hbytes].cuda(), d_in[int(offs[0]):int(offs[0]) + hbytes]))
        if not h_ok:
            raise SystemExit("bench.py: host-buffer decode differs from the input")
        dec = {"value": dbytes * world / float(dv[0]) / 1e6, "unit": "MB/s", "blocks_per_gpu": int(nd), "bytes_per_gpu": dbytes,
               "verified_identical": True, "sha256_of_sample_matches_input": True, "n_gpus": world,
               "e2e": {"value": hbytes * world / float(hv[0]) / 1e6, "unit": "MB/s", "blocks_per_gpu": int(nh), "h2d_bytes": cbytes, "d2h_bytes": hbytes,
                       "verified_identical": True},
               "note": "whole-job decode of every rank's shard of this step's archive: device-resident (value) and through pinned host buffers (e2e); "
                       "torch.equal against the input over every byte, SHA-256 over the first %d MiB" % (sn >> 20)}
        if rank == 0 and world == 1 and not args.no_cpu and os.path.exists(REF_LIB):
            threads = os.cpu_count() or 1
            nb_ = min(n, threads * 150)
            hb = d_out[:int(sizes[:nb_].astype(np.uint64).sum())].cpu().numpy().tobytes()
            blocks = []
            p = 0
            for i in range(nb_):
                blocks.append(hb[p:p + int(sizes[i])])
                p += int(sizes[i])
            r, dt = reference_cpu_decode_rate(blocks, int(lens[:nb_].astype(np.uint64).sum()) + nb_, threads)
            dec["cpu_baseline"] = {"value": r, "unit": "MB/s", "cores": threads, "kind": "reference",
                                   "sample": "first %d blocks, reference BlockCompressor::Read, one instance per thread, %.1f s" % (nb_, dt)}
        del h_c, h_f
    sampler.stop = True
    sampler.join(timeout=2)

    kser = None
    if not args.no_serial:
        L.dsrcgpu_release_workspace(ctx)
        kser = serialized_pass(L, _lib, local_rank, ds, cs, args.inflight, sh, d_out, out_cap, sizes)

    # roofline of the dominant kernel (CUDA-event time of its launches; isolated durations when the serialised pass ran)
    peak, peak_kind = load_peaks()
    ser_step_ms = kser.pop("_step_ms")[0] if kser else None
    kbase = {k: (v[0] * args.steps, v[1] * args.steps) for k, v in kser.items()} if kser else ktot
    syms = args.reads * 150
    alg = algorithmic_bytes(sh, syms, comp, comp_bytes)
    roofline = make_roofline(kbase, ktot, ser_step_ms, args.steps, alg, peak, peak_kind, payload, comp_bytes, syms, step_s, ncu_traffic)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and os.path.exists(REF_LIB):
        threads = os.cpu_count() or 1
        ns = min(n, threads * 1500)      # ~10-20 s of CPU work
        h_cpu = d_in[:int(offs[ns - 1]) + int(lens[ns - 1])].cpu().numpy()
        r, dt, tot = reference_cpu_rate(h_cpu.ctypes.data, offs, lens, ns, threads)
        nk = max(0, min(ns - 1, int(np.searchsorted(offs[:ns] + lens[:ns].astype(np.uint64), np.uint64(threads * (96 << 20)), side="right")) - 1))
        cpu = {"value": r, "unit": "MB/s", "cores": threads, "kind": "reference",
               "sample": "first %d blocks (%.1f MB) of the workload, reference BlockCompressor::Store, one instance per thread, %.1f s" % (ns, tot / 1e6, dt),
               "native_b8": reference_native_rate(h_cpu[:int(offs[nk] + lens[nk]) + 1], DNA_ORDER // 3, QUA_ORDER, threads),
               "note": "256 KB blocks are the reference's worst operating point (32 MB model Clear() per block, src/QualityEncoder.h:45-51); native_b8 is the "
                       "reference as it ships (DsrcCompressorMT, its own reader / pool / writer, 8 MB blocks) on the same data"}

    # the main workload's buffers make room for the other legs
    del d_out, d_in, sh
    L.dsrcgpu_destroy(ctx)
    torch.cuda.empty_cache()

    archive = None
    if world > 1 and not args.no_archive:
        archive = archive_leg(L, _lib, torch, dist, rank, local_rank, world)

    extras = None
    if world == 1 and not args.no_extras:
        extras = {}
        small = args.reads < 5_000_000                 # a reduced run keeps the extras proportionate
        for name, profile, d_o, q_o, nr in (
                ("illumina_41level_d2_q2: configs[1]'s shape with 41 quality levels (64-symbol quality models: the partition engine)", 1, 6, 2, 10_000_000),
                ("configs[2]: 454/Ion-Torrent shape, variable-length reads with IUPAC codes (N,R,W,S), -d3 -q2 (8-symbol order-7 DNA, 64-symbol quality)", 2, 9, 2, 4_000_000),
                ("configs[2] RLE path: configs[1]'s 4-level data at -d0 -q0 (2-bit DNA, QualityRLEModeler)", 0, 0, 0, 10_000_000)):
            nr = min(nr, max(20000, args.reads // 5)) if small else nr
            extras[name.split(":")[0]] = run_workload(L, _lib, torch, dist, args, rank, local_rank, world, name, profile, d_o, q_o, nr, 2, 1,
                                                      16, not args.no_e2e, not args.no_cpu, not args.no_serial)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config, "clocks": sampler.summary(), "gpu_launches": launches,
                "device_ms_per_step": dev_max * 1e3 / args.steps, "ratio": payload / max(1, comp_bytes),
                "blocks_per_gpu": int(n), "parity_checked_blocks": parity_min_per_rank, "parity_blocks": parity_where,
                "parity": "every rank: %d (block, cold|warm) encodes spread over its shard == oracle; e2e leg == resident leg byte for byte; decode == input" % parity_min_per_rank,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "decode": dec, "archive": archive, "extras": extras}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
