"""Developer tool: randomized parity fuzzing.
    python tools/fuzz_parity.py cpu [n] [seed]   oracle (C restatement) vs the unmodified reference, CPU only
    python tools/fuzz_parity.py gpu [n] [seed]   CUDA path vs oracle (encode + decode), needs a device
Inputs stay inside the envelope where the reference itself is well defined."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refbind  # noqa: E402

SEPS = b" ._,=:/-#"
ALNUM = b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789"


def rand_case(rng):
    n = int(rng.choice([1, 2, 3, 7, 40, 150, 400]))
    # ---- title template: a list of field generators
    nf = int(rng.integers(1, 9)) if rng.random() < 0.8 else int(rng.integers(9, 24))
    fields = []
    for f in range(nf):
        kind = rng.choice(["const", "counter", "randnum", "smallnum", "text", "vartext", "runnum", "zeronum", "bignum", "downcount", "longtext"])
        fields.append((kind, int(rng.integers(1, 9)), int(rng.integers(0, 1000))))
    seps = [SEPS[int(rng.integers(0, len(SEPS)))] for _ in range(nf - 1)]
    consts = [bytes(ALNUM[int(x)] for x in rng.integers(0, 52, size=int(rng.integers(1, 6)))) for _ in range(nf)]
    mixed = rng.random() < 0.08
    var_len = rng.random() < 0.4
    L = int(rng.integers(1, 200))
    qmode = rng.choice(["binned", "full", "decay", "const2"])
    iupac = rng.random() < 0.3
    plus_rep = rng.random() < 0.15
    out = []
    run_val = int(rng.integers(0, 50))
    for i in range(n):
        parts = []
        for f, (kind, a, b) in enumerate(fields):
            if kind == "const":
                parts.append(consts[f])
            elif kind == "counter":
                parts.append(b"%d" % (b + i * (a % 3 + 1)))
            elif kind == "randnum":
                parts.append(b"%d" % int(rng.integers(1, 10 ** a)))
            elif kind == "smallnum":
                parts.append(b"%d" % int(rng.integers(1, 2 + a * 4)))
            elif kind == "runnum":
                if rng.random() < 0.1:
                    run_val = int(rng.integers(0, 50))
                parts.append(b"%d" % (run_val + 1))
            elif kind == "zeronum":                    # leading zeros: not numeric for the reference (utils.h:163-175)
                parts.append(b"%04d" % int(rng.integers(0, 3000)))
            elif kind == "bignum":
                parts.append(b"%d" % int(rng.integers(1, 2 ** 29)))      # differences stay below 2^31: bit_length() returns 64 beyond, undefined upstream
            elif kind == "downcount":
                parts.append(b"%d" % max(1, 100000 - i * (a + 1) - int(rng.integers(0, 2))))
            elif kind == "longtext":
                parts.append(bytes(ALNUM[int(x)] for x in rng.integers(0, 4, size=int(rng.integers(100, 160)))))
            elif kind == "text":
                parts.append(bytes(ALNUM[int(x)] for x in rng.integers(0, 8 + a, size=a)))
            else:
                parts.append(bytes(ALNUM[int(x)] for x in rng.integers(0, 20, size=int(rng.integers(1, a + 1)))))
        title = b"@" + parts[0]
        for f in range(1, nf):
            title += bytes([seps[f - 1]]) + parts[f]
        if mixed and i > 0 and rng.random() < 0.3:
            title = b"@odd" + bytes([SEPS[int(rng.integers(0, len(SEPS)))]]) * int(rng.integers(0, 3)) + b"x%d" % i
        ln = int(np.clip(rng.normal(L, L / 3 + 1), 1, 400)) if var_len else L
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=ln)].copy()
        if qmode == "binned":
            q = np.array([2, 12, 23, 37], dtype=np.uint8)[rng.integers(0, 4, size=ln)]
        elif qmode == "full":
            q = rng.integers(2, 42, size=ln).astype(np.uint8)
        elif qmode == "decay":
            q = np.clip(np.linspace(40, 5, ln) + rng.normal(0, 3, ln), 2, 44).astype(np.uint8)
        else:
            q = np.where(rng.random(ln) < 0.9, 30, 2).astype(np.uint8)
        if rng.random() < 0.3:
            t = int(rng.integers(0, ln))
            q[ln - t:] = 2
        amb = rng.random(ln) < (0.02 if iupac else 0.003)
        for j in np.nonzero(amb)[0]:
            c = rng.choice(list(b"NNNRWS")) if iupac else ord("N")
            seq[j] = c
            q[j] = int(rng.integers(0, 7)) if (c == ord("N") and rng.random() < 0.8) else max(int(q[j]), 7)
        out.append(title + b"\n" + seq.tobytes() + b"\n+" + (title[1:] if plus_rep else b"") + b"\n" + (q + 33).astype(np.uint8).tobytes() + b"\n")
    data = b"".join(out)
    if rng.random() < 0.1 and not plus_rep:
        data = data.replace(b"\n", b"\r\n")
    d = int(rng.choice([0, 3, 6, 9]))
    qq = int(rng.choice([0, 1, 2]))
    # keep inside the envelope where the reference is defined
    syms = set(data.split(b"\n")[1::4][0]) if n else set()
    allseq = b"".join(data.split(b"\n")[1::4])
    allq = b"".join(data.split(b"\n")[3::4])
    kept = set(s for s, c in zip(allseq, allq) if not (s not in b"ACGT" and c - 33 < 7))
    if d == 0 and (kept - set(b"ACGT")) and ord("N") not in kept:
        d = 3
    if d == 0 and (kept & set(b"RWS")):
        d = 3                                   # -d0 Huffman needs a contiguous symbol prefix (a11)
    if qq == 0 and len(set(allq)) < 3:
        qq = 2                                  # single-symbol Huffman trees are undefined upstream (Q5)
    return data, d, qq, int(plus_rep), bool(rng.random() < 0.3)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "cpu"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(n):
        if it and it % 100 == 0:
            print("...", it, "cases, bad", bad, flush=True)
        data, d, q, pr, crc = rand_case(rng)
        chunk = data[:-2] if data.endswith(b"\r\n") else data[:-1]
        if data.endswith(b"\r\n"):
            data = data.replace(b"\r\n", b"\n")     # decode always writes LF
        try:
            o = refbind.Oracle(33, pr, d, q, crc=crc)
            a1, ra, ca = o.store(chunk)
            a2, _, _ = o.store(chunk)
        except RuntimeError as e:
            print("case", it, "oracle refused:", e)
            continue
        if mode == "cpu":
            # the reference has undefined behaviour on some inputs (it may crash): each case runs in a forked child
            pid = os.fork()
            if pid:
                _, st = os.waitpid(pid, 0)
                if os.WIFSIGNALED(st):
                    print("case", it, "reference crashed (signal %d): outside its defined envelope" % os.WTERMSIG(st))
                    continue
                code = os.WEXITSTATUS(st)
                if code:
                    bad += 1
                    path = os.path.join(ROOT, "gpurun_out", "fuzz_%s_%d_%d.fq" % (mode, seed, it))
                    os.makedirs(os.path.dirname(path), exist_ok=True)
                    open(path, "wb").write(data)
                    print("MISMATCH case", it, "d", d, "q", q, "plus_rep", pr, "crc", crc, "->", path)
                continue
            r = refbind.Ref(33, pr, d, q, crc=crc)
            b1, rb, cb = r.store(chunk)
            b2, _, _ = r.store(chunk)
            ok = a1 == b1 and a2 == b2 and ra == rb and ca == cb
            try:
                if r.read(b2) == data:          # some in-envelope-looking inputs do not round-trip through the reference itself
                    ok = ok and o.read(b2) == data
                else:
                    print('case', it, 'reference does not round-trip its own block; decode not compared')
            except RuntimeError as e:
                print('case', it, 'read failed:', e)
                ok = False
            sys.stdout.flush()
            os._exit(0 if ok else 1)
        else:
            from dsrc_b200 import BlockCompressor, DsrcGpuError
            bc = BlockCompressor(33, bool(pr), d, q, max_block_bytes=max(len(chunk) + 64, 1 << 16), calc_crc32=crc)
            try:
                g1, rg, cg = bc.store(chunk)
                g2, _, _ = bc.store(chunk)
                ok = g1 == a1 and g2 == a2 and rg == ra and cg == ca
                try:
                    rt = o.read(a2) == data
                except RuntimeError:
                    rt = False
                if rt:
                    ok = ok and bc.read(a2, out_cap=len(data) + 64) == data
            except DsrcGpuError as e:
                ok = e.code == -5          # UNSUPPORTED (outside the documented envelope) is reported, never a different bitstream
                if ok:
                    print("case", it, "unsupported:", e)
            bc.close()
        if not ok:
            bad += 1
            path = os.path.join(ROOT, "gpurun_out", "fuzz_%s_%d_%d.fq" % (mode, seed, it))
            os.makedirs(os.path.dirname(path), exist_ok=True)
            open(path, "wb").write(data)
            print("MISMATCH case", it, "d", d, "q", q, "plus_rep", pr, "crc", crc, "->", path)
    print("fuzz", mode, "cases", n, "bad", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
