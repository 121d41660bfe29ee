"""Generates tests/golden/blocks.json from the UNMODIFIED reference (oracle/_ref/libdsrcref.so, compiled from
/root/reference by oracle/Makefile). Run in the build container (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

For every case of tests/cases.py it records the SHA-256 of the input, and of the block the reference's
BlockCompressor::Store emits on its first call (cold TagStats::fields vector) and second call (warm) -- SURVEY 8-Q1 --
plus sizes, and of BlockCompressor::Read's output. Whole-archive fixtures (`dsrc c -t1` path, 3 block sizes) likewise.
Tiny cases also embed input and output bytes (hex) so the fixture is self-contained."""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases      # noqa: E402
import refbind    # noqa: E402
import synth      # noqa: E402


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    out = {"generator": "tests/golden/make_golden.py", "reference": "refresh-bio/DSRC 2.02 @ /root/reference, g++ -O2 -DNDEBUG -std=c++11",
           "cases": {}, "archives": {}}
    for name, data, d, q, pr in cases.all_cases():
        r = refbind.Ref(33, pr, d, q)
        cold, raw, cmp_ = r.store(data[:-1])
        warm, _, _ = r.store(data[:-1])
        back = r.read(warm)
        assert back == data, name
        e = {"dna_order": d, "quality_order": q, "plus_rep": pr, "input_sha256": sha(data), "input_bytes": len(data),
             "cold_sha256": sha(cold), "warm_sha256": sha(warm), "cold_bytes": len(cold), "warm_bytes": len(warm),
             "raw_streams": raw, "comp_streams": cmp_}
        if len(data) < 2500:
            e["input_hex"] = data.hex()
            e["cold_hex"] = cold.hex()
            e["warm_hex"] = warm.hex()
        out["cases"][name] = e
    R = refbind.Ref()
    big = synth.illumina(8000, seed=7)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "in.fq")
        open(src, "wb").write(big)
        for dl, ql in [(0, 0), (1, 1), (2, 2), (3, 2)]:
            dst = os.path.join(tmp, "o.dsrc")
            assert R.compress_file(src, dst, dl, ql, 1, 1, 0) == 0
            arc = open(dst, "rb").read()
            out["archives"]["illumina8000_seed7_d%d_q%d_b1" % (dl, ql)] = {
                "dna_level": dl, "quality_level": ql, "buf_mb": 1, "input_sha256": sha(big), "archive_sha256": sha(arc), "archive_bytes": len(arc)}
        dst = os.path.join(tmp, "c.dsrc")        # -c: CRC-32 words in every block header, crc flag in the footer
        assert R.compress_file(src, dst, 2, 2, 1, 1, 0, crc=True) == 0
        arc = open(dst, "rb").read()
        out["archives"]["illumina8000_seed7_d2_q2_b1_crc"] = {"dna_level": 2, "quality_level": 2, "buf_mb": 1, "crc": True, "input_sha256": sha(big),
                                                             "archive_sha256": sha(arc), "archive_bytes": len(arc)}
    json.dump(out, open(os.path.join(HERE, "blocks.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(out["cases"]), "cases,", len(out["archives"]), "archives")


if __name__ == "__main__":
    main()
