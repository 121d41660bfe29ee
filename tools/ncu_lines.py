"""Developer tool: per-source-line instruction counts / stall samples of one kernel from an .ncu-rep.
ncu's CSV source page is SASS-only, so the SASS is joined (by instruction order) with `nvdisasm -g` line info of the
cubin embedded in the built library.
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel-regex> <mangled-name> <cubin-stem> [top]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kre, mangled, stem = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "dsrc_b200", "libdsrc_b200.so")], cwd=tmp, capture_output=True)
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, stem + ".sm_100a.cubin")], capture_output=True, text=True).stdout
lines = []
infun = False
cur = ("?", 0)
for l in sass.splitlines():
    if l.startswith("//--------------------- .text."):
        infun = l.split(".text.")[1].split()[0] == mangled
        continue
    if not infun:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
data = []
sub = os.environ.get("KSUB", "")          # substring of the demangled name (ncu's regex only sees the base name)
take = False
for r in rows:
    if r and r[0] == "Kernel Name":
        if data:
            break          # first matching launch only
        take = sub in r[1]
        hdr = None
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if take and hdr and r and r[0].startswith("0x"):
        data.append(dict(zip(hdr, r)))
print("sass instrs: nvdisasm %d, ncu %d" % (len(lines), len(data)))
agg = {}
tot_i = tot_s = 0
for (off, key, txt), d in zip(lines, data):
    ie = int(d.get("Instructions Executed", "0") or 0)
    sm = int(d.get("# Samples", "0") or 0)
    a = agg.setdefault(key, [0, 0])
    a[0] += ie
    a[1] += sm
    tot_i += ie
    tot_s += sm
print("total warp-instr %d  samples %d" % (tot_i, tot_s))
src_cache = {}


def src(key):
    f, ln = key
    for d in ("dsrc_b200/csrc", "include"):
        p = os.path.join(ROOT, d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[ln - 1].strip()[:100] if 0 < ln <= len(L) else ""
    return ""


for key, (ie, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% instr %5.1f%% samples  %-16s:%4d  %s" % (100.0 * ie / max(1, tot_i), 100.0 * sm / max(1, tot_s), key[0], key[1], src(key)))
