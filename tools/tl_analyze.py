"""Developer tool: reads a DSRCGPU_TIMELINE dump, takes its last call and prints the kernels in start order plus how long each kind of
kernel ran alone / beside others.   python tools/tl_analyze.py gpurun_out/tl.csv"""
import csv
import sys

rows = [(int(r[0]), r[1], float(r[2]), float(r[3])) for r in csv.reader(open(sys.argv[1], "rb").read().replace(b"\x00", b"").decode().splitlines()) if len(r) == 4]
calls, cur = [], []
for r in rows:
    if r[1] in ("count_lines", "copy_h2d") and r[2] < 1.0 and cur and max(x[3] for x in cur) > 5:
        calls.append(cur)
        cur = []
    cur.append(r)
calls.append(cur)
c = sorted(calls[-1], key=lambda r: r[2])
if "-v" in sys.argv:
    for r in c:
        print("  slot %d %-14s %8.1f -> %8.1f (%6.1f)" % (r[0], r[1], r[2], r[3], r[3] - r[2]))
end = max(r[3] for r in c)
pts = sorted(set([r[2] for r in c] + [r[3] for r in c]))
alone, total, idle = {}, {}, 0.0
for a, b in zip(pts, pts[1:]):
    act = [r for r in c if r[2] <= a and r[3] >= b and not r[1].startswith("copy")]
    if not act:
        idle += b - a
    for r in act:
        total[r[1]] = total.get(r[1], 0) + (b - a)
        if len(act) == 1:
            alone[r[1]] = alone.get(r[1], 0) + (b - a)
print("call end %.1f ms, no kernel running %.1f ms" % (end, idle))
for k in sorted(total, key=lambda k: -total[k]):
    print("  %-14s busy %7.1f ms  of which alone %7.1f ms" % (k, total[k], alone.get(k, 0)))
