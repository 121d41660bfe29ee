"""Developer tool: turns the ncu outputs a gpurun call brought back into the summaries committed under profiles/.
    python tools/ncu_summary.py <launches.csv> <full.ncu-rep> <round tag, e.g. r01>
launch list  -> profiles/<tag>_ncu_launches_bench.csv (as captured) + <tag>_ncu_launches_summary.csv (per kernel: launches, ms, share)
full capture -> profiles/<tag>_ncu_bench_model_full.csv (one row per captured kernel, the metrics DESIGN.md / bench.py quote)"""
import csv
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1:4]
prof = os.path.join(ROOT, "profiles")
shutil.copy(launches, os.path.join(prof, tag + "_ncu_launches_bench.csv"))
rows = [r for r in csv.reader(open(launches)) if len(r) > 14 and r[0].isdigit()]
agg = {}
for r in rows:
    name = r[4].split("(Workspace")[0].split("(RcGroup")[0].split("(unsigned")[0]
    if name.startswith("void "):
        name = name[5:]
    if name.startswith("k_model<"):      # <QUALITY, PART>: walk / partition / classic engines are separate launches
        name = {"k_model<1, 1>": "k_model<quality, partition engine>", "k_model<1, 0>": "k_model<quality, tile engine>",
                "k_model<0, 0>": "k_model<dna, tile/direct engine>"}.get(name.replace("true", "1").replace("false", "0"), name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14]) / 1e6
tot = sum(v[1] for k, v in agg.items() if not k.startswith("k_synth") and not k.startswith("k_rcp_lut"))
with open(os.path.join(prof, tag + "_ncu_launches_summary.csv"), "w") as f:
    f.write("kernel,launches,total_ms,share_of_codec_kernels\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%s,%d,%.3f,%.4f\n" % (k, v[0], v[1], v[1] / tot if not (k.startswith("k_synth") or k.startswith("k_rcp_lut")) else 0.0))
print(open(os.path.join(prof, tag + "_ncu_launches_summary.csv")).read())

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h, units = r[0], r[1]
with open(os.path.join(prof, tag + "_ncu_bench_model_full.csv"), "w") as f:
    w = csv.writer(f)
    cols = [c for c in want if c in h]
    w.writerow(["Kernel Name"] + cols)
    w.writerow([""] + [units[h.index(c)] for c in cols])
    for row in r[2:]:
        w.writerow([row[h.index("Kernel Name")]] + [row[h.index(c)] for c in cols])
print(open(os.path.join(prof, tag + "_ncu_bench_model_full.csv")).read())
