"""Turns ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python scripts/summarize_profiles.py <tag> <launches.csv> [<full.ncu-rep>]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches = sys.argv[1], sys.argv[2]
rep = sys.argv[3] if len(sys.argv) > 3 else None
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

rows = list(csv.reader(open(launches)))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).split("::")[-1]
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    seq.append((name, v))
# one build = from a pyramidBaseKernel launch to the next one; take the last complete build on device-resident input
starts = [i for i, (n, _) in enumerate(seq) if n.startswith("pyramidBaseKernel")]
lines = ["# ncu launch list -- %s" % tag, "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline`",
         "(per-launch times are cold-cache and serialised: compare shares with bench.py's CUDA-event phases, not absolutes).", ""]
if len(starts) >= 2:
    one = seq[starts[-2]:starts[-1]]
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for n, v in one:
        tot[n] = tot.get(n, 0) + v
        cnt[n] += 1
    total = sum(tot.values())
    lines += ["## one 16K^2 terrain build (pyramid + create), %d launches, %.1f us serialised" % (len(one), total), "",
              "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for n, v in sorted(tot.items(), key=lambda x: -x[1]):
        lines.append("| %s | %d | %.1f | %.1f%% |" % (n, cnt[n], v, 100 * v / total))
lk = [(n, v) for n, v in seq if n in ("evaluateKernel", "buildSkipGridKernel")] + [(n, v) for n, v in seq if n == "lookupNdcKernel"][:4]
if lk:
    lines += ["", "## lookup kernels", "", "| kernel | us |", "|---|---|"] + ["| %s | %.1f |" % x for x in lk[:12]]
open(os.path.join(out_dir, "%s_launches.md" % tag), "w").write("\n".join(lines) + "\n")
import shutil
shutil.copy(launches, os.path.join(out_dir, "%s_launches.csv" % tag))

if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u = rr[0], rr[1]
    idx = {x: i for i, x in enumerate(h)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    md = ["# ncu --set full -- %s" % tag, "", "`ncu --set full --clock-control none --import-source on` on scripts/trace_run.py (16K^2 terrain);",
          "one warm launch per kernel. dram__bytes_* are the `traffic` of bench.py's roofline object.", ""]
    summary = {}
    # the longest launch of every kernel (the one on the biggest level)
    longest = {}
    for r in rr[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).split("::")[-1]
        t = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
        if name not in longest or t > longest[name][0]:
            longest[name] = (t, r)
    for name, (_, r) in longest.items():
        md += ["## %s" % name, "", "| metric | value | unit |", "|---|---|---|"]
        entry = {}
        for w in want:
            if w in idx:
                md.append("| %s | %s | %s |" % (w, r[idx[w]], u[idx[w]]))
                entry[w] = [r[idx[w]], u[idx[w]]]
        md.append("")
        summary[name] = entry
    open(os.path.join(out_dir, "%s_full.md" % tag), "w").write("\n".join(md) + "\n")

    def to_bytes(val, unit):
        v = float(val.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    traffic = {k: to_bytes(*e["dram__bytes_read.sum"]) + to_bytes(*e["dram__bytes_write.sum"]) for k, e in summary.items()
               if "dram__bytes_read.sum" in e}
    tpath = os.path.join(out_dir, "top_kernel_traffic.json")
    merged = json.load(open(tpath))["dram_bytes_per_launch"] if os.path.exists(tpath) else {}
    merged.update(traffic)  # kernels not in this capture keep their earlier figures
    json.dump({"tag": tag, "dram_bytes_per_launch": merged}, open(tpath, "w"), indent=1)
print("wrote profiles for", tag)
