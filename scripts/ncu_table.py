"""Dev helper: one line per kernel of an .ncu-rep (time, registers, occupancy, DRAM bytes, L2 hit rate, sectors, issue utilisation,
top stall reasons).    python scripts/ncu_table.py file.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
col = {name: i for i, name in enumerate(h)}
want = [("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("lts__t_sectors.sum", "sectors"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__inst_executed.sum", "winst")]
stalls = [c for c in h if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")]
print("%-34s" % "kernel" + "".join("%10s" % w[1] for w in want) + "  stalls(top3: warps per issue)")
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].split("::")[-1][:33]
    vals = []
    for key, _ in want:
        v = r[col[key]] if key in col else ""
        try:
            f = float(v.replace(",", ""))
            vals.append("%10.1f" % f if f < 1e6 else "%10.2e" % f)
        except ValueError:
            vals.append("%10s" % v[:9])
    st = sorted(((float(r[col[c]] or 0), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for c in stalls), reverse=True)[:3]
    print("%-34s" % name + "".join(vals) + "  " + ", ".join("%s %.1f" % (n, v) for v, n in st))
