"""Join the LAUNCH lines of scripts/gemm_traffic.py with an ncu --csv metrics log into a table.
    python scripts/ncu_traffic_table.py launches.log ncu.csv > table.md"""
import csv
import sys

labels = [ln.split() for ln in open(sys.argv[1]) if ln.startswith("LAUNCH")]
rows = {}
with open(sys.argv[2]) as fh:
    lines = [ln for ln in fh if ln.startswith('"')]
for r in csv.DictReader(lines):
    rows.setdefault(int(r["ID"]), {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
ids = sorted(rows)
print("| shape | variant | plan | time ms | DRAM read GB | DRAM write GB | total / algorithmic | L2 hit % |")
print("|---|---|---|---|---|---|---|---|")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
for lab, i in zip(labels, ids):
    m = rows[i]
    rd = m["dram__bytes_read.sum"][0] * scale[m["dram__bytes_read.sum"][1]]
    wr = m["dram__bytes_write.sum"][0] * scale[m["dram__bytes_write.sum"][1]]
    t = m["gpu__time_duration.sum"][0] * scale[m["gpu__time_duration.sum"][1]]
    hit = m.get("lts__t_sector_hit_rate.pct", (float("nan"), ""))[0]
    alg = float(lab[3].split("=")[1])
    print(f"| {lab[1]} | {lab[2]} | {lab[4].split('=')[1]} | {t:.3f} | {rd / 1e9:.3f} | {wr / 1e9:.3f} | {(rd + wr) / alg:.2f} | {hit:.1f} |")
