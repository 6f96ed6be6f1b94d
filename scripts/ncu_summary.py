"""Turn an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the small per-launch table
kept under profiles/.   python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv, io, subprocess, sys

KEYS = [("gpu__time_duration.sum", "time"), ("sm__cycles_elapsed.avg.per_second", "SM clock"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem TC %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [(hdr.index(k), lab) for k, lab in KEYS if k in hdr]
    print(f"# {path.split('/')[-1]}\n")
    print("`ncu --set full --clock-control none` (cold caches, serialised launches: compare shares and")
    print("ratios, not absolute times).  DRAM bytes are per launch.\n")
    print("| # | kernel | " + " | ".join(lab for _, lab in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    tot = 0.0
    for n, r in enumerate(rows[2:]):
        name = r[ki].split("(")[0].replace("void ", "").replace("mdb::", "")
        cells = []
        for i, lab in cols:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3f}".rstrip("0").rstrip(".") if abs(f) < 1e4 else f"{f:.0f}"
            except ValueError:
                pass
            cells.append(f"{v} {units[i]}".strip())
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            tot += float(r[i].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[i]]
        print(f"| {n} | `{name}` | " + " | ".join(cells) + " |")
    print(f"\nDRAM traffic (read + write) summed over these {len(rows) - 2} launches: {tot / 1e9:.3f} GB "
          f"= {tot / 1e9 / max(1, len(rows) - 2):.3f} GB per launch")


if __name__ == "__main__":
    main(sys.argv[1])
