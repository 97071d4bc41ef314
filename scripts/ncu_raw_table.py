"""`ncu -i X.ncu-rep --page raw --csv` -> the markdown table kept under profiles/ (one row per captured launch).

    python scripts/ncu_raw_table.py gpurun_out/r01_b32_step_raw.csv "title" > profiles/r01_b32_step_ncu.md
"""
import csv
import re
import sys

COLS = [
    ("dur", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"),
    ("dram rd", "dram__bytes_read.sum"),
    ("dram wr", "dram__bytes_write.sum"),
    ("dram % of peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("L1 LSU data pipe", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("SM thr", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit", "lts__t_sector_hit_rate.pct"),
    ("L2 sectors", "lts__t_sectors.sum"),
    ("warps active", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue active", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
]


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    name = name.replace("isdqn::", "").replace("(anonymous namespace)::", "")
    return name[:90]


def main() -> None:
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(label, key) for label, key in COLS if key in idx]
    print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}\n")
    print("`ncu --set full --clock-control none` (cold caches, one launch each, in launch order). Units: "
          + ", ".join(f"{label} [{units[idx[key]] or '-'}]" for label, key in cols) + "\n")
    print("| # | kernel | " + " | ".join(label for label, _ in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for n, r in enumerate(data):
        vals = []
        for _, key in cols:
            v = r[idx[key]].replace(",", "")
            try:
                f = float(v)
                vals.append(f"{f:.0f}" if f == int(f) and abs(f) >= 10 else (f"{f:.2f}" if abs(f) >= 1 or f == 0 else f"{f:.4g}"))
            except ValueError:
                vals.append(v)
        print(f"| {n} | `{short(r[idx['Kernel Name']])}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
