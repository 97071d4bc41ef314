"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` launch list into a markdown table.

    python scripts/ncu_summary.py gpurun_out/launches.csv "title line" > profiles/rNN_launches_summary.md

Kernel names are shortened to their base name + template arguments; grid/block of the first launch are shown so that
the table also documents the launch geometry (148-multiple persistent grids etc.).
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    return name[:110]


def main() -> None:
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") in ("us", "usecond"):
            ns *= 1e3
        rows.append((short(r["Kernel Name"]), ns, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"# {title}\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none`: cold-cache, serialised launches — compare SHARES, "
          "not absolutes.\n")
    print(f"total {total / 1e3:.1f} us over {len(rows)} launches\n")
    print("| share | total us | launches | avg us | grid | block | kernel |")
    print("|---|---|---|---|---|---|---|")
    for name, (n, ns, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * ns / total:.1f}% | {ns / 1e3:.1f} | {n} | {ns / 1e3 / n:.2f} | {grid} | {block} | `{name}` |")


if __name__ == "__main__":
    main()
