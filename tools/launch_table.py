"""Per-kernel table of an ncu launch list (--metrics gpu__time_duration.sum --csv): last launch's duration, grid, block.
Usage: python tools/launch_table.py launches.csv [more.csv ...]"""
import collections
import csv
import sys


def table(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hd = rows[h]
    ix = {k: i for i, k in enumerate(hd)}
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) != len(hd) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        k = r[ix["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        agg.setdefault(k, []).append((v, r[ix["Grid Size"]], r[ix["Block Size"]]))
    return agg


if __name__ == "__main__":
    tabs = [table(p) for p in sys.argv[1:]]
    keys = []
    for t in tabs:
        keys += [k for k in t if k not in keys]
    for k in keys:
        if k.startswith("void at"):
            continue
        cells = []
        for t in tabs:
            v = t.get(k)
            cells.append(f"n={len(v):3d} last={v[-1][0]:8.1f} us grid={v[-1][1]:14s}" if v else " " * 44)
        print(f"{k[:24]:24s} " + " | ".join(cells))
