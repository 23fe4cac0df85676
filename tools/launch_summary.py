"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list into a per-kernel table for the LAST
training step in the capture (delimited by the adam kernel).  Usage: python tools/launch_summary.py launches.csv"""
import collections
import csv
import re
import sys


def main(path, top=20):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        try:
            rows.append((row["Kernel Name"], float(row["Metric Value"].replace(",", ""))))
        except (KeyError, ValueError):
            pass
    adam = [i for i, (n, _) in enumerate(rows) if "adam_kernel" in n]
    step = rows[adam[-2] + 1: adam[-1] + 1] if len(adam) >= 2 else rows
    tot = sum(v for _, v in step)
    print(f"launches in step: {len(step)}; sum of kernel durations: {tot / 1e6:.3f} ms (ncu: cold-cache, serialised)")
    agg = collections.OrderedDict()
    for n, v in step:
        key = re.sub(r"\(.*", "", n).replace("void ", "").replace("b200::", "")
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| `{k[:70]}` | {c} | {v / 1e6:.3f} | {100 * v / tot:.1f}% | {v / c / 1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
