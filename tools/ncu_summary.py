"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"void |slk::", "", name)
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"{'total us':>12} {'share':>6} {'n':>5} {'avg us':>9}  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:12.1f} {100 * t / tot:5.1f}% {c:5d} {t / c:9.1f}  {k[:100]}")
    print(f"{tot:12.1f} 100.0%  all kernels")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
