"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=40):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        v = v / 1000.0 if unit in ("ns", "nsecond") else (v * 1000.0 if unit in ("ms", "msecond") else v)
        a = agg.setdefault(row["Kernel Name"], [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%7s %10s %6s  %-14s %s" % ("count", "avg_us", "share", "grid x block", "kernel"))
    for name, (c, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        short = re.sub(r"\(.*", "", name)[:110]
        print("%7d %10.1f %5.1f%%  %-14s %s" % (c, t / c, 100 * t / tot, g.replace(" ", "") + "x" + b.replace(" ", ""), short))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
