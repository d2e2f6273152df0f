"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def summarise(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, total = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("feba::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    out = [f"total device time {total:.3f} ms over {sum(a[0] for a in agg.values())} launches", "",
           "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {c} | {v:.3f} | {100 * v / total:.1f}% |")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarise(sys.argv[1]))
