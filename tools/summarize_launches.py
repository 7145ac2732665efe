"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share |")
    print("|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        if a[1] / tot < 0.0005:
            continue
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, a[0], a[1], 100 * a[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
