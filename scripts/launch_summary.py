"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into per-kernel totals and shares."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms of kernel time (cold-cache, serialised)")
    print(f"{'kernel':48s} {'launches':>8s} {'total_ms':>10s} {'mean_us':>10s} {'share':>7s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:48s} {a[0]:8d} {a[1]:10.3f} {1e3 * a[1] / a[0]:10.2f} {100 * a[1] / tot:6.1f}%")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
