"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share."""
import collections
import csv
import sys


def summarize(path, out=sys.stdout):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(r[ki][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms total (cold-cache, serialised)", file=out)
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{a[1]:12.1f} us {a[0]:5d} launches {a[1] / a[0]:10.1f} us/launch {100 * a[1] / tot:5.1f}%  {k}", file=out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        summarize(p)
