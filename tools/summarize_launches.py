"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: count, total, share."""
import csv
import collections
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv:
            continue
        name = r[kn].split('(')[0][:56]
        v = float(r[mv].replace(',', ''))
        v = v / 1000 if r[mu] == 'ns' else (v * 1000 if r[mu] == 'ms' else v)
        a = agg.setdefault(name, [0, 0.0, []])
        a[0] += 1
        a[1] += v
        a[2].append(v)
    tot = sum(a[1] for a in agg.values())
    print("# %s : %d launches, %.1f us total (cold-cache, serialised: compare SHARES)" % (path, sum(a[0] for a in agg.values()), tot))
    print("%-58s %5s %10s %8s %7s %7s %7s" % ("kernel", "n", "total_us", "avg_us", "share", "min", "max"))
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %5d %10.1f %8.1f %7.3f %7.1f %7.1f" % (n, a[0], a[1], a[1] / a[0], a[1] / tot, min(a[2]), max(a[2])))


if __name__ == "__main__":
    main(sys.argv[1])
