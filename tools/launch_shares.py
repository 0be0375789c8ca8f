"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_shares.py <launches.csv> [first_kernel_substring_of_the_last_step]"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = next(r for r in rows if "Kernel Name" in r)
    start = rows.index(hdr) + 1
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    data = []
    for r in rows[start:]:
        if len(r) > vi:
            try:
                data.append((r[ki], float(r[vi].replace(",", ""))))
            except ValueError:
                pass
    marker = sys.argv[2] if len(sys.argv) > 2 else None
    if marker:
        idx = [i for i, (n, _) in enumerate(data) if marker in n]
        data = data[idx[-1]:] if idx else data
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in data:
        n = re.sub(r"\(.*", "", n)[:80]
        agg[n][0] += 1
        agg[n][1] += t
    tot = sum(v[1] for v in agg.values())
    print("%-82s %5s %10s %6s" % ("kernel", "n", "ms", "%"))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-82s %5d %10.3f %6.1f" % (n, c, t / 1e6, 100 * t / tot))
    print("total %.3f ms over %d launches" % (tot / 1e6, len(data)))


if __name__ == "__main__":
    main()
