"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares, not absolutes:
ncu serialises launches and runs them cold-cache)."""
import collections
import csv
import re
import sys


def main(path, steps):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print("# %s: %d launches, %.1f ms total (%.2f ms per step over %d steps incl. warm-up/profiled/e2e steps)" %
          (path, sum(cnt.values()), T / 1e3, T / 1e3 / steps, steps))
    print("%-58s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("%-58s %6d %12.1f %10.1f %6.1f%%" % (k[:58], cnt[k], v, v / cnt[k], 100 * v / T))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
