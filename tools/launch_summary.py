"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares, not absolutes)."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for x in csv.DictReader(lines):
        k = x["Kernel Name"].replace("txh::<unnamed>::", "").replace("void ", "")[:52]
        v = float(x["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:52s} n={a[0]:4d} total={a[1]/1e3:10.1f} us  mean={a[1]/a[0]/1e3:9.1f} us  share={a[1]/tot:.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
