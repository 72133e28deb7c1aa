"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per kernel name launches, total, share, average."""
import collections
import csv
import gzip
import sys


def main(path, top=20):
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 2:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        agg[r[kn][:70]][0] += 1
        agg[r[kn][:70]][1] += v
    tot = sum(v[1] for v in agg.values())
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{n:70s} {c:6d} {t / 1e3:10.1f} us {t / tot * 100:5.1f}%  avg {t / c / 1e3:.2f} us")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20)
