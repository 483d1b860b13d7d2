"""Summarises an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"{'kernel':70s} {'n':>6s} {'total_ms':>10s} {'share':>7s} {'avg_us':>9s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{k:70s} {cnt[k]:6d} {v / 1e3:10.3f} {100 * v / T:6.1f}% {v / cnt[k]:9.1f}")
    print(f"{'TOTAL':70s} {sum(cnt.values()):6d} {T / 1e3:10.3f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
