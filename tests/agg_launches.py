"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) per kernel name."""
import collections
import csv
import re
import sys


def main(path, detail=False):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    rows = []
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = name.replace("mpu::<unnamed>::", "").replace("mpu::", "")
        t = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        t = t / 1e6 if unit in ("ns", "nsecond") else t / 1e3 if unit in ("us", "usecond") else t
        agg[name][0] += 1
        agg[name][1] += t
        tot += t
        rows.append((name, t, row.get("Grid Size", ""), row.get("ID", "")))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-48s n=%4d  ms=%9.3f  %5.1f%%" % (k[:48], v[0], v[1], 100 * v[1] / tot))
    print("total ms %.3f over %d launches" % (tot, len(rows)))
    if detail:
        for r in rows:
            print("  %-40s %9.3f ms grid %s" % (r[0][:40], r[1], r[2]))


if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) > 2)
