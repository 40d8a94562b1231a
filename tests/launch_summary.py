"""Per-kernel table (launches, serialised ms, share, DRAM GB, GB/s) from an ncu --csv launch list with the metrics
gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum (profiles/*_launches_trainstep_summary.txt)."""
import collections
import csv
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3,
        "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("mpu::<unnamed>::", "").replace("mpu::", "")
        v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        per[(row["ID"], name)][row["Metric Name"]] += v
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for (_, name), m in per.items():
        a = agg[name]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print("%-52s %4s %9s %6s %9s %12s" % ("kernel", "n", "ms", "share", "DRAM GB", "GB/s (ncu)"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-52s %4d %9.3f %5.1f%% %9.3f %12.0f" % (k[:52], a[0], a[1], 100 * a[1] / tot, a[2] / 1e9,
                                                       a[2] / 1e6 / max(a[1], 1e-9)))
    gem = [a for k, a in agg.items() if "mtgemm" in k]
    print("total %.3f ms over %d launches; tensor-core GEMMs %.3f ms (%.1f%%) in %d launches, %.2f GB DRAM per step" %
          (tot, sum(a[0] for a in agg.values()), sum(a[1] for a in gem), 100 * sum(a[1] for a in gem) / tot,
           sum(a[0] for a in gem), sum(a[2] for a in gem) / 1e9))


if __name__ == "__main__":
    main(sys.argv[1])
