"""Sum DRAM traffic of the tensor-core GEMM launches of one train step from an ncu --csv log
(metrics dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum) -> JSON for bench.py's
roofline.traffic (bytes per launch, averaged like roofline.achieved)."""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3,
        "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}


def main(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("mpu::<unnamed>::", "").replace("mpu::", "")
        v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        per[(row["ID"], name)][row["Metric Name"]] += v
    agg = collections.defaultdict(lambda: dict(launches=0, dram_bytes=0.0, ms=0.0))
    for (_, name), m in per.items():
        a = agg[name]
        a["launches"] += 1
        a["dram_bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        a["ms"] += m.get("gpu__time_duration.sum", 0.0)
    gem = [v for k, v in agg.items() if "mtgemm" in k]
    res = {"source": path, "kernels": {k: v for k, v in agg.items()},
           "gemm_launches": sum(v["launches"] for v in gem),
           "gemm_dram_bytes_per_step": sum(v["dram_bytes"] for v in gem),
           "gemm_ms_under_ncu": sum(v["ms"] for v in gem)}
    res["gemm_dram_bytes_per_launch"] = res["gemm_dram_bytes_per_step"] / max(1, res["gemm_launches"])
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "kernels"}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
