"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
launches, total device time and share.  Per-launch times under ncu are cold-cache and
serialised, so compare SHARES with bench.py's own event timings, not absolutes.

    python tools/summarize_launches.py gpurun_out/launches.csv [--skip N] > profiles/launches_rNN.md
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<.*", "", name)
    return name.split("::")[-1][:60]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
    rows = rows[skip:]
    agg = OrderedDict()
    for name, ns in rows:
        full = name
        key = short(name)
        # keep template arguments that tell our GEMM families apart
        m = re.search(r"(Epi\w+|ATaps|AConcat|ARows)", full)
        if key in ("simt_gemm_kernel", "tc_gemm_kernel"):
            epi = re.findall(r"Epi\w+", full)
            key += "<" + (epi[0] if epi else "?") + ">"
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% |" % (k, n, ns / 1e6, 100 * ns / total))
    print("| **total** | %d | %.3f | 100%% |" % (len(rows), total / 1e6))


if __name__ == "__main__":
    main()
