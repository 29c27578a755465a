"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share.

    python tools/summarize_launches.py gpurun_out/r1n_launches.csv > profiles/r1n_launches_summary.md
"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}[unit]
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
        rows.append((name, ns))
    agg = OrderedDict()
    for n, ns in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"# ncu launch list summary: {path}\n")
    print(f"{len(rows)} launches, {total / 1e6:.2f} ms serialised (cold-cache, per-launch replay; shares, not absolutes)\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
