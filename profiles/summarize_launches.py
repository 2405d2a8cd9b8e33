"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python profiles/summarize_launches.py gpurun_out/launches.csv [top_n]
Prints count, total ms, share of the captured GPU time per kernel (cold-cache, serialised
launches: compare SHARES, not absolutes)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=25):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0,
                 "msecond": 1.0, "second": 1e3, "s": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], v * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for name, ms in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"<.*", "", short) if len(short) > 70 else short
        agg[short][0] += 1
        agg[short][1] += ms
    total = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {total:.3f} ms total")
    print(f"{'kernel':70s} {'count':>6s} {'ms':>10s} {'share':>7s}")
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{name[:70]:70s} {cnt:6d} {ms:10.3f} {100 * ms / total:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
