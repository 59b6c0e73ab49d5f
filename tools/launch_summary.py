"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name.

    python tools/launch_summary.py gpurun_out/launches.csv [skip_first_n] > profiles/rNN_launches_bench.txt
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name if len(name) <= 110 else name[:107] + "..."


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"]) / 1e3))
    agg = collections.OrderedDict()
    for name, us in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {total:.1f} us of kernel time (per-launch times under ncu are cold-cache and serialised)")
    print(f"{'launches':>8} {'total_us':>10} {'avg_us':>9} {'share':>7}  kernel")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:8d} {us:10.1f} {us / n:9.2f} {100 * us / total:6.1f}%  {name}")


if __name__ == "__main__":
    main()
