"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import csv
import re
import sys
from collections import defaultdict


def main(path, title):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    tot = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "msecond": 1e3, "ms": 1e3, "nsecond": 1e-3}.get(unit, 1e-3)
        tot[name][0] += 1
        tot[name][1] += v * scale
    total = sum(v[1] for v in tot.values())
    print("# %s\n" % title)
    print("source: `%s` (ncu per-launch times are cold-cache and serialised: compare SHARES, not absolutes)\n" % path)
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (name[:90], n, us, 100 * us / total))
    print("| **total** | %d | %.1f | 100%% |" % (sum(v[0] for v in tot.values()), total))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "kernel launch list")
