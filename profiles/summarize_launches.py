"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (usage: python summarize_launches.py file.csv [top])."""
import collections
import csv
import re
import sys


def summarize(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, tot = {}, 0.0
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        t = t / 1e3 if unit == "ns" else (t * 1e3 if unit == "ms" else t)
        name = row["Kernel Name"]
        m = re.search(r"(\w+_kernel)\b(<[^(]*)?", name)
        short = (m.group(1) + (m.group(2) or "")[:44]) if m else name[:60]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    out = [f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f"{t:10.1f} us {100 * t / tot:5.1f}% {c:5d} launches  avg {t / c:8.1f} us  {k}")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30))
