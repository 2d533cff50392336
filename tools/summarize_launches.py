"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md
"""
import collections
import csv
import io
import sys


def main(path):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(row["Metric Unit"], 1e-6)
        name = row["Kernel Name"]
        agg[name][0] += 1
        agg[name][1] += v * scale
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:15]:
        print(f"| `{k[:90]}` | {v[0]} | {v[1]:.3f} | {1e3 * v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f} % |")
    print(f"\nTotal device time of the listed launches: {tot:.2f} ms over {sum(v[0] for v in agg.values())} launches.")


if __name__ == "__main__":
    main(sys.argv[1])
