"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the LAST frame/iteration.
usage: python tools/summarize_launches.py launches.csv [n_iterations]"""
import collections
import csv
import sys


def main():
    path, iters = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    cols, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
    per = len(data) // iters
    agg = collections.OrderedDict()
    for r in data[-per:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':44s} {'launches':>8s} {'us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {v[0]:8d} {v[1]:10.1f} {100 * v[1] / tot:6.1f}%")
    print(f"{'total (cold-cache, serialised under ncu)':44s} {sum(v[0] for v in agg.values()):8d} {tot:10.1f}")


if __name__ == "__main__":
    main()
