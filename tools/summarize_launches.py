"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel time per step."""
import collections
import csv
import sys


def main(path, steps):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for x in rows:
        v = float(x["Metric Value"].replace(",", ""))
        unit = x["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        k = x["Kernel Name"][:72]
        tot[k][0] += 1
        tot[k][1] += v
    T = sum(v[1] for v in tot.values())
    print(f"{len(rows)} launches over {steps} steps; {T / steps:.1f} us of kernel time per step (serialised, cold cache)")
    mine = sum(v[1] for k, v in tot.items() if "harp_ctx" in k or "<unnamed>::" in k and "at::" not in k)
    print(f"harp kernels: {100 * mine / T:.1f}% of kernel time; torch glue: {100 * (T - mine) / T:.1f}%")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{v[1] / steps:9.1f} us/step {100 * v[1] / T:5.1f}%  n/step={v[0] / steps:5.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
