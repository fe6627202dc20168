"""Per-source-line hot spots from `ncu -i rep --page source --csv --print-source cuda,sass` (needs -lineinfo builds).
Aggregates per kernel (Function Name) over (file, line)."""
import csv
import os
import subprocess
import sys
from collections import OrderedDict


def main(rep, top=28, only=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernels = OrderedDict()
    fpath, kernel, hdr = None, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = os.path.basename(r[1])
        elif r[0] == "Function Name":
            kernel = r[1]
            kernels.setdefault(kernel, {})
        elif r[0] == "Line No":
            hdr = r
            ii, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        elif hdr and kernel and r[0] not in ("",) and len(r) > ti:
            try:
                v = [float(r[ii] or 0), float(r[si] or 0), float(r[ti] or 0)]
            except ValueError:
                continue
            a = kernels[kernel].setdefault((fpath, r[0], r[1].strip()), [0.0, 0.0, 0.0])
            for k in range(3):
                a[k] += v[k]
    for kernel, agg in kernels.items():
        if only and only not in kernel:
            continue
        tot_i = sum(v[0] for v in agg.values()) or 1
        tot_s = sum(v[1] for v in agg.values()) or 1
        print(f"==== {kernel[:100]}\n     warp-instructions {tot_i:.3g}, stall samples {tot_s:.0f}")
        for (fp, ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
            print(f"  samples {100 * v[1] / tot_s:5.1f}%  inst {100 * v[0] / tot_i:5.1f}%  thr/inst {v[2] / max(v[0], 1):5.1f}  {fp}:{ln}: {src[:95]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 28, sys.argv[3] if len(sys.argv) > 3 else None)
