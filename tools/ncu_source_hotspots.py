"""Per-source-line hot spots from `ncu -i rep --page source --csv --print-source cuda,sass` (needs -lineinfo builds)."""
import csv
import subprocess
import sys


def main(rep, top=28):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernel, hdr, agg = None, None, {}

    def flush():
        if not agg:
            return
        tot_i = sum(v[0] for v in agg.values()) or 1
        tot_s = sum(v[1] for v in agg.values()) or 1
        print(f"==== {(kernel or "?")[:100]}\n     warp-instructions {tot_i:.3g}, stall samples {tot_s:.0f}")
        for (ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
            print(f"  samples {100 * v[1] / tot_s:5.1f}%  inst {100 * v[0] / tot_i:5.1f}%  thr/inst {v[2] / max(v[0], 1):5.1f}  L{ln}: {src[:105]}")

    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name":
            flush()
            kernel, agg, hdr = r[1], {}, None
        elif r[0] == "Line No":
            hdr = r
            ii, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        elif hdr and r[0] not in ("", "File Name") and len(r) > ti:
            try:
                agg[(r[0], r[1].strip())] = [float(r[ii] or 0), float(r[si] or 0), float(r[ti] or 0)]
            except ValueError:
                pass
    flush()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 28)
