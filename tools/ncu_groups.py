"""Instruction / stall-sample shares of source-line ranges for one kernel of an .ncu-rep:
   python tools/ncu_groups.py rep kernel_substring file:lo-hi[:label] ..."""
import csv, os, subprocess, sys

def load(rep, kern_sub):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    agg, kern, fp, hdr = {}, None, None, None
    for r in csv.reader(out.splitlines()):
        if not r: continue
        if r[0] == "File Path": fp = os.path.basename(r[1])
        elif r[0] == "Function Name": kern = r[1]
        elif r[0] == "Line No":
            hdr = r; ii, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        elif kern and kern_sub in kern and r[0] != "":
            try: v = [float(r[ii] or 0), float(r[si] or 0), float(r[ti] or 0)]
            except ValueError: continue
            a = agg.setdefault((fp, int(r[0])), [0, 0, 0, r[1].strip()])
            for k in range(3): a[k] += v[k]
    return agg

if __name__ == "__main__":
    agg = load(sys.argv[1], sys.argv[2])
    ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
    print(f"total warp-instructions {ti:.4g}, samples {ts:.0f}")
    used = set()
    for spec in sys.argv[3:]:
        f, rng, *lab = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        keys = [k for k in agg if k[0] == f and lo <= k[1] <= hi]
        used.update(keys)
        i = sum(agg[k][0] for k in keys); s = sum(agg[k][1] for k in keys); t = sum(agg[k][2] for k in keys)
        print(f"{(lab[0] if lab else spec):36s} inst {100*i/ti:5.1f}%  samples {100*s/ts:5.1f}%  thr/inst {t/max(i,1):5.1f}")
    rest = [k for k in agg if k not in used]
    i = sum(agg[k][0] for k in rest); s = sum(agg[k][1] for k in rest)
    print(f"{'(rest)':36s} inst {100*i/ti:5.1f}%  samples {100*s/ts:5.1f}%")
