"""Key metrics + top source-line hot spots from an .ncu-rep (run in the build container: ncu -i needs no GPU)."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        print("----", row[hdr.index("Kernel Name")][:90])
        for w in WANT:
            if w in hdr:
                j = hdr.index(w)
                print(f"   {w} = {row[j]} {units[j]}")


def source(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    blocks = out.split("\n\n")
    for blk in blocks:
        r = list(csv.reader(io.StringIO(blk)))
        if len(r) < 3:
            continue
        hdr = r[0]
        if "Source" not in hdr:
            continue
        si = hdr.index("Source")
        cols = [c for c in ("# Samples", "Samples", "Warp Stall Sampling (All Samples)", "Instructions Executed") if c in hdr]
        if not cols:
            continue
        key = hdr.index(cols[0])
        rows = []
        for row in r[1:]:
            try:
                rows.append((float(row[key].replace(",", "") or 0), row[si].strip()[:110]))
            except (ValueError, IndexError):
                pass
        tot = sum(x[0] for x in rows) or 1
        print(f"==== source hot spots by {cols[0]} (total {tot:.0f})")
        for v, s in sorted(rows, reverse=True)[:top]:
            print(f"  {100 * v / tot:5.1f}%  {s}")


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        source(sys.argv[1])
