"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_x.txt [units_per_launch_note]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"# ncu --set full summary of {rep}", f"# {note}", ""]
    ik = hdr.index("Kernel Name")
    for li, r in enumerate(data):
        lines.append(f"## launch {li}: {r[ik]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"{k:88s} {r[i]:>18s} {units[i]}")
        lines.append("")
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                          capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(sass)))
    # several kernels may be concatenated; only aggregate rows with numeric instruction counts
    hdr_row = next((r for r in srows if "Instructions Executed" in r), None)
    if hdr_row:
        iA, iN = hdr_row.index("Source"), hdr_row.index("Instructions Executed")
        mix = collections.Counter()
        for r in srows:
            if len(r) <= iN or not r[iN].isdigit():
                continue
            toks = r[iA].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            mix[op.split(".")[0]] += int(r[iN])
        tot = sum(mix.values())
        lines.append(f"## SASS instruction mix over all captured launches (warp-level instructions, total {tot})")
        for op, n in mix.most_common(24):
            lines.append(f"{op:12s} {n:14d} {100.0 * n / max(tot, 1):6.2f}%")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
