"""Development tool: condense ncu output into the text summaries kept under profiles/.
  python tools/ncu_summary.py raw <file.ncu-rep> "<header comment>"      -> key metrics of the first captured launch
  python tools/ncu_summary.py launches <launches.csv> "<header comment>" -> per-kernel totals / shares of a launch list
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def raw(path, header):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("# " + header)
    print("# kernel: " + vals[hdr.index("Kernel Name")])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%s %s %s" % (k, units[i], vals[i]))


def launches(path, header):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.OrderedDict()
    order = []
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        ns = float(r[vi].replace(",", ""))
        c, s = tot.get(name, (0, 0.0))
        tot[name] = (c + 1, s + ns)
        order.append((name, ns))
    total = sum(s for _, s in tot.values())
    print("# " + header)
    print("# per-launch device time is cold-cache and serialised under ncu: compare SHARES, not absolutes. unit=ns total=%d launches=%d"
          % (total, len(order)))
    print("%-60s %6s %14s %8s %12s" % ("kernel", "count", "sum_ns", "share", "mean_ns"))
    for name, (c, s) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %6d %14d %7.2f%% %12d" % (name, c, s, 100 * s / total, s / c))
    print("# launches in order:")
    for name, ns in order[:40]:
        print("#   %-56s %10d" % (name, ns))


if __name__ == "__main__":
    {"raw": raw, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
