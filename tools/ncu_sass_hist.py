"""Development tool: per-opcode histogram of executed warp instructions (and stall samples) from an ncu report's SASS page.
  python tools/ncu_sass_hist.py <file.ncu-rep> [units]   units = divisor for the per-unit column (e.g. number of tiles)
  python tools/ncu_sass_hist.py <file.ncu-rep> list MIN  -> the SASS lines executed at least MIN times, in address order"""
import collections
import csv
import subprocess
import sys


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    cols = {n: i for i, n in enumerate(hdr)}
    return [(r[ia], r[isrc].strip(), int(r[iex]), int(r[ismp]), r, cols) for r in rows[2:] if len(r) > iex]


def main():
    ins = load(sys.argv[1])
    if len(sys.argv) > 2 and sys.argv[2] == "list":
        lo = int(sys.argv[3])
        for a, s, ex, smp, r, cols in ins:
            if ex >= lo:
                extra = r[cols["L1 Wavefronts Shared"]] + "/" + r[cols["L1 Wavefronts Shared Ideal"]] if "L1 Wavefronts Shared" in cols else ""
                print("%s %10d %6d %-12s %s" % (a[-5:], ex, smp, extra, s))
        return
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    h = collections.Counter()
    hs = collections.Counter()
    for a, s, ex, smp, r, cols in ins:
        op = s.split()[0] if not s.startswith("@") else s.split()[1]
        op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "F2F", "DADD", "I2F", "F2I")) else op.split(".")[0]
        h[op] += ex
        hs[op] += smp
    tot, tots = sum(h.values()), sum(hs.values())
    print("total executed %d (%.1f per unit), samples %d" % (tot, tot / units, tots))
    for op, n in h.most_common(45):
        print("%-14s %12d %8.1f/unit %5.1f%%   samples %5.1f%%" % (op, n, n / units, 100.0 * n / tot, 100.0 * hs[op] / max(tots, 1)))


main()
