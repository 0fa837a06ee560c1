"""Development tool: stall samples / executed warp instructions per CUDA source line of an ncu report (needs -lineinfo and
--import-source on).   python tools/ncu_line_hist.py <file.ncu-rep> [top N]"""
import collections
import csv
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    smp = collections.Counter()
    exe = collections.Counter()
    text = {}
    cur = None
    fname = ""
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) < 5:
            continue
        if r[0].strip().isdigit():
            cur = (fname, int(r[0]))
            text.setdefault(cur, r[1].strip()[:110])
            continue
        if r[0] == "" and r[2].startswith("0x") and cur is not None:
            try:
                smp[cur] += int(r[4])
                exe[cur] += int(r[7])
            except (ValueError, IndexError):
                pass
    ts, te = sum(smp.values()), sum(exe.values())
    print("samples %d, executed %d" % (ts, te))
    for k, n in smp.most_common(top):
        print("%-16s:%5d  smp %5.1f%%  exe %5.1f%%  %s" % (k[0], k[1], 100.0 * n / ts, 100.0 * exe[k] / max(te, 1), text[k]))


main()
