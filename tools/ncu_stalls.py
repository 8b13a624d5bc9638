"""Warp-state (stall) samples per SASS instruction of one kernel, from a `--set full --import-source on` capture.

    ncu -i gpurun_out/X.ncu-rep --page source --csv --print-source sass > /tmp/x.csv
    python tools/ncu_stalls.py /tmp/x.csv [bucket=300] [top=30]

Prints the totals per stall reason, the samples summed over buckets of consecutive instructions (with the median execution
count of the bucket, which tells the loops apart), and the instructions with the most samples.  Runs here (no GPU); the
last kernel of the report is the one looked at.  Tuning aid, not part of the test suite."""
import csv
import statistics
import sys

REASONS = ["stall_barrier", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_mio", "stall_lg",
           "stall_not_selected", "stall_selected", "stall_no_inst", "stall_dispatch", "stall_branch_resolving", "stall_membar"]


def main():
    path = sys.argv[1]
    bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    s = starts[-1]
    print("#", rows[s][1])
    hdr = rows[s + 1]
    data = [r for r in rows[s + 2:] if len(r) == len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    n = lambda r, k: int(r[ix[k]])
    tot = sum(n(r, "# Samples") for r in data)
    print("instructions", len(data), "samples", tot)
    print("totals", {k[6:]: sum(n(r, k) for r in data) for k in REASONS})
    for a in range(0, len(data), bucket):
        sub = data[a:a + bucket]
        t = sum(n(r, "# Samples") for r in sub)
        ex = statistics.median(n(r, "Instructions Executed") for r in sub)
        d = {k[6:]: sum(n(r, k) for r in sub) for k in REASONS}
        print("instr %5d.. exec~%6d samples %5d (%.1f %%)" % (a, int(ex), t, 100.0 * t / max(tot, 1)), {x: y for x, y in d.items() if y >= 30})
    print("# the %d instructions with the most samples" % ntop)
    top = sorted(range(len(data)), key=lambda i: -n(data[i], "# Samples"))[:ntop]
    for i in sorted(top):
        r = data[i]
        print(i, r[ix["# Samples"]], r[ix["Source"]].strip()[:64], {k[6:]: n(r, k) for k in REASONS if n(r, k) >= 20}, "exec",
              r[ix["Instructions Executed"]])


if __name__ == "__main__":
    main()
