"""Warp-state samples per CUDA source line of one kernel, from a `--set full --import-source on` capture of a `-lineinfo` build.

    ncu -i gpurun_out/X.ncu-rep --page source --csv --print-source cuda,sass > /tmp/x.csv
    python tools/ncu_lines.py /tmp/x.csv [top=50]

Sums the SASS rows under every source line (all files the kernel inlines) and prints the lines with the most samples and their
stall reasons.  Runs here (no GPU).  Tuning aid, not part of the test suite."""
import collections
import csv
import sys

REASONS = ("stall_barrier", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_lg", "stall_math", "stall_no_inst", "stall_membar",
           "stall_selected", "stall_not_selected", "stall_branch_resolving")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    cur_file, hdr, cur, ix = None, None, None, {}
    agg = collections.defaultdict(collections.Counter)
    src = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {h: i for i, h in enumerate(hdr)}
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[0] not in ("", "-"):
            cur = (cur_file, int(r[0]))
            src[cur] = r[1].strip()
            continue
        if cur is None or not r[2].startswith("0x"):
            continue
        a = agg[cur]
        a["samples"] += int(r[6])
        for k in REASONS:
            a[k[6:]] += int(r[ix[k]])
    tot = sum(a["samples"] for a in agg.values())
    per_file = collections.Counter()
    for (f, _), a in agg.items():
        per_file[f] += a["samples"]
    print("samples", tot, dict(per_file))
    top = sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:ntop]
    for (f, l), a in sorted(top):
        print("%-22s %4d %5d (%.1f %%)" % (f[:22], l, a["samples"], 100.0 * a["samples"] / max(tot, 1)),
              {k: v for k, v in a.items() if k != "samples" and v >= 20}, "|", src[(f, l)][:90])


if __name__ == "__main__":
    main()
