#!/usr/bin/env python3
"""Reads an .ncu-rep captured with --set full --import-source on: key raw metrics, the executed-instruction mix by opcode, and the
source lines with the most stall samples.  usage: tools/ncu_hot.py report.ncu-rep [kernel-regex] [top-lines]"""
import collections, csv, io, re, subprocess, sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
       "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "lts__t_sector_hit_rate.pct"]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio")


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:3]:
        for k in RAW:
            if k in hdr:
                print(f"{k:70s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
        st = sorted(((float(r[i]), STALL.match(h).group(1)) for i, h in enumerate(hdr) if STALL.match(h)), reverse=True)
        print("stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in st[:9]))
    extra = ["--kernel-name", "regex:" + kre] if kre else []

    def page(view):
        src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", view] + extra + ["--launch-count", "1"]))))
        h = next(i for i, r in enumerate(src) if "Source" in r and "# Samples" in r)
        hd = src[h]
        iS, iE, iN = hd.index("Source"), hd.index("Instructions Executed"), hd.index("# Samples")
        out = []
        for r in src[h + 1:]:
            if len(r) <= max(iS, iE, iN):
                continue
            try:
                out.append((r[iS], int(r[iE]), int(r[iN]), r[0]))
            except ValueError:
                continue
        return out

    sass = page("sass")
    ops, samp = collections.Counter(), collections.Counter()
    tot = sum(n for _, n, _, _ in sass)
    tots = sum(s for _, _, s, _ in sass)
    for text, n, s, _ in sass:
        tk = text.split()
        if not tk:
            continue
        op = (tk[1] if tk[0].startswith("@") and len(tk) > 1 else tk[0]).split(".")[0]
        ops[op] += n; samp[op] += s
    print(f"warp-instructions executed {tot}, stall samples {tots}")
    for op, n in ops.most_common(24):
        print(f"  {op:10s} {n:11d} {100 * n / tot:5.1f}%   samples {100 * samp[op] / max(tots, 1):5.1f}%")
    try:
        both = page("sass,cuda")  # one record per SASS instruction, the Source column holding the CUDA line it came from
    except StopIteration:
        return
    lines_n, lines_s = collections.Counter(), collections.Counter()
    for text, n, s, _ in both:
        lines_n[text.strip()] += n; lines_s[text.strip()] += s
    print("source lines by stall samples (share of executed warp-instructions, share of samples, line):")
    for text, s in lines_s.most_common(top):
        print(f"  {100 * lines_n[text] / tot:5.1f}%  {100 * s / max(tots, 1):5.1f}%  {text[:150]}")

if __name__ == "__main__":
    main()
