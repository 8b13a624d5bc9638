#!/usr/bin/env python3
"""Condenses an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few per-launch numbers DESIGN.md and
bench.py's roofline quote: duration, DRAM bytes read/written, DRAM throughput %, registers, occupancy, pipe use.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.txt"""
import csv
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "occ_limit_regs_blocks"),
    ("launch__occupancy_limit_shared_mem", "occ_limit_smem_blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe_throttle"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}: ncu --set full --clock-control none (per launch; cold-cache, serialised replays)")
    for r in rows[2:]:
        print("-" * 100)
        for key, name in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"{name:28s} {r[i]} {units[i]}")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd *= scale[units[hdr.index("dram__bytes_read.sum")]]
            wr *= scale[units[hdr.index("dram__bytes_write.sum")]]
            t = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
            t *= {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[units[hdr.index("gpu__time_duration.sum")]]
            print(f"{'dram_traffic_GBps':28s} {(rd + wr) / t / 1e9:.1f}  (traffic {(rd + wr) / 1e6:.1f} MB)")
        except Exception:
            pass


if __name__ == "__main__":
    main()
