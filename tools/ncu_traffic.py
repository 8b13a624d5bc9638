#!/usr/bin/env python3
"""Records the ncu-measured DRAM traffic of one kernel in profiles/traffic.json, together with a fingerprint of the source
files that define the kernel.  bench.py prints `roofline.traffic` from that record only while the fingerprint still matches
(so a number is never carried over to a kernel that has changed since the capture).

usage: tools/ncu_traffic.py <report.ncu-rep> <kernel-regex> <name as bench.py prints it> <launch key, e.g. 65536x1024> \
           <profiles/summary file to cite> <source file> [<source file> ...]
The per-launch figure is the median over the launches of that kernel in the report."""
import csv
import hashlib
import json
import os
import re
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, rx, name, key, cite = sys.argv[1:6]
    files = sys.argv[6:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    vals = []
    for r in rows[2:]:
        if re.search(rx, r[ik]):
            vals.append(float(r[ir].replace(",", "")) * SCALE[units[ir]] + float(r[iw].replace(",", "")) * SCALE[units[iw]])
    if not vals:
        raise SystemExit(f"no launch of /{rx}/ in {rep}")
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(ROOT, f), "rb") as fh:
            h.update(fh.read())
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            db = json.load(f)
    except Exception:
        db = {}
    db.setdefault(name, {})[key] = {"bytes_per_launch": statistics.median(vals), "launches_in_capture": len(vals), "source": cite,
                                    "files": files, "sha16": h.hexdigest()[:16]}
    with open(path, "w") as f:
        json.dump(db, f, indent=1, sort_keys=True)
        f.write("\n")
    print(name, key, statistics.median(vals), "bytes per launch over", len(vals), "launches")


if __name__ == "__main__":
    main()
