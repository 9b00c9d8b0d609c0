#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view) per kernel: instruction mix, stall-reason totals and the
hottest SASS instructions.  Usage:  python tools/ncu_sass_summary.py gpurun_out/prof.ncu-rep [--top 25] [--kernel regex]
The text it prints is what gets committed under profiles/ (the .ncu-rep files are scratch)."""
import argparse
import collections
import csv
import io
import re
import subprocess
import sys


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    kernels, cur, hdr = [], None, None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            kernels.append(cur)
            hdr = None
        elif row[0] == "Address":
            hdr = row
            cur["hdr"] = hdr
        elif cur is not None and hdr is not None and len(row) == len(hdr):
            cur["rows"].append(row)
    return kernels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--kernel", default=".")
    ap.add_argument("--dump", action="store_true", help="print every SASS line with executed count + samples")
    a = ap.parse_args()
    for k in load(a.rep):
        if not re.search(a.kernel, k["name"]):
            continue
        h = k["hdr"]
        ci = {n: i for i, n in enumerate(h)}
        ex, smp = ci["Instructions Executed"], ci["# Samples"]
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        tot_ex = sum(int(r[ex]) for r in k["rows"])
        tot_s = sum(int(r[smp]) for r in k["rows"])
        print(f"=== {k['name']}\n    SASS lines {len(k['rows'])}, warp instructions executed {tot_ex}, stall samples {tot_s}")
        agg = collections.Counter()
        for n in stalls:
            agg[n] = sum(int(r[ci[n]]) for r in k["rows"])
        print("    stall reasons: " + ", ".join(f"{n[6:]} {100.0 * v / max(tot_s, 1):.1f}%" for n, v in agg.most_common(8)))
        ops = collections.Counter()
        for r in k["rows"]:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]])
            if m:
                ops[m.group(2).split(".")[0]] += int(r[ex])
        print("    opcode mix: " + ", ".join(f"{o} {100.0 * v / max(tot_ex, 1):.1f}%" for o, v in ops.most_common(16)))
        print(f"    top {a.top} SASS by stall samples:")
        for r in sorted(k["rows"], key=lambda r: -int(r[smp]))[:a.top]:
            why = max(stalls, key=lambda n: int(r[ci[n]]))
            print(f"      {int(r[smp]):7d} smp  {int(r[ex]):9d} exec  {why[6:]:12s} {r[ci['Source']].strip()}")
        if a.dump:
            for r in k["rows"]:
                print(f"{int(r[ex]):9d} {int(r[smp]):6d}  {r[ci['Source']]}")


if __name__ == "__main__":
    sys.exit(main())
