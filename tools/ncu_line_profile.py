#!/usr/bin/env python
"""Executed warp instructions (and stall samples) of one kernel per SOURCE LINE: joins the per-SASS-instruction counts of an
.ncu-rep (`--page source --csv`) with the line table of the object file the kernel was built from (`nvdisasm --print-line-info`);
the two listings have the same instruction order.

    python tools/ncu_line_profile.py gpurun_out/x_prof.ncu-rep variants/obj_x/cull_stream.cu.o 'stream_cull_kernelILi2ELi0ELi256ELi4ELi3ELb1'
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def ncu_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ie, ism = hdr.index("Instructions Executed"), hdr.index("# Samples")
    return [(r[1].strip(), int(r[ie]), int(r[ism])) for r in rows[2:] if len(r) == len(hdr)]


def line_table(obj, kernel_re):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    out, on, cur = [], False, None
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            on = re.search(kernel_re, ln) is not None
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m:
            out.append((m.group(1).strip(), cur))
    return out


def main():
    rep, obj, kre = sys.argv[1:4]
    a, b = ncu_rows(rep), line_table(obj, kre)
    if len(a) != len(b):
        sys.exit(f"instruction counts differ: ncu {len(a)} vs nvdisasm {len(b)}")
    tot = sum(x[1] for x in a)
    per = collections.OrderedDict()
    for (sa, ex, smp), (sb, loc) in zip(a, b):
        e = per.setdefault(loc, [0, 0, 0])
        e[0] += ex; e[1] += smp; e[2] += 1
    print(f"total warp instructions {tot}; per source line (>= 0.3 %):")
    src = {}
    for loc, (ex, smp, n) in sorted(per.items(), key=lambda kv: (kv[0] or ("", 0))):
        if ex < tot * 0.003:
            continue
        f, l = loc or ("?", 0)
        if f not in src:
            p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blitzen_b200", "csrc", f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][l - 1].strip()[:110] if 0 < l <= len(src[f]) else ""
        print(f"{100.0 * ex / tot:6.2f}%  {ex:10d} exec {smp:6d} smp {n:4d} sass  {f}:{l}  {text}")


if __name__ == "__main__":
    main()
