#!/usr/bin/env python
"""Aggregate an ncu source-page export (SASS view, CSV) by CUDA source line.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv
    python tools/sass_by_line.py src.csv path/to/object.o kernel_substring [top]

The SASS rows of the report are matched, in order, with `nvdisasm --print-line-info-inline`
of the same kernel in the object file; executed warp instructions, thread instructions and
stall samples are summed per (file, line) of the innermost inlined frame and per outermost
kernel line."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def disasm_lines(obj, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp,
                          stdout=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True,
                         text=True).stdout.splitlines()
    out, on = [], False
    pending = []
    inner = outer = None
    for ln in txt:
        if ln.startswith(".text."):
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            if pending:  # innermost frame first, the kernel's own line last
                inner, outer = pending[0], pending[-1]
                pending = []
            out.append((int(m.group(1), 16), m.group(2).strip(), inner, outer))
    return out


def main():
    src_csv, obj, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = disasm_lines(obj, kernel)
    rows = list(csv.reader(open(src_csv)))
    hdr = None
    data = []
    nk = 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            nk += 1
            if nk > 1:
                break
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    ci = hdr.index("Instructions Executed")
    ti = hdr.index("Thread Instructions Executed")
    si = hdr.index("# Samples")
    base = int(data[0][0], 16)
    by_addr = {a: (txt, i, o) for a, txt, i, o in dis}
    inner_tot, outer_tot = defaultdict(lambda: [0, 0, 0]), defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    miss = 0
    for r in data:
        off = int(r[0], 16) - base
        ent = by_addr.get(off)
        vals = (int(r[ci]), int(r[ti]), int(r[si]))
        for q in range(3):
            tot[q] += vals[q]
        if ent is None:
            miss += 1
            continue
        for key, dst in ((ent[1], inner_tot), (ent[2], outer_tot)):
            for q in range(3):
                dst[key][q] += vals[q]
    print("total warp-instr %d  thread-instr %d  samples %d  (unmatched rows %d of %d)"
          % (tot[0], tot[1], tot[2], miss, len(data)))
    for name, d in (("kernel line (outermost frame)", outer_tot), ("innermost frame", inner_tot)):
        print("== by", name)
        for key, v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
            print("  %-22s %6s  warp-instr %5.1f%%  lanes %4.1f  samples %5.1f%%"
                  % (key[0] if key else "?", key[1] if key else "?", 100.0 * v[0] / max(tot[0], 1),
                     v[1] / max(v[0], 1), 100.0 * v[2] / max(tot[2], 1)))


if __name__ == "__main__":
    main()
