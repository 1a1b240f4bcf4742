#!/usr/bin/env python
"""SASS of one kernel with the source line of every instruction in front of it.

    python tools/sass_listing.py object.o kernel_substring > listing.txt
"""
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, kernel = sys.argv[1:3]
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp,
                          stdout=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True,
                         text=True).stdout.splitlines()
    on, tag = False, ""
    for ln in txt:
        if ln.startswith(".text."):
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            tag = "%s:%s" % (os.path.basename(m.group(1)).split(".")[0][:8], m.group(2))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            print("%-14s %s  %s" % (tag, m.group(1), m.group(2).strip()))
        elif ln.startswith(".L_"):
            print(ln)


if __name__ == "__main__":
    main()
