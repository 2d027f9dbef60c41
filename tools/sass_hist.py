#!/usr/bin/env python
"""Opcode histogram of one kernel's SASS: python tools/sass_hist.py <object or .so> <substring of the mangled name>"""
import collections
import re
import subprocess
import sys


def main():
    path, key = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, hist, total = None, collections.Counter(), 0
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or key not in cur:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            base = op.split(".")[0]
            if base == "IMAD" and (".WIDE" in op or ".HI" in op or ".MOV" in op or ".SHL" in op or ".IADD" in op):
                base = "IMAD." + [x for x in ("WIDE", "HI", "MOV", "SHL", "IADD") if "." + x in op][0]
            hist[base] += 1
            total += 1
    print("total", total)
    for k, v in hist.most_common():
        print("%-14s %5d  %5.1f%%" % (k, v, 100.0 * v / total))


if __name__ == "__main__":
    main()
