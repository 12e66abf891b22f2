#!/usr/bin/env python
"""Per-function SASS opcode histogram of a cubin / shared library: instruction counts by mnemonic,
registers are in `nvcc -Xptxas -v`.  usage: sass_hist.py <file> [function-substring]"""
import collections
import re
import subprocess
import sys

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
fn = None
hist = collections.OrderedDict()
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and fn:
        toks = m.group(1).split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        hist[fn][op.split(".")[0]] += 1
for fn, h in hist.items():
    if want not in fn:
        continue
    print("%s: %d instructions" % (fn, sum(h.values())))
    print("   " + ", ".join("%s %d" % kv for kv in h.most_common(48)))
