#!/usr/bin/env python
"""Static SASS instruction count per source line for one kernel of a cubin built with -lineinfo.
usage: sass_lines.py <cubin> <kernel-name-substring> [top]"""
import collections
import re
import subprocess
import sys

cubin, pattern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
text = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
cur_fn, cur_line = None, None
counts = collections.Counter()
ops = collections.Counter()
for line in text.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m:
        cur_fn = m.group(1)
        continue
    if cur_fn is None or pattern not in cur_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m:
        counts[cur_line] += 1
        toks = m.group(1).split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ops[(cur_line, op.split(".")[0])] += 1
for k, v in sorted(counts.items(), key=lambda kv: -kv[1])[:top]:
    detail = ", ".join("%s %d" % (op, n) for (ln, op), n in sorted(ops.items(), key=lambda kv: -kv[1]) if ln == k)[:110]
    print("%5d  %s:%s   %s" % (v, k[0] if k else "?", k[1] if k else "?", detail))
print("total", sum(counts.values()))
