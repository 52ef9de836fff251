"""Compact view of a tools/trace.py log: `id:dt` pairs, one line per stage (ids >= 1000 start a line)."""
import re, sys
rows = [tuple(map(int, m.groups())) for m in (re.search(r"id=\s*(\d+)\s+t=\s*(\d+) clk\s+dt=\s*(\d+)", l) for l in open(sys.argv[1])) if m]
line, t0 = [], 0
for k, t, dt in rows:
    if 1000 <= k < 1100:
        if line:
            print(" ".join(line))
        line, t0 = ["[S%d @%d]" % (k - 1000, t)], t
    line.append("%d:%d" % (k, dt))
    if 1100 <= k < 1200:
        line.append("(stage %d)" % (t - t0))
print(" ".join(line))
