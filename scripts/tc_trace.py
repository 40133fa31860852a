#!/usr/bin/env python
"""Print the in-kernel timeline of CTA 0 (DP_TC_TRACE dump): per slot, event times relative to the slot's first event."""
import sys
import numpy as np
rows = [l.split() for l in open(sys.argv[1])]
data = {}
for r in rows:
    data[(int(r[0]), int(r[1]))] = np.array([int(v) for v in r[2:]], dtype=np.int64)
names = {0: "MMA ", 1: "cw0 ", 2: "cw4 ", 3: "cw15"}
t0 = min(v[v > 0].min() for v in data.values() if (v > 0).any())
for sl in range(2, 10):
    base = data[(0, sl)][0]
    print("slot %d (net %s tile %d)  start @%d" % (sl, "ba"[sl & 1], (sl >> 1) & 1, base - t0))
    for w in range(4):
        v = data[(w, sl)]
        print("   %s" % names[w], " ".join("%6d" % (x - base) if x > 0 else "     -" for x in v[:13]))
