#!/usr/bin/env python
"""Print the in-kernel timeline of CTA 0 of the two-pipeline kernel (DP_TC_TRACE dump): per chain, event times relative
to the chain's first event.  Events: 0 start, 1 L1 done, 2 published, 3 M1 issued, 4 M1 done, 5 C1 done, 6 M4 issued,
7 M4 done, 8 R4+seeds done, 9 M2 issued, 10 M2 done, 11 C2 done, 12 M3 done, 13 C3 done, 14 M5/M6 done, 15 R56 done;
16-18 (recorded into the NEXT chain's row): partial sums written, exchange barrier passed, Euler step done."""
import sys
import numpy as np
rows = [l.split() for l in open(sys.argv[1])]
data = {(int(r[0]), int(r[1])): np.array([int(v) for v in r[2:]], dtype=np.int64) for r in rows}
names = {0: "g0 w0", 1: "g0 w7", 2: "g1 w0", 3: "g1 w7"}
t0 = min(v[v > 0].min() for v in data.values() if (v > 0).any())
print("ev:        " + " ".join("%6d" % e for e in range(20)))
for ch in range(2, 10):
    for w in range(4):
        v = data[(w, ch)]
        base = v[0]
        print("chain %d %s @%7d: " % (ch, names[w], base - t0) + " ".join("%6d" % (x - base) if x > 0 else "     -" for x in v))
    print()
