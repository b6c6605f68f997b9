#!/usr/bin/env python
"""Text view of the stage timeline written by MM2GB_TIMELINE=1 (last end-to-end call in the log): one row per chunk."""
import re, sys
blocks, cur = [], []
for line in open(sys.argv[1]):
    if line.startswith("TL end"):
        blocks.append(cur); cur = []
    m = re.match(r"TL slot=(\d+) stage=(\w+) t0=([\d.]+) t1=([\d.]+)", line)
    if m:
        cur.append((int(m.group(1)), m.group(2), float(m.group(3)), float(m.group(4))))
b = blocks[-1]
chunks, open_ = [], {}
for slot, st, t0, t1 in b:
    if st == "h2d":
        open_[slot] = {"slot": slot}
        chunks.append(open_[slot])
    open_[slot][st] = (t0, t1)
chunks.sort(key=lambda c: c["h2d"][0])
print("chunk slot |  h2d            | range+units    | score          | backtrack      | d2h/drain      | total")
for i, c in enumerate(chunks):
    def f(k):
        return "%6.2f-%6.2f" % c[k] if k in c else "      -      "
    ru = (c["range"][0], c["units"][1]) if "range" in c else None
    print("%5d %4d | %s | %s | %s | %s | %s | %5.2f" % (i, c["slot"], f("h2d"), ("%6.2f-%6.2f" % ru) if ru else "-", f("score"), f("backtrack"), f("d2h"),
                                                     c.get("d2h", c["h2d"])[1] - c["h2d"][0]))
print("end of last stage: %.2f ms" % max(c.get("d2h", c["h2d"])[1] for c in chunks))
