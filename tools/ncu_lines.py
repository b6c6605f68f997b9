#!/usr/bin/env python
"""Per-CUDA-line stall samples / instruction counts of one kernel of an Nsight Compute report.

The report's SASS page has the counters per instruction but no line numbers (and `--print-source cuda` prints the file without
counters), so the line of every instruction is taken from `nvdisasm -g` of the SAME build of the library and matched by
instruction offset.  Usage:
  tools/ncu_lines.py REPORT.ncu-rep KERNEL_NAME [LAUNCH_INDEX] [--so mm2-gb_b200/libmm2gb_chain.so] [--top 40]"""
import argparse
import csv
import io
import os
import re
import subprocess
import tempfile


def line_table(so, kernel, depth=0):
    """offset -> (file line, source file) for the first .text section whose mangled name contains `kernel`"""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
        tab = {}
        for cub in sorted(os.listdir(d)):
            out = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
            inside, chain, fresh = False, [], True
            for ln in out.splitlines():
                if ln.lstrip().startswith(".section") and ".text." in ln:
                    if inside:
                        return tab
                    inside = kernel in ln
                    continue
                if not inside:
                    continue
                m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "[^"]+", line (\d+))?', ln)
                if m:   # a run of annotation lines = the inline chain of the next instruction, innermost first
                    if fresh:
                        chain, fresh = [], False
                    if not chain:
                        chain.append((int(m.group(2)), os.path.basename(m.group(1))))
                    if m.group(3):
                        chain.append((int(m.group(3)), ""))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
                if m:
                    fresh = True
                    inner = chain[0] if chain else (0, "?")
                    callers = tuple(c[0] for c in chain[1:1 + depth])
                    tab[int(m.group(1), 16)] = (inner[0], inner[1], callers)
            if tab:
                return tab
    return tab


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("launch", nargs="?", type=int, default=0, help="n-th launch of that kernel in the report")
    ap.add_argument("--so", default="mm2-gb_b200/libmm2gb_chain.so")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--src", default="mm2-gb_b200/csrc")
    ap.add_argument("--depth", type=int, default=0, help="split a line by its inline call chain, this many callers deep")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--kernel-name", a.kernel, "--launch-skip", str(a.launch),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "Address" in r and "# Samples" in r)
    h = rows[hi]
    iad, ismp, iex, isrc = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    data = [r for r in rows[hi + 1:] if len(r) > max(iad, ismp, iex) and r[iad].startswith("0x")]
    base = int(data[0][iad], 16)
    tab = line_table(a.so, a.kernel, a.depth)
    per = {}
    ts = te = 0
    for r in data:
        off = int(r[iad], 16) - base
        key = tab.get(off, (0, "?", ()))
        smp, ex = int(r[ismp] or 0), int(r[iex] or 0)
        ts += smp
        te += ex
        v = per.setdefault(key, [0, 0, 0])
        v[0] += smp
        v[1] += ex
        v[2] += 1
    text = {}
    print(f"# hottest CUDA lines of `{a.kernel}` (launch {a.launch} in `{a.rep}`): {ts} stall samples, {te} warp instructions, "
          f"{len(data)} SASS instructions\n\n| line | samples | share | instr share | SASS | source |\n|---:|---:|---:|---:|---:|---|")
    for (ln, fn, callers), (smp, ex, ns) in sorted(per.items(), key=lambda kv: -kv[1][0])[:a.top]:
        if fn not in text:
            try:
                text[fn] = open(os.path.join(a.src, fn)).read().splitlines()
            except OSError:
                text[fn] = []
        s = text[fn][ln - 1].strip() if 0 < ln <= len(text[fn]) else ""
        at = "".join(f" <{c}" for c in callers)
        print(f"| {fn}:{ln}{at} | {smp} | {100 * smp / max(ts, 1):.1f}% | {100 * ex / max(te, 1):.1f}% | {ns} | `{s[:100]}` |")


if __name__ == "__main__":
    main()
