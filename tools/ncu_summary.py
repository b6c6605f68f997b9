#!/usr/bin/env python
"""Summarise an Nsight Compute report (one kernel) into markdown: the roofline counters plus an instruction-count
breakdown by SASS region from the source page.  Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def cuda_lines(rep, kid, top=30):
    """hot CUDA-C lines (needs -lineinfo and --import-source on): by stall samples and by instructions executed"""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda", "--csv"] + (["--kernel-id", kid] if kid else []),
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next((i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r), None)
    if hi is None:
        return
    h = rows[hi]
    isrc, ismp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    iln = h.index("#") if "#" in h else None
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= max(isrc, ismp, iex):
            continue
        try:
            data.append((int(r[ismp] or 0), int(r[iex] or 0), r[iln] if iln is not None else "", r[isrc].strip()))
        except ValueError:
            continue
    ts, te = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
    print(f"\n## hottest CUDA lines (of {ts} stall samples, {te} warp instructions)\n\n| line | samples | share | instr share | source |\n|---:|---:|---:|---:|---|")
    for smp, ex, ln, src in sorted(data, reverse=True)[:top]:
        print(f"| {ln} | {smp} | {100 * smp / ts:.1f}% | {100 * ex / te:.1f}% | `{src[:110]}` |")


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 2 and sys.argv[2] == "--list":
        raw = page(rep, "raw")
        hdr = raw[0]
        for r in raw[2:]:
            print(r[hdr.index("ID")], r[hdr.index("Kernel Name")][:60], r[hdr.index("launch__grid_size")], r[hdr.index("gpu__time_duration.sum")])
        return
    raw = page(rep, "raw")
    if len(sys.argv) > 2:      # kernel id inside a multi-kernel report
        raw = [raw[0], raw[1]] + [r for r in raw[2:] if r[raw[0].index("ID")] == sys.argv[2]]
    hdr, units, vals = raw[0], raw[1], raw[2]
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"# ncu summary: `{rep}`\n\nkernel: `{kname}`\n\n| metric | value | unit |\n|---|---:|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} | {vals[i]} | {units[i]} |")
    kid = sys.argv[2] if len(sys.argv) > 2 else None
    if kid:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kid], capture_output=True, text=True).stdout
        src = list(csv.reader(io.StringIO(out)))
    else:
        src = page(rep, "source")
    h = src[1]
    ia, isrc, ismp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    data = src[2:]
    tot = sum(int(r[ia]) for r in data) or 1
    print(f"\n## warp-instruction count by SASS region (total {tot})\n\n| sass range | n instr | executed each | share | stall samples | first instruction |\n|---|---:|---:|---:|---:|---|")
    s = 0
    for k in range(1, len(data) + 1):
        if k == len(data) or not (0.8 < (int(data[k][ia]) + 1) / (int(data[s][ia]) + 1) < 1.25):
            cs = sum(int(data[i][ia]) for i in range(s, k))
            if cs * 200 > tot:
                print(f"| [{s}:{k}) | {k - s} | {cs // (k - s)} | {100 * cs / tot:.1f}% | {sum(int(data[i][ismp]) for i in range(s, k))} | `{data[s][isrc].strip()}` |")
            s = k
    cuda_lines(rep, kid)


if __name__ == "__main__":
    main()
