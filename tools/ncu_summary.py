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


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"# ncu summary: `{rep}`\n\nkernel: `{kname}`\n\n| metric | value | unit |\n|---|---:|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} | {vals[i]} | {units[i]} |")
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


if __name__ == "__main__":
    main()
