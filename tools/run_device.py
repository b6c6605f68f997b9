#!/usr/bin/env python
"""A few device-resident steps (whole mg_lchain_dp on the device) of a bench workload and nothing else -- the target of ncu runs:
    ncu --set full --import-source on --clock-control none -k regex:k_bt_walk_mid -s 9 -c 1 -o gpurun_out/x python tools/run_device.py long 2"""
import json, os, sys
import numpy as np
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import bench
import torch


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "ont"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    w = dict(bench.WORKLOADS[wl])
    if len(sys.argv) > 3:
        w["n_reads"] = int(sys.argv[3])
    pkg = entry.load_package()
    a, off = bench.make_workload(w, 0)
    n, n_reads = int(off[-1]), len(off) - 1
    ctx = pkg.ChainContext(pkg.map_ont_misc(), device=0, max_anchors=max(n, 1 << 20), max_reads=n_reads + 1, n_slots=1, flags=pkg.ChainContext.DEVICE_ONLY)
    d_a = torch.from_numpy(a.view(np.int64)).cuda()
    d_off = torch.from_numpy(off).cuda()
    d_f = torch.empty(n, dtype=torch.int32, device="cuda"); d_p = torch.empty(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for _ in range(min(3, steps)):      # warm-up (not timed)
        ctx.chain_device(d_a, d_off, off, n_reads, n, d_f, d_p)
    ctx.sync()
    ctx.profile(True)
    for _ in range(steps):
        ctx.chain_device(d_a, d_off, off, n_reads, n, d_f, d_p)
    ctx.sync()
    prof = ctx.profile_read()
    # a digest of f / p so that variants of a kernel (MM2GB_LIB, MM2GB_RANGE_TMA ...) can be compared across processes
    dig = int((d_f.to(torch.int64) * 1000003 + d_p.to(torch.int64)).sum().item())
    print(json.dumps({"workload": wl, "lib": os.path.basename(os.environ.get("MM2GB_LIB", "libmm2gb_chain.so")), "range_tma": os.environ.get("MM2GB_RANGE_TMA", "0"),
                      "kernel_ms": {k: round(v[0] / max(1, v[1]), 4) for k, v in prof.items() if v[1]}, "steps": steps, "anchors": n, "reads": n_reads,
                      "fp_digest": dig}))
    ctx.close()


if __name__ == "__main__":
    main()
