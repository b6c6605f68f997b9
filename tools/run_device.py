#!/usr/bin/env python
"""A few device-resident steps (whole mg_lchain_dp on the device) of a bench workload and nothing else -- the target of ncu runs:
    ncu --set full --import-source on --clock-control none -k regex:k_bt_walk_mid -s 9 -c 1 -o gpurun_out/x python tools/run_device.py long 2"""
import os, sys
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
    ctx.profile(True)
    for _ in range(steps):
        ctx.chain_device(d_a, d_off, off, n_reads, n, d_f, d_p)
    ctx.sync()
    prof = ctx.profile_read()
    print({k: round(v[0] / max(1, v[1]), 3) for k, v in prof.items() if v[1]}, "anchors", n, "reads", n_reads)
    ctx.close()


if __name__ == "__main__":
    main()
