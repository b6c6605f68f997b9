#!/usr/bin/env python
"""BASELINE.json configs[3]: chaining-only anchor-array bench -- N (default 500 M) synthetic anchors resident in HBM,
max_dist = 5000, bw = 500, a 6000 bp gap every S anchors (segment-length sweep).  No PCIe in the timed region.

The anchors are generated ON the device with torch (input generation only; distribution as mm2gb_b200.synth.
chaining_only_array / SURVEY.md 8d: one strand, cumulative Geometric(25 bp) gaps, qpos = rpos + random-walk drift, 10 %
off-diagonal noise anchors, q_span 15).  The DP runs through the C ABI (mm2gb_chain_dp_device) on a context created with
MM2GB_CTX_DEVICE_ONLY | MM2GB_CTX_NO_CHAINS; f / p of a 60 k-anchor prefix are checked against the oracle for every S, one whole segment for every S >= 8192 and every
segment of the first 64 M anchors for S = 8192.

    python tools/chainonly_sweep.py [--n 500000000] [--segs 32,128,512,2048,8192,32768,131072,1048576] [--reps 3]
prints one JSON line per S and a final summary line."""
import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def gen_device(n, seg_len, seed, dev):
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    n_noise = int(n * 0.10 / 1.10)
    nb = n - n_noise
    u = torch.rand(nb, generator=g, device=dev, dtype=torch.float32).clamp_min_(1e-30)
    gaps = (torch.log(u) / math.log(1.0 - 1.0 / 25.0)).floor_().to(torch.int64).add_(1)
    del u
    gaps[::seg_len] += 6000
    rpos = torch.cumsum(gaps, 0)
    del gaps
    mag = torch.randint(1, 21, (nb,), generator=g, device=dev, dtype=torch.int64)
    sign = torch.randint(0, 2, (nb,), generator=g, device=dev, dtype=torch.int64).mul_(2).sub_(1)
    move = torch.rand(nb, generator=g, device=dev) < 0.1
    drift = torch.cumsum(torch.where(move, mag * sign, torch.zeros_like(mag)), 0)
    del mag, sign, move
    mask = (1 << 30) - 1
    rid = rpos >> 30
    rp = rpos & mask
    del rpos
    x = (rid << 32) | rp
    y = (15 << 32) | ((rp + drift) & mask)
    del rid, rp, drift
    if n_noise:
        idx = torch.randint(0, nb, (n_noise,), generator=g, device=dev, dtype=torch.int64)
        xn = x[idx]
        yn = (15 << 32) | torch.randint(0, 1 << 30, (n_noise,), generator=g, device=dev, dtype=torch.int64)
        del idx
        x = torch.cat([x, xn]); y = torch.cat([y, yn])
        del xn, yn
        x, perm = torch.sort(x, stable=True)
        y = y[perm]
        del perm
    a = torch.stack([x, y], dim=1).contiguous()
    return a


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=int(os.environ.get("MM2GB_SWEEP_N", "500000000")))
    ap.add_argument("--segs", default="32,128,512,2048,8192,32768,131072,1048576")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", type=int, default=60000)
    ap.add_argument("--full-len", type=int, default=8192, help="segment length whose segments are ALL checked (over the first --full-n anchors)")
    ap.add_argument("--full-n", type=int, default=64_000_000)
    args = ap.parse_args()
    pkg = entry.load_package()
    po = entry.load_oracle()
    dev = torch.device("cuda", 0)
    n = args.n
    misc = pkg.map_ont_misc()
    ctx = pkg.ChainContext(misc, device=0, max_anchors=n, max_reads=4, n_slots=1, flags=pkg.ChainContext.DEVICE_ONLY | pkg.ChainContext.NO_CHAINS)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr(0), device=0)
    off = np.array([0, n], np.int64)
    d_off = torch.from_numpy(off).to(dev)
    d_f = torch.empty(n, dtype=torch.int32, device=dev)
    d_p = torch.empty(n, dtype=torch.int32, device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    prm = po.map_ont_params()
    rows = []
    for S in [int(x) for x in args.segs.split(",")]:
        a = gen_device(n, S, 3, dev)
        torch.cuda.synchronize()
        ctx.chain_dp_device(a, d_off, 1, n, d_f, d_p)
        ctx.sync()
        st = ctx.device_stats()
        # parity on a prefix: f[i] / p[i] depend on a[0..i] only
        m = min(args.check, n)
        ah = a[:m].cpu().numpy().view(np.uint64)
        fo, pq, _ = po.oracle_dp(prm, ah)
        ok = bool(np.array_equal(d_f[:m].cpu().numpy(), fo) and np.array_equal(d_p[:m].cpu().numpy().astype(np.int64), pq))
        # ... and on WHOLE segments: the 6000 bp gaps make the segments independent (no window crosses one), so a segment can be
        # chained by the oracle on its own (p shifted by the segment's start).  One full segment from the middle of the array for
        # every length >= 8192 -- a 1 M-anchor segment exercises the global-window path of k_score_long end to end -- and, for
        # S = args.full_len, every segment of the first args.full_n anchors (oracle calls spread over the host threads).
        seg_checked, seg_bad, seg_anchors = 0, 0, 0

        def check_segments(bounds):
            nonlocal seg_checked, seg_bad, seg_anchors
            lo, hi = int(bounds[0]), int(bounds[-1])
            ah2 = a[lo:hi].cpu().numpy().view(np.uint64)
            fh, ph = d_f[lo:hi].cpu().numpy(), d_p[lo:hi].cpu().numpy().astype(np.int64)

            def one(k):
                s0, s1 = int(bounds[k]) - lo, int(bounds[k + 1]) - lo
                fo2, pq2, _ = po.oracle_dp(prm, ah2[s0:s1])
                pq2 = np.where(pq2 >= 0, pq2 + (s0 + lo), -1)
                return bool(np.array_equal(fh[s0:s1], fo2) and np.array_equal(ph[s0:s1], pq2))
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
                res = list(ex.map(one, range(len(bounds) - 1)))
            seg_checked += len(res); seg_bad += sum(1 for r in res if not r); seg_anchors += hi - lo
        if S >= 8192 or S == args.full_len:
            x = a[:, 0]
            lim = n if S != args.full_len else min(n, args.full_n)
            starts = torch.nonzero(x[1:lim] - x[:lim - 1] >= 6000).flatten().add_(1).cpu().numpy()   # first anchor of every later segment
            if S == args.full_len and len(starts) >= 2:
                check_segments(np.concatenate([[0], starts]))
            elif len(starts) >= 2:
                k = len(starts) // 2
                check_segments(starts[k:k + 2])
            del x
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.profile(True)
        e0.record(stream)
        for _ in range(args.reps):
            ctx.chain_dp_device(a, d_off, 1, n, d_f, d_p)
        e1.record(stream)
        ctx.sync()
        ms = e0.elapsed_time(e1) / args.reps
        prof = ctx.profile_read()
        ctx.profile(False)
        row = {"seg_len": S, "anchors": n, "pairs": int(st.n_pairs), "pairs_per_anchor": st.n_pairs / n, "units": int(st.n_units),
               "units_exact": int(st.n_units_exact), "ms": ms, "pairs_per_s": st.n_pairs / (ms / 1e3), "anchors_per_s": n / (ms / 1e3),
               "hbm_algorithmic_gbs": 24.0 * n / (ms / 1e3) / 1e9, "hbm_frac": 24.0 * n / (ms / 1e3) / 1e9 / hbm_peak,
               "kernel_ms": {k: v[0] / max(1, v[1]) for k, v in prof.items() if v[1]}, "prefix_parity": ok, "prefix": m,
               "whole_segments_checked": seg_checked, "whole_segment_anchors": seg_anchors, "whole_segment_mismatches": seg_bad}
        print(json.dumps(row), flush=True)
        rows.append(row)
        del a
        torch.cuda.empty_cache()
    if any(r["whole_segment_mismatches"] or not r["prefix_parity"] for r in rows):
        print(json.dumps({"error": "PARITY FAILURE"}))
        sys.exit(1)
    print(json.dumps({"workload": "chaining-only anchor array (BASELINE.json configs[3])", "n": n, "max_dist": 5000, "bw": 500,
                      "hbm_peak_gbs": hbm_peak, "all_prefix_parity": all(r["prefix_parity"] for r in rows),
                      "whole_segments_checked": sum(r["whole_segments_checked"] for r in rows), "whole_segment_mismatches": sum(r["whole_segment_mismatches"] for r in rows),
                      "best_pairs_per_s": max(r["pairs_per_s"] for r in rows)}))
    ctx.close()


if __name__ == "__main__":
    main()
