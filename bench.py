#!/usr/bin/env python
"""bench.py -- chaining anchor-pairs/s of the B200 chaining path (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ont|long|chainonly]

step      = one pass of the hot path over one batch of synthetic reads, everything on the device: range -> units -> score
            kernels (f, p), then chain extraction + compaction (k_bt_sort / k_bt_walk: u, a') -- the whole mg_lchain_dp.
workload  = BASELINE.json configs[1]: 100 Mb random reference, 10k simulated ONT-like reads 10-100 kb at ~10 % error,
            map-ont chaining parameters; anchors come from the package's own minimizer seeder (mm2-gb_b200/csrc/
            synth_seed.cpp).  At N GPUs every rank chains its own 10k reads (same reference, different reads): weak scaling,
            no data-path collective (reads are independent).
value     = pairs chained per second with the anchors already resident in HBM (CUDA events on the launching stream,
            max over ranks); pairs = sum_i (i - st_i) = the reference's n_iter (lchain.c:177), counted by the device and
            cross-checked against the oracle in the tests.
e2e       = the same metric through the reference-facing boundary itself -- init_stream_gpu / chain_stream_gpu / finish_stream_gpu
            (include/mm2gb_plchain.h) -- called the way `minimap2 -t T --gpu-chain` calls it (tests/fake_host.c: fake_drive): T worker
            threads, per-read kmalloc'd pageable anchor arrays, host gather + upload + all kernels + download + compact_a's gather
            into kmalloc'd results inside the timed region.  e2e.core_abi is the same batch through mm2gb_chain_host_index with
            caller-pinned buffers (no host pass at all).
parity    = every read of the timed e2e run (n_u, chain-anchor count, digests of u[] and a'[]) against the reference's own lchain.c;
            a mismatch makes the run exit non-zero.
clocks    = SM clock and throttle reasons sampled through NVML every 20 ms over both timed regions.
roofline  = the score kernel (dominant): algorithmic HBM bytes (24 B/anchor) over its CUDA-event time vs the measured copy
            peak, plus the issue-slot view that actually bounds it (SASS thread-instructions per pair / SM issue rate).
seed_chain = (row N2) the fused device step on a sample of the same kind of reads: read SEQUENCES in (host memory), minimizers /
            index lookups / seed filters / anchors / x-sort on the device (mm_map_seed, map.c:355-391), the anchors handed to the
            chaining kernels in HBM, chains + compacted anchors out; next to the reference's mm_map_seed + mg_lchain_dp on the host
            threads (oracle/_ref/libref_seed.so), with a per-read parity check of n_u, u[] and a'[].
cpu_baseline / --impl reference = the reference's own lchain.c (oracle/_ref/libref_lchain.so, compiled from the
            reference sources in the build container; falls back to the oracle port) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# one hardware work queue per CUDA stream (the slots + the chain-extraction size classes); must be set before CUDA starts
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

INSTR_PER_PAIR = 20.0      # algorithmic integer-pipe operations per anchor pair of comput_sc + max/argmax
HBM_BYTES_PER_ANCHOR = 24  # read 16 B mm128_t, write 4 B f + 4 B p

WORKLOADS = {
    # BASELINE.json configs[1]
    "ont": dict(name="synthetic 100 Mb random reference, 10k ONT-like reads 10-100 kb @10% error, map-ont",
                ref_len=100_000_000, contig_len=25_000_000, n_reads=10_000, lo=10_000, hi=100_000, err=0.10),
    # configs[2]: super-long reads, lower error + planted repeats so that reads exceed 50k anchors
    "long": dict(name="synthetic 100 Mb reference with 4000 planted 3 kb repeat copies, 2000 super-long reads 100-300 kb @2% error, map-ont",
                 ref_len=100_000_000, contig_len=25_000_000, n_reads=2000, lo=100_000, hi=300_000, err=0.02,
                 n_repeat_copies=4000, repeat_unit=3000),
    # configs[4]-shaped: the hit mix of a 3 Gb, 24-contig reference (4 real contigs of 25 Mb that the reads come from + the chance hits
    # of 20 virtual contigs of 50-250 Mb, 2.9 Gb in total: ~0.5 per read minimizer) -- ~2.5x the anchors per read of configs[1] at a
    # lower pairs/anchor.  configs[4]'s 1 M reads are processed as steps of 20 k reads per GPU (every rank draws its own reads).
    "hg": dict(name="synthetic human-scale hit mix: 100 Mb real + 2.9 Gb virtual reference in 24 contigs, 20k ONT-like reads 10-100 kb @10% error per GPU and step, map-ont",
               ref_len=100_000_000, contig_len=25_000_000, n_reads=20_000, lo=10_000, hi=100_000, err=0.10, bg_len=2_900_000_000, bg_contigs=20),
    # reads from a tandem array with the occurrence filter off (minimap2 -f 10000): every window is clipped by
    # max_iter, so every unit takes the exact max_ii path (score_unit_exact)
    "tandem": dict(name="tandem-repeat reads: 300 x 200 bp array @2% divergence, 48 reads of 15 kb @5% error, occurrence filter off (-f 10000 style), map-ont",
                   ref_len=20_000_000, contig_len=5_000_000, n_reads=48, lo=15_000, hi=15_000, err=0.05, tandem_copies=300, tandem_unit=200,
                   tandem_div=0.02, tandem_read_frac=1.0, mid_occ=100_000),
    # small variant for quick checks
    "mini": dict(name="synthetic 5 Mb random reference, 400 ONT-like reads 10-100 kb @10% error, map-ont",
                 ref_len=5_000_000, contig_len=0, n_reads=400, lo=10_000, hi=100_000, err=0.10),
}


def make_workload(w, rank):
    from mm2gb_b200 import synth
    kw = {k: w[k] for k in ("n_repeat_copies", "repeat_unit", "bg_len", "bg_contigs", "tandem_copies", "tandem_unit", "tandem_div", "tandem_read_frac",
                            "mid_occ") if k in w}
    return synth.seeded_workload(1, w["ref_len"], w["n_reads"], w["lo"], w["hi"], read_seed=1000 + rank, err=w["err"],
                                 contig_len=w["contig_len"], **kw)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._stop_evt = gpu_index, [], threading.Event()

    def _nvml_loop(self):
        """fast path: NVML in-process (a sample every 20 ms instead of one nvidia-smi process every ~0.3 s)"""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._stop_evt.is_set():
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
            self.rows.append([str(sm), str(mx), "%.1f" % pw] + ["Active" if rs & bits[k] else "Not Active"
                                                                 for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
            self._stop_evt.wait(0.02)

    def run(self):
        try:
            self._nvml_loop()
            return
        except Exception:
            pass
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def workload_shape(a, off, sample=64):
    """anchors per read, and (on a sample of the longest reads) the independent segments = runs of anchors no chaining window
    crosses (st_i == i starts a segment; lchain.c:172 with max_dist_x = 5000) -- what SURVEY.md 8d asks configs[2] to print"""
    rn = np.diff(off)
    out = {"anchors_per_read": {"mean": float(rn.mean()), "min": int(rn.min()), "max": int(rn.max()), "median": float(np.median(rn))}}
    seg_max, seg_over_10k, n_seg = 0, 0, 0
    for r in np.argsort(-rn)[:sample]:
        x = a[int(off[r]):int(off[r + 1]), 0]
        lo32 = x & np.uint64(0xffffffff)
        lower = (x & np.uint64(0xffffffff00000000)) | np.where(lo32 > 5000, lo32 - np.uint64(5000), np.uint64(0))
        st = np.searchsorted(x, lower, side="left")
        cuts = np.flatnonzero(st == np.arange(len(x)))
        seg = np.diff(np.append(cuts, len(x)))
        n_seg += len(seg)
        seg_max = max(seg_max, int(seg.max()))
        seg_over_10k += int((seg > 10_000).sum())
    out["independent_segments_of_the_%d_longest_reads" % min(sample, len(rn))] = {"segments": n_seg, "largest": seg_max, "over_10k_anchors": seg_over_10k}
    return out


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(a, off, budget_s=20.0):
    """The reference's own lchain.c (whole mg_lchain_dp: DP + backtracking) on all host threads, on a bounded sample of the
    reads of this workload.  Pairs are counted once with the oracle port."""
    po = entry.load_oracle()
    prm = po.map_ont_params()
    cores = cpu_threads()
    kind = "reference" if po.ref_available() else "port"
    n_reads = len(off) - 1
    # calibrate on a few reads, then size the sample for ~budget_s of CPU work (all of it if it fits)
    probe = min(n_reads, 4 * cores)
    t0 = time.perf_counter()
    po.lchain_batch(prm, a, off, 0, probe, n_threads=cores, use_ref=(kind == "reference"))
    dt = max(time.perf_counter() - t0, 1e-4)
    per_read_cpu = dt * cores / probe
    n_sample = int(min(n_reads, max(probe, budget_s / per_read_cpu)))
    t0 = time.perf_counter()
    po.lchain_batch(prm, a, off, 0, n_sample, n_threads=cores, use_ref=(kind == "reference"))
    dt = time.perf_counter() - t0
    pairs, _, _ = po.lchain_batch(prm, a, off, 0, n_sample, n_threads=cores)
    return {"value": pairs / dt, "unit": "pairs/s", "cores": cores, "kind": kind,
            "sample": f"first {n_sample} of {n_reads} reads ({int(off[n_sample])} anchors, {pairs} pairs), whole mg_lchain_dp, {dt:.2f} s wall",
            "reads_per_s": n_sample / dt}, pairs, dt, n_sample


def host_stage_alone(pkg, misc, a, off, f, p, n_threads):
    """Wall time of the threaded host stage (backtracking + compaction) alone, f/p already on the host."""
    import ctypes as C
    L = pkg.lib()
    n_reads = len(off) - 1
    n = int(off[-1])
    u = np.empty(max(n, 1), np.uint64)
    b = np.empty((max(n, 1), 2), np.uint64)
    n_u = np.zeros(n_reads, np.int32)
    n_b = np.zeros(n_reads, np.int64)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        rc = L.mm2gb_backtrack_batch(C.byref(misc), a.ctypes.data, off.ctypes.data, n_reads, f.ctypes.data, p.ctypes.data, u.ctypes.data,
                                     n_u.ctypes.data, b.ctypes.data, n_b.ctypes.data, n_threads)
        dt = time.perf_counter() - t0
        assert rc == 0
        best = dt if best is None else min(best, dt)
    return best


_REAL_STDOUT = None


def seed_chain_leg(pkg, torch, dist, w, rank, local_rank, world, host_threads, n_reads=3000, steps=5):
    """Device seeding + chaining (include/mm2gb_seed.h) on `n_reads` reads of the workload's reference: end to end from host
    sequences, device-resident, stage timers; the reference's mm_map_seed + mg_lchain_dp on the host threads beside it."""
    from mm2gb_b200 import seed, synth
    os.environ.setdefault("MM2GB_STAGE_THREADS", str(max(1, min(8, host_threads))))   # the ranks share the host's cores
    ref = synth.simulate_reference(w["ref_len"], seed=1, n_repeat_copies=w.get("n_repeat_copies", 0), repeat_unit=w.get("repeat_unit", 3000))
    reads = synth.simulate_reads(ref, n_reads, w["lo"], w["hi"], seed=100 + rank, err=w["err"])
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    buf = synth._NT[np.concatenate(reads)]
    ref_b = synth._NT[ref]
    del ref, reads
    t0 = time.perf_counter()
    ix = seed.Index((ref_b, np.array([0, len(ref_b)], dtype=np.int64)), w=10, k=15, device=local_rank)
    index_s = time.perf_counter() - t0
    mid_occ = ix.mid_occ()
    prm = seed.map_ont_seed_params(mid_occ)
    misc = pkg.map_ont_misc()
    bases = int(off[-1])
    probe = seed.Seeder(ix, max_bases=bases + 4096, max_reads=n_reads + 8, max_anchors=max(1 << 20, bases))
    _, a_off, _, _, _ = probe.seed(prm, buf, off, want_mini_pos=False)
    n_a = int(a_off[-1])
    probe.close()
    cap = n_a + 4096
    ctx = pkg.ChainContext(misc, device=local_rank, max_anchors=cap, max_reads=n_reads + 8, n_slots=1, flags=pkg.ChainContext.DEVICE_ONLY)
    sd = seed.Seeder(ix, max_bases=bases + 4096, max_reads=n_reads + 8, max_anchors=cap)
    for _ in range(2):
        res = sd.seed_chain(ctx, prm, buf, off, copy=False)
    pairs = int(res.stats.n_pairs)
    torch.cuda.synchronize()     # (no barrier here: a rank whose leg fails must not leave the others waiting; the job figure takes the max time)
    t0 = time.perf_counter()
    for _ in range(steps):
        res = sd.seed_chain(ctx, prm, buf, off, copy=False)     # pageable host sequences in, chains + compacted anchors in host memory out
    e2e_s = (time.perf_counter() - t0) / steps
    stage_ms, n_mv, n_m = sd.profile()
    d_seq = torch.from_numpy(buf).cuda()
    sd.seed_chain_device(ctx, prm, d_seq.data_ptr(), off); ctx.sync(0)
    t0 = time.perf_counter()
    for _ in range(steps):
        sd.seed_chain_device(ctx, prm, d_seq.data_ptr(), off)
        ctx.sync(0)
    dev_s = (time.perf_counter() - t0) / steps
    full = sd.seed_chain(ctx, prm, buf, off)
    out = {"what": "read sequences -> minimizers -> index lookups -> seed filters -> anchors -> x-sort (mm_map_seed, map.c:355-391) on the device, "
                   "anchors handed to the chaining kernels in HBM, chains + compacted anchors back (include/mm2gb_seed.h: mm2gb_seed_chain)",
           "sample": {"reads": n_reads, "bases": bases, "minimizers": n_mv, "seeds": n_m, "anchors": n_a, "pairs": pairs,
                      "chains": int(full["n_chains"]), "chain_anchors": int(full["n_chain_anchors"]), "mid_occ": mid_occ},
           "index": {"keys": ix.n_keys, "occurrences": ix.n_occ, "build_s": index_s,
                     "note": "reference sketched by the same device kernel, grouped on the host once, open-addressing table in HBM"},
           "e2e": {"ms_per_step": 1e3 * e2e_s, "reads_per_s": n_reads / e2e_s, "pairs_per_s": pairs / e2e_s, "bases_per_s": bases / e2e_s,
                   "h2d_bytes_per_step": int(full["h2d_bytes"]), "d2h_bytes_per_step": int(full["d2h_bytes"])},
           "device_resident": {"ms_per_step": 1e3 * dev_s, "reads_per_s": n_reads / dev_s, "pairs_per_s": pairs / dev_s},
           "seed_stage_ms": stage_ms}
    # the reference's own seeding + chaining on the host threads, and parity per read
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyrefseed as rs
        if rank == 0 and rs.available():
            rix = rs.RefIndex([ref_b.tobytes()], w=10, k=15)
            rix.field("max_chain_skip", 2147483647)
            ref_mid = int(rix.field("mid_occ"))
            nc = min(n_reads, 1000)
            sub = off[:nc + 1].copy()
            t0 = time.perf_counter()
            r_na, r_nu, r_dig, _ = rix.seed_batch(buf[:sub[-1]], sub, chain=True, threads=host_threads)
            cpu_s = time.perf_counter() - t0
            t0 = time.perf_counter()
            rix.seed_batch(buf[:sub[-1]], sub, chain=False, threads=host_threads)
            cpu_seed_s = time.perf_counter() - t0
            bad = int(ref_mid != mid_occ)
            for r in range(nc):
                u = full["u"][full["u_pos"][r]:full["u_pos"][r] + full["n_u"][r]]
                b = full["b"][full["b_pos"][r]:full["b_pos"][r] + full["n_b"][r]]
                ok = int(full["n_u"][r]) == int(r_nu[r]) and int(full["a_off"][r + 1] - full["a_off"][r]) == int(r_na[r])
                ok = ok and rs.chain_digest(u, b) == int(r_dig[r])
                bad += 0 if ok else 1
            out["cpu_reference"] = {"kind": "reference", "cores": host_threads, "sample": "first %d reads" % nc,
                                    "seed_chain_reads_per_s": nc / cpu_s, "seed_only_reads_per_s": nc / cpu_seed_s,
                                    "what": "mm_map_seed + mg_lchain_dp (max-chain-skip = inf) per read on the host threads"}
            out["parity"] = {"reads_checked": nc, "mismatches": bad, "against": "reference mm_map_seed + mg_lchain_dp (oracle/_ref/libref_seed.so)",
                             "what": "per read: anchors seeded, n_u, digest of u[] and of the compacted anchors a'[]"}
            rix.close()
    except Exception as e:  # noqa: BLE001
        out["cpu_reference"] = {"error": repr(e)}
    sd.close(); ctx.close(); ix.close()
    return out, (n_reads, pairs), e2e_s


def emit(obj):
    """the ONE JSON line of the contract, on the process's original stdout"""
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # libraries print to stdout behind our back (NCCL: "NCCL version ..." at the first collective): everything but the result
    # line goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MM2GB_BENCH_WORKLOAD", "ont"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = WORKLOADS[args.workload]
    pkg = entry.load_package()
    cfg = {"workload": w["name"], "reads_per_gpu": w["n_reads"], "chaining": "map-ont: bw=500 max_dist=5000 max_iter=5000 max_skip=inf k=15",
           "sharding": f"reads sharded by rank, {world} rank(s), no collective", "l2": "inputs_larger_than_l2"}

    # ---- reference arm: the reference's CPU lchain.c on the host cores (rank 0 only) -------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        a, off = make_workload(w, 0)
        vals = []
        info = None
        for i in range(args.warmup + args.steps):
            info, pairs, dt, n_sample = cpu_baseline(a, off, budget_s=8.0)
            if i >= args.warmup:
                vals.append((pairs, dt, n_sample))
        pairs = sum(v[0] for v in vals); dt = sum(v[1] for v in vals); nr = sum(v[2] for v in vals)
        val = pairs / dt
        info["value"] = val
        emit(({"impl": "reference", "metric": "chaining anchor-pairs/s", "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                          "config": cfg, "reads_per_s": nr / dt, "cpu_baseline": info,
                          "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---- our arm ----------------------------------------------------------------------------------------------------
    # a rank runs next to its GPU: on a box with several NUMA nodes the process is confined to the CPUs of the GPU's node before
    # anything is allocated (mm2-gb_b200/sharding.py: place_rank); a no-op for one rank or without exposed topology.  Reported as
    # the line's "host_placement" (not inside "config": both arms print the same config)
    from mm2gb_b200 import sharding as _sharding
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    try:
        placement = {"pinned": False, "why": "MM2GB_BENCH_NO_PIN"} if os.environ.get("MM2GB_BENCH_NO_PIN") else _sharding.place_rank(local_rank, local_world)
    except Exception as e:      # placement is an optimisation: never the reason a run fails
        placement = {"pinned": False, "why": "place_rank failed: %s" % e}
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the chaining path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    a, off = make_workload(w, rank)
    n, n_reads = int(off[-1]), len(off) - 1
    misc = pkg.map_ont_misc()
    # whole batch resident; two slots (stream + scratch each) so that two batches can be in flight
    in_flight = max(1, min(2, int(os.environ.get("MM2GB_BENCH_IN_FLIGHT", "2"))))
    ctx = pkg.ChainContext(misc, device=local_rank, max_anchors=max(n, 1 << 20), max_reads=n_reads + 1, n_slots=in_flight,
                           flags=pkg.ChainContext.DEVICE_ONLY)
    streams = [torch.cuda.ExternalStream(ctx.stream_ptr(k), device=local_rank) for k in range(in_flight)]

    # device-resident inputs
    h_a = torch.from_numpy(a.view(np.int64)).pin_memory()
    h_off = torch.from_numpy(off).pin_memory()
    d_a = h_a.cuda(non_blocking=True)
    d_off = h_off.cuda(non_blocking=True)
    d_f = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(in_flight)]
    d_p = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(in_flight)]
    torch.cuda.synchronize()

    def step(slot=0):
        # the whole device side of mg_lchain_dp: range -> units -> score (f, p) -> chain extraction + compaction (u, a')
        ctx.chain_device(d_a, d_off, off, n_reads, n, d_f[slot], d_p[slot], slot=slot)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(max(3, args.warmup)):
        step(k % in_flight)
    for k in range(in_flight):
        ctx.sync(k)
    st = ctx.device_stats()
    pairs = int(st.n_pairs)
    sampler = ClockSampler(local_rank)
    # (1) one batch at a time on one stream, with the per-kernel CUDA-event timers on: the step's latency and the kernel times
    #     the roofline is computed from (a kernel timed while another batch's kernels share the GPU would not be its own time)
    ctx.profile(True)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for _ in range(args.steps):
        step(0)
    e1.record(streams[0])
    ctx.sync(0)
    barrier()
    ms_seq = e0.elapsed_time(e1)
    prof = ctx.profile_read()
    ctx.profile(False)
    # (2) the timed region of `value`: the same K steps with two batches in flight, alternating between the two slots -- batches
    #     are independent, and the latency-bound chain extraction of one overlaps the score kernels of the next.  Every step
    #     does all of its work; the clock is a pair of CUDA events on the launching streams around all K steps.
    ms = ms_seq
    if in_flight > 1:
        ej = [torch.cuda.Event(enable_timing=False) for _ in range(in_flight)]
        barrier()
        e0.record(streams[0])
        for k in range(1, in_flight):
            streams[k].wait_event(e0)           # no slot starts before the clock does
        for k in range(args.steps):
            step(k % in_flight)
        for k in range(1, in_flight):
            ej[k].record(streams[k])
            streams[0].wait_event(ej[k])        # the clock stops when every slot is done
        e1.record(streams[0])
        for k in range(in_flight):
            ctx.sync(k)
        barrier()
        ms = e0.elapsed_time(e1)

    # ---- end to end through the reference-facing boundary: init / chain / finish_stream_gpu (include/mm2gb_plchain.h), called
    #      the way `minimap2 -t T --gpu-chain` calls them (tests/fake_host.c: fake_drive): T worker threads, every read's anchors
    #      in its own kmalloc'd (pageable) array, one batch per thread and mini-batch, flush at the end of every mini-batch.
    #      Inside the timed region, per step: the gather pass into pinned staging (packed 8-byte wire format), H2D, all kernels,
    #      D2H of chains + chain-anchor indices, and on the calling threads kmalloc of u / a', compact_a's gather, kfree of the
    #      input arrays, post_chaining_helper.  The seeded reads of every step are built before the clock starts.
    import ctypes as C
    host_threads = max(1, cpu_threads() // max(1, placement["ranks_sharing"] if placement.get("pinned") else world))
    drv_threads = int(os.environ.get("MM2GB_BENCH_THREADS", "0")) or host_threads
    e2e_steps = max(1, min(args.steps, 8))
    D = C.CDLL(os.path.join(ROOT, "tests", "_build", "libdropin_test.so"))
    D.init_stream_gpu.argtypes = [C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, pkg.Misc]
    D.fake_set_misc.argtypes = [C.POINTER(pkg.Misc)]
    D.free_stream_gpu.argtypes = [C.c_int]
    D.fake_drive.restype = C.c_double
    D.fake_drive.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int] + [C.c_void_p] * 4
    D.mm2gb_dropin_traffic.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    os.environ["MM2GB_GPU_BASE"] = str(local_rank)     # this rank's worker threads all drive this rank's GPU
    os.environ["MM2GB_N_GPUS"] = "1"
    os.environ["MM2GB_THREADS_PER_GPU"] = str(drv_threads)
    cfg_path = os.path.join(ROOT, "mm2-gb_b200", "b200_config.json")
    D.fake_set_misc(C.byref(misc))
    mx, mr, mn = C.c_size_t(0), C.c_int(0), C.c_int(-1)
    D.init_stream_gpu(C.byref(mx), C.byref(mr), C.byref(mn), cfg_path.encode(), misc)
    dr = {"n_u": np.zeros(n_reads, np.int32), "n_b": np.zeros(n_reads, np.int64), "hu": np.zeros(n_reads, np.uint64), "hb": np.zeros(n_reads, np.uint64)}

    def drive(steps, want=True):
        outs = [dr[k].ctypes.data if want else None for k in ("n_u", "n_b", "hu", "hb")]
        return D.fake_drive(a.ctypes.data, off.ctypes.data, n_reads, drv_threads, 0, steps, int(mx.value), 1, *outs)
    drive(2, want=False)                                # warm-up: contexts are created, buffers pinned
    traffic = (C.c_longlong * 2)()
    D.mm2gb_dropin_traffic(traffic, 1)
    barrier()
    e2e_s = drive(e2e_steps)
    torch.cuda.synchronize()
    D.mm2gb_dropin_traffic(traffic, 1)
    dropin_h2d, dropin_d2h = traffic[0] / e2e_steps, traffic[1] / e2e_steps
    D.free_stream_gpu(drv_threads)

    # ---- the same through the core C ABI (mm2gb_chain_host_index) with pinned HOST buffers: no host pass at all -- the anchors
    #      are DMA'd from the caller's pinned array (16 B each), the indices land in the caller's pinned array (k_drain)
    e2e_cap = max(1 << 20, n // 8 + int(np.diff(off).max()) + 1)
    ctx_e2e = pkg.ChainContext(misc, device=local_rank, max_anchors=e2e_cap, max_reads=n_reads + 1, n_slots=6)
    out = {"u": np.empty(n, np.uint64), "v": torch.empty(n, dtype=torch.int32).pin_memory(),
           "n_u": np.zeros(n_reads, np.int32), "n_b": np.zeros(n_reads, np.int64), "v_pos": np.zeros(n_reads, np.int64)}
    for _ in range(2):
        ctx_e2e.chain(h_a, off, out=out, packed=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = ctx_e2e.chain(h_a, off, out=out, packed=True)
    torch.cuda.synchronize()
    core_s = time.perf_counter() - t0
    clocks = sampler.stop()
    n_chains = int(res["n_u"].sum())
    n_chain_anchors = int(res["n_b"].sum())
    core_abi = {"value": e2e_steps * pairs / core_s, "unit": "pairs/s", "ms_per_step": 1e3 * core_s / e2e_steps,
                "h2d_bytes_per_step": int(res["h2d_anchor_bytes"]) + 8 * (n_reads + 1), "d2h_bytes_per_step": 4 * n_chain_anchors + 8 * n_chains + 16 * (n_reads + 64),
                "note": "mm2gb_chain_host_index, one synchronous caller, caller-pinned anchors (raw 16 B DMA, no host pass) and caller-pinned index output; 6 slots"}
    # the packed upload through the same call: a pageable source is packed (8 B/anchor) by one host thread per call
    t0 = time.perf_counter()
    resn = ctx_e2e.chain(a, off, out=out, packed=True)
    core_abi["pageable_source_packed_upload"] = {"ms_per_step": 1e3 * (time.perf_counter() - t0), "h2d_anchor_bytes": int(resn["h2d_anchor_bytes"]),
                                                 "note": "single caller thread does the 16 -> 8 byte pack pass"}
    # where the time goes: DP only (upload, kernels, f/p download) and, as a diagnostic, the host-stage variant of the same call
    # (f/p downloaded, chain extraction on host threads -- the reference's arrangement)
    outh = {"f": torch.empty(n, dtype=torch.int32).pin_memory(), "p": torch.empty(n, dtype=torch.int32).pin_memory(),
            "u": out["u"], "b": np.empty((n, 2), np.uint64), "n_u": np.zeros(n_reads, np.int32), "n_b": np.zeros(n_reads, np.int64)}
    ctx_e2e.chain_dp(h_a, off, f=outh["f"], p=outh["p"])
    t0 = time.perf_counter()
    for _ in range(3):
        ctx_e2e.chain_dp(h_a, off, f=outh["f"], p=outh["p"])
    dp_only_s = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    resh = ctx_e2e.chain(h_a, off, n_threads=host_threads, out=outh)
    hostvar_s = time.perf_counter() - t0
    host_only_s = host_stage_alone(pkg, misc, a, off, outh["f"].numpy(), outh["p"].numpy(), host_threads)

    # ---- parity at benchmark scale (outside every timed region): the chains and compacted anchors the boundary returned for
    #      this very workload against the reference's own lchain.c (oracle/_ref; the restatement where that was not built)
    po = entry.load_oracle()
    use_ref = po.ref_available()
    order = np.argsort(-np.diff(off), kind="stable")
    if world == 1:
        sel = np.arange(n_reads, dtype=np.int64)
    else:   # the 50 longest reads + every k-th of the rest
        rest = order[50:]
        sel = np.sort(np.concatenate([order[:50], rest[::max(1, len(rest) // 1950)]])).astype(np.int64)
    r_nu, r_nb, r_hu, r_hb = po.lchain_digests(po.map_ont_params(), a, off, sel, n_threads=host_threads, use_ref=use_ref)
    bad_dropin = int(np.count_nonzero((dr["n_u"][sel] != r_nu) | (dr["n_b"][sel] != r_nb) | (dr["hu"][sel] != r_hu) | (dr["hb"][sel] != r_hb)))
    bad_core = 0
    vv = res["v"].numpy()
    for k, r in enumerate(sel):
        s0, q = int(off[r]), int(res["v_pos"][r])
        nu_r, nb_r = int(res["n_u"][r]), int(res["n_b"][r])
        ok = nu_r == int(r_nu[k]) and nb_r == int(r_nb[k])
        ok = ok and po.digest(res["u"][s0:s0 + nu_r]) == int(r_hu[k])
        ok = ok and po.digest(a[s0 + vv[q:q + nb_r].astype(np.int64)]) == int(r_hb[k])
        bad_core += 0 if ok else 1
    bad_hostvar = int(not (np.array_equal(resh["n_u"], res["n_u"]) and np.array_equal(resh["n_b"], res["n_b"])))
    parity = {"reads_checked": int(len(sel)), "mismatches": bad_dropin + bad_core + bad_hostvar, "against": "reference lchain.c (oracle/_ref)" if use_ref else "oracle port",
              "includes_longest_reads": 50, "paths": {"dropin_chain_stream_gpu": bad_dropin, "core_abi_chain_host_index": bad_core, "host_stage_variant_counts": bad_hostvar},
              "what": "per read: n_u, number of chain anchors, digest of u[], digest of the compacted anchors a'[]"}

    # ---- row N2: device seeding + chaining fused (sequences in, chains out) ------------------------------------------------
    seed_leg = None
    from mm2gb_b200 import sharding
    if args.workload in ("ont", "mini") and not os.environ.get("MM2GB_BENCH_NO_SEED"):
        try:
            seed_leg, (s_reads, s_pairs), s_sec = seed_chain_leg(pkg, torch, dist, w, rank, local_rank, world, host_threads,
                                                                n_reads=min(3000, w["n_reads"]))
        except Exception as e:  # noqa: BLE001 -- the leg must not take the headline line down with it; the failure is reported in the line
            sys.stderr.write("bench.py: seed_chain leg failed: %r\n" % (e,))
            seed_leg, s_reads, s_pairs, s_sec = {"error": repr(e)}, 0, 0, 0.0
        # every rank takes part in the reduction, whether its leg worked or not
        (j_reads, j_pairs, j_ok), (j_sec,) = sharding.reduce_job(dist, [s_reads, s_pairs, 0 if "error" in seed_leg else 1], [s_sec], device="cuda")
        if "error" not in seed_leg and int(j_ok) == world and j_sec > 0:
            seed_leg["job"] = {"n_gpus": world, "reads_per_s": j_reads / j_sec, "pairs_per_s": j_pairs / j_sec,
                               "note": "every rank seeds + chains its own reads; sum of work / max over ranks of the end-to-end time"}

    # ---- reduce over ranks: max time, sum of work -------------------------------------------------------------------
    (tot_pairs, tot_anchors, tot_reads), (ms_max, e2e_max) = sharding.reduce_job(dist, [pairs, n, n_reads], [ms, e2e_s], device="cuda")
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    sec = ms_max / 1e3
    value = tot_pairs * args.steps / sec
    # kernels launched per step (mirrors enqueue_kernels / enqueue_backtrack in chain_core.cu): 6 range/unit kernels (k_block_reads,
    # k_range, k_scan, k_units, k_unit_clip, k_order), k_score_long + the two k_score_units instantiations + k_score_exact, then a
    # sort + a walk kernel per non-empty chain-extraction size class (7 shared-memory classes, the 9 mid classes up to 196608
    # anchors, the global-memory class) and the overflow pass
    rn = np.diff(off)
    small = [1024, 1536, 2048, 3072, 4096, 6144, 8192]
    mid = [10048, 13952, 19776, 29504, 37248, 48640, 65536, 98304, 196608]
    mid_min = max(8192, int(os.environ.get("MM2GB_BT_MID_MIN", "8192")))

    def bt_class(x):
        if x <= small[-1]:
            return int(np.searchsorted(small, x))
        if x <= mid_min or x > mid[-1]:
            return 99
        return 7 + int(np.searchsorted(mid, x))
    bt_set = {bt_class(int(x)) for x in rn}
    bt_classes = len(bt_set) + (1 if any(c < 7 for c in bt_set) else 0)
    dp_ms = sum(prof[k][0] / max(1, prof[k][1]) for k in ("range", "units", "score"))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # DRAM bytes of one k_score_units launch on this workload from the committed `ncu --set full` capture (profiles/traffic.json)
    traffic, sass_per_pair = None, 15.54
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(args.workload, {}).get("dram_bytes_per_launch")
        sass_per_pair = float(tj.get("ont", {}).get("thread_instr_per_pair", sass_per_pair))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    score_ms, score_n = prof["score"]
    score_avg_s = score_ms / 1e3 / max(1, score_n)
    hbm_gbs = HBM_BYTES_PER_ANCHOR * n / score_avg_s / 1e9
    sm_mhz = clocks["sm_mhz"] or float(peaks.get("sm_max_mhz", 1965.0))
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
    # integer / issue ceilings MEASURED on this GPU type by tools/int_peak.cu (profiles/int_peak.json): both integer pipes kept
    # busy by independent add + mad chains at the score kernel's own residency -> thread-instructions/s; and the speed of light
    # of a kernel made of nothing but the MID loop of score_unit_packed.  Scaled by the SM clock of this run.
    ip = {}
    try:
        ip = json.load(open(os.path.join(ROOT, "profiles", "int_peak.json")))
    except Exception:
        pass
    nominal_issue = n_sm * 4 * 32 * sm_mhz * 1e6          # 4 schedulers x 32 lanes per SM
    clk_scale = sm_mhz / 1965.0
    issue_peak = float(ip.get("issue_mix", {}).get("thread_instr_per_s", 0.0)) * clk_scale or nominal_issue
    issue_src = "measured (tools/int_peak.cu issue_mix, profiles/int_peak.json)" if ip else "computed (148 SM x 128 lanes x clock)"
    mid_peak = float(ip.get("mid_mix", {}).get("pairs_per_s", 0.0)) * clk_scale
    issue_ach = INSTR_PER_PAIR * pairs / score_avg_s
    kernel_pairs_s = pairs / score_avg_s
    roofline = {"bound": "int-issue", "kernel": "k_score_units", "achieved": issue_ach, "peak": issue_peak, "unit": "int-ops/s", "frac": issue_ach / issue_peak,
                "peak_source": issue_src, "int_ops_per_pair": INSTR_PER_PAIR, "pairs_per_s_kernel": kernel_pairs_s, "sm_mhz": sm_mhz, "n_sm": n_sm,
                "kernel_ms": score_avg_s * 1e3, "kernel_share_of_step": score_ms / ms_seq,
                "traffic": traffic, "algorithmic_bytes_per_launch": HBM_BYTES_PER_ANCHOR * n,
                "note": "achieved = %.0f algorithmic integer ops per pair (SURVEY.md 8d) x pairs / kernel time; the kernel executes %.2f SASS thread-instructions per pair "
                        "on configs[1] (profiles/traffic.json, from the ncu capture named there), i.e. %.2f of the measured issue peak" % (INSTR_PER_PAIR, sass_per_pair, sass_per_pair * kernel_pairs_s / issue_peak),
                "mid_loop_speed_of_light": {"pairs_per_s": mid_peak or None, "frac": (kernel_pairs_s / mid_peak) if mid_peak else None,
                                            "note": "a kernel of MID pairs only (6.6 instr/pair, LSU-bound); the score kernel also pays the in-tile triangle, GEN / FAR pairs and tile set-up"},
                "hbm": {"bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak, "peak_source": peak_src, "traffic": traffic,
                        "note": "not the binding bound: ~%d pairs per 24-byte anchor" % round(pairs / max(1, n))}}
    line = {"metric": "chaining anchor-pairs/s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_max / args.steps, "batches_in_flight": in_flight,
            "one_batch_at_a_time": {"ms_per_step": ms_seq / args.steps, "value": pairs * args.steps / (ms_seq / 1e3), "unit": "pairs/s",
                                    "note": "the same steps strictly one after the other on one stream (latency of a step; the per-kernel timers and the roofline come from this run)"},
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": cfg, "host_placement": placement, "reads_per_s": tot_reads * args.steps / sec, "anchors_per_s": tot_anchors * args.steps / sec,
            "batch": {"reads": n_reads, "anchors": n, "pairs": pairs, "pairs_per_anchor": pairs / max(1, n), "units": int(st.n_units),
                      "units_exact": int(st.n_units_exact), "units_long_kernel": int(st.n_long), "chains": n_chains, **workload_shape(a, off)},
            "kernel_ms_per_step": {k: v[0] / max(1, v[1]) for k, v in prof.items() if v[1]},
            "dp_only": {"value": pairs / (dp_ms / 1e3), "unit": "pairs/s", "ms": dp_ms,
                        "note": "range + units + score kernels only (f, p); `value` also includes the device chain extraction + compaction"},
            "roofline": roofline,
            "e2e": {"value": tot_pairs * e2e_steps / e2e_max, "unit": "pairs/s", "h2d_bytes_per_step": int(dropin_h2d),
                    "d2h_bytes_per_step": int(dropin_d2h),
                    "reads_per_s": tot_reads * e2e_steps / e2e_max, "ms_per_step": 1e3 * e2e_max / e2e_steps,
                    "through": "init_stream_gpu / chain_stream_gpu / finish_stream_gpu (the reference's plugin boundary, gpu/plutils.h:98-104), driven like `minimap2 -t %d --gpu-chain`: %d worker thread(s), per-read kmalloc'd pageable anchor arrays, one batch per thread and mini-batch, flush per mini-batch" % (drv_threads, drv_threads),
                    "includes": "host gather into pinned staging (packed 8-byte wire format), H2D, k_expand + range + unit + score kernels, device chain extraction + compaction, D2H of chains + chain-anchor indices (k_drain), kmalloc of u / a', compact_a's gather on the calling threads, kfree of the input arrays, post_chaining_helper (= whole mg_lchain_dp as the driver sees it)",
                    "driver_threads": drv_threads, "chain_anchors": n_chain_anchors, "steps": e2e_steps, "batch_limit_anchors": int(mx.value),
                    "bytes_per_anchor": {"h2d": dropin_h2d / max(1, n), "d2h": dropin_d2h / max(1, n)},
                    "core_abi": core_abi,
                    "breakdown_ms": {"dp_only_upload_kernels_fp_download": 1e3 * dp_only_s, "host_stage_variant_same_call": 1e3 * hostvar_s,
                                     "host_stage_alone": 1e3 * host_only_s, "host_threads": host_threads}},
            "parity": parity, "seed_chain": seed_leg,
            "gpu_launches": (10 + 2 * bt_classes) * args.steps, "clocks": clocks}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(a, off)[0]
    emit(line)
    ctx.close()
    ctx_e2e.close()
    if dist is not None:
        dist.destroy_process_group()
    if seed_leg is not None and seed_leg.get("parity", {}).get("mismatches"):
        sys.stderr.write("bench.py: PARITY FAILURE (seed_chain): %r\n" % (seed_leg["parity"],))
        sys.exit(1)
    if parity["mismatches"]:
        sys.stderr.write("bench.py: PARITY FAILURE: %r\n" % (parity,))
        sys.exit(1)


if __name__ == "__main__":
    main()
