"""Device seeding + chaining on a configs[1]-shaped workload (N2): index build time, stage timers of the seeding kernels,
the fused sequences-in / chains-out step (end to end from host memory, and with the reads resident in HBM), next to the
reference's mm_map_seed (+ mg_lchain_dp) on the host threads.  Prints one JSON line.

    python tools/seed_bench.py [--ref-len 100000000] [--reads 10000] [--lo 10000] [--hi 100000] [--steps 5] [--cpu-reads 600]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-len", type=int, default=100_000_000)
    ap.add_argument("--repeats", type=int, default=6000)
    ap.add_argument("--reads", type=int, default=10000)
    ap.add_argument("--lo", type=int, default=10000)
    ap.add_argument("--hi", type=int, default=100000)
    ap.add_argument("--err", type=float, default=0.10)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-reads", type=int, default=600)
    ap.add_argument("--check-reads", type=int, default=100)
    args = ap.parse_args()
    pkg = entry.load_package()
    from mm2gb_b200 import seed, synth
    import torch
    nt = synth._NT
    t0 = time.time()
    ref = synth.simulate_reference(args.ref_len, seed=1, n_repeat_copies=args.repeats, repeat_unit=3000)
    reads = synth.simulate_reads(ref, args.reads, args.lo, args.hi, seed=2, err=args.err)
    ref_b = nt[ref]
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    buf = nt[np.concatenate(reads)]
    gen_s = time.time() - t0
    out = {"workload": f"{args.ref_len} bp random reference + {args.repeats} x 3 kb repeat copies, {args.reads} reads U[{args.lo},{args.hi}] @ {args.err}",
           "bases": int(off[-1]), "generation_s": round(gen_s, 2)}
    t0 = time.time()
    ix = seed.Index((ref_b, np.array([0, len(ref_b)], dtype=np.int64)), w=10, k=15)
    out["index_build_s"] = round(time.time() - t0, 2)
    out["index"] = {"keys": ix.n_keys, "occurrences": ix.n_occ, "mid_occ": ix.mid_occ()}
    prm = seed.map_ont_seed_params(ix.mid_occ())
    misc = pkg.map_ont_misc()
    # size the buffers from one seeding pass
    probe = seed.Seeder(ix, max_bases=int(off[-1]) + 4096, max_reads=len(reads) + 8, max_anchors=int(off[-1]) * 2)
    a, a_off, rep, _, _ = probe.seed(prm, buf, off, want_mini_pos=False)
    n_a = int(a_off[-1])
    prof, n_mv, n_m = probe.profile()
    out["batch"] = {"reads": len(reads), "minimizers": n_mv, "seeds": n_m, "anchors": n_a, "anchors_per_read": n_a / len(reads)}
    out["seed_stage_ms"] = {k: round(v, 3) for k, v in prof.items()}
    del a
    probe.close()
    cap = n_a + 4096
    ctx = pkg.ChainContext(misc, device=0, max_anchors=cap, max_reads=len(reads) + 8, n_slots=1, flags=pkg.ChainContext.DEVICE_ONLY)
    sd = seed.Seeder(ix, max_bases=int(off[-1]) + 4096, max_reads=len(reads) + 8, max_anchors=cap)
    pin = torch.from_numpy(buf).pin_memory()
    # end to end: pinned host sequences in, chains + compacted anchors in host memory out
    res = None
    times = []
    for it in range(args.steps + 2):
        t0 = time.perf_counter()
        res = sd.seed_chain(ctx, prm, pin.data_ptr(), off, copy=False)
        times.append(time.perf_counter() - t0)
    e2e = float(np.median(times[2:]))
    pairs = int(res.stats.n_pairs)
    out["fused_e2e"] = {"ms_per_step": 1e3 * e2e, "reads_per_s": len(reads) / e2e, "pairs_per_s": pairs / e2e, "bases_per_s": int(off[-1]) / e2e,
                        "h2d_bytes": int(res.h2d_bytes), "d2h_bytes": int(res.d2h_bytes), "pairs": pairs, "chains": int(res.n_chains),
                        "chain_anchors": int(res.n_chain_anchors)}
    prof, _, _ = sd.profile()
    out["fused_e2e"]["seed_stage_ms"] = {k: round(v, 3) for k, v in prof.items()}
    # device-resident: reads already in HBM, results stay there
    d_seq = torch.from_numpy(buf).cuda()
    times = []
    for it in range(args.steps + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sd.seed_chain_device(ctx, prm, d_seq.data_ptr(), off)
        ctx.sync(0)
        times.append(time.perf_counter() - t0)
    dev = float(np.median(times[2:]))
    out["fused_device"] = {"ms_per_step": 1e3 * dev, "reads_per_s": len(reads) / dev, "pairs_per_s": pairs / dev}
    # two callers (host threads), each with its own seeder + chaining context and half of the reads: the latency-bound kernels of one
    # batch (x-sort, chain extraction) overlap the issue-bound ones (sketch, score) of the other
    import threading
    half = len(reads) // 2
    parts = []
    for lo, hi in ((0, half), (half, len(reads))):
        o = (off[lo:hi + 1] - off[lo]).copy()
        b = buf[off[lo]:off[hi]].copy()
        c2 = pkg.ChainContext(misc, device=0, max_anchors=cap, max_reads=len(reads) + 8, n_slots=1, flags=pkg.ChainContext.DEVICE_ONLY)
        s2 = seed.Seeder(ix, max_bases=int(o[-1]) + 4096, max_reads=len(reads) + 8, max_anchors=cap)
        parts.append((s2, c2, torch.from_numpy(b).pin_memory(), o))
    def run(part, n):
        s2, c2, pb, o = part
        for _ in range(n):
            s2.seed_chain(c2, prm, pb.data_ptr(), o, copy=False)
    for part in parts:
        run(part, 2)
    t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(part, args.steps)) for part in parts]
    for t in th:
        t.start()
    for t in th:
        t.join()
    two = (time.perf_counter() - t0) / args.steps
    out["fused_e2e_two_callers"] = {"ms_per_step": 1e3 * two, "reads_per_s": len(reads) / two, "pairs_per_s": pairs / two}
    for s2, c2, _, _ in parts:
        s2.close(); c2.close()
    # parity + CPU baseline: the reference's mm_map_seed + mg_lchain_dp on the host threads
    try:
        po_dir = os.path.join(ROOT, "oracle")
        sys.path.insert(0, po_dir)
        import pyrefseed as rs
        if rs.available():
            t0 = time.time()
            rix = rs.RefIndex([ref_b.tobytes()], w=10, k=15)
            out["reference_index_build_s"] = round(time.time() - t0, 2)
            rix.field("max_chain_skip", 2147483647)
            assert int(rix.field("mid_occ")) == ix.mid_occ(), (rix.field("mid_occ"), ix.mid_occ())
            nc = min(args.cpu_reads, len(reads))
            threads = os.cpu_count() or 1
            sub_off = off[:nc + 1].copy()
            t0 = time.perf_counter()
            n_a_ref, n_u_ref, dig, _ = rix.seed_batch(buf[:sub_off[-1]], sub_off, chain=True, threads=threads)
            cpu_s = time.perf_counter() - t0
            t0 = time.perf_counter()
            rix.seed_batch(buf[:sub_off[-1]], sub_off, chain=False, threads=threads)
            cpu_seed_s = time.perf_counter() - t0
            out["cpu_reference"] = {"reads": nc, "threads": threads, "seed_chain_reads_per_s": nc / cpu_s, "seed_only_reads_per_s": nc / cpu_seed_s,
                                    "seed_chain_s": cpu_s, "seed_only_s": cpu_seed_s}
            full = sd.seed_chain(ctx, prm, buf, off)
            mism = 0
            ncheck = min(args.check_reads, nc)
            for r in range(ncheck):
                u = full["u"][full["u_pos"][r]:full["u_pos"][r] + full["n_u"][r]]
                b = full["b"][full["b_pos"][r]:full["b_pos"][r] + full["n_b"][r]]
                h = rs.chain_digest(u, b)
                if h != int(dig[r]) or int(full["n_u"][r]) != int(n_u_ref[r]) or int(full["a_off"][r + 1] - full["a_off"][r]) != int(n_a_ref[r]):
                    mism += 1
            out["parity"] = {"reads_checked": ncheck, "mismatches": mism, "against": "reference mm_map_seed + mg_lchain_dp (oracle/_ref/libref_seed.so)"}
    except Exception as e:  # noqa: BLE001
        out["cpu_reference"] = {"error": repr(e)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
