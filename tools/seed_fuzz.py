"""Randomised differential test of the device seeding path against the reference's mm_map_seed (oracle/_ref/libref_seed.so):
random references (repeat families, tandem arrays of random period, low-complexity stretches, N runs, several contigs, lower
case), random supported (w, k), random seeding parameters, reads from the reference with errors plus junk reads.  Every read's
anchors (order included), rep_len and mini_pos must agree.  Prints one JSON line; exits non-zero on a mismatch.

    python tools/seed_fuzz.py [--rounds 12] [--seed 1]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as entry  # noqa: E402

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def make_ref(rng, n):
    ref = bytearray(ACGT[rng.integers(0, 4, n)].tobytes())
    for _ in range(int(rng.integers(0, 4))):                       # repeat families
        ul = int(rng.integers(200, 3000))
        unit = ACGT[rng.integers(0, 4, ul)].tobytes()
        for p in rng.integers(0, n - ul - 1, int(rng.integers(3, 60))):
            u = bytearray(unit)
            for q in rng.integers(0, ul, max(1, int(ul * rng.uniform(0, 0.05)))):
                u[q] = ACGT[rng.integers(0, 4)]
            ref[p:p + ul] = u
    for _ in range(int(rng.integers(0, 4))):                       # tandem arrays
        per = int(rng.integers(1, 120))
        cnt = int(rng.integers(5, 200))
        p = int(rng.integers(0, max(1, n - per * cnt - 1)))
        ref[p:p + per * cnt] = (ACGT[rng.integers(0, 4, per)].tobytes() * cnt)[:max(0, min(per * cnt, n - p))]
    for _ in range(int(rng.integers(0, 6))):                       # N runs
        p = int(rng.integers(0, n - 1)); ln = int(rng.integers(1, 80))
        ref[p:p + ln] = b"N" * min(ln, n - p)
    ref = bytes(ref[:n])
    if rng.random() < 0.3:
        ref = ref.lower()
    return ref


def make_reads(rng, refs, n_reads):
    reads = []
    for i in range(n_reads):
        ref = refs[int(rng.integers(0, len(refs)))]
        kind = rng.random()
        if kind < 0.08:
            reads.append(ACGT[rng.integers(0, 4, int(rng.integers(1, 400)))].tobytes())        # junk
            continue
        ln = int(min(len(ref) - 1, rng.integers(1, 60) if kind < 0.15 else rng.integers(200, 40000)))
        st = int(rng.integers(0, len(ref) - ln))
        r = bytearray(ref[st:st + ln].upper())
        err = rng.uniform(0, 0.15)
        for q in rng.integers(0, ln, int(ln * err)):
            r[q] = ACGT[rng.integers(0, 4)]
        if rng.random() < 0.05 and ln > 10:
            r[int(rng.integers(0, ln))] = ord("N")
        r = bytes(r)
        reads.append(r.translate(COMP)[::-1] if rng.random() < 0.5 else r)
    return reads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=12)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--reads", type=int, default=150)
    ap.add_argument("--chain", action="store_true", help="also run the fused seed + chain step and compare chains / compacted anchors with mm_map_seed + mg_lchain_dp")
    args = ap.parse_args()
    pkg = entry.load_package()
    from mm2gb_b200 import seed
    import pyrefseed as rs
    rng = np.random.default_rng(args.seed)
    total, bad, detail = 0, 0, []
    for rnd in range(args.rounds):
        w, k = [(10, 15), (5, 15), (19, 19), (11, 21), (10, 13), (3, 11), (32, 27), (1, 15)][int(rng.integers(0, 8))]
        refs = [make_ref(rng, int(rng.integers(20000, 400000))) for _ in range(int(rng.integers(1, 4)))]
        reads = make_reads(rng, refs, args.reads)
        mid = int(rng.choice([1, 2, 3, 5, 10, 20, 50, 200]))
        dist = int(rng.choice([0, 50, 100, 500, 2000]))
        frac = float(rng.choice([0.0, 0.002, 0.01, 0.05]))
        mmo = int(rng.choice([mid, mid + 3, 100, 4095]))
        hpc = bool(rng.random() < 0.3)
        rix = rs.RefIndex(refs, w=w, k=k, hpc=hpc)
        rix.field("mid_occ", mid); rix.field("occ_dist", dist); rix.field("q_occ_frac", frac); rix.field("max_max_occ", mmo)
        buf, off = seed.pack_seqs(reads)
        with seed.Index(refs, w=w, k=k, hpc=hpc) as ix, seed.Seeder(ix, max_bases=int(off[-1]) + 1024, max_reads=len(reads) + 8, max_anchors=1 << 26) as sd:
            prm = seed.map_ont_seed_params(mid, occ_dist=dist, q_occ_frac=frac, max_max_occ=mmo)
            a, a_off, rep, mp, mp_off = sd.seed(prm, buf, off)
        if args.chain:
            rix.field("max_chain_skip", 2147483647)
            n_a, n_u, dig, _ = rix.seed_batch(buf, off, chain=True, threads=8)
            misc = pkg.Misc.from_buffer_copy(rix.misc())
            cap = int(n_a.sum()) + 1024
            with seed.Index(refs, w=w, k=k, hpc=hpc) as ix, pkg.ChainContext(misc, device=0, max_anchors=cap, max_reads=len(reads) + 8, n_slots=1,
                                                                    flags=pkg.ChainContext.DEVICE_ONLY) as ctx, \
                    seed.Seeder(ix, max_bases=int(off[-1]) + 1024, max_reads=len(reads) + 8, max_anchors=cap) as sd:
                res = sd.seed_chain(ctx, prm, buf, off)
            for r in range(len(reads)):
                u = res["u"][res["u_pos"][r]:res["u_pos"][r] + res["n_u"][r]]
                b = res["b"][res["b_pos"][r]:res["b_pos"][r] + res["n_b"][r]]
                ok = int(res["n_u"][r]) == int(n_u[r]) and int(res["a_off"][r + 1] - res["a_off"][r]) == int(n_a[r]) and rs.chain_digest(u, b) == int(dig[r])
                total += 1
                if not ok:
                    bad += 1
                    if len(detail) < 10:
                        detail.append({"round": rnd, "read": r, "len": len(reads[r]), "w": w, "k": k, "mid_occ": mid, "chain": True,
                                       "n_u": [int(res["n_u"][r]), int(n_u[r])], "n_a": [int(res["a_off"][r + 1] - res["a_off"][r]), int(n_a[r])]})
        for r, read in enumerate(reads):
            ea, erep, emp = rix.seed(read)
            ga, gmp = a[a_off[r]:a_off[r + 1]], mp[mp_off[r]:mp_off[r + 1]]
            total += 1
            if not (ga.shape == ea.shape and np.array_equal(ga, ea) and int(rep[r]) == erep and np.array_equal(gmp, emp)):
                bad += 1
                if len(detail) < 10:
                    same_set = ga.shape == ea.shape and np.array_equal(ga[np.lexsort((ga[:, 1], ga[:, 0]))], ea[np.lexsort((ea[:, 1], ea[:, 0]))])
                    detail.append({"round": rnd, "read": r, "len": len(read), "w": w, "k": k, "hpc": hpc, "mid_occ": mid, "occ_dist": dist, "q_occ_frac": frac,
                                   "max_max_occ": mmo, "anchors": [int(len(ga)), int(len(ea))], "rep": [int(rep[r]), erep],
                                   "mini_pos": [int(len(gmp)), int(len(emp))], "same_anchor_set": bool(same_set)})
        rix.close()
    print(json.dumps({"reads": total, "mismatches": bad, "rounds": args.rounds, "seed": args.seed, "detail": detail}))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
