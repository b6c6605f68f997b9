"""Writes tests/golden/seed_golden.npz from the reference's own seeding code (oracle/_ref/libref_seed.so = sketch.c, index.c,
seed.c, map.c compiled where they lie): minimizers, anchors, rep_len and mini_pos of the deterministic cases of
tests/seed_cases.py under several parameter sets.  Run in the build container (needs /root/reference for `make -C oracle ref`);
the fixture travels with the repository, the reference does not."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import pyrefseed as rs  # noqa: E402
import seed_cases  # noqa: E402

PARAM_SETS = [  # (mid_occ or None = what mm_mapopt_update derives, occ_dist, q_occ_frac)
    (None, 500, 0.01), (3, 500, 0.01), (3, 0, 0.01), (5, 100, 0.0), (2, 50, 0.002)]


def main():
    refs = seed_cases.make_reference()
    reads = seed_cases.make_reads(refs)
    out = {}
    ix = rs.RefIndex(refs, w=10, k=15)
    out["mid_occ_default"] = np.int64(ix.field("mid_occ"))
    out["cal_max_occ_2e-4"] = np.int64(ix.mid_occ_of(2e-4))
    out["cal_max_occ_1e-2"] = np.int64(ix.mid_occ_of(1e-2))
    mv = [rs.sketch(r, 10, 15) for r in reads]
    out["mv"] = np.concatenate(mv)
    out["mv_off"] = np.cumsum([0] + [len(m) for m in mv]).astype(np.int64)
    for pi, (mid, dist, frac) in enumerate(PARAM_SETS):
        ix.field("mid_occ", out["mid_occ_default"] if mid is None else mid)
        ix.field("occ_dist", dist)
        ix.field("q_occ_frac", frac)
        a, rep, mp = [], [], []
        for r in reads:
            ai, ri, mi = ix.seed(r) if len(r) else (np.zeros((0, 2), np.uint64), 0, np.zeros(0, np.uint64))
            a.append(ai); rep.append(ri); mp.append(mi)
        out[f"a_{pi}"] = np.concatenate(a)
        out[f"a_off_{pi}"] = np.cumsum([0] + [len(x) for x in a]).astype(np.int64)
        out[f"rep_{pi}"] = np.array(rep, dtype=np.int32)
        out[f"mp_{pi}"] = np.concatenate(mp)
        out[f"mp_off_{pi}"] = np.cumsum([0] + [len(x) for x in mp]).astype(np.int64)
    # a second index geometry (map-hifi / asm: k = 19, w = 19)
    ix2 = rs.RefIndex(refs, w=19, k=19)
    ix2.field("mid_occ", 10)
    mv2 = [rs.sketch(r, 19, 19) for r in reads]
    out["mv_k19"] = np.concatenate(mv2)
    out["mv_off_k19"] = np.cumsum([0] + [len(m) for m in mv2]).astype(np.int64)
    a = [ix2.seed(r)[0] for r in reads]
    out["a_k19"] = np.concatenate(a)
    out["a_off_k19"] = np.cumsum([0] + [len(x) for x in a]).astype(np.int64)
    # homopolymer-compressed minimizers (map-pb: MM_I_HPC, k = 19, w = 10): minimizers and anchors carry their own spans
    ix3 = rs.RefIndex(refs, w=10, k=19, hpc=True, preset="map-pb")
    ix3.field("mid_occ", 10)
    mv3 = [rs.sketch(r, 10, 19, hpc=True) for r in reads]
    out["mv_hpc"] = np.concatenate(mv3)
    out["mv_off_hpc"] = np.cumsum([0] + [len(m) for m in mv3]).astype(np.int64)
    res3 = [ix3.seed(r) for r in reads]
    out["a_hpc"] = np.concatenate([x[0] for x in res3])
    out["a_off_hpc"] = np.cumsum([0] + [len(x[0]) for x in res3]).astype(np.int64)
    out["rep_hpc"] = np.array([x[1] for x in res3], dtype=np.int32)
    out["mp_hpc"] = np.concatenate([x[2] for x in res3])
    path = os.path.join(HERE, "..", "tests", "golden", "seed_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
