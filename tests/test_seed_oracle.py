"""CPU-side tests of the seeding row: the checker (oracle/_ref/libref_seed.so = the reference's own sketch.c /
index.c / seed.c / map.c) against the committed golden vectors, the pure-Python restatement of the seeding stage in the data-parallel form the
kernels use (oracle/seed_model.py) against the reference, and the exported C ABI.  No GPU needed."""
import os
import re
import sys

import numpy as np
import pytest

import seed_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyrefseed as rs  # noqa: E402

needs_ref = pytest.mark.skipif(not rs.available(), reason="oracle/_ref/libref_seed.so not built (needs /root/reference; make -C oracle ref)")

PARAM_SETS = [(None, 500, 0.01), (3, 500, 0.01), (3, 0, 0.01), (5, 100, 0.0), (2, 50, 0.002)]


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "seed_golden.npz"))


@pytest.fixture(scope="module")
def cases():
    refs = seed_cases.make_reference()
    return refs, seed_cases.make_reads(refs)


@needs_ref
def test_reference_matches_golden(golden, cases):
    """The checker reproduces the committed vectors (they were generated from it: guards against a drifting build)."""
    refs, reads = cases
    ix = rs.RefIndex(refs, w=10, k=15)
    assert int(ix.field("mid_occ")) == int(golden["mid_occ_default"])
    assert ix.mid_occ_of(2e-4) == int(golden["cal_max_occ_2e-4"])
    mv = np.concatenate([rs.sketch(r, 10, 15) for r in reads])
    assert np.array_equal(mv, golden["mv"])
    for pi, (mid, dist, frac) in enumerate(PARAM_SETS):
        ix.field("mid_occ", int(golden["mid_occ_default"]) if mid is None else mid)
        ix.field("occ_dist", dist)
        ix.field("q_occ_frac", frac)
        off = golden[f"a_off_{pi}"]
        for r, read in enumerate(reads):
            a, rep, mp = ix.seed(read)
            assert np.array_equal(a, golden[f"a_{pi}"][off[r]:off[r + 1]]), (pi, r)
            assert rep == int(golden[f"rep_{pi}"][r])


@needs_ref
def test_sketch_model_vs_reference(cases):
    """The per-position form of mm_sketch's loop (what k_sketch evaluates) equals the sequential reference."""
    import seed_model
    refs, reads = cases
    seqs = list(reads) + [refs[0][59900:60200], refs[0][119900:120200], refs[1][:3000], b"ACGTN" * 50, b"acgu" * 40]
    for w, k in ((10, 15), (5, 15), (19, 19), (1, 15), (11, 21), (32, 27)):
        for s in seqs:
            if len(s) > 4000:
                s = s[:4000]
            assert np.array_equal(rs.sketch(s, w, k), seed_model.sketch_model(s, w, k)), (w, k, len(s))


@needs_ref
def test_sort_replay_model_vs_reference():
    """Replaying the American-flag passes on (digit, index) words + ranking buckets of <= 64 = radix_sort_128x, tie order included."""
    import seed_model
    rng = np.random.default_rng(5)
    for n in (2, 64, 65, 300, 2000):
        for spread in (2, 40, 1 << 18, 1 << 44):
            x = rng.integers(0, spread, n).astype(np.uint64) * np.uint64(0x0101010101) + (rng.integers(0, 2, n).astype(np.uint64) << np.uint64(63))
            xy = np.stack([x, np.arange(n, dtype=np.uint64)], axis=1)
            ref = rs.radix_sort_128x(xy)
            assert np.array_equal(ref[:, 1], np.array(seed_model.flag_sort_model([int(v) for v in x]), dtype=np.uint64)), (n, spread)


def test_pass_forms_agree():
    """one flag pass: the step-by-step replay on (digit, index) words, the replay on the original digits alone (destinations), its closed
    form for two buckets, and the walk with run jumps all give the same permutation"""
    import seed_model
    rng = np.random.default_rng(12)
    for n in (2, 7, 100, 1500):
        for nb in (2, 3, 9, 256):
            for skew in (False, True):
                pool = rng.choice(256, nb, replace=False)
                p = None
                if skew and nb > 1:
                    p = np.full(nb, 0.1 / (nb - 1)); p[0] = 0.9
                dig = [int(v) for v in rng.choice(pool, n, p=p)]
                a, _ = seed_model.pass_dest_walk(dig)
                assert seed_model.pass_dest_runs(dig) == a, (n, nb, skew)
                assert seed_model.pass_dest_foreign(dig) == a, (n, nb, skew)
                if len(set(dig)) == 2:
                    assert seed_model.pass_dest_two(dig) == a, (n, nb, skew)


@needs_ref
def test_seed_model_vs_reference(cases):
    """Closed forms of mm_seed_mz_flt / mm_seed_select / rep_len + the sort replay = mm_map_seed."""
    import seed_model
    refs, reads = cases
    ix = rs.RefIndex(refs, w=10, k=15)
    for mid, dist, frac in ((3, 500, 0.01), (2, 50, 0.002), (5, 100, 0.0)):
        ix.field("mid_occ", mid); ix.field("occ_dist", dist); ix.field("q_occ_frac", frac)
        for r in (0, 2, 4, 5, 8, 9, 13, 14, 15):
            a0, rep0, mp0 = ix.seed(reads[r])
            a1, rep1, mp1 = seed_model.seed_model(ix, reads[r], 10, 15, mid, 4095, dist, np.float32(frac))
            assert np.array_equal(a0, a1) and rep0 == rep1 and np.array_equal(mp0, mp1), (mid, dist, frac, r)


def test_seed_abi_symbols(pkg):
    """Every function include/mm2gb_seed.h declares is exported by the library (no compute calls here)."""
    import ctypes
    text = open(os.path.join(ROOT, "include", "mm2gb_seed.h")).read()
    names = set(re.findall(r"\b(mm2gb_[a-z0-9_]+)\s*\(", text))
    assert {"mm2gb_index_build", "mm2gb_seed_host", "mm2gb_seed_chain", "mm2gb_sketch_host"} <= names
    L = ctypes.CDLL(pkg.LIB_PATH)
    for n in sorted(names):
        assert hasattr(L, n), n
    for n in ("mm2gb_chain_device_fetch", "mm2gb_chain_device_results"):
        assert hasattr(L, n), n


def test_seed_params_layout(pkg):
    from mm2gb_b200 import seed
    import ctypes
    assert ctypes.sizeof(seed.SeedParams) == 32
    assert seed.SeedParams.flag.offset == 16 and seed.SeedParams.max_qlen.offset == 28
