"""GPU parity suite (-m gpu): the CUDA chaining path, called through the C ABI (include/mm2gb_chain.h), against the
oracle on the same inputs.  Bar: bit-exact f[], p[] (integer work) and identical chains after the host stage."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check_batch(pkg, po, ctx, a, off, prm, misc=None, chains=True):
    f, p, st = ctx.chain_dp(a, off)
    pairs = 0
    for r in range(len(off) - 1):
        s, e = int(off[r]), int(off[r + 1])
        fo, pq, npairs = po.oracle_dp(prm, a[s:e])
        pairs += npairs
        bad = np.flatnonzero((f[s:e] != fo) | (p[s:e].astype(np.int64) != pq))
        assert bad.size == 0, f"read {r} (n={e - s}): first mismatch at {bad[0]}: gpu f/p=({f[s + bad[0]]},{p[s + bad[0]]}) oracle=({fo[bad[0]]},{pq[bad[0]]}); {bad.size} bad"
        if chains:
            u, b = pkg.backtrack(misc if misc is not None else ctx.misc, a[s:e], f[s:e], p[s:e])
            uo, bo = po.oracle_backtrack(prm, a[s:e], fo, pq)
            assert np.array_equal(u, uo) and np.array_equal(b, bo), f"read {r}: chains differ"
    assert st.n_pairs == pairs, (st.n_pairs, pairs)
    assert st.n_anchors == int(off[-1])
    return st


@pytest.mark.parametrize("name", ["fixtures.npz", "synth_reads.npz"])
def test_golden_reads(pkg, po, ctx, golden_dir, name):
    """anchors seeded by the reference driver; expected f/p/u/b produced by the reference's own mg_lchain_dp"""
    g = np.load(os.path.join(golden_dir, name), allow_pickle=False)
    a, off = g["a"], g["off"]
    f, p, st = ctx.chain_dp(a, off)
    assert np.array_equal(f, g["f"]) and np.array_equal(p, g["p"])
    for r in range(len(off) - 1):
        s, e = int(off[r]), int(off[r + 1])
        u, b = pkg.backtrack(ctx.misc, a[s:e], f[s:e], p[s:e])
        assert np.array_equal(u, g["u"][int(g["u_off"][r]):int(g["u_off"][r + 1])])
        assert np.array_equal(b, g["b"][int(g["b_off"][r]):int(g["b_off"][r + 1])])
    if name == "fixtures.npz":
        assert st.n_pairs == 30829 + 22140 + 243060  # MT, inv:0, inv:1 (t2 has no anchors)


def test_adversarial_suite(pkg, po, synth):
    """every edge case of lchain.c, each with its own parameter set (general float path, narrow band, small max_iter,
    clipped windows + max_ii fallback, ties, equal x, empty / single reads)"""
    for name, (a, over) in synth.adversarial_suite().items():
        misc = pkg.map_ont_misc(**over)
        prm = po.map_ont_params(**over)
        with pkg.ChainContext(misc, max_anchors=1 << 16, max_reads=16, n_slots=1) as c:
            off = np.array([0, len(a)], np.int64)
            st = _check_batch(pkg, po, c, a, off, prm, misc)
            if name in ("dense", "max_ii", "small_iter"):
                assert st.n_units_exact > 0, name      # the exact max_ii path really ran
            if name == "skip_pen":
                assert st.general_path == 1
    a, over = synth.adversarial_suite()["max_ii"]
    with pkg.ChainContext(pkg.map_ont_misc(**over), max_anchors=1 << 14, max_reads=4, n_slots=1) as c:
        f, p, _ = c.chain_dp(a, np.array([0, len(a)], np.int64))
    assert int(f[-1]) == 30 and int(p[-1]) == 0


def test_adversarial_as_one_batch(pkg, po, synth, ctx):
    """the map-ont members of the suite concatenated into one ragged batch, with empty reads in between"""
    reads = [a for _, (a, over) in synth.adversarial_suite().items() if not over]
    reads = [reads[0][:0]] + reads + [reads[0][:0], reads[1], reads[0][:0]]
    off = np.zeros(len(reads) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    _check_batch(pkg, po, ctx, np.concatenate(reads), off, po.map_ont_params())


@pytest.mark.parametrize("seed,n_reads,lo,hi", [(1, 64, 1, 40), (2, 40, 30, 600), (3, 24, 500, 5000), (4, 3, 9000, 14000)])
def test_random_batches(pkg, po, synth, ctx, seed, n_reads, lo, hi):
    a, off = synth.ont_like_batch(seed, n_reads, lo, hi)
    _check_batch(pkg, po, ctx, a, off, po.map_ont_params())


def test_repeat_dense_reads(pkg, po, synth, ctx):
    """planted repeats: windows of thousands of predecessors (beyond the shared-memory ring -> global path)"""
    rng = np.random.default_rng(21)
    reads = [synth.ont_like_anchors(rng, 2500, mean_gap=4.0, repeat_copies=3, repeat_len=400),
             synth.ont_like_anchors(rng, 4000, mean_gap=2.0, noise_frac=0.3)]
    off = np.zeros(len(reads) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    _check_batch(pkg, po, ctx, np.concatenate(reads), off, po.map_ont_params())


@pytest.mark.parametrize("long_min", [0, 2048, 8192])
def test_long_units(pkg, po, synth, long_min):
    """units of thousands of anchors: the CTA-cooperative pipelined kernel (k_score_long) against the one-warp kernel
    (long_min = 0) and the oracle.  Covers chain scores beyond 2^18 (the packed keys of the one-warp kernel overflow),
    windows longer than the shared ring (global-memory part of the window), cuts inside a long unit and a strand switch."""
    rng = np.random.default_rng(77)
    two_strands = np.concatenate([synth.ont_like_anchors(rng, 2600, rev=0, rpos0=5000), synth.ont_like_anchors(rng, 2600, rev=1, rpos0=5000)])
    two_strands = two_strands[np.argsort(two_strands[:, 0], kind="stable")]
    reads = [synth.ont_like_anchors(rng, 50000, mean_gap=9.0, noise_frac=0.2),      # ~60k anchors, one unit, f beyond 2^18
             synth.ont_like_anchors(rng, 9000, mean_gap=1.2, noise_frac=0.05),      # ~4100 predecessors per anchor
             synth.ont_like_anchors(rng, 2100), two_strands, synth.ont_like_anchors(rng, 300),
             synth.chaining_only_array(5, 20000, 4500)]                              # independent segments of 4500 anchors
    off = np.zeros(len(reads) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    os.environ["MM2GB_LONG_MIN"] = str(long_min)
    try:
        with pkg.ChainContext(pkg.map_ont_misc(), max_anchors=1 << 18, max_reads=16, n_slots=1) as c:
            st = _check_batch(pkg, po, c, np.concatenate(reads), off, po.map_ont_params(), chains=False)
    finally:
        del os.environ["MM2GB_LONG_MIN"]
    assert (st.n_long > 0) == (long_min > 0), st.as_dict() if hasattr(st, "as_dict") else st.n_long
    if long_min == 2048:
        assert st.n_long >= 8


@pytest.mark.parametrize("over", [dict(chn_pen_skip=0.03), dict(is_cdna=1), dict(n_seg=2), dict(bw=2000, max_dist_x=10000, max_dist_y=10000),
                                  dict(bw=100000, max_dist_x=100000, max_dist_y=100000), dict(max_dist_x=100, max_dist_y=100, bw=500)])
def test_parameter_variants(pkg, po, synth, over):
    a, off = synth.ont_like_batch(31, 10, 100, 1500)
    rng = np.random.default_rng(1)
    if over.get("n_seg", 1) > 1 or over.get("is_cdna"):
        sid = rng.integers(0, 2, len(a)).astype(np.uint64)   # mixed segment ids (paired-end style)
        a = a.copy()
        a[:, 1] |= sid << np.uint64(48)
    with pkg.ChainContext(pkg.map_ont_misc(**over), max_anchors=1 << 18, max_reads=64, n_slots=2) as c:
        _check_batch(pkg, po, c, a, off, po.map_ont_params(**over), c.misc)


def test_mixed_segment_ids_force_general_path(pkg, po, synth, ctx):
    """map-ont parameters but a read that mixes segment ids: the table path must step aside"""
    a, off = synth.ont_like_batch(33, 4, 100, 800)
    a = a.copy()
    a[5:60, 1] |= np.uint64(1) << np.uint64(48)
    st = _check_batch(pkg, po, ctx, a, off, po.map_ont_params())
    assert st.general_path == 1


def test_chunked_and_async_paths_agree(pkg, po, synth):
    a, off = synth.ont_like_batch(41, 60, 100, 3000)
    os.environ["MM2GB_CHUNK"] = "20000"   # force many chunks through the slots
    try:
        with pkg.ChainContext(pkg.map_ont_misc(), max_anchors=1 << 18, max_reads=256, n_slots=3) as c:
            f1, p1, st1 = c.chain_dp(a, off)
    finally:
        del os.environ["MM2GB_CHUNK"]
    with pkg.ChainContext(pkg.map_ont_misc(), max_anchors=1 << 18, max_reads=256, n_slots=2) as c:
        f2, p2, st2 = c.chain_dp(a, off)
        half = 30
        c.submit(0, a[:off[half]], off[:half + 1])
        c.submit(1, a[off[half]:], off[half:] - off[half])
        fa, pa, sa = c.wait(0)
        fa, pa = fa.copy(), pa.copy()
        fb, pb, sb = c.wait(1)
        fb, pb = fb.copy(), pb.copy()       # views into the slot's pinned buffers die with the context
        with pytest.raises(pkg.Mm2gbError):
            c.wait(0)                       # idle slot
    assert np.array_equal(f1, f2) and np.array_equal(p1, p2) and st1.n_pairs == st2.n_pairs
    assert np.array_equal(np.concatenate([fa, fb]), f2) and np.array_equal(np.concatenate([pa, pb]), p2)
    assert sa.n_pairs + sb.n_pairs == st2.n_pairs
    _, fo, po_ = po.lchain_batch(po.map_ont_params(), a, off, want_fp=True)
    assert np.array_equal(f2, fo) and np.array_equal(p2.astype(np.int64), po_)


def test_device_resident_path(pkg, po, synth):
    import torch
    a, off = synth.ont_like_batch(51, 30, 100, 2500)
    n = int(off[-1])
    d_a = torch.from_numpy(a.view(np.int64)).cuda()
    d_off = torch.from_numpy(off).cuda()
    d_f = torch.empty(n, dtype=torch.int32, device="cuda")
    d_p = torch.empty(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    with pkg.ChainContext(pkg.map_ont_misc(), max_anchors=1 << 18, max_reads=64, n_slots=1) as c:
        c.profile(True)
        for _ in range(2):  # idempotent
            c.chain_dp_device(d_a, d_off, len(off) - 1, n, d_f, d_p)
            c.sync()
        st = c.device_stats()
        prof = c.profile_read()
    _, fo, pq = po.lchain_batch(po.map_ont_params(), a, off, want_fp=True)
    assert np.array_equal(d_f.cpu().numpy(), fo) and np.array_equal(d_p.cpu().numpy().astype(np.int64), pq)
    assert st.n_anchors == n and st.n_pairs > 0
    assert prof["score"][1] == 2 and prof["score"][0] > 0 and prof["range"][1] == 2


def test_capacity_and_argument_errors(pkg, synth):
    a, off = synth.ont_like_batch(61, 4, 100, 200)
    with pkg.ChainContext(pkg.map_ont_misc(), max_anchors=64, max_reads=8, n_slots=1) as c:
        with pytest.raises(pkg.Mm2gbError):
            c.chain_dp(a, off)              # a read larger than the context capacity
    with pytest.raises(pkg.Mm2gbError):
        pkg.ChainContext(pkg.map_ont_misc(), n_slots=9)
    with pytest.raises(pkg.Mm2gbError):
        pkg.ChainContext(pkg.map_ont_misc(), device=99)


def test_large_batch_properties(pkg, po, synth):
    """~1.2M anchors: size-independent invariants + an oracle check on a sample of reads + pair count from numpy"""
    a, off = synth.ont_like_batch(71, 500, 800, 4000)
    n = int(off[-1])
    with pkg.ChainContext(pkg.map_ont_misc(), max_anchors=1 << 21, max_reads=1024, n_slots=2) as c:
        f, p, st = c.chain_dp(a, off)
        f2, p2, _ = c.chain_dp(a, off)
    assert np.array_equal(f, f2) and np.array_equal(p, p2)            # deterministic / idempotent
    qs = ((a[:, 1] >> np.uint64(32)) & np.uint64(0xff)).astype(np.int32)
    assert np.all(f >= qs) and np.all((p == -1) == (f == qs) | (p == -1))
    idx = np.arange(n) - np.repeat(off[:-1], np.diff(off))
    assert np.all(p < idx) and np.all(p >= -1)                        # predecessors precede, inside the read
    # window sizes by searchsorted == device pair counter
    pairs = 0
    prm = po.map_ont_params()
    for r in range(len(off) - 1):
        x = a[off[r]:off[r + 1], 0]
        lower = np.maximum(x - np.uint64(prm.max_dist_x), x & np.uint64(0xffffffff00000000)) if len(x) else x
        lower = np.where((x & np.uint64(0xffffffff)) > np.uint64(prm.max_dist_x), x - np.uint64(prm.max_dist_x), x & np.uint64(0xffffffff00000000))
        st_i = np.searchsorted(x, lower, side="left")
        st_i = np.maximum(st_i, np.arange(len(x)) - prm.max_iter)
        pairs += int((np.arange(len(x)) - st_i).sum())
    assert st.n_pairs == pairs
    for r in np.random.default_rng(0).choice(len(off) - 1, 25, replace=False):
        s, e = int(off[r]), int(off[r + 1])
        fo, pq, _ = po.oracle_dp(prm, a[s:e])
        assert np.array_equal(f[s:e], fo) and np.array_equal(p[s:e].astype(np.int64), pq), r


def test_smoke_entry():
    import __graft_entry__ as entry
    entry.smoke()


def test_range_kernel_tma_variant(pkg, po, synth, monkeypatch):
    """k_range_tma (x history staged by cp.async.bulk + mbarrier, MM2GB_RANGE_TMA=1) against the plain k_range and the oracle:
    ragged batches with empty reads, reads that start on and next to a 256-anchor block boundary, windows longer than the staged
    history (dense reads: global-memory fallback), clipped windows, the golden reads"""
    rng = np.random.default_rng(77)
    reads = [a for _, (a, over) in synth.adversarial_suite().items() if not over]
    reads += [synth.ont_like_anchors(rng, n) for n in (255, 256, 257, 511, 513, 1, 2, 3000, 9000)]
    reads += [synth.ont_like_anchors(rng, 4000, mean_gap=3.0)]           # ~1700 anchors per window: beyond the 1024 staged
    reads = [reads[0][:0]] + reads + [reads[0][:0]]
    off = np.zeros(len(reads) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    a = np.concatenate(reads)
    prm = po.map_ont_params()
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("MM2GB_RANGE_TMA", mode)
        with pkg.ChainContext(pkg.map_ont_misc(), max_anchors=1 << 18, max_reads=256, n_slots=1) as c:
            st = _check_batch(pkg, po, c, a, off, prm, chains=False)
            f, p, _ = c.chain_dp(a, off)
            out[mode] = (f.copy(), p.copy(), st.n_pairs, st.n_units, st.n_units_exact)
    assert np.array_equal(out["0"][0], out["1"][0]) and np.array_equal(out["0"][1], out["1"][1])
    assert out["0"][2:] == out["1"][2:]
