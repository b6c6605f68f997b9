"""Device seeding against the reference's own mm_sketch / mm_idx_get / mm_map_seed (oracle/_ref/libref_seed.so)
and against the committed golden vectors: minimizers, index lists, anchors in the reference's exact order (tie order of its
unstable radix sort included), rep_len, mini_pos; and the fused seed + chain step against mm_map_seed + mg_lchain_dp."""
import os
import sys

import numpy as np
import pytest

import seed_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyrefseed as rs  # noqa: E402

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not rs.available(), reason="oracle/_ref/libref_seed.so not built")

PARAM_SETS = [(None, 500, 0.01), (3, 500, 0.01), (3, 0, 0.01), (5, 100, 0.0), (2, 50, 0.002)]


@pytest.fixture(scope="module")
def sd(pkg):
    from mm2gb_b200 import seed
    return seed


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "seed_golden.npz"))


@pytest.fixture(scope="module")
def cases():
    refs = seed_cases.make_reference()
    return refs, seed_cases.make_reads(refs)


@pytest.fixture(scope="module")
def index(sd, cases):
    ix = sd.Index(cases[0], w=10, k=15)
    yield ix
    ix.close()


@pytest.fixture(scope="module")
def seeder(sd, index):
    s = sd.Seeder(index, max_bases=1 << 20, max_reads=4096, max_anchors=1 << 20)
    yield s
    s.close()


def test_sketch_golden(sd, seeder, cases, golden):
    buf, off = sd.pack_seqs(cases[1])
    mv, mvo = seeder.sketch(buf, off)
    assert np.array_equal(mvo, golden["mv_off"])
    assert np.array_equal(mv, golden["mv"])


@needs_ref
@pytest.mark.parametrize("w,k", [(10, 15), (5, 15), (19, 19), (1, 15), (11, 21), (32, 27)])
def test_sketch_vs_reference(sd, cases, w, k):
    refs, reads = cases
    rng = np.random.default_rng(w * 100 + k)
    seqs = list(reads) + [refs[0][59900:61200], refs[0][119000:121500], refs[1][:5000], b"ACGTN" * 500, b"acgu" * 400, refs[0][:70000]]
    seqs += [seed_cases.ACGT[rng.integers(0, 4, n)].tobytes() for n in (1, 2, k - 1, k, k + 1, w + k - 2, w + k - 1, w + k, 1023, 1024, 1025, 2047, 2048, 2049)]
    with sd.Index([refs[0][:5000]], w=w, k=k) as ix, sd.Seeder(ix, max_bases=1 << 19, max_reads=256, max_anchors=1024) as s:
        buf, off = sd.pack_seqs(seqs)
        mv, mvo = s.sketch(buf, off, rid_is_seq=True)
    for i, q in enumerate(seqs):
        exp = rs.sketch(q, w, k, rid=i)
        got = mv[mvo[i]:mvo[i + 1]]
        assert got.shape == exp.shape and np.array_equal(got, exp), (w, k, i, len(q), got.shape, exp.shape)


@needs_ref
def test_index_vs_reference(sd, index, cases, golden):
    refs, _ = cases
    ix = rs.RefIndex(refs, w=10, k=15)
    assert index.cal_max_occ(2e-4) == ix.mid_occ_of(2e-4) == int(golden["cal_max_occ_2e-4"])
    assert index.cal_max_occ(1e-2) == ix.mid_occ_of(1e-2) == int(golden["cal_max_occ_1e-2"])
    assert index.mid_occ() == int(ix.field("mid_occ"))
    keys = set()
    for rid, ref in enumerate(refs):
        mv = rs.sketch(ref, 10, 15, rid=rid)
        keys.update(int(x) >> 8 for x in mv[:, 0])
    assert index.n_keys == len(keys)
    rng = np.random.default_rng(1)
    sample = list(keys)
    big = sorted(sample, key=lambda m: -len(ix.get(m)))[:50]
    for m in big + [sample[i] for i in rng.integers(0, len(sample), 2000)] + [12345, 0, (1 << 30) - 1]:
        assert np.array_equal(index.get(m), ix.get(m)), m


def test_index_golden(index, golden):
    assert index.mid_occ() == int(golden["mid_occ_default"])
    assert index.cal_max_occ(1e-2) == int(golden["cal_max_occ_1e-2"])


@pytest.mark.parametrize("pi", range(len(PARAM_SETS)))
def test_seed_golden(sd, seeder, cases, golden, pi):
    mid, dist, frac = PARAM_SETS[pi]
    prm = sd.map_ont_seed_params(int(golden["mid_occ_default"]) if mid is None else mid, occ_dist=dist, q_occ_frac=frac)
    buf, off = sd.pack_seqs(cases[1])
    a, a_off, rep, mp, mp_off = seeder.seed(prm, buf, off)
    assert np.array_equal(a_off, golden[f"a_off_{pi}"])
    assert np.array_equal(rep, golden[f"rep_{pi}"])
    assert np.array_equal(mp_off, golden[f"mp_off_{pi}"]) and np.array_equal(mp, golden[f"mp_{pi}"])
    exp = golden[f"a_{pi}"]
    if not np.array_equal(a, exp):
        for r in range(len(a_off) - 1):
            g, e = a[a_off[r]:a_off[r + 1]], exp[a_off[r]:a_off[r + 1]]
            if not np.array_equal(g, e):
                same_set = np.array_equal(g[np.lexsort((g[:, 1], g[:, 0]))], e[np.lexsort((e[:, 1], e[:, 0]))])
                bad = int(np.argmax(np.any(g != e, axis=1)))
                pytest.fail(f"set {pi} read {r}: anchors differ at {bad} of {len(g)} ({'same set, order differs' if same_set else 'different set'})")


def test_seed_golden_k19(sd, cases, golden):
    with sd.Index(cases[0], w=19, k=19) as ix, sd.Seeder(ix, max_bases=1 << 19, max_reads=256, max_anchors=1 << 16) as s:
        buf, off = sd.pack_seqs(cases[1])
        mv, mvo = s.sketch(buf, off)
        assert np.array_equal(mv, golden["mv_k19"]) and np.array_equal(mvo, golden["mv_off_k19"])
        a, a_off, _, _, _ = s.seed(sd.map_ont_seed_params(10), buf, off)
        assert np.array_equal(a_off, golden["a_off_k19"]) and np.array_equal(a, golden["a_k19"])


def test_hpc_golden(sd, cases, golden):
    """homopolymer-compressed minimizers (MM_I_HPC; map-pb: k = 19, w = 10): minimizers with their spans, anchors, rep_len, mini_pos"""
    with sd.Index(cases[0], w=10, k=19, hpc=True) as ix, sd.Seeder(ix, max_bases=1 << 19, max_reads=256, max_anchors=1 << 16) as s:
        buf, off = sd.pack_seqs(cases[1])
        mv, mvo = s.sketch(buf, off)
        assert np.array_equal(mvo, golden["mv_off_hpc"]) and np.array_equal(mv, golden["mv_hpc"])
        a, a_off, rep, mp, _ = s.seed(sd.map_ont_seed_params(10), buf, off)
        assert np.array_equal(a_off, golden["a_off_hpc"]) and np.array_equal(a, golden["a_hpc"])
        assert np.array_equal(rep, golden["rep_hpc"]) and np.array_equal(mp, golden["mp_hpc"])


@needs_ref
@pytest.mark.parametrize("w,k", [(10, 19), (5, 15), (19, 19)])
def test_hpc_sketch_vs_reference(sd, cases, w, k):
    refs, reads = cases
    rng = np.random.default_rng(k)
    def hp(n):      # homopolymer-rich
        out = bytearray()
        while len(out) < n:
            out += bytes([seed_cases.ACGT[rng.integers(0, 4)]]) * int(rng.geometric(0.35))
        return bytes(out[:n])
    seqs = list(reads) + [hp(n) for n in (1, 30, 600, 5000, 40000)] + [b"A" * 400 + hp(300) + b"C" * 300 + hp(2500), refs[0][59000:62000], refs[0][:30000]]
    s2 = bytearray(hp(3000))
    for p in (10, 11, 500, 1500, 1501, 2999):
        s2[p] = ord("N")
    seqs.append(bytes(s2))
    with sd.Index([refs[0][:5000]], w=w, k=k, hpc=True) as ix, sd.Seeder(ix, max_bases=1 << 19, max_reads=256, max_anchors=1024) as s:
        buf, off = sd.pack_seqs(seqs)
        mv, mvo = s.sketch(buf, off, rid_is_seq=True)
    for i, q in enumerate(seqs):
        exp = rs.sketch(q, w, k, rid=i, hpc=True)
        got = mv[mvo[i]:mvo[i + 1]]
        assert got.shape == exp.shape and np.array_equal(got, exp), (w, k, i, len(q), got.shape, exp.shape)


@needs_ref
def test_hpc_seed_batch_vs_reference(sd, cases):
    refs, _ = cases
    reads = make_batch(refs, 120, seed=77)
    buf, off = sd.pack_seqs(reads)
    ix = rs.RefIndex(refs, w=10, k=19, hpc=True, preset="map-pb")
    mid = int(ix.field("mid_occ"))
    n_a, _, dig, _ = ix.seed_batch(buf, off, chain=False, threads=4)
    with sd.Index(refs, w=10, k=19, hpc=True) as dix:
        assert dix.mid_occ() == mid
        with sd.Seeder(dix, max_bases=int(off[-1]) + 1024, max_reads=256, max_anchors=int(n_a.sum()) + 1024) as s:
            a, a_off, _, _, _ = s.seed(sd.map_ont_seed_params(mid), buf, off, want_mini_pos=False)
    assert np.array_equal(np.diff(a_off), n_a)
    for r in range(len(reads)):
        assert rs.word_digest(a[a_off[r]:a_off[r + 1]]) == int(dig[r]), r


def test_sort_words_in_hbm(sd, index, cases, golden, monkeypatch):
    """Reads whose digit bytes do not fit shared memory keep them in HBM (same replay)."""
    monkeypatch.setenv("MM2GB_SEED_SORT_CAP", "64")
    with sd.Seeder(index, max_bases=1 << 20, max_reads=256, max_anchors=1 << 20) as s:
        buf, off = sd.pack_seqs(cases[1])
        a, a_off, _, _, _ = s.seed(sd.map_ont_seed_params(3), buf, off)
    assert np.array_equal(a_off, golden["a_off_1"]) and np.array_equal(a, golden["a_1"])


def make_batch(refs, n_reads, seed, lo=2000, hi=30000):
    rng = np.random.default_rng(seed)
    ref = refs[0]
    reads = []
    for i in range(n_reads):
        ln = int(rng.integers(lo, hi))
        st = int(rng.integers(0, len(ref) - ln))
        r = bytearray(ref[st:st + ln])
        for q in rng.integers(0, ln, ln // 10):
            r[q] = seed_cases.ACGT[rng.integers(0, 4)]
        r = bytes(r)
        reads.append(seed_cases.revcomp(r) if i & 1 else r)
    return reads


@needs_ref
def test_seed_batch_vs_reference(sd, index, cases):
    """300 reads of 2-30 kb in one batch against mm_map_seed read by read (digest of the whole anchor array)."""
    refs, _ = cases
    reads = make_batch(refs, 300, seed=21)
    buf, off = sd.pack_seqs(reads)
    ix = rs.RefIndex(refs, w=10, k=15)
    for mid in (int(ix.field("mid_occ")), 4):
        ix.field("mid_occ", mid)
        n_a, _, dig, _ = ix.seed_batch(buf, off, chain=False, threads=4)
        with sd.Seeder(index, max_bases=int(off[-1]) + 1024, max_reads=512, max_anchors=int(n_a.sum()) + 1024) as s:
            a, a_off, rep, _, _ = s.seed(sd.map_ont_seed_params(mid), buf, off, want_mini_pos=False)
        assert np.array_equal(np.diff(a_off), n_a)
        for r in range(len(reads)):
            assert rs.word_digest(a[a_off[r]:a_off[r + 1]]) == int(dig[r]), (mid, r)


@needs_ref
def test_seed_chain_fused_vs_reference(pkg, sd, index, cases):
    """Sequences in, chains out: n_u, u[] and the compacted anchors equal mm_map_seed + mg_lchain_dp (max-chain-skip = inf)."""
    refs, fixed = cases
    reads = make_batch(refs, 200, seed=33) + list(fixed)
    buf, off = sd.pack_seqs(reads)
    ix = rs.RefIndex(refs, w=10, k=15)
    ix.field("max_chain_skip", 2147483647)
    mid = 5
    ix.field("mid_occ", mid)
    n_a, n_u, dig, _ = ix.seed_batch(buf, off, chain=True, threads=4)
    misc = pkg.Misc.from_buffer_copy(ix.misc())
    cap = int(n_a.sum()) + 1024
    with pkg.ChainContext(misc, device=0, max_anchors=cap, max_reads=1024, n_slots=1, flags=pkg.ChainContext.DEVICE_ONLY) as ctx, \
            sd.Seeder(index, max_bases=int(off[-1]) + 1024, max_reads=1024, max_anchors=cap) as s:
        res = s.seed_chain(ctx, sd.map_ont_seed_params(mid), buf, off)
        res2 = s.seed_chain(ctx, sd.map_ont_seed_params(mid), buf, off)       # buffers are reusable
        mp, mp_off = s.last_mini_pos(len(reads), int(off[-1]))
    for r in (0, 7, 150, len(reads) - 1, len(reads) - 5):
        exp = ix.seed(reads[r])[2]
        assert np.array_equal(mp[mp_off[r]:mp_off[r + 1]], exp), r
    assert np.array_equal(np.diff(res["a_off"]), n_a)
    assert np.array_equal(res["n_u"], n_u)
    for r in range(len(reads)):
        u = res["u"][res["u_pos"][r]:res["u_pos"][r] + res["n_u"][r]]
        b = res["b"][res["b_pos"][r]:res["b_pos"][r] + res["n_b"][r]]
        assert int(res["n_b"][r]) == int((u & np.uint64(0xffffffff)).sum())
        assert rs.chain_digest(u, b) == int(dig[r]), r
    for r in range(len(reads)):      # the packed positions depend on the order the reads finish in; the per-read contents do not
        for key, pos, cnt in (("u", "u_pos", "n_u"), ("b", "b_pos", "n_b")):
            x1 = res[key][res[pos][r]:res[pos][r] + res[cnt][r]]
            x2 = res2[key][res2[pos][r]:res2[pos][r] + res2[cnt][r]]
            assert np.array_equal(x1, x2), (key, r)
    assert res["h2d_bytes"] >= int(off[-1]) and res["d2h_bytes"] >= 16 * int(res["n_chain_anchors"])


def test_refused_settings(pkg, sd, index, seeder, cases):
    buf, off = sd.pack_seqs(cases[1][:2])
    for kw in (dict(flag=0x001), dict(flag=0x400000), dict(sdust_thres=20), dict(flag=0x100000000)):
        with pytest.raises(pkg.Mm2gbError):
            seeder.seed(sd.map_ont_seed_params(10, **kw), buf, off)
    with pytest.raises(pkg.Mm2gbError):
        sd.Index(cases[0], w=10, k=16)
    with pytest.raises(pkg.Mm2gbError):     # capacity: fails loudly, no partial result
        with sd.Seeder(index, max_bases=1 << 20, max_reads=64, max_anchors=100) as s:
            s.seed(sd.map_ont_seed_params(10), *sd.pack_seqs(cases[1][:4]))
