"""GPU tests (-m gpu) of the device chain-extraction stage (k_bt_sort / k_bt_walk and their _mid / _big variants): mg_chain_backtrack + compact_a (lchain.c:27-111)
including the tie order of the reference's unstable radix sort.  Checked against the oracle on real DP output, and against
the host implementation (itself pinned to the reference by test_oracle / test_gpu_parity) on synthetic score / predecessor
forests that stress the sort emulation (many equal scores, deep radix recursion, many chains)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _host_chains(pkg, misc, a, off, f, p):
    out = []
    for r in range(len(off) - 1):
        s, e = int(off[r]), int(off[r + 1])
        out.append(pkg.backtrack(misc, a[s:e], f[s:e], p[s:e]))
    return out


def _compare(pkg, misc, a, off, f, p, u, n_u, b, n_b):
    for r, (uh, bh) in enumerate(_host_chains(pkg, misc, a, off, f, p)):
        s = int(off[r])
        assert n_u[r] == len(uh) and n_b[r] == len(bh), f"read {r}: counts ({n_u[r]},{n_b[r]}) vs host ({len(uh)},{len(bh)})"
        assert np.array_equal(u[s:s + n_u[r]], uh), f"read {r}: chains differ"
        assert np.array_equal(b[s:s + n_b[r]], bh), f"read {r}: compacted anchors differ"


def _forest(rng, n, kind):
    """random (f, p): p[i] < i or -1; f > 0.  `kind` picks the score distribution."""
    i = np.arange(n)
    back = rng.integers(1, 6, n)
    far = rng.random(n) < 0.02                                  # links of 255 anchors or more (the escape of the one-byte links)
    back = np.where(far, rng.integers(255, 3000, n), back)
    p = np.where((rng.random(n) < 0.85) & (i - back >= 0), i - back, -1).astype(np.int32)
    if kind == "few":          # a handful of distinct scores: huge tie groups, one radix level
        f = rng.choice([40, 41, 55, 70, 300], n)
    elif kind == "narrow":     # everything inside one 256-bucket
        f = rng.integers(40, 200, n)
    elif kind == "wide":       # three radix levels
        f = rng.integers(40, 400000, n)
    elif kind == "cluster":    # large buckets at the top level, ties below
        f = rng.choice([1000, 70000, 140000], n) + rng.integers(0, 3, n) * 256 + rng.integers(0, 4, n)
    elif kind == "huge":       # scores beyond 2^19: do not pack into the 32-bit keys, four radix levels in the 64-bit kernels
        f = rng.integers(40, 1 << 30, n)
        f[rng.random(n) < 0.3] = 1 << 29                      # and a big tie group
    elif kind == "chainlike":  # scores growing along the read with side branches, like real DP output
        f = 15 * (i + 1) - rng.integers(0, 40, n)
        side = rng.random(n) < 0.15
        f = np.where(side, rng.integers(15, 200, n), f)
    else:
        raise ValueError(kind)
    return f.astype(np.int32), p


@pytest.mark.parametrize("kind", ["few", "narrow", "wide", "cluster", "chainlike", "huge"])
@pytest.mark.parametrize("min_cnt,min_score", [(3, 40), (1, 1), (2, 60)])
def test_forests_vs_host(pkg, synth, kind, min_cnt, min_score, monkeypatch):
    if min_cnt != 3:    # two of the three parameter sets send every read above 8192 anchors to the mid kernels (default: above 32768)
        monkeypatch.setenv("MM2GB_BT_MID_MIN", "8192")
    rng = np.random.default_rng(hash((kind, min_cnt)) % (1 << 31))
    # every size class: shared-memory kernels (<= 8192), the nine mid classes (caps 10048 .. 196608 anchors, both sides of several
    # class boundaries), global-memory kernels
    sizes = [0, 1, 2, 3, 31, 32, 33, 64, 65, 66, 200, 1000, 1024, 1025, 2048, 3000, 4096, 5000, 8192, 8193, 9000, 10048, 10049, 13952, 16384,
             20000, 29504, 33000, 40000, 48640, 48641, 60000, 65536, 70001, 98304, 140000, 196608, 196609]
    reads, fs, ps = [], [], []
    for n in sizes:
        a = synth.ont_like_anchors(rng, max(n, 1), noise_frac=0.0)[:n]
        if kind == "few" and n:
            a[:, 0] = (a[:, 0] & ~np.uint64(0xffffffff)) | (a[:, 0] & np.uint64(0xff))   # many chains start at the same x
            a = a[np.argsort(a[:, 0], kind="stable")]
        f, p = _forest(rng, n, kind)
        reads.append(a); fs.append(f); ps.append(p)
    off = np.zeros(len(sizes) + 1, np.int64)
    off[1:] = np.cumsum(sizes)
    a, f, p = np.concatenate(reads), np.concatenate(fs), np.concatenate(ps)
    misc = pkg.map_ont_misc(min_cnt=min_cnt, min_score=min_score)
    with pkg.ChainContext(misc, max_anchors=1 << 21, max_reads=64, n_slots=1) as c:
        u, n_u, b, n_b, nd = c.backtrack_device(a, off, f, p)
    _compare(pkg, misc, a, off, f, p, u, n_u, b, n_b)
    # reads above 8192 anchors (k_bt_*_mid up to 196608, k_bt_*_big beyond), scores >= 2^19 and reads with more chains than the
    # shared-memory key buffer holds (device overflow list -> k_bt_*_big) all stay on the device: nothing goes to the host code
    assert nd == 0


def test_drop_and_negative_links(pkg, synth):
    """score drops along a chain larger than max_drop (lchain.c:21 break) and links with negative gain"""
    rng = np.random.default_rng(5)
    n = 3000
    a = synth.ont_like_anchors(rng, n, noise_frac=0.0)[:n]
    f = (15 * (np.arange(n) + 1)).astype(np.int32)
    f[1000:1040] += 900          # a bump: walking back from the end, the score first drops by > bw = 500
    f[2000:2005] -= 700
    p = (np.arange(n) - 1).astype(np.int32)
    p[rng.random(n) < 0.02] = -1
    off = np.array([0, n], np.int64)
    misc = pkg.map_ont_misc()
    with pkg.ChainContext(misc, max_anchors=1 << 14, max_reads=4, n_slots=1) as c:
        u, n_u, b, n_b, nd = c.backtrack_device(a, off, f, p)
    assert nd == 0
    _compare(pkg, misc, a, off, f, p, u, n_u, b, n_b)


@pytest.mark.parametrize("seed,n_reads,lo,hi", [(11, 64, 1, 60), (12, 40, 30, 900), (13, 16, 1500, 7000)])
def test_chain_end_to_end_vs_oracle(pkg, po, synth, ctx, seed, n_reads, lo, hi):
    """DP + chain extraction both on the device (mm2gb_chain_host, n_threads = 0) against the oracle's whole mg_lchain_dp"""
    a, off = synth.ont_like_batch(seed, n_reads, lo, hi, repeat_copies=2, repeat_len=40)
    res = ctx.chain(a, off)
    res_h = ctx.chain(a, off, n_threads=2)
    assert ctx.backtrack_device(a, off, res["f"], res["p"])[4] == 0   # real DP output of reads below 8192 anchors: nothing declined
    prm = po.map_ont_params()
    for r in range(n_reads):
        s, e = int(off[r]), int(off[r + 1])
        fo, pq, _ = po.oracle_dp(prm, a[s:e])
        uo, bo = po.oracle_backtrack(prm, a[s:e], fo, pq)
        assert np.array_equal(res["f"][s:e], fo)
        for rr in (res, res_h):
            assert rr["n_u"][r] == len(uo) and rr["n_b"][r] == len(bo), r
            assert np.array_equal(rr["u"][s:s + len(uo)], uo) and np.array_equal(rr["b"][s:s + len(bo)], bo), r


def test_more_overflowing_reads_than_the_device_list_holds(pkg, synth):
    """300 short reads whose scores do not pack into 32 bits: every one goes through the device-side overflow list to the
    global-memory kernels (the list holds one entry per read of the batch; there is no host path)"""
    rng = np.random.default_rng(77)
    n_reads, n = 300, 100
    a = np.concatenate([synth.ont_like_anchors(rng, n, noise_frac=0.0)[:n] for _ in range(n_reads)])
    off = (np.arange(n_reads + 1) * n).astype(np.int64)
    f, p = _forest(rng, n_reads * n, "huge")
    for r in range(n_reads):   # predecessors must stay inside the read
        q = p[r * n:(r + 1) * n]
        q[:] = np.where(q >= r * n, q - r * n, -1)
    misc = pkg.map_ont_misc()
    with pkg.ChainContext(misc, max_anchors=1 << 16, max_reads=512, n_slots=1) as c:
        u, n_u, b, n_b, nd = c.backtrack_device(a, off, f, p)
    assert nd == 0
    _compare(pkg, misc, a, off, f, p, u, n_u, b, n_b)


@pytest.mark.parametrize("seed,n_reads,lo,hi", [(31, 3, 9000, 30000), (32, 1, 70000, 70001)])
def test_long_reads_end_to_end_vs_oracle(pkg, po, synth, seed, n_reads, lo, hi):
    """reads above 8192 anchors: DP + chain extraction on the device (global-memory kernels), packed output"""
    a, off = synth.ont_like_batch(seed, n_reads, lo, hi, repeat_copies=2, repeat_len=40)
    with pkg.ChainContext(pkg.map_ont_misc(), max_anchors=1 << 18, max_reads=16, n_slots=2) as c:
        res = c.chain(a, off)
        resp = c.chain(a, off, packed=True)
        assert c.backtrack_device(a, off, res["f"], res["p"])[4] == 0
    prm = po.map_ont_params()
    for r in range(n_reads):
        s, e = int(off[r]), int(off[r + 1])
        fo, pq, _ = po.oracle_dp(prm, a[s:e])
        uo, bo = po.oracle_backtrack(prm, a[s:e], fo, pq)
        assert np.array_equal(res["f"][s:e], fo)
        assert res["n_u"][r] == len(uo) and res["n_b"][r] == len(bo), r
        assert np.array_equal(res["u"][s:s + len(uo)], uo) and np.array_equal(res["b"][s:s + len(bo)], bo), r
        q = int(resp["v_pos"][r])
        assert resp["n_u"][r] == len(uo) and resp["n_b"][r] == len(bo), r
        assert np.array_equal(resp["u"][s:s + len(uo)], uo), r
        assert np.array_equal(pkg.gather_anchors(a[s:e], resp["v"][q:q + len(bo)]), bo), r   # index wire format: a'[k] = a[v[k]]


def test_chain_without_fp_and_pinned_output(pkg, po, synth):
    import torch
    a, off = synth.ont_like_batch(21, 50, 100, 3000)
    n = int(off[-1])
    out = {"b": torch.empty((n, 2), dtype=torch.int64).pin_memory()}
    h_a = torch.from_numpy(a.view(np.int64)).pin_memory()
    import os
    os.environ["MM2GB_CHUNK"] = "30000"   # several chunks through the slots
    try:
        with pkg.ChainContext(pkg.map_ont_misc(), max_anchors=1 << 18, max_reads=256, n_slots=3) as c:
            res = c.chain(h_a, off, out=out, want_fp=False)
            ref = c.chain(a, off, n_threads=4)
            outp = {"v": torch.zeros(n, dtype=torch.int32).pin_memory()}
            resp = c.chain(h_a, off, out=outp, packed=True)       # device writes the packed indices into the pinned buffer
            assert resp["h2d_anchor_bytes"] == 16 * n             # pinned source: raw DMA, no host pass
            respn = c.chain(a, off, packed=True)                 # pageable buffers: packed 8-byte upload, indices via the slot's pinned memory
            assert 8 * n <= respn["h2d_anchor_bytes"] < 9 * n
    finally:
        del os.environ["MM2GB_CHUNK"]
    assert res["f"] is None
    b = res["b"].numpy().view(np.uint64)
    assert np.array_equal(res["n_u"], ref["n_u"]) and np.array_equal(res["n_b"], ref["n_b"])
    for r in range(len(off) - 1):
        s = int(off[r])
        assert np.array_equal(res["u"][s:s + res["n_u"][r]], ref["u"][s:s + ref["n_u"][r]])
        assert np.array_equal(b[s:s + res["n_b"][r]], ref["b"][s:s + ref["n_b"][r]])
    assert int(resp["n_b"].sum()) == int(ref["n_b"].sum())
    for rr, vv in ((resp, resp["v"].numpy()), (respn, respn["v"])):
        assert np.array_equal(rr["n_u"], ref["n_u"]) and np.array_equal(rr["n_b"], ref["n_b"])
        spans = []
        for r in range(len(off) - 1):
            s, q = int(off[r]), int(rr["v_pos"][r])
            assert np.array_equal(rr["u"][s:s + rr["n_u"][r]], ref["u"][s:s + ref["n_u"][r]])
            assert np.array_equal(pkg.gather_anchors(a[s:int(off[r + 1])], vv[q:q + rr["n_b"][r]]), ref["b"][s:s + ref["n_b"][r]])
            if rr["n_b"][r]:
                spans.append((q, q + int(rr["n_b"][r])))
        spans.sort()
        assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1)), "packed regions overlap"


def test_golden_chains_on_device(pkg, ctx, golden_dir):
    """chains of the reference's own mg_lchain_dp on the fixture reads"""
    import os
    for name in ("fixtures.npz", "synth_reads.npz"):
        g = np.load(os.path.join(golden_dir, name), allow_pickle=False)
        a, off = g["a"], g["off"]
        res = ctx.chain(a, off)
        for r in range(len(off) - 1):
            s = int(off[r])
            assert np.array_equal(res["u"][s:s + res["n_u"][r]], g["u"][int(g["u_off"][r]):int(g["u_off"][r + 1])]), (name, r)
            assert np.array_equal(res["b"][s:s + res["n_b"][r]], g["b"][int(g["b_off"][r]):int(g["b_off"][r + 1])]), (name, r)
