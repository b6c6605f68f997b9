"""CPU suite (-m "not gpu"): the oracle against the reference's golden vectors, the host stage of the product
(backtracking) against the oracle, and the C-ABI surface of the built library."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _reads(g):
    for r in range(len(g["off"]) - 1):
        s, e = int(g["off"][r]), int(g["off"][r + 1])
        us, ue = int(g["u_off"][r]), int(g["u_off"][r + 1])
        bs, be = int(g["b_off"][r]), int(g["b_off"][r + 1])
        yield r, g["a"][s:e], g["f"][s:e], g["p"][s:e], g["u"][us:ue], g["b"][bs:be], g["f25"][s:e], g["p25"][s:e]


@pytest.mark.parametrize("name", ["fixtures.npz", "synth_reads.npz"])
def test_oracle_matches_reference_golden(po, golden_dir, name):
    """oracle (restatement) == outputs of the reference's own mg_lchain_dp stored by oracle/gen_golden.py"""
    g = _load(golden_dir, name)
    prm, prm25 = po.map_ont_params(), po.map_ont_params(max_skip=25)
    n = 0
    for r, a, f, p, u, b, f25, p25 in _reads(g):
        o = po.oracle_lchain(prm, a)
        assert np.array_equal(o.f, f) and np.array_equal(o.p, p.astype(np.int64)), r
        assert np.array_equal(o.u, u) and np.array_equal(o.b, b), r
        o25 = po.oracle_lchain(prm25, a)
        assert np.array_equal(o25.f, f25) and np.array_equal(o25.p, p25.astype(np.int64)), r
        n += len(a)
    assert n > 0


def test_fixture_known_answers(po, golden_dir):
    """The one PAF line the reference documents (README.md:85-96): MT-orang vs MT-human has
    346 anchors, 30829 pairs and a 342-anchor chain of score 3187 (s1:i:3187, cm:i:342)."""
    g = _load(golden_dir, "fixtures.npz")
    assert list(g["names"]) == ["MT:0", "inv:0", "inv:1", "t2:0"]
    a = g["a"][int(g["off"][0]):int(g["off"][1])]
    assert len(a) == 346
    o = po.oracle_lchain(po.map_ont_params(), a)
    assert o.n_pairs == 30829
    assert [(int(x >> 32), int(x & 0xffffffff)) for x in o.u] == [(3187, 342)]
    paf = open(os.path.join(golden_dir, "MT.paf")).read().split("\t")
    assert paf[0] == "MT_orang" and "s1:i:3187" in paf and "cm:i:342" in paf
    assert open(os.path.join(golden_dir, "t2.paf")).read() == ""  # t2/q2: no hit, no output


def test_adversarial_digests(po, synth, golden_dir):
    """Edge cases of lchain.c (ties, equal x, band edge, q_span != 15, clipped windows + max_ii, multi rid/strand ...):
    oracle output digests == digests of the reference's outputs."""
    dig = json.load(open(os.path.join(golden_dir, "adversarial.json")))
    suite = synth.adversarial_suite()
    assert set(suite) == set(dig)
    for name, (a, over) in suite.items():
        o = po.oracle_lchain(po.map_ont_params(**over), a)
        h = hashlib.sha256()
        for arr in (o.f, o.p, o.u, o.b):
            h.update(np.ascontiguousarray(arr).tobytes())
        assert h.hexdigest() == dig[name]["sha256"], name
        assert len(a) == dig[name]["n"] and len(o.u) == dig[name]["n_u"], name
    # the max_ii fallback wins, predecessor 6001 anchors back
    a, over = suite["max_ii"]
    o = po.oracle_lchain(po.map_ont_params(**over), a)
    assert int(o.f[-1]) == 30 and int(o.p[-1]) == 0


def test_oracle_vs_live_reference(po, synth):
    """When oracle/_ref was built (build container only), compare on fresh random inputs too."""
    if not po.ref_available():
        pytest.skip("oracle/_ref not built here")
    a, off = synth.ont_like_batch(seed=123, n_reads=12, lo=20, hi=2500)
    for skip in (po.INT32_MAX, 25):
        prm = po.map_ont_params(max_skip=skip)
        for r in range(len(off) - 1):
            ar = a[off[r]:off[r + 1]]
            assert po.oracle_lchain(prm, ar).same(po.ref_lchain(prm, ar)), (skip, r)
    rng = np.random.default_rng(3)
    for n in (1, 2, 63, 64, 65, 300, 5000):
        z = np.stack([rng.integers(0, 50, n).astype(np.uint64), np.arange(n, dtype=np.uint64)], axis=1)
        assert np.array_equal(po.radix_sort_128x(z), po.radix_sort_128x(z, use_ref=True)), n


def test_radix_sort_is_a_sort(po):
    rng = np.random.default_rng(5)
    for n in (0, 1, 64, 65, 1000, 20000):
        z = np.stack([rng.integers(0, 1 << 40, n).astype(np.uint64), np.arange(n, dtype=np.uint64)], axis=1)
        s = po.radix_sort_128x(z)
        assert np.all(np.diff(s[:, 0].astype(np.int64)) >= 0)
        assert sorted(s[:, 1].tolist()) == list(range(n))


def test_gap_penalty_matches_pair_score(po):
    """orc_gap_penalty (the table the device builds) == the penalty inside comput_sc for every dd in 0..bw"""
    prm = po.map_ont_params()
    for dd in range(0, prm.bw + 1):
        ai = np.array([[10000 + 100 + dd, (15 << 32) | 5100]], np.uint64)
        aj = np.array([[10000, (15 << 32) | 5000]], np.uint64)
        sc = po.lib().orc_pair_score(ai.ctypes.data, aj.ctypes.data, C.byref(prm))
        assert sc == 15 - po.lib().orc_gap_penalty(dd, 100, C.byref(prm)), dd
    assert po.lib().orc_gap_penalty(2, 0, C.byref(prm)) == 1  # SURVEY.md trap T2: CPU gives 1 where mm2-gb's kernel gives 0


# ---- host stage of the product (no GPU needed) ------------------------------------------------------------

def test_host_backtrack_matches_oracle(pkg, po, synth, golden_dir):
    """mm2gb_backtrack (csrc/backtrack.cpp) == mg_chain_backtrack + compact_a on golden f/p and on the adversarial
    suite (tie order of the unstable radix sort included)."""
    misc = pkg.map_ont_misc()
    for name in ("fixtures.npz", "synth_reads.npz"):
        g = _load(golden_dir, name)
        for r, a, f, p, u, b, *_ in _reads(g):
            uu, bb = pkg.backtrack(misc, a, f, p)
            assert np.array_equal(uu, u) and np.array_equal(bb, b), (name, r)
    for name, (a, over) in synth.adversarial_suite().items():
        prm = po.map_ont_params(**over)
        o = po.oracle_lchain(prm, a)
        m = pkg.map_ont_misc(**over)
        uu, bb = pkg.backtrack(m, a, o.f, o.p.astype(np.int32))
        assert np.array_equal(uu, o.u) and np.array_equal(bb, o.b), name


def test_abi_exports_every_declared_symbol(pkg):
    """The C-ABI library loads and exports every function include/mm2gb_chain.h declares (no compute calls)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "mm2gb_chain.h")).read()
    names = set(re.findall(r"\b(mm2gb_[a-z_0-9]+)\s*\(", hdr))
    assert {"mm2gb_ctx_create", "mm2gb_chain_dp_host", "mm2gb_submit", "mm2gb_wait", "mm2gb_chain_dp_device",
            "mm2gb_backtrack"} <= names
    L = pkg.lib()
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/mm2gb_chain.h but not exported"
    assert C.sizeof(pkg.Misc) == 44  # sizeof(Misc), SURVEY.md 8b


def test_no_cpu_fallback_without_device(pkg):
    """Without a CUDA device the product path must fail loudly."""
    if pkg.lib().mm2gb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.Mm2gbError):
        pkg.ChainContext(pkg.map_ont_misc())
