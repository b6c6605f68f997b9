"""GPU tests (-m gpu) of the two wire formats of the boundary, against the oracle and against each other:
   upload   raw 16-byte anchors  vs  the packed 8-byte format + k_expand (csrc/wire.h, chain_kernels.cuh)
   download indices of the chain anchors (4 B each, gathered on the host)  -- compared with the reference's a'[]."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(pkg, a, off, mode, monkeypatch, chunk=None, pinned=False):
    import torch
    monkeypatch.setenv("MM2GB_WIRE", mode)
    if chunk:
        monkeypatch.setenv("MM2GB_CHUNK", str(chunk))
    src = torch.from_numpy(a.view(np.int64)).pin_memory() if pinned else a
    with pkg.ChainContext(pkg.map_ont_misc(), max_anchors=1 << 18, max_reads=512, n_slots=3) as c:
        res = c.chain(src, off)
        up = c.upload_bytes()
        idx = c.chain(src, off, packed=True)
    monkeypatch.delenv("MM2GB_WIRE")
    if chunk:
        monkeypatch.delenv("MM2GB_CHUNK")
    return res, up, idx


def _same(x, y, off):
    assert np.array_equal(x["f"], y["f"]) and np.array_equal(x["p"], y["p"])
    assert np.array_equal(x["n_u"], y["n_u"]) and np.array_equal(x["n_b"], y["n_b"])
    for r in range(len(off) - 1):
        s = int(off[r])
        assert np.array_equal(x["u"][s:s + x["n_u"][r]], y["u"][s:s + y["n_u"][r]]), r
        assert np.array_equal(x["b"][s:s + x["n_b"][r]], y["b"][s:s + y["n_b"][r]]), r


def _vs_oracle(pkg, po, a, off, res, idx):
    prm = po.map_ont_params()
    for r in range(len(off) - 1):
        s, e = int(off[r]), int(off[r + 1])
        o = po.oracle_lchain(prm, a[s:e])
        assert res["n_u"][r] == len(o.u) and res["n_b"][r] == len(o.b), r
        assert np.array_equal(res["u"][s:s + len(o.u)], o.u) and np.array_equal(res["b"][s:s + len(o.b)], o.b), r
        q = int(idx["v_pos"][r])
        assert idx["n_u"][r] == len(o.u) and idx["n_b"][r] == len(o.b), r
        assert np.array_equal(idx["u"][s:s + len(o.u)], o.u), r
        assert np.array_equal(pkg.gather_anchors(a[s:e], idx["v"][q:q + len(o.b)]), o.b), r


@pytest.mark.parametrize("seed,n_reads,lo,hi,chunk", [(41, 60, 1, 4000, 40000), (42, 300, 1, 300, 5000), (43, 6, 9000, 30000, None)])
def test_packed_upload_equals_raw_upload_equals_oracle(pkg, po, synth, monkeypatch, seed, n_reads, lo, hi, chunk):
    a, off = synth.ont_like_batch(seed, n_reads, lo, hi, repeat_copies=2, repeat_len=40)
    n = int(off[-1])
    raw, up_raw, idx_raw = _run(pkg, a, off, "raw", monkeypatch, chunk)
    pk, up_pk, idx_pk = _run(pkg, a, off, "packed", monkeypatch, chunk)
    auto, up_auto, _ = _run(pkg, a, off, "auto", monkeypatch, chunk)
    pin, up_pin, _ = _run(pkg, a, off, "auto", monkeypatch, chunk, pinned=True)
    pkpin, up_pkpin, _ = _run(pkg, a, off, "packed", monkeypatch, chunk, pinned=True)
    assert up_raw == 16 * n and up_pin == 16 * n            # raw: staged copy / direct DMA from the caller's pinned buffer
    assert 8 * n <= up_pk <= 8 * n + 64 * n_reads + 8192 * 8  # 8 B/anchor + block index + a few runs per read and chunk
    assert up_auto == up_pk and up_pkpin == up_pk           # pageable sources are packed by default
    for other in (pk, auto, pin, pkpin):
        _same(raw, other, off)
    _vs_oracle(pkg, po, a, off, pk, idx_pk)
    _vs_oracle(pkg, po, a, off, raw, idx_raw)


def test_high_words_that_change_every_anchor_fall_back_to_raw(pkg, po, synth, monkeypatch):
    """a q_span per anchor (what HPC seeding produces) and seed flags sprinkled over the reads: the run list does not fit the
    staging buffer, the batch goes up raw, results are the oracle's either way"""
    a, off = synth.ont_like_batch(44, 20, 50, 2500)
    n = int(off[-1])
    i = np.arange(n, dtype=np.uint64)
    a = a.copy()
    a[:, 1] = (a[:, 1] & ~(np.uint64(0xff) << np.uint64(32))) | ((np.uint64(11) + i % np.uint64(9)) << np.uint64(32))
    a[::37, 1] |= np.uint64(1) << np.uint64(42)              # MM_SEED_TANDEM on some seeds
    res, up, idx = _run(pkg, a, off, "packed", monkeypatch)
    assert up == 16 * n
    _vs_oracle(pkg, po, a, off, res, idx)
    # flags on a few seeds only: still packed, still exact
    a2, _ = synth.ont_like_batch(44, 20, 50, 2500)
    a2 = a2.copy()
    a2[::37, 1] |= np.uint64(1) << np.uint64(42)
    res2, up2, idx2 = _run(pkg, a2, off, "packed", monkeypatch)
    assert up2 < 12 * n
    _vs_oracle(pkg, po, a2, off, res2, idx2)


def test_negative_min_score_is_chained_on_the_device(pkg, po, synth):
    """min_score < 0 selects what 0 selects (scores are never negative): no host path is involved"""
    a, off = synth.ont_like_batch(45, 12, 20, 1500)
    misc = pkg.map_ont_misc(min_score=-7, min_cnt=1)
    prm = po.map_ont_params(min_score=-7, min_cnt=1)
    with pkg.ChainContext(misc, max_anchors=1 << 16, max_reads=64, n_slots=2) as c:
        res = c.chain(a, off)
    for r in range(len(off) - 1):
        s, e = int(off[r]), int(off[r + 1])
        o = po.oracle_lchain(prm, a[s:e])
        assert res["n_u"][r] == len(o.u) and res["n_b"][r] == len(o.b), r
        assert np.array_equal(res["u"][s:s + len(o.u)], o.u) and np.array_equal(res["b"][s:s + len(o.b)], o.b), r
