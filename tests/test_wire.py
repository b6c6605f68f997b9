"""The packed upload format (mm2-gb_b200/csrc/wire.h): 8 bytes per anchor + one record per run of equal high words.
Host-only tests of the packer against its inverse (the rule k_expand applies on the device); the GPU side is covered by
tests/test_gpu_wire.py.  No CUDA device needed: the library loads and these entry points do no device work."""
import numpy as np
import pytest


def _roundtrip(pkg, a, off):
    n = int(off[-1])
    buf, nbytes, n_runs = pkg.wire_pack(a, off)
    assert nbytes >= 8 * n
    back = pkg.wire_unpack(buf, n, n_runs)
    assert np.array_equal(back, a[:n].reshape(-1, 2))
    return nbytes, n_runs


def test_ont_like_batch_roundtrip(pkg, synth):
    a, off = synth.ont_like_batch(7, 40, 1, 3000)
    n = int(off[-1])
    nbytes, n_runs = _roundtrip(pkg, a, off)
    # one run per (rid, strand) stretch of a read: a handful per read, so the upload is 8 B/anchor plus a sliver
    assert n_runs <= 16 * (len(off) - 1)
    assert nbytes < 8.2 * n + 4096


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 31, 255, 256, 257, 511, 512, 513, 1000])
def test_sizes_around_block_and_vector_boundaries(pkg, n):
    rng = np.random.default_rng(n)
    a = np.zeros((n, 2), np.uint64)
    a[:, 0] = np.sort(rng.integers(0, 1 << 31, n).astype(np.uint64)) | (np.uint64(3) << np.uint64(32))
    a[:, 1] = rng.integers(0, 1 << 31, n).astype(np.uint64) | (np.uint64(15) << np.uint64(32))
    _roundtrip(pkg, a, np.array([0, n], np.int64))


def test_every_high_bit_survives(pkg):
    """rev bit 63, rid, seg_id << 48, flags << 40, q_span << 32; low words with the top bit set"""
    rng = np.random.default_rng(3)
    n = 5000
    a = np.zeros((n, 2), np.uint64)
    xh = np.repeat(rng.integers(0, 1 << 32, 25, dtype=np.uint64), 200)
    yh = np.repeat(rng.integers(0, 1 << 32, 50, dtype=np.uint64), 100)
    a[:, 0] = (xh << np.uint64(32)) | rng.integers(0, 1 << 32, n, dtype=np.uint64)
    a[:, 1] = (yh << np.uint64(32)) | rng.integers(0, 1 << 32, n, dtype=np.uint64)
    off = np.array([0, 1, 1, 777, 4096, n], np.int64)       # an empty read, reads that start inside a vector group
    nbytes, n_runs = _roundtrip(pkg, a, off)
    assert 50 <= n_runs <= 76


def test_run_boundaries_at_every_offset_of_a_vector_group(pkg):
    """a run that starts at anchor k for every k in a window: the 4-anchor vector body must hand over to the scalar path"""
    base = np.zeros((64, 2), np.uint64)
    base[:, 0] = np.arange(64, dtype=np.uint64) + (np.uint64(1) << np.uint64(32))
    base[:, 1] = np.arange(64, dtype=np.uint64) + (np.uint64(15) << np.uint64(32))
    for k in range(1, 40):
        a = base.copy()
        a[k:, 0] += np.uint64(1) << np.uint64(32)
        nbytes, n_runs = _roundtrip(pkg, a, np.array([0, 64], np.int64))
        assert n_runs == 2
        a[k:, 1] += np.uint64(1) << np.uint64(40)
        a[k + 1:, 1] += np.uint64(1) << np.uint64(48)
        _, n_runs = _roundtrip(pkg, a, np.array([0, 64], np.int64))
        assert n_runs == 3


def test_run_list_that_does_not_fit_is_refused(pkg):
    """a q_span per anchor (HPC seeds): one run per anchor -- the packer says so and the upload path sends raw anchors"""
    n = 4096
    a = np.zeros((n, 2), np.uint64)
    a[:, 0] = np.arange(n, dtype=np.uint64)
    a[:, 1] = np.arange(n, dtype=np.uint64) | ((np.arange(n, dtype=np.uint64) % np.uint64(7) + np.uint64(10)) << np.uint64(32))
    off = np.array([0, n], np.int64)
    buf, nbytes, _ = pkg.wire_pack(a, off, cap_bytes=16 * n)       # the size of the staging buffer of an n-anchor slot
    assert nbytes == -1
    buf, nbytes, n_runs = pkg.wire_pack(a, off, cap_bytes=32 * n)   # with room for the runs it still round-trips
    assert nbytes > 0 and n_runs > n // 2
    assert np.array_equal(pkg.wire_unpack(buf, n, n_runs), a)


def test_gather_is_compact_a(pkg):
    rng = np.random.default_rng(1)
    a = rng.integers(0, 1 << 63, (1000, 2), dtype=np.uint64)
    for m in (0, 1, 3, 4, 5, 999):
        v = rng.integers(0, 1000, m).astype(np.int32)
        assert np.array_equal(pkg.gather_anchors(a, v), a[v])
