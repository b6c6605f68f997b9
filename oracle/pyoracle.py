"""ctypes front-end of the parity checker.  TEST INFRASTRUCTURE ONLY.

Loads oracle/liboracle.so (the C restatement, chain_oracle.c) and, when it has been built in the
container that holds /root/reference, oracle/_ref/libref_lchain.so (the reference's own lchain.c,
see ref_shim.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.

Anchors travel as numpy uint64 arrays of shape (n, 2): column 0 = mm128_t.x, column 1 = mm128_t.y
(minimap.h:72).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
INT32_MAX = 2**31 - 1


class Params(C.Structure):
    """gpu/plutils.h:33-37 (Misc), same field order."""

    _fields_ = [(k, C.c_int32) for k in
                ("max_iter", "max_dist_x", "max_dist_y", "max_skip", "bw", "min_cnt", "min_score", "is_cdna", "n_seg")] + \
               [("chn_pen_gap", C.c_float), ("chn_pen_skip", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def chain_pen(scale: float, k: int) -> float:
    """map.c:409-410: `opt->chain_gap_scale * 0.01 * mi->k` -- float * double * int, rounded to float on store."""
    return float(np.float32(np.float64(np.float32(scale)) * 0.01 * k))


def map_ont_params(k: int = 15, max_skip: int = INT32_MAX, **over) -> Params:
    """options.c:24-36 defaults (= -x map-ont) pushed through build_misc (map.c:393-426)."""
    p = Params(max_iter=5000, max_dist_x=5000, max_dist_y=5000, max_skip=max_skip, bw=500, min_cnt=3, min_score=40,
               is_cdna=0, n_seg=1, chn_pen_gap=chain_pen(0.8, k), chn_pen_skip=chain_pen(0.0, k))
    for key, val in over.items():
        setattr(p, key, val)
    return p


def _build(target: str | None = None):
    cmd = ["make", "-s", "-C", HERE] + ([target] if target else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(HERE, "chain_oracle.c")):
            _build()
        _lib = C.CDLL(path)
        _lib.orc_lchain.restype = C.c_int32
        _lib.orc_lchain.argtypes = [C.POINTER(Params), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64),
                                    C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        _lib.orc_chain_dp.restype = C.c_int64
        _lib.orc_chain_dp.argtypes = [C.POINTER(Params), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_backtrack.restype = C.c_int32
        _lib.orc_backtrack.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        _lib.orc_compact.restype = None
        _lib.orc_compact.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_radix_sort_128x.restype = None
        _lib.orc_radix_sort_128x.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_gap_penalty.restype = C.c_int32
        _lib.orc_gap_penalty.argtypes = [C.c_int32, C.c_int32, C.POINTER(Params)]
        _lib.orc_pair_score.restype = C.c_int32
        _lib.orc_pair_score.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Params)]
        _lib.orc_lchain_batch.restype = C.c_int64
        _lib.orc_lchain_batch.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                          C.c_void_p, C.c_void_p]
    return _lib


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_lchain.so"))


def ref():
    """The reference's own lchain.c (oracle/_ref/libref_lchain.so); None when it was never built."""
    global _ref
    if _ref is None:
        if not ref_available():
            return None
        _ref = C.CDLL(os.path.join(HERE, "_ref", "libref_lchain.so"))
        _ref.ref_lchain.restype = C.c_int32
        _ref.ref_lchain.argtypes = [C.POINTER(Params), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64),
                                    C.c_void_p, C.c_void_p]
        _ref.ref_radix_sort_128x.restype = None
        _ref.ref_radix_sort_128x.argtypes = [C.c_void_p, C.c_int64]
        _ref.ref_backtrack_compact.restype = C.c_int32
        _ref.ref_backtrack_compact.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        _ref.ref_lchain_batch.restype = None
        _ref.ref_lchain_batch.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int]
    return _ref


@dataclass
class ChainResult:
    f: np.ndarray        # int32[n]   chain score ending at each anchor (lchain.c:202)
    p: np.ndarray        # int64[n]   predecessor index, -1 = none
    u: np.ndarray        # uint64[n_u] score<<32 | count, ordered by chain start (lchain.c:100-106)
    b: np.ndarray        # uint64[n_b,2] compacted anchors
    n_pairs: int         # evaluated pairs (n_iter, lchain.c:177); -1 when the reference ran (it does not expose it)

    def same(self, other: "ChainResult") -> bool:
        return (np.array_equal(self.f, other.f) and np.array_equal(self.p, other.p)
                and np.array_equal(self.u, other.u) and np.array_equal(self.b, other.b))


def _anchors(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.ndim == 2 and a.shape[1] == 2
    return a


def oracle_lchain(prm: Params, a) -> ChainResult:
    """Whole mg_lchain_dp via the restatement (chain_oracle.c:orc_lchain)."""
    a = _anchors(a)
    n = a.shape[0]
    f = np.empty(n, np.int32); p = np.empty(n, np.int64)
    u = np.empty(max(n, 1), np.uint64); b = np.empty((max(n, 1), 2), np.uint64)
    nb = C.c_int64(0); npairs = C.c_int64(0)
    n_u = lib().orc_lchain(C.byref(prm), n, a.ctypes.data, u.ctypes.data, b.ctypes.data, C.byref(nb),
                           f.ctypes.data, p.ctypes.data, C.byref(npairs))
    return ChainResult(f, p, u[:n_u].copy(), b[:nb.value].copy(), npairs.value)


def oracle_dp(prm: Params, a):
    """Forward DP only: (f int32[n], p int64[n], n_pairs)."""
    a = _anchors(a)
    n = a.shape[0]
    f = np.empty(n, np.int32); p = np.empty(n, np.int64)
    pairs = lib().orc_chain_dp(C.byref(prm), n, a.ctypes.data, f.ctypes.data, p.ctypes.data, None)
    return f, p, int(pairs)


def oracle_backtrack(prm: Params, a, f, p):
    """Backtrack + compact on given f/p: (u uint64[n_u], b uint64[n_b,2])."""
    a = _anchors(a)
    n = a.shape[0]
    f = np.ascontiguousarray(f, np.int32); p = np.ascontiguousarray(p, np.int64)
    u = np.empty(max(n, 1), np.uint64); v = np.empty(max(n, 1), np.int32); t = np.empty(max(n, 1), np.int32)
    b = np.empty((max(n, 1), 2), np.uint64)
    nv = C.c_int32(0)
    max_drop = INT32_MAX if prm.is_cdna else prm.bw
    n_u = lib().orc_backtrack(n, f.ctypes.data, p.ctypes.data, prm.min_cnt, prm.min_score, max_drop, u.ctypes.data,
                              v.ctypes.data, t.ctypes.data, C.byref(nv))
    if n_u > 0:
        lib().orc_compact(n_u, u.ctypes.data, nv.value, v.ctypes.data, a.ctypes.data, b.ctypes.data)
    return u[:n_u].copy(), b[:nv.value if n_u > 0 else 0].copy()


def ref_lchain(prm: Params, a) -> ChainResult:
    """Whole mg_lchain_dp via the reference's own compiled lchain.c (needs oracle/_ref)."""
    r = ref()
    assert r is not None, "oracle/_ref/libref_lchain.so not built (make -C oracle ref, needs /root/reference)"
    a = _anchors(a)
    n = a.shape[0]
    f = np.empty(n, np.int32); p = np.empty(n, np.int64)
    u = np.empty(max(n, 1), np.uint64); b = np.empty((max(n, 1), 2), np.uint64)
    nb = C.c_int64(0)
    n_u = r.ref_lchain(C.byref(prm), n, a.ctypes.data, u.ctypes.data, b.ctypes.data, C.byref(nb), f.ctypes.data, p.ctypes.data)
    return ChainResult(f, p, u[:n_u].copy(), b[:nb.value].copy(), -1)


def ref_backtrack(prm: Params, a, f, p):
    r = ref()
    assert r is not None
    a = _anchors(a)
    n = a.shape[0]
    f = np.ascontiguousarray(f, np.int32); p = np.ascontiguousarray(p, np.int64)
    u = np.empty(max(n, 1), np.uint64); b = np.empty((max(n, 1), 2), np.uint64)
    nb = C.c_int64(0)
    max_drop = INT32_MAX if prm.is_cdna else prm.bw
    n_u = r.ref_backtrack_compact(n, f.ctypes.data, p.ctypes.data, a.ctypes.data, prm.min_cnt, prm.min_score, max_drop,
                                  u.ctypes.data, b.ctypes.data, C.byref(nb))
    return u[:n_u].copy(), b[:nb.value].copy()


def radix_sort_128x(a, use_ref=False) -> np.ndarray:
    a = _anchors(a).copy()
    if use_ref:
        ref().ref_radix_sort_128x(a.ctypes.data, a.shape[0])
    else:
        lib().orc_radix_sort_128x(a.ctypes.data, a.ctypes.data + a.nbytes)
    return a


def lchain_batch(prm: Params, a, off, r0=0, r1=None, n_threads=1, want_fp=False, use_ref=False):
    """Chain reads [r0, r1) of a concatenated anchor array on `n_threads` host threads (CPU baseline timing).
    Returns (n_pairs, f, p); f/p are None unless want_fp.  With use_ref the reference's own lchain.c runs
    (n_pairs = -1)."""
    a = _anchors(a)
    off = np.ascontiguousarray(off, np.int64)
    r1 = len(off) - 1 if r1 is None else r1
    if use_ref:
        ref().ref_lchain_batch(C.byref(prm), a.ctypes.data, off.ctypes.data, r0, r1, n_threads)
        return -1, None, None
    f = np.zeros(a.shape[0], np.int32) if want_fp else None
    p = np.zeros(a.shape[0], np.int64) if want_fp else None
    pairs = lib().orc_lchain_batch(C.byref(prm), a.ctypes.data, off.ctypes.data, r0, r1, n_threads,
                                   f.ctypes.data if want_fp else None, p.ctypes.data if want_fp else None)
    return int(pairs), f, p


DIGEST_C = 0x9E3779B97F4A7C15


def digest(w) -> int:
    """order-sensitive 64-bit digest of a word array: C (m + 1) + sum_k w[k] (2k + 1) C  mod 2^64 (chain_oracle.c: orc_digest)"""
    w = np.ascontiguousarray(w, np.uint64).reshape(-1)
    k = np.arange(len(w), dtype=np.uint64)
    c = np.uint64(DIGEST_C)
    with np.errstate(over="ignore"):
        return int(c * np.uint64(len(w) + 1) + (w * ((np.uint64(2) * k + np.uint64(1)) * c)).sum(dtype=np.uint64))


def lchain_digests(prm: Params, a, off, sel, n_threads=1, use_ref=False):
    """Whole mg_lchain_dp of the reads `sel` on n_threads host threads; per read (n_u, n_b, digest(u), digest(b)).
    use_ref: the reference's own lchain.c (oracle/_ref), else the restatement."""
    a = _anchors(a)
    off = np.ascontiguousarray(off, np.int64)
    sel = np.ascontiguousarray(sel, np.int64)
    m = len(sel)
    nu = np.zeros(m, np.int32); nb = np.zeros(m, np.int64); hu = np.zeros(m, np.uint64); hb = np.zeros(m, np.uint64)
    if use_ref:
        r = ref()
        r.ref_lchain_digest_batch.restype = None
        r.ref_lchain_digest_batch.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 4
        fn = r.ref_lchain_digest_batch
    else:
        L = lib()
        L.orc_lchain_digest_batch.restype = None
        L.orc_lchain_digest_batch.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 4
        fn = L.orc_lchain_digest_batch
    fn(C.byref(prm), a.ctypes.data, off.ctypes.data, sel.ctypes.data, m, n_threads, nu.ctypes.data, nb.ctypes.data, hu.ctypes.data, hb.ctypes.data)
    return nu, nb, hu, hb


def read_dump(path: str):
    """Anchor dump written by oracle/dump_stub.c: (Params, [ (n_seg, qlen_sum, anchors uint64[n,2]) ... ])."""
    with open(path, "rb") as fh:
        raw = fh.read()
    assert raw[:8] == b"MM2GBAD1"
    prm = Params.from_buffer_copy(raw[8:8 + C.sizeof(Params)])
    pos = 8 + C.sizeof(Params)
    reads = []
    while pos < len(raw):
        n = int(np.frombuffer(raw, np.int64, 1, pos)[0]); pos += 8
        n_seg, qlen = (int(x) for x in np.frombuffer(raw, np.int32, 2, pos)); pos += 8
        a = np.frombuffer(raw, np.uint64, 2 * n, pos).reshape(n, 2).copy(); pos += 16 * n
        reads.append((n_seg, qlen, a))
    return prm, reads
