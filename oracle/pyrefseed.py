"""ctypes binding of oracle/_ref/libref_seed.so -- the reference's own seeding stage (mm_sketch, mm_idx_get, mm_map_seed)
behind oracle/seed_shim.c.  TEST INFRASTRUCTURE ONLY: the checker of the device seeding path.  Only
tests/, __graft_entry__.smoke() and the reference / cpu_baseline legs of bench.py may import this module."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref_seed.so")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.refseed_index_build.restype = C.c_void_p
        L.refseed_index_build.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
        L.refseed_index_destroy.argtypes = [C.c_void_p]
        L.refseed_mid_occ.restype = C.c_int
        L.refseed_mid_occ.argtypes = [C.c_void_p, C.c_float]
        L.refseed_index_get.restype = C.c_int64
        L.refseed_index_get.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64]
        L.refseed_sketch.restype = C.c_int64
        L.refseed_sketch.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_int64]
        L.refseed_opt_new.restype = C.c_void_p
        L.refseed_opt_new.argtypes = [C.c_char_p, C.c_void_p]
        L.refseed_opt_free.argtypes = [C.c_void_p]
        L.refseed_opt_field.restype = C.c_double
        L.refseed_opt_field.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_double]
        L.refseed_seed.restype = C.c_int64
        L.refseed_seed.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int),
                                   C.c_void_p, C.c_int64, C.POINTER(C.c_int)]
        L.refseed_seed_batch.restype = C.c_int
        L.refseed_seed_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refseed_misc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.refseed_radix_sort_128x.argtypes = [C.c_void_p, C.c_int64]
        _lib = L
    return _lib


def sketch(seq: bytes, w: int, k: int, rid: int = 0, hpc: bool = False) -> np.ndarray:
    """mm_sketch (sketch.c:77): (n, 2) uint64 array of (x = hash << 8 | span, y = rid << 32 | last_pos << 1 | strand)."""
    cap = len(seq) + 16
    while True:
        out = np.empty((cap, 2), dtype=np.uint64)
        n = lib().refseed_sketch(seq, len(seq), w, k, rid, int(hpc), out.ctypes.data, cap)
        if n <= cap:
            return out[:n].copy()
        cap = int(n)


DIGEST_C = 0x9E3779B97F4A7C15


def word_digest(w) -> int:
    """seed_shim.c word_digest: C (m + 1) + sum_k w[k] (2k + 1) C  mod 2^64 over the 64-bit words of `w`."""
    w = np.ascontiguousarray(w, np.uint64).reshape(-1)
    k = np.arange(len(w), dtype=np.uint64)
    c = np.uint64(DIGEST_C)
    with np.errstate(over="ignore"):
        return int(c * np.uint64(len(w) + 1) + (w * ((np.uint64(2) * k + np.uint64(1)) * c)).sum(dtype=np.uint64))


def chain_digest(u, b) -> int:
    """what seed_batch(chain=True) reports per read: digest(u[]) + 31 * digest(compacted anchors)"""
    return (word_digest(u) + 31 * word_digest(b)) & 0xFFFFFFFFFFFFFFFF


def radix_sort_128x(xy: np.ndarray) -> np.ndarray:
    """The reference's unstable in-place MSD radix sort of mm128_t by x (ksort.h:98-151) on a copy of the (n, 2) array."""
    out = np.ascontiguousarray(xy, dtype=np.uint64).copy()
    lib().refseed_radix_sort_128x(out.ctypes.data, len(out))
    return out


class RefIndex:
    """mm_idx_str over in-memory sequences + map options set up the way the driver does it."""

    def __init__(self, seqs, w=10, k=15, hpc=False, bucket_bits=14, preset="map-ont", names=None):
        L = lib()
        n = len(seqs)
        self._seqs = [s if isinstance(s, bytes) else bytes(s) for s in seqs]
        names = names or [b"ref%d" % i for i in range(n)]
        sa = (C.c_char_p * n)(*self._seqs)
        na = (C.c_char_p * n)(*names)
        self.mi = L.refseed_index_build(w, k, int(hpc), bucket_bits, n, sa, na)
        if not self.mi:
            raise RuntimeError("mm_idx_str failed")
        self.w, self.k = w, k
        self.opt = L.refseed_opt_new(preset.encode(), self.mi)
        if not self.opt:
            raise RuntimeError("unknown preset")

    def close(self):
        if getattr(self, "opt", None):
            lib().refseed_opt_free(self.opt)
            self.opt = None
        if getattr(self, "mi", None):
            lib().refseed_index_destroy(self.mi)
            self.mi = None

    __del__ = close

    def field(self, name: str, value=None):
        v = lib().refseed_opt_field(self.opt, name.encode(), 0 if value is None else 1, 0.0 if value is None else float(value))
        if v == -1e300:
            raise KeyError(name)
        return v

    def mid_occ_of(self, frac: float) -> int:
        return lib().refseed_mid_occ(self.mi, frac)

    def get(self, minier: int) -> np.ndarray:
        cap = 64
        while True:
            out = np.empty(cap, dtype=np.uint64)
            n = lib().refseed_index_get(self.mi, minier, out.ctypes.data, cap)
            if n <= cap:
                return out[:n].copy()
            cap = int(n)

    def misc(self) -> bytes:
        buf = C.create_string_buffer(44)
        lib().refseed_misc(self.mi, self.opt, buf)
        return buf.raw

    def seed(self, seq: bytes):
        """mm_map_seed (map.c:355-391) of one read -> anchors (n, 2) uint64, rep_len, mini_pos."""
        cap = 1 << 16
        while True:
            out = np.empty((cap, 2), dtype=np.uint64)
            mp = np.empty(len(seq) + 16, dtype=np.uint64)
            rep, nmp = C.c_int(0), C.c_int(0)
            n = lib().refseed_seed(self.mi, self.opt, seq, len(seq), out.ctypes.data, cap, C.byref(rep), mp.ctypes.data, len(mp),
                                   C.byref(nmp))
            if n <= cap:
                return out[:n].copy(), rep.value, mp[:nmp.value].copy()
            cap = int(n)

    def seed_batch(self, seqs: np.ndarray, seq_off: np.ndarray, chain=False, threads=1, out_off=None):
        """n_a (and with chain: n_u, digest of u[] + compacted anchors; without: digest of the anchors) per read; anchors themselves
        into a (total, 2) array when out_off (room per read) is given."""
        n = len(seq_off) - 1
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
        n_a = np.zeros(n, dtype=np.int64)
        n_u = np.zeros(n, dtype=np.int32)
        dig = np.zeros(n, dtype=np.uint64)
        out = None
        if out_off is not None:
            out_off = np.ascontiguousarray(out_off, dtype=np.int64)
            out = np.zeros((int(out_off[-1]), 2), dtype=np.uint64)
        lib().refseed_seed_batch(self.mi, self.opt, seqs.ctypes.data_as(C.c_char_p), seq_off.ctypes.data, n, int(chain), threads,
                                 out.ctypes.data if out is not None else None, out_off.ctypes.data if out is not None else None,
                                 n_a.ctypes.data, n_u.ctypes.data, dig.ctypes.data)
        return n_a, n_u, dig, out
