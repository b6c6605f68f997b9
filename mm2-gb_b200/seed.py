"""ctypes mirror of include/mm2gb_seed.h -- device seeding (mm_map_seed, map.c:355-391) and the fused seed + chain step.

Thin binding for the tests and bench.py; no seeding logic lives here and there is NO CPU fallback: without the library or a
CUDA device every call raises."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import Mm2gbError, Stats, lib as _chain_lib

N_TIMERS = 6
TIMER_NAMES = ("sketch", "qocc_filter", "lookup", "select", "expand", "sort")


class SeedParams(C.Structure):
    """mm2gb_seed_params_t: the fields of mm_mapopt_t the seeding stage reads (options.c:18-38 defaults for map-ont)."""

    _fields_ = [("mid_occ", C.c_int32), ("max_max_occ", C.c_int32), ("occ_dist", C.c_int32), ("q_occ_frac", C.c_float),
                ("flag", C.c_int64), ("sdust_thres", C.c_int32), ("max_qlen", C.c_int32)]


def map_ont_seed_params(mid_occ: int, **over) -> SeedParams:
    p = SeedParams(mid_occ=mid_occ, max_max_occ=4095, occ_dist=500, q_occ_frac=0.01, flag=0, sdust_thres=0, max_qlen=0)
    for k, v in over.items():
        setattr(p, k, v)
    return p


class SeedChainResult(C.Structure):
    _fields_ = [("n_reads", C.c_int), ("n_anchors", C.c_int64), ("n_chain_anchors", C.c_int64), ("n_chains", C.c_int64),
                ("a_off", C.POINTER(C.c_int64)), ("rep_len", C.POINTER(C.c_int32)),
                ("n_u", C.POINTER(C.c_int32)), ("u_pos", C.POINTER(C.c_int32)), ("n_b", C.POINTER(C.c_int32)),
                ("b_pos", C.POINTER(C.c_int32)), ("u", C.POINTER(C.c_uint64)), ("b", C.c_void_p), ("stats", Stats),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


_bound = False


def lib():
    global _bound
    L = _chain_lib()
    if not _bound:
        vp = C.c_void_p
        L.mm2gb_index_build.argtypes = [C.POINTER(vp), C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.mm2gb_index_destroy.argtypes = [vp]
        L.mm2gb_index_destroy.restype = None
        L.mm2gb_index_cal_max_occ.argtypes = [vp, C.c_float]
        L.mm2gb_index_cal_max_occ.restype = C.c_int32
        L.mm2gb_index_get.argtypes = [vp, C.c_uint64, vp, C.c_int64]
        L.mm2gb_index_get.restype = C.c_int64
        L.mm2gb_index_n_keys.argtypes = [vp]
        L.mm2gb_index_n_keys.restype = C.c_int64
        L.mm2gb_index_n_occ.argtypes = [vp]
        L.mm2gb_index_n_occ.restype = C.c_int64
        L.mm2gb_seeder_create.argtypes = [C.POINTER(vp), vp, C.c_int64, C.c_int, C.c_int64]
        L.mm2gb_seeder_destroy.argtypes = [vp]
        L.mm2gb_seeder_destroy.restype = None
        L.mm2gb_sketch_host.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, C.c_int64, vp]
        L.mm2gb_seed_host.argtypes = [vp, C.POINTER(SeedParams), vp, vp, C.c_int, vp, C.c_int64, vp, vp, vp, C.c_int64, vp]
        L.mm2gb_seed_chain.argtypes = [vp, vp, C.POINTER(SeedParams), vp, vp, C.c_int, C.POINTER(SeedChainResult)]
        L.mm2gb_seed_chain_device.argtypes = [vp, vp, C.POINTER(SeedParams), vp, vp, C.c_int, C.POINTER(C.c_int64)]
        L.mm2gb_seed_last_mini_pos.argtypes = [vp, C.c_int, vp, C.c_int64, vp]
        L.mm2gb_seed_profile.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _bound = True
    return L


def _ck(rc: int):
    if rc != 0:
        raise Mm2gbError(f"mm2gb error {rc}: {lib().mm2gb_last_error().decode()}")


def pack_seqs(seqs):
    """list of bytes -> (uint8 array of the concatenated bases, int64 offsets)."""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    for i, s in enumerate(seqs):
        off[i + 1] = off[i] + len(s)
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, dtype=np.uint8)
    return buf, off


class Index:
    """mm2gb_index_t: the minimizer index of in-memory sequences, built on and resident in the device."""

    def __init__(self, seqs, w: int = 10, k: int = 15, hpc: bool = False, device: int = 0):
        buf, off = seqs if isinstance(seqs, tuple) else pack_seqs(seqs)
        self._h = C.c_void_p()
        self.w, self.k = w, k
        _ck(lib().mm2gb_index_build(C.byref(self._h), device, buf.ctypes.data, off.ctypes.data, len(off) - 1, w, k, int(hpc), 14))

    def close(self):
        if getattr(self, "_h", None) and self._h:
            lib().mm2gb_index_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def cal_max_occ(self, frac: float) -> int:
        return lib().mm2gb_index_cal_max_occ(self._h, frac)

    def mid_occ(self, frac: float = 2e-4, lo: int = 10, hi: int = 1000000) -> int:
        """mm_mapopt_update (options.c:72-77)."""
        m = self.cal_max_occ(frac)
        m = max(m, lo)
        if hi > lo and m > hi:
            m = hi
        return m

    def get(self, minier: int) -> np.ndarray:
        n = lib().mm2gb_index_get(self._h, minier, None, 0)
        out = np.empty(max(n, 1), dtype=np.uint64)
        lib().mm2gb_index_get(self._h, minier, out.ctypes.data, n)
        return out[:n]

    @property
    def n_keys(self) -> int:
        return lib().mm2gb_index_n_keys(self._h)

    @property
    def n_occ(self) -> int:
        return lib().mm2gb_index_n_occ(self._h)


class Seeder:
    def __init__(self, index: Index, max_bases: int, max_reads: int, max_anchors: int):
        self.index = index
        self.max_anchors = max_anchors
        self._h = C.c_void_p()
        _ck(lib().mm2gb_seeder_create(C.byref(self._h), index._h, max_bases, max_reads, max_anchors))

    def close(self):
        if getattr(self, "_h", None) and self._h:
            lib().mm2gb_seeder_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def sketch(self, buf: np.ndarray, off: np.ndarray, rid_is_seq: bool = False):
        """mm_sketch of every sequence: (n, 2) uint64 minimizers and the per-sequence offsets."""
        n = len(off) - 1
        cap = max(16, int(off[-1]) + 64 * n)
        out = np.zeros((cap, 2), dtype=np.uint64)
        mvo = np.zeros(n + 1, dtype=np.int64)
        _ck(lib().mm2gb_sketch_host(self._h, buf.ctypes.data, off.ctypes.data, n, int(rid_is_seq), out.ctypes.data, cap, mvo.ctypes.data))
        return out[:int(mvo[-1])], mvo

    def seed(self, prm: SeedParams, buf: np.ndarray, off: np.ndarray, want_mini_pos: bool = True):
        """mm_map_seed of a batch -> anchors (n, 2) uint64, a_off, rep_len, mini_pos, mp_off."""
        n = len(off) - 1
        a = np.zeros((self.max_anchors, 2), dtype=np.uint64)
        a_off = np.zeros(n + 1, dtype=np.int64)
        rep = np.zeros(max(n, 1), dtype=np.int32)
        mp_cap = int(off[-1]) + 16
        mp = np.zeros(mp_cap, dtype=np.uint64)
        mp_off = np.zeros(n + 1, dtype=np.int64)
        _ck(lib().mm2gb_seed_host(self._h, C.byref(prm), buf.ctypes.data, off.ctypes.data, n, a.ctypes.data, self.max_anchors,
                                  a_off.ctypes.data, rep.ctypes.data, mp.ctypes.data if want_mini_pos else None, mp_cap, mp_off.ctypes.data))
        return a[:int(a_off[-1])], a_off, rep[:n], mp[:int(mp_off[-1])], mp_off

    def seed_chain(self, ctx, prm: SeedParams, buf: np.ndarray, off: np.ndarray, copy: bool = True):
        """Fused step: sequences in, chains + compacted anchors out.  Returns a dict of numpy arrays (copies unless copy=False)."""
        n = len(off) - 1
        res = SeedChainResult()
        ptr = buf if isinstance(buf, int) else buf.ctypes.data
        _ck(lib().mm2gb_seed_chain(self._h, ctx._h, C.byref(prm), ptr, off.ctypes.data, n, C.byref(res)))
        if not copy:
            return res
        def arr(p, cnt, dt):
            if cnt == 0:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(p, shape=(cnt,)).astype(dt, copy=True)
        out = {
            "n_anchors": res.n_anchors, "n_chains": res.n_chains, "n_chain_anchors": res.n_chain_anchors,
            "a_off": arr(res.a_off, n + 1, np.int64), "rep_len": arr(res.rep_len, n, np.int32),
            "n_u": arr(res.n_u, n, np.int32), "u_pos": arr(res.u_pos, n, np.int32), "n_b": arr(res.n_b, n, np.int32),
            "b_pos": arr(res.b_pos, n, np.int32), "u": arr(res.u, int(res.n_chains), np.uint64),
            "stats": res.stats.as_dict(), "h2d_bytes": res.h2d_bytes, "d2h_bytes": res.d2h_bytes,
        }
        nb = int(res.n_chain_anchors)
        if nb:
            raw = (C.c_uint64 * (2 * nb)).from_address(res.b)
            out["b"] = np.frombuffer(raw, dtype=np.uint64).reshape(nb, 2).copy()
        else:
            out["b"] = np.zeros((0, 2), dtype=np.uint64)
        return out

    def last_mini_pos(self, n_reads: int, cap: int):
        """mini_pos of the last batch (what mm_est_err needs): (entries, per-read offsets)."""
        mp = np.zeros(max(cap, 1), dtype=np.uint64)
        mp_off = np.zeros(n_reads + 1, dtype=np.int64)
        _ck(lib().mm2gb_seed_last_mini_pos(self._h, n_reads, mp.ctypes.data, cap, mp_off.ctypes.data))
        return mp[:int(mp_off[-1])], mp_off

    def seed_chain_device(self, ctx, prm: SeedParams, d_seqs: int, off: np.ndarray) -> int:
        n_a = C.c_int64(0)
        _ck(lib().mm2gb_seed_chain_device(self._h, ctx._h, C.byref(prm), d_seqs, off.ctypes.data, len(off) - 1, C.byref(n_a)))
        return n_a.value

    def profile(self):
        ms = (C.c_float * N_TIMERS)()
        nmv, nm = C.c_int64(0), C.c_int64(0)
        _ck(lib().mm2gb_seed_profile(self._h, ms, C.byref(nmv), C.byref(nm)))
        return dict(zip(TIMER_NAMES, [float(x) for x in ms])), nmv.value, nm.value
