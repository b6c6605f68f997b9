"""mm2gb_b200 -- B200-native anchor chaining (minimap2-v2.24 mg_lchain_dp at max-chain-skip = infinity).

The product is the C-ABI shared library `libmm2gb_chain.so` (hand-written sm_100a CUDA, see csrc/ and
include/mm2gb_chain.h).  This Python module is only a thin ctypes mirror of that ABI for the tests and bench.py;
it holds no chaining logic and has NO CPU fallback: if the library or a CUDA device is missing every call raises.

The directory is named `mm2-gb_b200` (not importable as is); load it with `__graft_entry__.load_package()`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MM2GB_LIB: an alternative build of the library (kernel experiments, see csrc/Makefile EXTRA); the product is libmm2gb_chain.so
LIB_PATH = os.environ.get("MM2GB_LIB") or os.path.join(HERE, "libmm2gb_chain.so")
INT32_MAX = 2**31 - 1
N_TIMERS = 6
TIMER_NAMES = ("range", "units", "score", "backtrack", "h2d", "d2h")


class Mm2gbError(RuntimeError):
    pass


class Misc(C.Structure):
    """gpu/plutils.h:33-37 (Misc) == mm2gb_misc_t."""

    _fields_ = [(k, C.c_int) for k in
                ("max_iter", "max_dist_x", "max_dist_y", "max_skip", "bw", "min_cnt", "min_score", "is_cdna", "n_seg")] + \
               [("chn_pen_gap", C.c_float), ("chn_pen_skip", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("n_anchors", C.c_int64), ("n_pairs", C.c_int64), ("n_units", C.c_int32), ("n_units_exact", C.c_int32),
                ("n_long", C.c_int32), ("general_path", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def chain_pen(scale: float, k: int) -> float:
    """map.c:409-410 `opt->chain_gap_scale * 0.01 * mi->k` (float * double * int, stored to float)."""
    return float(np.float32(np.float64(np.float32(scale)) * 0.01 * k))


def map_ont_misc(k: int = 15, **over) -> Misc:
    """build_misc (map.c:393-426) for -x map-ont defaults (options.c:24-36); max_skip is ignored by the device path."""
    m = Misc(max_iter=5000, max_dist_x=5000, max_dist_y=5000, max_skip=INT32_MAX, bw=500, min_cnt=3, min_score=40,
             is_cdna=0, n_seg=1, chn_pen_gap=chain_pen(0.8, k), chn_pen_skip=chain_pen(0.0, k))
    for key, val in over.items():
        setattr(m, key, val)
    return m


def build(verbose: bool = False) -> str:
    """Compile libmm2gb_chain.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], capture_output=True, text=True)
    if out.returncode != 0:
        raise Mm2gbError("building libmm2gb_chain.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout + out.stderr)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Mm2gbError(f"{LIB_PATH} is missing: run __graft_entry__.build() (make -C mm2-gb_b200/csrc)")
        L = C.CDLL(LIB_PATH)
        vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
        L.mm2gb_last_error.restype = C.c_char_p
        L.mm2gb_device_count.restype = C.c_int
        L.mm2gb_ctx_create.argtypes = [C.POINTER(vp), C.c_int, C.c_size_t, C.c_int, C.c_int, C.POINTER(Misc)]
        L.mm2gb_ctx_create_ex.argtypes = [C.POINTER(vp), C.c_int, C.c_size_t, C.c_int, C.c_int, C.POINTER(Misc), C.c_uint]
        L.mm2gb_ctx_destroy.argtypes = [vp]
        L.mm2gb_ctx_destroy.restype = None
        L.mm2gb_ctx_set_misc.argtypes = [vp, C.POINTER(Misc)]
        L.mm2gb_chain_dp_host.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.POINTER(Stats)]
        L.mm2gb_chain_host.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, C.POINTER(Stats)]
        L.mm2gb_chain_host_index.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, C.POINTER(Stats)]
        L.mm2gb_gather_anchors.argtypes = [vp, vp, C.c_int64, vp]
        L.mm2gb_gather_anchors.restype = None
        L.mm2gb_last_batch_upload_bytes.argtypes = [vp]
        L.mm2gb_last_batch_upload_bytes.restype = C.c_int64
        L.mm2gb_device_memory.argtypes = [C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.mm2gb_wire_pack.argtypes = [vp, vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_int32)]
        L.mm2gb_wire_pack.restype = C.c_int64
        L.mm2gb_wire_unpack.argtypes = [vp, C.c_size_t, C.c_int64, C.c_int32, vp]
        L.mm2gb_submit.argtypes = [vp, C.c_int, vp, vp, C.c_int]
        L.mm2gb_submit_gather.argtypes = [vp, C.c_int, vp, vp, C.c_int]
        L.mm2gb_wait.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(Stats)]
        L.mm2gb_submit_gather_chains.argtypes = [vp, C.c_int, vp, vp, C.c_int]
        L.mm2gb_wait_chains.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(Stats)]
        L.mm2gb_slot_busy.argtypes = [vp, C.c_int]
        L.mm2gb_chain_dp_device.argtypes = [vp, vp, vp, C.c_int, C.c_int64, vp, vp]
        L.mm2gb_chain_device.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int64, vp, vp]
        L.mm2gb_chain_device_slot.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int64, vp, vp]
        L.mm2gb_sync.argtypes = [vp, C.c_int]
        L.mm2gb_stream.argtypes = [vp, C.c_int]
        L.mm2gb_stream.restype = vp
        L.mm2gb_device_stats.argtypes = [vp, C.POINTER(Stats)]
        L.mm2gb_profile.argtypes = [vp, C.c_int]
        L.mm2gb_profile_read.argtypes = [vp, C.POINTER(C.c_float), i64p]
        L.mm2gb_backtrack_batch.argtypes = [C.POINTER(Misc), vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int]
        L.mm2gb_backtrack_device.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int32)]
        L.mm2gb_backtrack.restype = C.c_int32
        L.mm2gb_backtrack.argtypes = [C.c_int64, vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, i64p]
        _lib = L
    return _lib


def _ck(rc: int):
    if rc != 0:
        raise Mm2gbError(f"mm2gb error {rc}: {lib().mm2gb_last_error().decode()}")


def _ptr(x):
    """Raw address of a numpy array, a torch tensor, an int, or None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return int(x)


class ChainContext:
    """One chaining context = (GPU, capacity, chaining parameters); mirrors mm2gb_ctx_t."""

    DEVICE_ONLY, NO_CHAINS, NO_FP_STAGING = 1, 2, 4

    def __init__(self, misc: Misc | None = None, device: int = 0, max_anchors: int = 1 << 22, max_reads: int = 1 << 16, n_slots: int = 2,
                 flags: int = 0):
        self.misc = misc if misc is not None else map_ont_misc()
        self._h = C.c_void_p()
        if lib().mm2gb_device_count() <= device:
            raise Mm2gbError(f"no CUDA device {device}: the chaining path is GPU-only (no CPU fallback)")
        _ck(lib().mm2gb_ctx_create_ex(C.byref(self._h), device, max_anchors, max_reads, n_slots, C.byref(self.misc), flags))
        self.device, self.max_anchors, self.max_reads, self.n_slots = device, max_anchors, max_reads, n_slots

    def close(self):
        if self._h:
            lib().mm2gb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_misc(self, misc: Misc):
        _ck(lib().mm2gb_ctx_set_misc(self._h, C.byref(misc)))
        self.misc = misc

    # ---- host-buffer path -------------------------------------------------------------------------------
    def chain_dp(self, a, off, f=None, p=None):
        """a: uint64[n,2] host anchors (numpy or pinned torch), off: int64[n_reads+1].  Returns (f, p, stats);
        f/p int32[n] (p = predecessor index inside the read, -1 none).  Host<->device copies are inside the call."""
        n_reads = len(off) - 1
        n = int(off[-1])
        if f is None:
            f = np.empty(n, np.int32)
        if p is None:
            p = np.empty(n, np.int32)
        st = Stats()
        _ck(lib().mm2gb_chain_dp_host(self._h, _ptr(a), _ptr(off), n_reads, _ptr(f), _ptr(p), C.byref(st)))
        return f, p, st

    def chain(self, a, off, n_threads: int = 0, out=None, want_fp: bool = True, packed: bool = False):
        """Whole mg_lchain_dp for a batch.  n_threads <= 0: chain extraction + compaction on the device (k_bt_sort/k_bt_walk);
        n_threads >= 1: that stage on n_threads host threads.  Returns dict with f, p (None unless want_fp), u, n_u, b,
        n_b, stats; read r's chains are u[off[r]:off[r]+n_u[r]], its compacted anchors b[off[r]:off[r]+n_b[r]].
        packed=True (device stage only, no f/p): mm2gb_chain_host_index, the wire format of the results -- read r's compacted
        anchors are a[off[r] + v[v_pos[r]:v_pos[r]+n_b[r]]] (`v` int32 indices inside the read); with a pinned `v` the device
        writes exactly the bytes produced straight into it.  `out` may carry preallocated (e.g. pinned) buffers under the
        same keys."""
        n_reads = len(off) - 1
        n = int(off[-1])
        out = dict(out or {})
        if packed:
            out["f"] = out["p"] = None
            out.setdefault("u", np.empty(n, np.uint64)); out.setdefault("v", np.empty(n, np.int32))
            out.setdefault("n_u", np.zeros(n_reads, np.int32)); out.setdefault("n_b", np.zeros(n_reads, np.int64))
            out.setdefault("v_pos", np.zeros(n_reads, np.int64))
            st = Stats()
            _ck(lib().mm2gb_chain_host_index(self._h, _ptr(a), _ptr(off), n_reads, _ptr(out["u"]), _ptr(out["n_u"]), _ptr(out["v"]),
                                             _ptr(out["v_pos"]), _ptr(out["n_b"]), C.byref(st)))
            out["stats"] = st
            out["h2d_anchor_bytes"] = int(lib().mm2gb_last_batch_upload_bytes(self._h))
            return out
        if want_fp or n_threads >= 1:
            out.setdefault("f", np.empty(n, np.int32)); out.setdefault("p", np.empty(n, np.int32))
        else:
            out["f"] = out["p"] = None
        out.setdefault("u", np.empty(n, np.uint64)); out.setdefault("b", np.empty((n, 2), np.uint64))
        out.setdefault("n_u", np.zeros(n_reads, np.int32)); out.setdefault("n_b", np.zeros(n_reads, np.int64))
        st = Stats()
        _ck(lib().mm2gb_chain_host(self._h, _ptr(a), _ptr(off), n_reads, _ptr(out["f"]), _ptr(out["p"]), _ptr(out["u"]),
                                   _ptr(out["n_u"]), _ptr(out["b"]), _ptr(out["n_b"]), n_threads, C.byref(st)))
        out["stats"] = st
        return out

    def upload_bytes(self) -> int:
        """anchor bytes the last pipelined batch moved host -> device (16 B/anchor raw, ~8 B/anchor in the packed wire format)"""
        return int(lib().mm2gb_last_batch_upload_bytes(self._h))

    def submit(self, slot: int, a, off):
        _ck(lib().mm2gb_submit(self._h, slot, _ptr(a), _ptr(off), len(off) - 1))

    def wait(self, slot: int):
        f, p, off = C.c_void_p(), C.c_void_p(), C.c_void_p()
        st = Stats()
        _ck(lib().mm2gb_wait(self._h, slot, C.byref(f), C.byref(p), C.byref(off), C.byref(st)))
        n = st.n_anchors
        fa = np.ctypeslib.as_array(C.cast(f, C.POINTER(C.c_int32)), (n,)) if n else np.zeros(0, np.int32)
        pa = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), (n,)) if n else np.zeros(0, np.int32)
        return fa, pa, st

    def backtrack_device(self, a, off, f, p):
        """Chain extraction + compaction on the device for given f / p.  Returns (u, n_u, b, n_b, n_declined) in the layout
        of chain()."""
        a = np.ascontiguousarray(a, np.uint64); off = np.ascontiguousarray(off, np.int64)
        f = np.ascontiguousarray(f, np.int32); p = np.ascontiguousarray(p, np.int32)
        n_reads, n = len(off) - 1, int(off[-1])
        u = np.zeros(max(n, 1), np.uint64); b = np.zeros((max(n, 1), 2), np.uint64)
        n_u = np.zeros(max(n_reads, 1), np.int32); n_b = np.zeros(max(n_reads, 1), np.int64)
        nd = C.c_int32(0)
        _ck(lib().mm2gb_backtrack_device(self._h, _ptr(a), _ptr(off), n_reads, _ptr(f), _ptr(p), _ptr(u), _ptr(n_u), _ptr(b), _ptr(n_b), C.byref(nd)))
        return u, n_u[:n_reads], b, n_b[:n_reads], nd.value

    # ---- device-resident path ------------------------------------------------------------------------------
    def chain_dp_device(self, d_a, d_off, n_reads: int, n_total: int, d_f, d_p):
        """Enqueue the kernels on slot 0's stream; all arguments are device tensors/pointers.  Asynchronous."""
        _ck(lib().mm2gb_chain_dp_device(self._h, _ptr(d_a), _ptr(d_off), n_reads, n_total, _ptr(d_f), _ptr(d_p)))

    def chain_device(self, d_a, d_off, off, n_reads: int, n_total: int, d_f, d_p, slot: int = 0):
        """chain_dp_device + chain extraction on the device (off = host copy of the offsets), on slot `slot`.  Asynchronous."""
        _ck(lib().mm2gb_chain_device_slot(self._h, slot, _ptr(d_a), _ptr(d_off), _ptr(off), n_reads, n_total, _ptr(d_f), _ptr(d_p)))

    def sync(self, slot: int = 0):
        _ck(lib().mm2gb_sync(self._h, slot))

    def stream_ptr(self, slot: int = 0) -> int:
        return int(lib().mm2gb_stream(self._h, slot) or 0)

    def device_stats(self) -> Stats:
        st = Stats()
        _ck(lib().mm2gb_device_stats(self._h, C.byref(st)))
        return st

    def profile(self, enable: bool = True):
        _ck(lib().mm2gb_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        ms = (C.c_float * N_TIMERS)()
        n = (C.c_int64 * N_TIMERS)()
        _ck(lib().mm2gb_profile_read(self._h, ms, n))
        return {TIMER_NAMES[i]: (float(ms[i]), int(n[i])) for i in range(N_TIMERS)}


def gather_anchors(a, v):
    """compact_a's gather (lchain.c:100-105): b[k] = a[v[k]] through the library's own routine (what the drop-in runs)."""
    a = np.ascontiguousarray(a, np.uint64)
    v = np.ascontiguousarray(v, np.int32)
    b = np.empty((len(v), 2), np.uint64)
    if len(v):
        lib().mm2gb_gather_anchors(a.ctypes.data, v.ctypes.data, len(v), b.ctypes.data)
    return b


def wire_pack(a, off, cap_bytes=None):
    """The packed upload format (csrc/wire.h) of a batch, packed on the host exactly as the upload path does.
    Returns (buffer uint8[cap], bytes_to_upload, n_runs); bytes_to_upload == -1 if the run list does not fit cap_bytes."""
    a = np.ascontiguousarray(a, np.uint64)
    off = np.ascontiguousarray(off, np.int64)
    n = int(off[-1])
    cap = int(cap_bytes if cap_bytes is not None else 16 * max(n, 1) + 4096)
    raw = np.zeros(cap + 64, np.uint8)
    shift = (-raw.ctypes.data) % 32
    buf = raw[shift:shift + cap]
    nr = C.c_int32(0)
    nbytes = lib().mm2gb_wire_pack(a.ctypes.data, off.ctypes.data, len(off) - 1, buf.ctypes.data, cap, C.byref(nr))
    return buf, int(nbytes), int(nr.value)


def wire_unpack(buf, n, n_runs):
    out = np.zeros((max(n, 1), 2), np.uint64)
    _ck(lib().mm2gb_wire_unpack(buf.ctypes.data, len(buf), n, n_runs, out.ctypes.data))
    return out[:n]


def device_memory(device: int = 0):
    f, t = C.c_size_t(0), C.c_size_t(0)
    _ck(lib().mm2gb_device_memory(device, C.byref(f), C.byref(t)))
    return f.value, t.value


def backtrack(misc: Misc, a, f, p):
    """Host stage (lchain.c:27-111) on one read: returns (u uint64[n_u], b uint64[n_b,2])."""
    a = np.ascontiguousarray(a, np.uint64)
    f = np.ascontiguousarray(f, np.int32)
    p = np.ascontiguousarray(p, np.int32)
    n = a.shape[0]
    u = np.empty(max(n, 1), np.uint64)
    b = np.empty((max(n, 1), 2), np.uint64)
    nb = C.c_int64(0)
    max_drop = INT32_MAX if misc.is_cdna else misc.bw
    n_u = lib().mm2gb_backtrack(n, f.ctypes.data, p.ctypes.data, a.ctypes.data, misc.min_cnt, misc.min_score, max_drop,
                                u.ctypes.data, b.ctypes.data, C.byref(nb))
    return u[:n_u].copy(), b[:nb.value].copy()
