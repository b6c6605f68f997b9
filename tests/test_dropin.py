"""The drop-in boundary (include/mm2gb_plchain.h): the four entry points minimap2's driver calls, exercised through ctypes
with a fake host (tests/fake_host.c) that supplies kmalloc/kfree/build_misc/post_chaining_helper."""
import ctypes as C
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "_build", "libdropin_test.so")


class SeqMeta(C.Structure):   # gpu/plutils.h:19-31
    _fields_ = [("i", C.c_long), ("seg_id", C.c_int), ("name", C.c_char * 200), ("len", C.c_uint32), ("n_alt", C.c_int),
                ("is_alt", C.c_int), ("qlen_sum", C.c_int)]


class ChainRead(C.Structure):  # gpu/plutils.h:45-73, release layout
    _fields_ = [("seq", SeqMeta), ("qseqs", C.c_void_p), ("qlens", C.c_void_p), ("n_seg", C.c_int), ("rep_len", C.c_int),
                ("frag_gap", C.c_int), ("mini_pos", C.c_void_p), ("n_mini_pos", C.c_int), ("a", C.c_void_p), ("n", C.c_int64),
                ("u", C.c_void_p), ("n_u", C.c_int)]


def test_struct_layout_matches_reference_probe():
    """sizeof(chain_read_t)=312, a@280 n@288 u@296 n_u@304 rep_len@252 frag_gap@256"""
    assert C.sizeof(ChainRead) == 312
    assert (ChainRead.a.offset, ChainRead.n.offset, ChainRead.u.offset, ChainRead.n_u.offset) == (280, 288, 296, 304)
    assert (ChainRead.rep_len.offset, ChainRead.frag_gap.offset) == (252, 256)


def test_dropin_exports_reference_symbols():
    """no compute: the boundary library loads and exports exactly the reference's four symbols (plutils.h:98-104)"""
    assert os.path.exists(SO), "run __graft_entry__.build()"
    L = C.CDLL(SO)
    for name in ("init_stream_gpu", "chain_stream_gpu", "finish_stream_gpu", "free_stream_gpu"):
        assert hasattr(L, name), name
    L.free_stream_gpu(1)   # never initialised: must be a no-op (main.c:466 calls it unconditionally)


def test_gpu_config_is_parsed_as_json():
    """--gpu-cfg is real JSON (the reference parses it with cJSON, gpu/plmem.cu:373-451): only numeric members of the TOP-LEVEL
    object count; same-named keys inside nested objects, comment keys with string values and strings are skipped"""
    L = C.CDLL(SO)
    L.mm2gb_dropin_parse_config_key.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_double)]

    def get(text, key):
        v = C.c_double(-1.0)
        rc = L.mm2gb_dropin_parse_config_key(text.encode(), key.encode(), C.byref(v))
        return rc, v.value

    ref_cfg = """{
        "num_streams": 1, "//num_streams": "Must set to 1", "min_n": 512,
        "//min_n": "queries with less anchors will be handled on cpu: \\"max_total_n\\": 7",
        "max_total_n": 500000000, "max_read": 500000,
        "range_kernel": {"blockdim": 512, "max_total_n": 1, "nested": {"max_read": 2, "list": [1, 2, {"min_n": 3}]}},
        "score_kernel": {"micro_batch": 4, "//micro_batch": "text", "mid_blockdim": 512},
        "flags": [true, false, null, -1.5e3], "n_gpus": 2
    }"""
    assert get(ref_cfg, "max_total_n") == (1, 500000000.0)
    assert get(ref_cfg, "max_read") == (1, 500000.0)
    assert get(ref_cfg, "min_n") == (1, 512.0)
    assert get(ref_cfg, "n_gpus") == (1, 2.0)
    assert get(ref_cfg, "blockdim")[0] == 0 and get(ref_cfg, "micro_batch")[0] == 0      # nested: not ours
    assert get(ref_cfg, "//min_n")[0] == 0                                                # a string, not a number
    assert get("{}", "max_total_n")[0] == 0
    for bad in ("{", '{"a": }', '{"a": 1,}', '{"a" 1}', '{"a": 1} x', '["max_total_n", 5'):
        assert get(bad, "a")[0] == -1, bad
    # the reference's own config file format, as shipped (copied as a fixture string: keys with "//" comments, nested kernels)
    here = os.path.join(ROOT, "mm2-gb_b200", "b200_config.json")
    assert get(open(here).read(), "max_total_n") == (1, 4194304.0)


def _lib(pkg):
    L = C.CDLL(SO)
    L.init_stream_gpu.argtypes = [C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, pkg.Misc]
    L.chain_stream_gpu.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(ChainRead)), C.POINTER(C.c_int), C.c_int, C.c_void_p]
    L.finish_stream_gpu.argtypes = L.chain_stream_gpu.argtypes
    L.free_stream_gpu.argtypes = [C.c_int]
    L.fake_set_misc.argtypes = [C.POINTER(pkg.Misc)]
    L.fake_live_blocks.restype = C.c_long
    L.fake_helper_calls.restype = C.c_long
    L.fake_alloc_anchors.restype = C.c_void_p
    L.fake_alloc_anchors.argtypes = [C.c_void_p, C.c_int64]
    L.kfree.argtypes = [C.c_void_p, C.c_void_p]
    return L


def _make_batch(L, a, off):
    n = len(off) - 1
    arr = (ChainRead * max(n, 1))()
    for r in range(n):
        s, e = int(off[r]), int(off[r + 1])
        arr[r].n = e - s
        arr[r].n_seg = 1
        arr[r].seq.i = r
        arr[r].a = L.fake_alloc_anchors(a[s:e].ctypes.data, e - s) if e > s else None
    return arr


def _check_and_free(L, po, pkg, arr, a, off, prm):
    for r in range(len(off) - 1):
        s, e = int(off[r]), int(off[r + 1])
        o = po.oracle_lchain(prm, a[s:e])
        rd = arr[r]
        assert rd.n_u == len(o.u), (r, rd.n_u, len(o.u))
        assert rd.frag_gap == prm.max_dist_x
        if rd.n_u:
            u = np.ctypeslib.as_array(C.cast(rd.u, C.POINTER(C.c_uint64)), (rd.n_u,))
            nb = int((u & np.uint64(0xffffffff)).sum())
            b = np.ctypeslib.as_array(C.cast(rd.a, C.POINTER(C.c_uint64)), (nb * 2,)).reshape(nb, 2)
            assert np.array_equal(u, o.u) and np.array_equal(b, o.b), r
        else:
            assert not rd.a and not rd.u
        L.kfree(None, rd.a)
        L.kfree(None, rd.u)


@pytest.mark.gpu
def test_batch_handoff_protocol(pkg, po, synth, tmp_path):
    """init -> chain (returns NULL) -> chain (returns batch 1) -> chain(empty) -> finish (drains) -> free, two thread ids
    interleaved; results == oracle's whole mg_lchain_dp; every arena block is accounted for."""
    L = _lib(pkg)
    cfg = tmp_path / "cfg.json"
    cfg.write_text('{"num_streams": 1, "min_n": 512, "max_total_n": 300000, "max_read": 500, "host_threads": 4,\n'
                   ' "range_kernel": {"blockdim": 512, "max_total_n": 5}, "score_kernel": {"micro_batch": 4}}')
    os.environ["MM2GB_SUB_MIN"] = "20000"     # small batches are cut into sub-batches too (default: 512 k anchors and 64 reads each at least)
    os.environ["MM2GB_SUB_MIN_READS"] = "2"
    misc = pkg.map_ont_misc()
    prm = po.map_ont_params()
    L.fake_set_misc(C.byref(misc))
    mx, mr, mn = C.c_size_t(0), C.c_int(0), C.c_int(-1)
    L.init_stream_gpu(C.byref(mx), C.byref(mr), C.byref(mn), str(cfg).encode(), misc)
    assert (mx.value, mr.value, mn.value) == (300000, 500, 512)
    batches = {}
    for tid in (0, 1):
        batches[tid] = []
        for k in range(3):
            a, off = synth.ont_like_batch(100 + 10 * tid + k, 12, 1, 1500)
            batches[tid].append((a, off, _make_batch(L, a, off)))
        # a batch larger than max_total_n (the context must grow, never hand reads back for CPU chaining)
        a, off = synth.ont_like_batch(200 + tid, 90, 3000, 5000)
        assert off[-1] > 300000
        batches[tid].append((a, off, _make_batch(L, a, off)))
    returned = {0: [], 1: []}
    for k in range(4):
        for tid in (0, 1):
            a, off, arr = batches[tid][k]
            ptr = C.cast(arr, C.POINTER(ChainRead))
            n = C.c_int(len(off) - 1)
            L.chain_stream_gpu(None, None, C.byref(ptr), C.byref(n), tid, None)
            if k == 0:
                assert not ptr and n.value == 0            # first call: nothing to hand back (plchain.cu:293-305)
            else:
                assert C.addressof(ptr.contents) == C.addressof(batches[tid][k - 1][2]) and n.value == len(batches[tid][k - 1][1]) - 1
                returned[tid].append(k - 1)
    for tid in (0, 1):
        ptr, n = C.POINTER(ChainRead)(), C.c_int(-1)
        L.finish_stream_gpu(None, None, C.byref(ptr), C.byref(n), tid, None)
        assert C.addressof(ptr.contents) == C.addressof(batches[tid][3][2]) and n.value == len(batches[tid][3][1]) - 1
        L.finish_stream_gpu(None, None, C.byref(ptr), C.byref(n), tid, None)
        assert not ptr and n.value == 0                     # idle (plchain.cu:524-528)
    n_reads = sum(len(off) - 1 for tid in (0, 1) for _, off, _ in batches[tid])
    assert L.fake_helper_calls() == n_reads
    for tid in (0, 1):
        for a, off, arr in batches[tid]:
            _check_and_free(L, po, pkg, arr, a, off, prm)
    assert L.fake_live_blocks() == 0                        # no arena leak, no double free
    L.free_stream_gpu(2)
    L.free_stream_gpu(2)
    del os.environ["MM2GB_SUB_MIN"]
    del os.environ["MM2GB_SUB_MIN_READS"]


def _digest(w):
    """fake_digest of tests/fake_host.c"""
    w = np.ascontiguousarray(w, np.uint64).reshape(-1)
    k = np.arange(len(w), dtype=np.uint64)
    c = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        return int((c * np.uint64(len(w) + 1) + (w * ((np.uint64(2) * k + np.uint64(1)) * c)).sum(dtype=np.uint64)) & np.uint64(0xFFFFFFFFFFFFFFFF))


@pytest.mark.gpu
@pytest.mark.parametrize("n_threads,batch_anchors,sync", [(1, 1 << 30, 1), (3, 150000, 1), (4, 60000, 0)])
def test_driver_call_pattern_many_threads(pkg, po, synth, tmp_path, n_threads, batch_anchors, sync):
    """fake_drive = the call pattern of `minimap2 -t T --gpu-chain` (what bench.py's end-to-end leg times): T worker threads,
    per-read kmalloc'd anchor arrays, chain_stream_gpu per batch, finish_stream_gpu per mini-batch.  Every read's chains and
    compacted anchors (digests) equal the oracle's whole mg_lchain_dp."""
    L = _lib(pkg)
    L.fake_drive.restype = C.c_double
    L.fake_drive.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_void_p]
    cfg = tmp_path / "cfg.json"
    cfg.write_text('{"max_total_n": 400000, "max_read": 5000}')
    os.environ["MM2GB_SUB_MIN"] = "30000"
    os.environ["MM2GB_SUB_MIN_READS"] = "2"
    misc = pkg.map_ont_misc()
    prm = po.map_ont_params()
    L.fake_set_misc(C.byref(misc))
    mx, mr, mn = C.c_size_t(0), C.c_int(0), C.c_int(-1)
    L.init_stream_gpu(C.byref(mx), C.byref(mr), C.byref(mn), str(cfg).encode(), misc)
    a, off = synth.ont_like_batch(300 + n_threads, 150, 1, 4000, repeat_copies=2, repeat_len=40)
    n_reads = len(off) - 1
    nu = np.zeros(n_reads, np.int32); nb = np.zeros(n_reads, np.int64)
    hu = np.zeros(n_reads, np.uint64); hb = np.zeros(n_reads, np.uint64)
    helper0, live0 = L.fake_helper_calls(), L.fake_live_blocks()
    dt = L.fake_drive(a.ctypes.data, off.ctypes.data, n_reads, n_threads, 0, 3, batch_anchors, sync, nu.ctypes.data, nb.ctypes.data,
                      hu.ctypes.data, hb.ctypes.data)
    assert dt > 0
    assert L.fake_helper_calls() - helper0 == 3 * n_reads and L.fake_live_blocks() == live0
    for r in range(n_reads):
        o = po.oracle_lchain(prm, a[int(off[r]):int(off[r + 1])])
        assert nu[r] == len(o.u) and nb[r] == len(o.b), r
        assert int(hu[r]) == _digest(o.u) and int(hb[r]) == _digest(o.b), r
    L.free_stream_gpu(n_threads)
    del os.environ["MM2GB_SUB_MIN"]
    del os.environ["MM2GB_SUB_MIN_READS"]


@pytest.mark.gpu
def test_threads_spread_over_gpus(pkg, po, synth, tmp_path):
    """thread_id t drives GPU t % n_gpus (all visible GPUs by default).  Six worker threads on a box with at least two
    GPUs: every GPU chains batches and every read's result equals the oracle's.  Skipped on a one-GPU box."""
    n_gpus = pkg.lib().mm2gb_device_count()
    if n_gpus < 2:
        pytest.skip("needs at least two GPUs")
    L = _lib(pkg)
    L.fake_drive.restype = C.c_double
    L.fake_drive.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_void_p]
    L.mm2gb_dropin_device_batches.restype = C.c_longlong
    L.mm2gb_dropin_device_batches.argtypes = [C.c_int]
    cfg = tmp_path / "cfg.json"
    cfg.write_text('{"max_total_n": 400000, "max_read": 5000, "n_gpus": 0}')
    for k in ("MM2GB_N_GPUS", "MM2GB_GPU_BASE"):
        os.environ.pop(k, None)
    misc = pkg.map_ont_misc()
    prm = po.map_ont_params()
    L.fake_set_misc(C.byref(misc))
    mx, mr, mn = C.c_size_t(0), C.c_int(0), C.c_int(-1)
    L.init_stream_gpu(C.byref(mx), C.byref(mr), C.byref(mn), str(cfg).encode(), misc)
    n_threads = 6
    a, off = synth.ont_like_batch(411, 240, 1, 4000, repeat_copies=2, repeat_len=40)
    n_reads = len(off) - 1
    nu = np.zeros(n_reads, np.int32); nb = np.zeros(n_reads, np.int64)
    hu = np.zeros(n_reads, np.uint64); hb = np.zeros(n_reads, np.uint64)
    dt = L.fake_drive(a.ctypes.data, off.ctypes.data, n_reads, n_threads, 0, 2, 60000, 1, nu.ctypes.data, nb.ctypes.data, hu.ctypes.data, hb.ctypes.data)
    assert dt > 0
    used = [L.mm2gb_dropin_device_batches(d) for d in range(n_gpus)]
    assert all(u > 0 for u in used[:min(n_gpus, n_threads)]), used
    for r in range(n_reads):
        o = po.oracle_lchain(prm, a[int(off[r]):int(off[r + 1])])
        assert nu[r] == len(o.u) and nb[r] == len(o.b), r
        assert int(hu[r]) == _digest(o.u) and int(hb[r]) == _digest(o.b), r
    L.free_stream_gpu(n_threads)


_GUARD = """
import ctypes as C, os, sys
sys.path.insert(0, {root!r})
import numpy as np
import __graft_entry__ as entry
sys.path.insert(0, os.path.join({root!r}, "tests"))
import test_dropin as T
pkg = entry.load_package()
L = T._lib(pkg)
L.fake_set_misc_depends_on_qlen.argtypes = [C.c_int]
misc = pkg.map_ont_misc()
L.fake_set_misc(C.byref(misc))
mx, mr, mn = C.c_size_t(0), C.c_int(0), C.c_int(-1)
L.init_stream_gpu(C.byref(mx), C.byref(mr), C.byref(mn), b"", misc)
from mm2gb_b200 import synth
a, off = synth.ont_like_batch(1, 3, 50, 200)
arr = T._make_batch(L, a, off)
if {mode!r} == "qlen":
    L.fake_set_misc_depends_on_qlen(1)
else:
    arr[1].n_seg = 2
ptr = C.cast(arr, C.POINTER(T.ChainRead)); n = C.c_int(3)
L.chain_stream_gpu(None, None, C.byref(ptr), C.byref(n), 0, None)
print("NOT REACHED")
"""


@pytest.mark.gpu
@pytest.mark.parametrize("mode,msg", [("qlen", "depend on the read length"), ("nseg", "single-segment reads only")])
def test_unsupported_modes_fail_loudly(mode, msg):
    """plchain.cu:497 guards the single-segment assumption with an assert that release builds drop; here a preset whose
    chaining parameters depend on the read (short-read / fragment mode) or a multi-segment read ends the run with a message"""
    out = subprocess.run([sys.executable, "-c", textwrap.dedent(_GUARD.format(root=ROOT, mode=mode))], capture_output=True, text=True, timeout=300)
    assert out.returncode == 1, out.stderr[-400:]
    assert msg in out.stderr and "NOT REACHED" not in out.stdout
