"""The drop-in boundary (include/mm2gb_plchain.h): the four entry points minimap2's driver calls, exercised through ctypes
with a fake host (tests/fake_host.c) that supplies kmalloc/kfree/build_misc/post_chaining_helper."""
import ctypes as C
import os

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
                   ' "range_kernel": {"blockdim": 512}, "score_kernel": {"micro_batch": 4}}')
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
