#!/usr/bin/env python
"""PCIe probe for the result path: k_drain (SM stores into mapped pinned memory) vs the copy engine, GB/s."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
L = pkg.lib()
L.mm2gb_debug_drain.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_float)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
with pkg.ChainContext(pkg.map_ont_misc(), max_anchors=n, max_reads=16, n_slots=1) as c:
    for blocks in (4, 8, 16, 32, 64, 128, 296, 592, 1184):
        ms = (C.c_float * 5)()
        rc = L.mm2gb_debug_drain(c._h, n, blocks, ms)
        assert rc == 0, L.mm2gb_last_error()
        gb = n * 16 / 1e9
        print(json.dumps({"anchors": n, "blocks": blocks, "drain_gbs": gb / (ms[0] / 1e3), "memcpy_gbs": gb / (ms[1] / 1e3),
                          "drain_with_h2d_ms": ms[2], "memcpy_with_h2d_ms": ms[3], "h2d_alone_ms": ms[4], "drain_ms": ms[0], "memcpy_ms": ms[1]}))
