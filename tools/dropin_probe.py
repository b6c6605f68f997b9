#!/usr/bin/env python
"""End-to-end time of the plugin boundary (tests/fake_host.c: fake_drive = the call pattern of `minimap2 -t T --gpu-chain`)
for several thread counts / batch limits on a bench workload.  One JSON line per configuration."""
import ctypes as C, json, os, sys, time
import numpy as np
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import bench


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "ont"
    pkg = entry.load_package()
    a, off = bench.make_workload(bench.WORKLOADS[wl], 0)
    n, n_reads = int(off[-1]), len(off) - 1
    D = C.CDLL(os.path.join(ROOT, "tests", "_build", "libdropin_test.so"))
    D.init_stream_gpu.argtypes = [C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, pkg.Misc]
    D.fake_set_misc.argtypes = [C.POINTER(pkg.Misc)]
    D.free_stream_gpu.argtypes = [C.c_int]
    D.fake_drive.restype = C.c_double
    D.fake_drive.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int] + [C.c_void_p] * 4
    misc = pkg.map_ont_misc()
    D.fake_set_misc(C.byref(misc))
    cfgs = [(int(t), int(b), int(s)) for t, b, s in (x.split(":") for x in (sys.argv[2:] or ["16:16777216:1", "16:2097152:1", "16:1048576:1", "16:524288:1", "16:2097152:0", "8:2097152:1", "4:2097152:1", "1:4194304:1"]))]
    for T, batch, sync in cfgs:
        os.environ["MM2GB_THREADS_PER_GPU"] = str(T)
        mx, mr, mn = C.c_size_t(0), C.c_int(0), C.c_int(-1)
        D.init_stream_gpu(C.byref(mx), C.byref(mr), C.byref(mn), b"", misc)
        D.fake_drive(a.ctypes.data, off.ctypes.data, n_reads, T, 0, 2, batch, sync, None, None, None, None)
        steps = 5
        dt = D.fake_drive(a.ctypes.data, off.ctypes.data, n_reads, T, 0, steps, batch, sync, None, None, None, None)
        D.free_stream_gpu(T)
        print(json.dumps({"workload": wl, "threads": T, "batch_anchors": batch, "sync_steps": sync, "ms_per_step": 1e3 * dt / steps, "anchors": n}), flush=True)


if __name__ == "__main__":
    main()
