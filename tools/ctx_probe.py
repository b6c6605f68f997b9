#!/usr/bin/env python
"""How long the chaining contexts of T driver threads take to create (each thread creates its own, concurrently, the way the
drop-in does on the first batch of every thread).  MM2GB_VERBOSE=3 prints the phases of every creation.
    python tools/ctx_probe.py [threads] [slot_anchors] [slots]"""
import ctypes as C, json, os, sys, threading, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    cap = int(sys.argv[2]) if len(sys.argv) > 2 else 820_000
    slots = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    pkg = entry.load_package()
    L = pkg.lib()
    misc = pkg.map_ont_misc()
    t00 = time.perf_counter()
    L.mm2gb_device_count()
    t_init = time.perf_counter() - t00
    out = [None] * T
    ctxs = [C.c_void_p() for _ in range(T)]

    def work(t):
        t0 = time.perf_counter()
        rc = L.mm2gb_ctx_create_ex(C.byref(ctxs[t]), 0, cap, 200001, slots, C.byref(misc), 4)   # MM2GB_CTX_NO_FP_STAGING
        out[t] = (rc, time.perf_counter() - t0)
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    for x in th: x.start()
    for x in th: x.join()
    wall = time.perf_counter() - t0
    t1 = time.perf_counter()
    for c in ctxs:
        if c: L.mm2gb_ctx_destroy(c)
    print(json.dumps({"threads": T, "slot_anchors": cap, "slots": slots, "device_count_call_s": round(t_init, 3), "wall_s": round(wall, 3),
                      "per_thread_s": [round(o[1], 3) for o in out], "rc": [o[0] for o in out], "destroy_s": round(time.perf_counter() - t1, 3)}))


if __name__ == "__main__":
    main()
