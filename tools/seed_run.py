"""ncu target for the seeding kernels: index of a 20 Mb reference, then `--iters` seeding passes over 2000 reads (tools/seed_bench.py
workload at reduced size).  ncu -k regex:k_sketch -s 1 -c 1 skips the launch that sketches the reference."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ref-len", type=int, default=20_000_000)
ap.add_argument("--reads", type=int, default=2000)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--lo", type=int, default=10000)
ap.add_argument("--hi", type=int, default=100000)
ap.add_argument("--err", type=float, default=0.10)
ap.add_argument("--repeats", type=int, default=-1)
ap.add_argument("--pinned", action="store_true", help="sequences in pinned host memory (DMA as they are): the sketch stage is then bound by the kernel, not by staging")
args = ap.parse_args()
pkg = entry.load_package()
from mm2gb_b200 import seed, synth  # noqa: E402
ref = synth.simulate_reference(args.ref_len, seed=1, n_repeat_copies=(args.ref_len // 16667 if args.repeats < 0 else args.repeats), repeat_unit=3000)
reads = synth.simulate_reads(ref, args.reads, args.lo, args.hi, seed=2, err=args.err)
off = np.zeros(len(reads) + 1, dtype=np.int64)
off[1:] = np.cumsum([len(r) for r in reads])
buf = synth._NT[np.concatenate(reads)]
if args.pinned:
    import torch
    buf = torch.from_numpy(buf).pin_memory().numpy()
ix = seed.Index((synth._NT[ref], np.array([0, len(ref)], dtype=np.int64)), w=10, k=15)
prm = seed.map_ont_seed_params(ix.mid_occ())
sd = seed.Seeder(ix, max_bases=int(off[-1]) + 4096, max_reads=len(reads) + 8, max_anchors=int(off[-1]))
for _ in range(args.iters):
    a, a_off, rep, _, _ = sd.seed(prm, buf, off, want_mini_pos=False)
import hashlib
print("anchors", int(a_off[-1]), "sha1", hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16], "bases", int(off[-1]), sd.profile())
