"""Probe of the x-sort's process-to-process bimodality (profiles/r8f_sort_streams.txt): one process, one index, several seeders created
one after the other (with other allocations in between, so that their buffers land elsewhere), three seeding passes each; prints the
sort stage time of every pass and the SM clock NVML reports right after it."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=300)
ap.add_argument("--lo", type=int, default=100000)
ap.add_argument("--hi", type=int, default=300000)
ap.add_argument("--err", type=float, default=0.02)
ap.add_argument("--repeats", type=int, default=2400)
ap.add_argument("--seeders", type=int, default=5)
args = ap.parse_args()
pkg = entry.load_package()
import torch  # noqa: E402
import pynvml  # noqa: E402
from mm2gb_b200 import seed, synth  # noqa: E402
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
ref = synth.simulate_reference(20_000_000, seed=1, n_repeat_copies=args.repeats, repeat_unit=3000)
reads = synth.simulate_reads(ref, args.reads, args.lo, args.hi, seed=2, err=args.err)
off = np.zeros(len(reads) + 1, dtype=np.int64)
off[1:] = np.cumsum([len(r) for r in reads])
buf = torch.from_numpy(synth._NT[np.concatenate(reads)]).pin_memory().numpy()
ix = seed.Index((synth._NT[ref], np.array([0, len(ref)], dtype=np.int64)), w=10, k=15)
prm = seed.map_ont_seed_params(ix.mid_occ())
pad = []
for s in range(args.seeders):
    sd = seed.Seeder(ix, max_bases=int(off[-1]) + 4096, max_reads=len(reads) + 8, max_anchors=int(off[-1]))
    out = []
    for _ in range(3):
        a, a_off, rep, _, _ = sd.seed(prm, buf, off, want_mini_pos=False)
        out.append((round(sd.profile()[0]["sort"], 2), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
    print("seeder", s, "anchors", int(a_off[-1]), "max per read", int(np.diff(a_off).max()), "sort ms / sm MHz:", out, flush=True)
    sd.close()
    pad.append(torch.empty((97 + 61 * s) << 20, dtype=torch.uint8, device="cuda"))   # shifts where the next seeder's buffers land
