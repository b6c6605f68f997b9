#!/usr/bin/env python
"""Generate the known-answer vectors under tests/golden/ from the REFERENCE itself.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference, i.e. `make -C oracle ref`):

    python oracle/gen_golden.py

What it writes (all small, committed):
  tests/golden/fixtures.npz    anchors seeded by the reference driver (oracle/_ref/minimap2_ref --gpu-chain with
                               MM2GB_DUMP, see dump_stub.c) for the reference's own test FASTA pairs
                               (test/MT-human.fa x MT-orang.fa, t-inv.fa x q-inv.fa, t2.fa x q2.fa) and the f / p / u /
                               compacted-anchor outputs of the reference's mg_lchain_dp (oracle/_ref/libref_lchain.so)
                               at max_chain_skip = INT32_MAX (true infinity, SURVEY.md trap T1) and = 25.
  tests/golden/synth_reads.npz the same for simulated ONT-like reads on a random reference with planted repeats.
  tests/golden/adversarial.json sha256 of the reference outputs for mm2-gb_b200/synth.py:adversarial_suite()
                               (inputs are regenerated from the seed, only digests are stored).
  tests/golden/*.paf           PAF ground truth of the CPU driver for the same inputs
                               (`minimap2_ref -t 1 --max-chain-skip=2147483647`).
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "mm2-gb_b200"))
import pyoracle as po  # noqa: E402
import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MM2 = os.path.join(HERE, "_ref", "minimap2_ref")
FIX = os.path.join(HERE, "_ref", "fixtures")


def run_driver(ref_fa, qry_fa, extra=(), dump=None, gpu=False):
    env = dict(os.environ)
    cmd = [MM2, "-t", "1", "--max-chain-skip=2147483647", *extra]
    if gpu:
        cmd.append("--gpu-chain")
    if dump:
        env["MM2GB_DUMP"] = dump
    out = subprocess.run(cmd + [ref_fa, qry_fa], check=True, capture_output=True, env=env)
    return out.stdout.decode()


def pack(reads, prm_inf, prm_25):
    """reads: list of anchor arrays -> dict of concatenated arrays with reference outputs."""
    d = {"a": [], "off": [0], "f": [], "p": [], "u": [], "u_off": [0], "b": [], "b_off": [0], "f25": [], "p25": []}
    for a in reads:
        r = po.ref_lchain(prm_inf, a)
        r25 = po.ref_lchain(prm_25, a)
        o = po.oracle_lchain(prm_inf, a)
        assert o.same(r), "restatement differs from the reference"
        d["a"].append(a); d["off"].append(d["off"][-1] + len(a))
        d["f"].append(r.f); d["p"].append(r.p.astype(np.int32))
        d["u"].append(r.u); d["u_off"].append(d["u_off"][-1] + len(r.u))
        d["b"].append(r.b); d["b_off"].append(d["b_off"][-1] + len(r.b))
        d["f25"].append(r25.f); d["p25"].append(r25.p.astype(np.int32))
    cat = lambda k, dt, shp: (np.concatenate(d[k]) if d[k] else np.zeros(shp, dt)).astype(dt)
    return dict(a=cat("a", np.uint64, (0, 2)), off=np.array(d["off"], np.int64), f=cat("f", np.int32, (0,)),
                p=cat("p", np.int32, (0,)), u=cat("u", np.uint64, (0,)), u_off=np.array(d["u_off"], np.int64),
                b=cat("b", np.uint64, (0, 2)), b_off=np.array(d["b_off"], np.int64), f25=cat("f25", np.int32, (0,)),
                p25=cat("p25", np.int32, (0,)))


def digest(res: po.ChainResult) -> str:
    h = hashlib.sha256()
    for arr in (res.f, res.p, res.u, res.b):
        h.update(np.ascontiguousarray(arr).tobytes())
    return h.hexdigest()


def main():
    assert po.ref_available() and os.path.exists(MM2), "run `make -C oracle ref` first (needs /root/reference)"
    os.makedirs(GOLD, exist_ok=True)
    prm_inf, prm_25 = po.map_ont_params(), po.map_ont_params(max_skip=25)

    # (i) the reference's own fixtures
    reads, names = [], []
    with tempfile.TemporaryDirectory() as tmp:
        for tag, (t, q) in {"MT": ("MT-human.fa", "MT-orang.fa"), "inv": ("t-inv.fa", "q-inv.fa"), "t2": ("t2.fa", "q2.fa")}.items():
            dump = os.path.join(tmp, tag + ".dump")
            paf_gpu = run_driver(os.path.join(FIX, t), os.path.join(FIX, q), dump=dump, gpu=True)
            paf_cpu = run_driver(os.path.join(FIX, t), os.path.join(FIX, q))
            assert paf_gpu == paf_cpu
            with open(os.path.join(GOLD, tag + ".paf"), "w") as fh:
                fh.write(paf_cpu)
            prm, rs = po.read_dump(dump)
            assert prm.as_dict() == prm_inf.as_dict(), prm.as_dict()
            for i, (_, _, a) in enumerate(rs):
                reads.append(a); names.append(f"{tag}:{i}")
    np.savez_compressed(os.path.join(GOLD, "fixtures.npz"), names=np.array(names), **pack(reads, prm_inf, prm_25))
    print("fixtures:", names, [len(r) for r in reads])

    # (ii) simulated ONT-like reads through the reference's own seeding
    ref = synth.simulate_reference(600_000, seed=1, n_repeat_copies=40, repeat_unit=2000)
    rds = synth.simulate_reads(ref, 8, 4000, 16000, seed=2)
    with tempfile.TemporaryDirectory() as tmp:
        # the FASTA inputs are regenerated from the seeds by whoever needs them (tests/test_driver.py); not committed
        synth.write_fasta(os.path.join(tmp, "synth_ref.fa"), [ref], prefix="ref")
        synth.write_fasta(os.path.join(tmp, "synth_reads.fa"), rds, prefix="read")
        dump = os.path.join(tmp, "s.dump")
        args = (os.path.join(tmp, "synth_ref.fa"), os.path.join(tmp, "synth_reads.fa"))
        paf_gpu = run_driver(*args, extra=("-x", "map-ont"), dump=dump, gpu=True)
        paf_cpu = run_driver(*args, extra=("-x", "map-ont"))
        assert paf_gpu == paf_cpu
        with open(os.path.join(GOLD, "synth.paf"), "w") as fh:
            fh.write(paf_cpu)
        prm, rs = po.read_dump(dump)
    sreads = [a for _, _, a in rs]
    np.savez_compressed(os.path.join(GOLD, "synth_reads.npz"), **pack(sreads, prm_inf, prm_25))
    print("synthetic reads:", [len(r) for r in sreads])

    # (iii) adversarial suite: digests only
    dig = {}
    for name, (a, over) in synth.adversarial_suite().items():
        prm = po.map_ont_params(**over)
        r = po.ref_lchain(prm, a)
        assert po.oracle_lchain(prm, a).same(r), name
        dig[name] = {"n": int(len(a)), "n_u": int(len(r.u)), "sha256": digest(r),
                     "f_last": int(r.f[-1]) if len(a) else None, "p_last": int(r.p[-1]) if len(a) else None}
    with open(os.path.join(GOLD, "adversarial.json"), "w") as fh:
        json.dump(dig, fh, indent=1, sort_keys=True)
    print("adversarial:", {k: v["n"] for k, v in dig.items()})


if __name__ == "__main__":
    main()
