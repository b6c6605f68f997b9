#!/usr/bin/env python
"""The reference's UNMODIFIED host driver on a BASELINE.json config, CPU path against --gpu-chain (the B200 drop-in):

    python tools/driver_run.py [ont|ont40k|long|mini] [threads] [gpu_threads]

generates the reference / read FASTA of the workload (numpy simulator of mm2-gb_b200/synth.py, fixed seeds), runs
  oracle/_ref/minimap2_ref_timed  -t T -x map-ont --max-chain-skip=2147483647        (CPU ground truth, SURVEY.md trap T1)
  oracle/_ref/minimap2_b200_timed -t G -x map-ont --gpu-chain --gpu-cfg b200_config.json
and prints one JSON line: PAF md5 of both, number of differing lines, wall seconds, and the per-thread stage seconds of the
driver's own timers (map.c:390,1087,630; switched back on at build time, see oracle/Makefile).  North-star bar: empty diff."""
import hashlib, json, os, re, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

WORK = {"ont": dict(ref_len=100_000_000, n_contigs=4, n_reads=10_000, lo=10_000, hi=100_000, err=0.10, repeats=0),
        "ont40k": dict(ref_len=100_000_000, n_contigs=4, n_reads=40_000, lo=10_000, hi=100_000, err=0.10, repeats=0),
        "long": dict(ref_len=100_000_000, n_contigs=4, n_reads=1000, lo=100_000, hi=300_000, err=0.03, repeats=3000),
        "mini": dict(ref_len=5_000_000, n_contigs=2, n_reads=300, lo=10_000, hi=100_000, err=0.10, repeats=0)}


def timers(stderr):
    out = {}
    for name in ("Seed", "Chain", "Align"):
        m = re.search(r"^%s\s*=\s*([\d.]+)\s+([\d.]+)\s+([\d.]+)" % name, stderr, re.M)
        if m:
            out[name.lower()] = {"min_s": float(m.group(1)), "max_s": float(m.group(2)), "avg_s": float(m.group(3))}
    m = re.search(r"Real time: ([\d.]+) sec; CPU: ([\d.]+) sec; Peak RSS: ([\d.]+) GB", stderr)
    if m:
        out["real_s"], out["cpu_s"], out["peak_rss_gb"] = float(m.group(1)), float(m.group(2)), float(m.group(3))
    return out


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "ont"
    T = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    G = int(sys.argv[3]) if len(sys.argv) > 3 else T
    w = WORK[wl]
    pkg = entry.load_package()
    from mm2gb_b200 import synth
    d = os.environ.get("MM2GB_RUN_DIR", "/tmp/mm2gb_driver_%s" % wl)
    os.makedirs(d, exist_ok=True)
    ref_fa, reads_fa = os.path.join(d, "ref.fa"), os.path.join(d, "reads.fa")
    t0 = time.time()
    if not (os.path.exists(ref_fa) and os.path.exists(reads_fa)):
        ref = synth.simulate_reference(w["ref_len"], seed=1, n_repeat_copies=w["repeats"], repeat_unit=3000)
        rds = synth.simulate_reads(ref, w["n_reads"], w["lo"], w["hi"], seed=2, err=w["err"])
        cl = w["ref_len"] // w["n_contigs"]
        synth.write_fasta(ref_fa, [ref[i * cl:(i + 1) * cl] for i in range(w["n_contigs"])], prefix="ref")
        synth.write_fasta(reads_fa, rds, prefix="read")
    gen_s = time.time() - t0
    REF = os.path.join(ROOT, "oracle", "_ref")
    cfg = os.path.join(ROOT, "mm2-gb_b200", "b200_config.json")
    env = dict(os.environ, MM2GB_THREADS_PER_GPU=str(G), MM2GB_VERBOSE=os.environ.get("MM2GB_VERBOSE", "1"))
    if os.environ.get("MM2GB_MAX_TOTAL_N"):     # a batch limit other than the shipped config's
        cfg = os.path.join(d, "cfg.json")
        open(cfg, "w").write('{"max_total_n": %d, "max_read": 200000}' % int(os.environ["MM2GB_MAX_TOTAL_N"]))

    def run(binary, args):
        t1 = time.time()
        p = subprocess.run([os.path.join(REF, binary)] + args + [ref_fa, reads_fa], capture_output=True, cwd=d, env=env)
        dt = time.time() - t1
        err = p.stderr.decode(errors="replace")
        if os.environ.get("MM2GB_LOG_DIR"):     # the driver's own log (its [M::...] lines carry wall-clock stamps) next to the JSON
            open(os.path.join(os.environ["MM2GB_LOG_DIR"], "%s_%s.log" % (binary, wl)), "w").write(err)
        if p.returncode != 0:
            raise SystemExit("%s failed (%d):\n%s" % (binary, p.returncode, err[-3000:]))
        return p.stdout, dt, err
    cpu_paf, cpu_s, cpu_err = run("minimap2_ref_timed", ["-t", str(T), "-x", "map-ont", "--max-chain-skip=2147483647"])
    gpu_paf, gpu_s, gpu_err = run("minimap2_b200_timed", ["-t", str(G), "-x", "map-ont", "--gpu-chain", "--gpu-cfg", cfg])
    a, b = cpu_paf.splitlines(), gpu_paf.splitlines()
    ndiff = sum(1 for x, y in zip(a, b) if x != y) + abs(len(a) - len(b))
    warn = [ln for ln in gpu_err.splitlines() if "WARNING" in ln or "ERROR" in ln][:5]
    per_thread = [ln for ln in gpu_err.splitlines() if ln.startswith("[mm2gb] thread")]
    gpus = {}
    for ln in per_thread:   # "[mm2gb] thread 3 (GPU 1): 2 batches, 1989314 anchors; ..."
        m = re.match(r"\[mm2gb\] thread \d+ \(GPU (\d+)\): (\d+) batches, (\d+) anchors", ln)
        if m:
            g = gpus.setdefault("gpu%s" % m.group(1), {"threads": 0, "batches": 0, "anchors": 0})
            g["threads"] += 1; g["batches"] += int(m.group(2)); g["anchors"] += int(m.group(3))
    # row N2: the driver with seeding + chaining on the device (integration/: three-line change of map.c + the glue); no --gpu-chain,
    # MM2GB_GPU_SEED=1 routes every batch of a worker thread through mm2gb_seed_chain
    seed_res = None
    if os.path.exists(os.path.join(REF, "minimap2_b200_seed")) and not os.environ.get("MM2GB_SKIP_SEED_DRIVER"):
        S = int(os.environ.get("MM2GB_SEED_THREADS", "0")) or G
        env_seed = dict(env, MM2GB_GPU_SEED="1")
        t1 = time.time()
        p = subprocess.run([os.path.join(REF, "minimap2_b200_seed"), "-t", str(S), "-x", "map-ont", "--max-chain-skip=2147483647", ref_fa, reads_fa],
                           capture_output=True, cwd=d, env=env_seed)
        seed_s = time.time() - t1
        seed_err = p.stderr.decode(errors="replace")
        if p.returncode != 0:
            raise SystemExit("minimap2_b200_seed failed (%d):\n%s" % (p.returncode, seed_err[-3000:]))
        c = p.stdout.splitlines()
        seed_res = {"binary": "MM2GB_GPU_SEED=1 minimap2_b200_seed -t %d (batches of %s reads per thread)" % (S, os.environ.get("MM2GB_SEED_BATCH_READS", "512")),
                    "wall_s": seed_s, "paf_md5": hashlib.md5(p.stdout).hexdigest(), "paf_lines": len(c),
                    "paf_lines_differing_from_cpu": sum(1 for x, y in zip(a, c) if x != y) + abs(len(a) - len(c)), "paf_identical": p.stdout == cpu_paf,
                    "timers": timers(seed_err), "fused_call_per_thread": [ln for ln in seed_err.splitlines() if ln.startswith("[mm2gb] seed+chain")][:4]}
    print(json.dumps({"workload": wl, "reads": w["n_reads"], "ref_len": w["ref_len"], "fasta_generation_s": gen_s,
                      "cpu": {"binary": "minimap2_ref_timed -t %d --max-chain-skip=2147483647" % T, "wall_s": cpu_s, "paf_md5": hashlib.md5(cpu_paf).hexdigest(),
                              "paf_lines": len(a), "timers": timers(cpu_err)},
                      "gpu": {"binary": "minimap2_b200_timed -t %d --gpu-chain" % G, "wall_s": gpu_s, "paf_md5": hashlib.md5(gpu_paf).hexdigest(),
                              "paf_lines": len(b), "timers": timers(gpu_err), "messages": warn, "work_per_gpu": gpus,
                              "boundary_per_thread": per_thread[:4] + (["... %d more" % (len(per_thread) - 4)] if len(per_thread) > 4 else [])},
                      "gpu_seed": seed_res,
                      "paf_lines_differing": ndiff, "paf_identical": cpu_paf == gpu_paf}))
    if seed_res is not None and not seed_res["paf_identical"]:
        sys.exit(1)
    if cpu_paf != gpu_paf:
        open(os.path.join(d, "cpu.paf"), "wb").write(cpu_paf); open(os.path.join(d, "gpu.paf"), "wb").write(gpu_paf)
        sys.exit(1)


if __name__ == "__main__":
    main()
