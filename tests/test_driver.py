"""End-to-end through the reference's UNMODIFIED host driver (main.c / map.c, --gpu-chain, PAF output) linked against the
drop-in: oracle/_ref/minimap2_b200 (built by `make -C oracle ref` in the build container; the binary travels to the GPU
box) against the CPU ground truth oracle/_ref/minimap2_ref --max-chain-skip=2147483647 (SURVEY.md trap T1) and the
committed golden PAFs.  Bar: empty diff."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
B200 = os.path.join(REF, "minimap2_b200")
CPU = os.path.join(REF, "minimap2_ref")
FIX = os.path.join(REF, "fixtures")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (os.path.exists(B200) and os.path.exists(CPU)),
                                                  reason="oracle/_ref binaries not built (needs /root/reference at build time)")]


def run(binary, args, cwd):
    out = subprocess.run([binary] + args, capture_output=True, cwd=cwd, timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    return out.stdout.decode()


def write_cfg(path, max_total_n=2_000_000, max_read=2000):
    with open(path, "w") as fh:
        fh.write('{"max_total_n": %d, "max_read": %d, "host_threads": 4}' % (max_total_n, max_read))


@pytest.mark.parametrize("tag,t,q", [("MT", "MT-human.fa", "MT-orang.fa"), ("inv", "t-inv.fa", "q-inv.fa"), ("t2", "t2.fa", "q2.fa")])
def test_reference_fixtures_paf(tmp_path, golden_dir, tag, t, q):
    """the reference's own accuracy recipe (README.md:85-96) on its own test FASTA"""
    write_cfg(tmp_path / "gpu_config.json")
    gpu = run(B200, ["-t", "1", "--gpu-chain", "--gpu-cfg", str(tmp_path / "gpu_config.json"), os.path.join(FIX, t), os.path.join(FIX, q)], tmp_path)
    cpu = run(CPU, ["-t", "1", "--max-chain-skip=2147483647", os.path.join(FIX, t), os.path.join(FIX, q)], tmp_path)
    assert gpu == cpu
    assert gpu == open(os.path.join(golden_dir, tag + ".paf")).read()


def _synth_fasta(synth, tmp_path, ref_len, n_reads, lo, hi, repeats=0, seed=1):
    ref = synth.simulate_reference(ref_len, seed=seed, n_repeat_copies=repeats, repeat_unit=2000)
    rds = synth.simulate_reads(ref, n_reads, lo, hi, seed=seed + 1)
    synth.write_fasta(str(tmp_path / "ref.fa"), [ref], prefix="ref")
    synth.write_fasta(str(tmp_path / "reads.fa"), rds, prefix="read")
    return str(tmp_path / "ref.fa"), str(tmp_path / "reads.fa")


def test_golden_synthetic_reads_paf(synth, tmp_path, golden_dir):
    """the 8 simulated reads of tests/golden/synth.paf (regenerated from their seeds)"""
    ref, reads = _synth_fasta(synth, tmp_path, 600_000, 8, 4000, 16000, repeats=40)
    write_cfg(tmp_path / "cfg.json")
    gpu = run(B200, ["-t", "1", "-x", "map-ont", "--gpu-chain", "--gpu-cfg", str(tmp_path / "cfg.json"), ref, reads], tmp_path)
    assert gpu == open(os.path.join(golden_dir, "synth.paf")).read()


@pytest.mark.parametrize("threads,max_total_n", [(1, 2_000_000), (4, 150_000), (3, 40_000)])
def test_ont_like_reads_paf_multithread_small_batches(synth, tmp_path, threads, max_total_n):
    """120 ONT-like reads on a 3 Mb reference with repeats; several driver threads and batch limits small enough that
    batches rotate many times (and some single reads exceed the limit)"""
    ref, reads = _synth_fasta(synth, tmp_path, 3_000_000, 120, 5000, 40000, repeats=150, seed=11)
    write_cfg(tmp_path / "cfg.json", max_total_n=max_total_n, max_read=64)
    gpu = run(B200, ["-t", str(threads), "-x", "map-ont", "--gpu-chain", "--gpu-cfg", str(tmp_path / "cfg.json"), ref, reads], tmp_path)
    cpu = run(CPU, ["-t", "4", "-x", "map-ont", "--max-chain-skip=2147483647", ref, reads], tmp_path)
    assert len(cpu.splitlines()) >= 120
    assert gpu == cpu


def test_max_chain_skip_spelling_is_ignored_by_gpu_path(synth, tmp_path):
    """trap T1: `--max-chain-skip=infinity` parses to 0 on the CPU path; the GPU path is true infinity whatever the flag"""
    ref, reads = _synth_fasta(synth, tmp_path, 1_000_000, 30, 5000, 30000, repeats=60, seed=21)
    write_cfg(tmp_path / "cfg.json")
    base = ["-t", "2", "-x", "map-ont", "--gpu-chain", "--gpu-cfg", str(tmp_path / "cfg.json")]
    a = run(B200, base + ["--max-chain-skip=infinity", ref, reads], tmp_path)
    b = run(B200, base + [ref, reads], tmp_path)
    cpu = run(CPU, ["-t", "2", "-x", "map-ont", "--max-chain-skip=2147483647", ref, reads], tmp_path)
    assert a == b == cpu


# ---- row N2: seeding + chaining on the device inside the reference's driver ------------------------------------------------
# oracle/_ref/minimap2_b200_seed = the reference's host sources + the three-line change of integration/map_gpu_seed.sed + the
# reference-side binding integration/mm2gb_seed_glue.c.  With MM2GB_GPU_SEED=1 no read is seeded or chained on the host; the PAF
# (alignment, mapq, dv: everything downstream of seeds, chains, rep_len and mini_pos) must equal the CPU driver's.
SEED = os.path.join(REF, "minimap2_b200_seed")
needs_seed_driver = pytest.mark.skipif(not os.path.exists(SEED), reason="oracle/_ref/minimap2_b200_seed not built")


def run_env(binary, args, cwd, **env):
    out = subprocess.run([binary] + args, capture_output=True, cwd=cwd, timeout=900, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    return out.stdout.decode(), out.stderr.decode()


@needs_seed_driver
@pytest.mark.parametrize("tag,t,q", [("MT", "MT-human.fa", "MT-orang.fa"), ("inv", "t-inv.fa", "q-inv.fa"), ("t2", "t2.fa", "q2.fa")])
def test_gpu_seed_driver_fixtures_paf(tmp_path, golden_dir, tag, t, q):
    gpu, _ = run_env(SEED, ["-t", "1", "--max-chain-skip=2147483647", os.path.join(FIX, t), os.path.join(FIX, q)], tmp_path, MM2GB_GPU_SEED="1")
    assert gpu == open(os.path.join(golden_dir, tag + ".paf")).read()


@needs_seed_driver
@pytest.mark.parametrize("threads", [1, 4])
def test_gpu_seed_driver_ont_like_reads_paf(synth, tmp_path, threads):
    """300 ONT-like reads on a 3 Mb reference with repeats (batches of 64 reads per driver thread, map.c:23)"""
    ref, reads = _synth_fasta(synth, tmp_path, 3_000_000, 300, 5000, 40000, repeats=150, seed=11)
    gpu, err = run_env(SEED, ["-t", str(threads), "-x", "map-ont", "--max-chain-skip=2147483647", ref, reads], tmp_path, MM2GB_GPU_SEED="1", MM2GB_VERBOSE="1")
    off, _ = run_env(SEED, ["-t", "4", "-x", "map-ont", "--max-chain-skip=2147483647", ref, reads], tmp_path)     # same binary, host path
    cpu = run(CPU, ["-t", "4", "-x", "map-ont", "--max-chain-skip=2147483647", ref, reads], tmp_path)
    assert len(cpu.splitlines()) >= 300
    assert off == cpu
    assert gpu == cpu
    assert "seed+chain thread" in err          # the device path really ran


@needs_seed_driver
def test_gpu_seed_driver_threads_spread_over_gpus(pkg, synth, tmp_path):
    """worker thread t drives GPU t % n_gpus (one device index per GPU); skipped on a one-GPU box"""
    if pkg.lib().mm2gb_device_count() < 2:
        pytest.skip("needs two GPUs")
    ref, reads = _synth_fasta(synth, tmp_path, 3_000_000, 300, 5000, 40000, repeats=150, seed=11)
    gpu, err = run_env(SEED, ["-t", "4", "-x", "map-ont", "--max-chain-skip=2147483647", ref, reads], tmp_path, MM2GB_GPU_SEED="1", MM2GB_VERBOSE="1",
                       MM2GB_SEED_BATCH_READS="32")
    cpu = run(CPU, ["-t", "4", "-x", "map-ont", "--max-chain-skip=2147483647", ref, reads], tmp_path)
    assert gpu == cpu
    assert err.count("seed+chain thread") >= 2


@needs_seed_driver
def test_gpu_seed_driver_map_pb_hpc_paf(synth, tmp_path):
    """-x map-pb: homopolymer-compressed minimizers (k = 19) seeded and chained on the device inside the driver"""
    ref, reads = _synth_fasta(synth, tmp_path, 3_000_000, 200, 5000, 40000, repeats=150, seed=13)
    gpu, err = run_env(SEED, ["-t", "3", "-x", "map-pb", "--max-chain-skip=2147483647", ref, reads], tmp_path, MM2GB_GPU_SEED="1", MM2GB_VERBOSE="1")
    cpu = run(CPU, ["-t", "4", "-x", "map-pb", "--max-chain-skip=2147483647", ref, reads], tmp_path)
    assert len(cpu.splitlines()) >= 200
    assert gpu == cpu
    assert "seed+chain thread" in err
