"""CPU-side checks of the integration build (oracle/_ref/minimap2_b200_seed: the reference's host sources + the edits under
integration/): without MM2GB_GPU_SEED it is the plain CPU driver, and `--max-chain-skip=infinity` means infinity
(integration/main_inf.sed; SURVEY.md trap T1: the reference parses that spelling with atoi and gets 0)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
SEED = os.path.join(REF, "minimap2_b200_seed")
CPU = os.path.join(REF, "minimap2_ref")

pytestmark = pytest.mark.skipif(not (os.path.exists(SEED) and os.path.exists(CPU)), reason="oracle/_ref binaries not built (needs /root/reference)")


def run(binary, args, cwd):
    out = subprocess.run([binary] + args, capture_output=True, cwd=cwd, timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    return out.stdout.decode()


def test_infinity_spelling_and_plain_cpu_path(synth, tmp_path):
    ref = synth.simulate_reference(1_500_000, seed=5, n_repeat_copies=120, repeat_unit=2000)
    rds = synth.simulate_reads(ref, 40, 5000, 30000, seed=6)
    synth.write_fasta(str(tmp_path / "ref.fa"), [ref], prefix="ref")
    synth.write_fasta(str(tmp_path / "reads.fa"), rds, prefix="read")
    args = ["-t", "2", "-x", "map-ont", str(tmp_path / "ref.fa"), str(tmp_path / "reads.fa")]
    truth = run(CPU, ["--max-chain-skip=2147483647"] + args, tmp_path)
    assert run(SEED, ["--max-chain-skip=2147483647"] + args, tmp_path) == truth      # the edits change nothing while switched off
    assert run(SEED, ["--max-chain-skip=infinity"] + args, tmp_path) == truth        # T1 fixed at the source
    assert run(SEED, ["--max-chain-skip=inf"] + args, tmp_path) == truth
    assert run(SEED, ["--max-chain-skip=25"] + args, tmp_path) == run(CPU, ["--max-chain-skip=25"] + args, tmp_path)
    assert len(truth.splitlines()) >= 40
