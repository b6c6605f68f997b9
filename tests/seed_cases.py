"""Deterministic inputs for the seeding tests (shared by the golden generator oracle/gen_seed_golden.py and the tests):
a small reference with repeat families, a tandem array, ambiguous bases and two contigs, plus reads drawn from it."""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def revcomp(s: bytes) -> bytes:
    return s.translate(COMP)[::-1]


def make_reference(seed: int = 3, n: int = 240000):
    rng = np.random.default_rng(seed)
    ref = bytearray(ACGT[rng.integers(0, 4, n)].tobytes())
    unit = ACGT[rng.integers(0, 4, 900)].tobytes()
    for p in rng.integers(0, n - 1000, 50):               # a repeat family, 2-3 % diverged copies
        u = bytearray(unit)
        for q in rng.integers(0, 900, 22):
            u[q] = ACGT[rng.integers(0, 4)]
        ref[p:p + 900] = u
    tu = ACGT[rng.integers(0, 4, 47)].tobytes()
    ref[80000:80000 + 47 * 90] = tu * 90                  # tandem array (period 47)
    ref[150000:150000 + 400] = b"AC" * 200                # dinucleotide repeat
    for p in (1234, 60000, 60001, 60002, 200000):
        ref[p] = ord("N")
    ref[120000:120050] = b"N" * 50
    ref = bytes(ref)
    contig2 = revcomp(ref[20000:70000]).lower()           # second contig: reverse strand copy, lower case
    return [ref, contig2]


def make_reads(refs, seed: int = 4, error: float = 0.08):
    rng = np.random.default_rng(seed)
    ref = refs[0]
    reads = []
    spans = [(500, 4000), (78000, 9000), (80100, 2500), (30000, 25000), (0, 250), (149000, 3000), (119000, 3000), (200000, 12000),
             (5, 14), (7, 15), (9, 40), (100000, 1)]
    for i, (st, ln) in enumerate(spans):
        r = bytearray(ref[st:st + ln])
        for q in rng.integers(0, max(ln, 1), int(ln * error)):
            r[q] = ACGT[rng.integers(0, 4)]
        r = bytes(r)
        if i % 3 == 1:
            r = revcomp(r)
        reads.append(r)
    reads.append(b"N" * 100)
    reads.append(b"A" * 700)
    reads.append((ACGT[rng.integers(0, 4, 31)].tobytes()) * 120)     # a read that is one tandem array not in the reference
    reads.append(ref[80000:80000 + 47 * 60])                         # ... and one that is (every seed high-occurrence)
    return reads
