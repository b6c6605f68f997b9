"""Synthetic workloads for the chaining path (seeded, numpy only).

Three generators:
  * simulate_reference / simulate_reads -- FASTA-level: random reference with planted repeat families and
    ONT-like reads.  Used to drive the minimap2 host driver end to end.
  * ont_like_anchors -- anchor-level: what `collect_seed_hits` (map.c:295-331) hands to the chaining stage for
    such reads: x-sorted mm128_t with a collinear backbone, indel drift, off-diagonal noise, both strands and
    several reference ids.  Used by the parity tests and the chaining-only bench (config 4).
  * adversarial_* -- hand-shaped arrays for the edge cases of lchain.c (ties, equal x, clipped windows ...).

Anchor packing follows minimap.h:72 / lchain.c:140-147:
    x = rev<<63 | rid<<32 | rpos          y = seg_id<<48 | flags<<40 | q_span<<32 | qpos
"""
from __future__ import annotations

import numpy as np

_COMP = np.frombuffer(b"TGCA", dtype=np.uint8)  # complement in 2-bit code order A,C,G,T
_NT = np.frombuffer(b"ACGT", dtype=np.uint8)


def pack_anchors(rid, rev, rpos, qpos, q_span=15, seg_id=0, flags=0) -> np.ndarray:
    """Build uint64[n,2] anchors (unsorted) from component arrays."""
    rid = np.asarray(rid, np.uint64); rev = np.asarray(rev, np.uint64)
    rpos = np.asarray(rpos, np.uint64); qpos = np.asarray(qpos, np.uint64)
    n = rpos.shape[0]
    x = (rev << np.uint64(63)) | (rid << np.uint64(32)) | (rpos & np.uint64(0xffffffff))
    y = (np.broadcast_to(np.asarray(seg_id, np.uint64), (n,)) << np.uint64(48)) \
        | (np.broadcast_to(np.asarray(flags, np.uint64), (n,)) << np.uint64(40)) \
        | (np.broadcast_to(np.asarray(q_span, np.uint64), (n,)) << np.uint64(32)) | (qpos & np.uint64(0xffffffff))
    return np.stack([x, y], axis=1)


def sort_by_x(a: np.ndarray, rng: np.random.Generator | None = None) -> np.ndarray:
    """Sort by x only.  minimap2 uses an unstable radix sort (map.c:329), so the order among equal x is
    arbitrary; with `rng` the ties are shuffled first to exercise that (SURVEY.md trap T5)."""
    if rng is not None:
        a = a[rng.permutation(a.shape[0])]
    return a[np.argsort(a[:, 0], kind="stable")]


def ont_like_anchors(rng: np.random.Generator, n_backbone: int, *, mean_gap: float = 25.0, noise_frac: float = 0.10,
                     drift_every: int = 10, drift_max: int = 20, rid: int = 0, rev: int = 0, rpos0: int = 100000,
                     qlen: int | None = None, q_span: int = 15, n_rid_noise: int = 1, repeat_copies: int = 0,
                     repeat_len: int = 0) -> np.ndarray:
    """One read's seeded anchors.

    Backbone: rpos = cumulative Geometric(mean_gap) gaps (about 5000/mean_gap anchors per max_dist window),
    qpos = rpos - rpos0 + drift, drift a random walk of +-1..drift_max every ~drift_every anchors (indels).
    noise_frac * n_backbone extra anchors land anywhere on (rid..rid+n_rid_noise-1, both strands).
    repeat_copies > 0 adds shifted copies of a backbone stretch of repeat_len anchors (a tandem-ish repeat):
    these create dense windows and equal-score ties.
    """
    gaps = rng.geometric(1.0 / mean_gap, size=n_backbone).astype(np.int64)
    rpos = rpos0 + np.cumsum(gaps)
    steps = np.where(rng.random(n_backbone) < 1.0 / drift_every,
                     rng.integers(1, drift_max + 1, n_backbone) * rng.choice([-1, 1], n_backbone), 0)
    drift = np.cumsum(steps)
    qpos = (rpos - rpos0) + drift
    qpos -= min(0, int(qpos.min())) - 20
    span = int(qpos.max()) + 100 if qlen is None else qlen
    parts = [pack_anchors(np.full(n_backbone, rid), np.full(n_backbone, rev), rpos, qpos, q_span)]
    n_noise = int(noise_frac * n_backbone)
    if n_noise:
        nr = rng.integers(rid, rid + n_rid_noise, n_noise)
        nrev = rng.integers(0, 2, n_noise)
        # half of the noise stays near the backbone in x (so it lands inside real windows), half anywhere
        near = rng.random(n_noise) < 0.5
        nx = np.where(near, rng.integers(int(rpos[0]), int(rpos[-1]) + 1, n_noise),
                      rng.integers(0, 2 * int(rpos[-1]) + 1000, n_noise))
        nr = np.where(near, rid, nr); nrev = np.where(near & (rng.random(n_noise) < 0.7), rev, nrev)
        ny = rng.integers(0, span, n_noise)
        parts.append(pack_anchors(nr, nrev, nx, ny, q_span))
    if repeat_copies and repeat_len:
        s = int(rng.integers(0, max(1, n_backbone - repeat_len)))
        for c in range(1, repeat_copies + 1):
            shift = c * int(rpos[min(s + repeat_len, n_backbone) - 1] - rpos[s] + mean_gap)
            parts.append(pack_anchors(np.full(repeat_len, rid)[: n_backbone - s], np.full(repeat_len, rev)[: n_backbone - s],
                                      rpos[s:s + repeat_len] + shift, qpos[s:s + repeat_len], q_span))
            parts.append(pack_anchors(np.full(repeat_len, rid)[: n_backbone - s], np.full(repeat_len, rev)[: n_backbone - s],
                                      rpos[s:s + repeat_len], qpos[s:s + repeat_len] + shift, q_span))
    return sort_by_x(np.concatenate(parts), rng)


def ont_like_batch(seed: int, n_reads: int, lo: int = 200, hi: int = 4000, **kw):
    """A batch of reads: (anchors uint64[N,2], offsets int64[n_reads+1])."""
    rng = np.random.default_rng(seed)
    reads = [ont_like_anchors(rng, int(rng.integers(lo, hi + 1)), rid=int(rng.integers(0, 24)),
                              rev=int(rng.integers(0, 2)), rpos0=int(rng.integers(0, 1 << 27)), **kw) for _ in range(n_reads)]
    off = np.zeros(n_reads + 1, np.int64)
    off[1:] = np.cumsum([r.shape[0] for r in reads])
    return (np.concatenate(reads) if reads else np.zeros((0, 2), np.uint64)), off


def chaining_only_array(seed: int, n: int, seg_len: int, *, mean_gap: float = 25.0, noise_frac: float = 0.10,
                        gap_bp: int = 6000) -> np.ndarray:
    """Config 4: one rid/strand, cumulative Geometric(mean_gap) gaps, qpos = rpos + drift,
    a gap of `gap_bp` (> max_dist) every `seg_len` anchors so the array splits into independent segments."""
    rng = np.random.default_rng(seed)
    n_noise = int(n * noise_frac / (1 + noise_frac))
    nb = n - n_noise
    gaps = rng.geometric(1.0 / mean_gap, size=nb).astype(np.int64)
    gaps[::seg_len] += gap_bp
    rpos = np.cumsum(gaps)
    steps = np.where(rng.random(nb) < 0.1, rng.integers(1, 21, nb) * rng.choice([-1, 1], nb), 0)
    # positions wrap inside 31 bits per "contig": bump rid every 2^30 bp
    rid = (rpos >> 30).astype(np.int64)
    rp = rpos & ((1 << 30) - 1)
    qpos = (rp + np.cumsum(steps)) & ((1 << 30) - 1)
    a = pack_anchors(rid, np.zeros(nb), rp, qpos)
    if n_noise:
        idx = rng.integers(0, nb, n_noise)
        noise = pack_anchors(rid[idx], np.zeros(n_noise), rp[idx], rng.integers(0, 1 << 30, n_noise))
        a = np.concatenate([a, noise])
    return a[np.argsort(a[:, 0], kind="stable")]


# ---------------------------------------------------------------------------------------------------------
# read-level workload through an (independent, minimap2-shaped) minimizer seeder -- csrc/synth_seed.cpp

def seeded_workload(seed: int, ref_len: int, n_reads: int, len_lo: int, len_hi: int, *, read_seed: int | None = None, err: float = 0.10,
                    contig_len: int = 0, n_repeat_copies: int = 0, repeat_unit: int = 3000, repeat_div: float = 0.03,
                    k: int = 15, w: int = 10, n_threads: int = 0, bg_len: int = 0, bg_contigs: int = 0, tandem_copies: int = 0,
                    tandem_unit: int = 0, tandem_div: float = 0.0, tandem_read_frac: float = 0.0, mid_occ: int = 0):
    """Random reference of `ref_len` bp (+ optional planted repeats), `n_reads` ONT-like reads U[len_lo,len_hi] at
    `err` error (40% sub / 30% del / 30% ins, half reverse-complemented), (w,k)-minimizer seeding with a mid-occ
    filter.  Returns (anchors uint64[N,2], offsets int64[n_reads+1]) -- x-sorted per read, as collect_seed_hits
    (map.c:295-331) would hand them to the chaining stage.

    bg_len / bg_contigs: the chance hits of a (virtual) reference of bg_len more bp in bg_contigs more contigs -- the hit mix of a
    human-scale index without building one.  tandem_*: a tandem array in contig 0 that tandem_read_frac of the reads come from;
    mid_occ > 0 fixes the occurrence cut-off (minimap2 -f <large>) so repetitive seeds survive."""
    import ctypes as C
    import os
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmm2gb_synth.so"))

    class Extra(C.Structure):
        _fields_ = [("bg_len", C.c_int64), ("bg_contigs", C.c_int), ("tandem_copies", C.c_int), ("tandem_unit", C.c_int),
                    ("tandem_div", C.c_double), ("tandem_read_frac", C.c_double), ("mid_occ", C.c_int)]
    lib.mm2gb_synth_create_ex.restype = C.c_void_p
    lib.mm2gb_synth_create_ex.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                          C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(Extra)]
    lib.mm2gb_synth_n_anchors.restype = C.c_int64
    lib.mm2gb_synth_n_anchors.argtypes = [C.c_void_p]
    lib.mm2gb_synth_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mm2gb_synth_free.argtypes = [C.c_void_p]
    if n_threads <= 0:
        n_threads = min(32, os.cpu_count() or 1)
    ex = Extra(bg_len, bg_contigs, tandem_copies, tandem_unit, tandem_div, tandem_read_frac, mid_occ)
    h = lib.mm2gb_synth_create_ex(seed, seed + 1 if read_seed is None else read_seed, ref_len, contig_len, n_repeat_copies, repeat_unit, repeat_div, n_reads,
                                  len_lo, len_hi, err, k, w, n_threads, C.byref(ex))
    if not h:
        raise ValueError("bad workload parameters")
    try:
        n = lib.mm2gb_synth_n_anchors(h)
        a = np.empty((n, 2), np.uint64)
        off = np.empty(n_reads + 1, np.int64)
        lib.mm2gb_synth_copy(h, a.ctypes.data, off.ctypes.data)
    finally:
        lib.mm2gb_synth_free(h)
    return a, off


# ---------------------------------------------------------------------------------------------------------
# adversarial vectors (SURVEY.md 8c iii)

def adversarial_max_ii() -> np.ndarray:
    """6002 anchors where the max_ii fallback of lchain.c:189-205 wins.
    Use with min_cnt=1, min_score=1.  Expected f[-1]=30, p[-1]=0."""
    k = np.arange(6000)
    a = np.concatenate([pack_anchors([0], [0], [1000], [1000]),
                        pack_anchors(np.zeros(6000), np.zeros(6000), 1001 + k // 2, 200000 - k),
                        pack_anchors([0], [0], [4500], [4500])])
    return a  # already x-sorted, ties in the given order


def adversarial_ties(rng: np.random.Generator, n: int = 600) -> np.ndarray:
    """Perfect diagonal with duplicated anchors and equal gaps: many equal-score candidates per row, equal x."""
    base = np.arange(n // 2) * 20 + 5000
    rpos = np.concatenate([base, base])           # every x twice
    qpos = np.concatenate([base - 4000, base - 4000 + rng.integers(0, 2, n // 2) * 20])
    return sort_by_x(pack_anchors(np.zeros(n), np.zeros(n), rpos, qpos), rng)


def adversarial_dd_sweep(bw: int = 500) -> np.ndarray:
    """Pairs whose |dr-dq| sweeps 0..bw+2: touches every entry of the gap-penalty table and the band edge."""
    parts = []
    for dd in range(bw + 3):
        x0 = 10000 + dd * 20000
        parts.append(pack_anchors([0, 0, 0], [0, 0, 0], [x0, x0 + 600, x0 + 1300 + dd], [100, 700 + dd, 1400 + dd]))
    return sort_by_x(np.concatenate(parts))


def adversarial_qspan(rng: np.random.Generator, n: int = 800) -> np.ndarray:
    """Anchors with q_span != 15 (k != 15 / HPC seeds): spans 11..40 mixed."""
    a = ont_like_anchors(rng, n, noise_frac=0.05)
    span = rng.integers(11, 41, a.shape[0]).astype(np.uint64)
    a[:, 1] = (a[:, 1] & ~np.uint64(0xff << 32)) | (span << np.uint64(32))
    return a


def adversarial_dense(rng: np.random.Generator, n: int = 7000, width: int = 3000) -> np.ndarray:
    """> max_iter anchors inside one max_dist window (tandem repeat): every window is clipped (lchain.c:173)."""
    rpos = np.sort(rng.integers(50000, 50000 + width, n))
    qpos = np.sort(rng.integers(1000, 1000 + width, n)) + rng.integers(-30, 31, n)
    return sort_by_x(pack_anchors(np.zeros(n), np.zeros(n), rpos, qpos), rng)


def adversarial_multi(rng: np.random.Generator) -> np.ndarray:
    """Several reference ids and both strands in one read, including runs that straddle rid/strand changes."""
    parts = [ont_like_anchors(rng, int(rng.integers(20, 400)), rid=r, rev=s, rpos0=int(rng.integers(0, 5000)), noise_frac=0.2)
             for r in (0, 1, 7) for s in (0, 1)]
    return sort_by_x(np.concatenate(parts), rng)


def adversarial_suite(seed: int = 7):
    """name -> (anchors, params overrides)."""
    rng = np.random.default_rng(seed)
    return {
        "empty": (np.zeros((0, 2), np.uint64), {}),
        "single": (pack_anchors([3], [1], [12345], [77]), {}),
        "pair": (pack_anchors([0, 0], [0, 0], [100, 130], [10, 40]), {"min_cnt": 1, "min_score": 1}),
        "max_ii": (adversarial_max_ii(), {"min_cnt": 1, "min_score": 1}),
        "ties": (adversarial_ties(rng), {}),
        "dd_sweep": (adversarial_dd_sweep(), {"min_cnt": 1, "min_score": 1}),
        "qspan": (adversarial_qspan(rng), {}),
        "dense": (adversarial_dense(rng), {}),
        "multi": (adversarial_multi(rng), {}),
        "repeat": (ont_like_anchors(rng, 3000, repeat_copies=6, repeat_len=300), {}),
        "skip_pen": (ont_like_anchors(rng, 1500), {"chn_pen_skip": 0.05}),
        "narrow": (ont_like_anchors(rng, 1500), {"bw": 100, "max_dist_x": 800, "max_dist_y": 600}),
        "small_iter": (ont_like_anchors(rng, 2500, mean_gap=8.0), {"max_iter": 50}),
    }


# ---------------------------------------------------------------------------------------------------------
# FASTA level

def simulate_reference(length: int, seed: int = 1, n_repeat_copies: int = 0, repeat_unit: int = 3000,
                       repeat_div: float = 0.03) -> np.ndarray:
    """Uniform-random ACGT (2-bit codes, uint8) with `n_repeat_copies` diverged copies of one repeat unit."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, length, dtype=np.uint8)
    if n_repeat_copies:
        unit = rng.integers(0, 4, repeat_unit, dtype=np.uint8)
        for pos in rng.integers(0, length - repeat_unit, n_repeat_copies):
            cp = unit.copy()
            m = rng.random(repeat_unit) < repeat_div
            cp[m] = (cp[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
            ref[pos:pos + repeat_unit] = cp
    return ref


def simulate_reads(ref: np.ndarray, n_reads: int, lo: int, hi: int, seed: int = 2, err: float = 0.10):
    """ONT-like reads: length U[lo,hi], error `err` split 40% sub / 30% del / 30% ins, 50% reverse-complemented.
    Returns a list of uint8 code arrays."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_reads):
        ln = int(rng.integers(lo, hi + 1))
        ln = min(ln, ref.shape[0] - 1)
        st = int(rng.integers(0, ref.shape[0] - ln))
        s = ref[st:st + ln].copy()
        r = rng.random(ln)
        sub = r < 0.4 * err
        dele = (r >= 0.4 * err) & (r < 0.7 * err)
        ins = (r >= 0.7 * err) & (r < err)
        s[sub] = (s[sub] + rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)) & 3
        keep = ~dele
        reps = np.where(ins, 2, 1)[keep]
        s = np.repeat(s[keep], reps)
        # the duplicated base of an insertion becomes a random base
        dup = np.zeros(s.shape[0], bool)
        ends = np.cumsum(reps) - 1
        dup[ends[reps == 2]] = True
        s[dup] = rng.integers(0, 4, int(dup.sum()), dtype=np.uint8)
        if rng.random() < 0.5:
            s = (3 - s)[::-1]
        out.append(np.ascontiguousarray(s))
    return out


def write_fasta(path: str, seqs, prefix: str = "s", width: int = 0):
    with open(path, "wb") as fh:
        for i, s in enumerate(seqs):
            fh.write(b">%s%d\n" % (prefix.encode(), i))
            fh.write(_NT[s].tobytes())
            fh.write(b"\n")
