"""seed_model.py -- CPU restatement (pure Python, small cases) of the reference's seeding stage in the DATA-PARALLEL form the device
kernels use (mm2-gb_b200/csrc/seed_kernels.cuh), pinned against the reference itself through oracle/pyrefseed.py.

TEST INFRASTRUCTURE ONLY (tests/test_seed_oracle.py); nothing in the product imports it.  What it restates, and from where:
  sketch_model      mm_sketch, sketch.c:77-143, as per-position rules P1-P4 over the hashes of the last w + 1 positions and the
                    run length of unambiguous bases (odd k, no HPC: the loop never takes its `continue`, sketch.c:104)
  flag_sort_model   radix_sort_128x, ksort.h:98-151: MSD American-flag passes on (digit, index) words, insertion sort of buckets
                    of <= 64 elements as a stable rank -- reproduces the tie order the unstable sort leaves among equal keys
  pass_dest_walk /  one flag pass (ksort.h:125-138) replayed on the ORIGINAL digits alone, emitting a destination per moved
  pass_dest_two     element, and its closed form for passes with exactly two non-empty buckets (what k_seed_sort runs)
  seed_model        mm_map_seed, map.c:355-391: mm_seed_mz_flt (seed.c:5-29), mm_seed_collect_all (:31-53), mm_seed_select in
                    closed form (:57-96: the heap keeps the k smallest (n, j) of a streak), rep_len (:117-121,128), anchors
                    (map.c:303-325) and the sort
Parity pinned: every function is checked against oracle/_ref/libref_seed.so (the reference's own sketch.c / seed.c / map.c / ksort.h
compiled where they lie) in tests/test_seed_oracle.py and by `python oracle/seed_model.py`."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

NONE = (1 << 64) - 1
M64 = (1 << 64) - 1

def hash64(key, mask):
    key = (~key + (key << 21)) & mask
    key = key ^ key >> 24
    key = ((key + (key << 3)) + (key << 8)) & mask
    key = key ^ key >> 14
    key = ((key + (key << 2)) + (key << 4)) & mask
    key = key ^ key >> 28
    key = (key + (key << 31)) & mask
    return key

NT4 = {ord('A'): 0, ord('a'): 0, ord('C'): 1, ord('c'): 1, ord('G'): 2, ord('g'): 2, ord('T'): 3, ord('t'): 3, ord('U'): 3, ord('u'): 3}

def sketch_model(seq: bytes, w: int, k: int, rid: int = 0):
    n = len(seq)
    code = [NT4.get(c, 4) for c in seq]
    mask = (1 << (2 * k)) - 1
    ix = [NONE] * n
    iz = [0] * n
    run = 0
    l = [0] * n
    for i in range(n):
        run = run + 1 if code[i] < 4 else 0
        l[i] = run
        if run >= k:
            f = r = 0
            for t in range(k - 1, -1, -1):
                c = code[i - t]
                f = (f << 2 | c) & mask
                r = (r >> 2) | (3 ^ c) << (2 * (k - 1))
            if f != r:
                z = 0 if f < r else 1
                ix[i] = hash64(r if z else f, mask) << 8 | k
                iz[i] = z
    def X(p):
        return ix[p] if p >= 0 else NONE
    out = []
    T1 = w + k - 1
    for i in range(n):
        cur = ix[i]
        mprev_x, mprev_p = NONE, -1
        for d in range(w, 0, -1):
            x = X(i - d)
            if x <= mprev_x:
                mprev_x, mprev_p = x, i - d
        mode = 0
        mx, mp = NONE, -1
        if cur <= mprev_x:
            mode = 2
        elif mprev_p == i - w:
            mode = 3
            for d in range(w - 1, -1, -1):
                x = X(i - d)
                if x <= mx:
                    mx, mp = x, i - d
        def emit(p):
            out.append((ix[p], rid << 32 | p << 1 | iz[p]))
        if l[i] == T1 and mprev_x != NONE:
            for d in range(w - 1, 0, -1):
                if X(i - d) == mprev_x and i - d != mprev_p:
                    emit(i - d)
        if mode == 2:
            if l[i] >= T1 + 1 and mprev_x != NONE:
                emit(mprev_p)
        elif mode == 3:
            if l[i] >= T1:
                emit(mprev_p)
            if l[i] >= T1 and mx != NONE:
                for d in range(w - 1, -1, -1):
                    if X(i - d) == mx and i - d != mp:
                        emit(i - d)
        if i == n - 1:
            fx, fp = (cur, i) if mode == 2 else (mx, mp) if mode == 3 else (mprev_x, mprev_p)
            if fx != NONE:
                emit(fp)
    return np.array(out, dtype=np.uint64).reshape(-1, 2)


def flag_sort_model(keys):
    """radix_sort_128x (ksort.h:98-151) replayed the way k_seed_sort does it: returns the permutation (indices into keys)."""
    n = len(keys)
    W = list(range(n))
    def small(beg, end):
        seg = W[beg:end]
        seg = [x for _, x in sorted(zip([keys[i] for i in seg], range(len(seg))), key=lambda t: (t[0], t[1]))]
        W[beg:end] = [W[beg + j] for j in seg]
    if n <= 64:
        small(0, n)
        return W
    level = [(0, n)]
    sh = 56
    while sh >= 0 and level:
        nxt = []
        for beg, end in level:
            m = end - beg
            dig = {i: (keys[W[i]] >> sh) & 255 for i in range(beg, end)}
            cnt = [0] * 256
            for i in range(beg, end):
                cnt[dig[i]] += 1
            if max(cnt) == m:
                if sh > 0:
                    nxt.append((beg, end))
                continue
            cur = [0] * 256
            en = [0] * 256
            run = beg
            for d in range(256):
                cur[d] = run
                run += cnt[d]
                en[d] = run
            D = [0] * n
            for i in range(beg, end):
                D[i] = dig[i]
            for kk in range(256):
                kb = cur[kk]
                while kb != en[kk]:
                    cw, cd = W[kb], D[kb]
                    if cd == kk:
                        kb += 1
                        continue
                    while True:
                        pos = cur[cd]
                        cur[cd] = pos + 1
                        ew, ed = W[pos], D[pos]
                        W[pos], D[pos] = cw, cd
                        cw, cd = ew, ed
                        if cd == kk:
                            break
                    W[kb], D[kb] = cw, cd
                    kb += 1
            if sh > 0:
                for kk in range(256):
                    bb = en[kk - 1] if kk else beg
                    be = en[kk]
                    if be - bb > 64:
                        nxt.append((bb, be))
                    elif be - bb > 1:
                        small(bb, be)
        level = nxt
        sh -= 8
    return W


if __name__ == "__main__":
    import pyrefseed as rs
    rng = np.random.default_rng(7)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    cases = []
    for n in (1, 5, 14, 15, 16, 24, 25, 26, 40, 300, 3000):
        cases.append(acgt[rng.integers(0, 4, n)].tobytes())
    s = bytearray(acgt[rng.integers(0, 4, 4000)].tobytes())
    for p in (10, 30, 31, 100, 101, 102, 500, 523, 524, 525, 526, 2000, 3999):
        s[p] = ord('N')
    cases.append(bytes(s))
    cases.append(b"A" * 500)
    cases.append(b"AC" * 400)
    cases.append(b"ACG" * 300 + b"N" + b"ACGTT" * 100)
    unit = acgt[rng.integers(0, 4, 37)].tobytes()
    cases.append(unit * 60)
    cases.append(b"acgtn" * 100 + unit * 5)
    bad = 0
    for w, k in ((10, 15), (5, 15), (19, 19), (1, 15), (11, 21), (32, 27)):
        for c in cases:
            a = rs.sketch(c, w, k)
            b = sketch_model(c, w, k)
            if a.shape != b.shape or not np.array_equal(a, b):
                bad += 1
                print("MISMATCH w", w, "k", k, "len", len(c), a.shape, b.shape)
    print("sketch model mismatches:", bad)


def seed_model(ix, seq: bytes, w, k, mid_occ, max_max_occ, occ_dist, q_occ_frac):
    """mm_map_seed the way the kernels compute it; index lookups through the reference index object."""
    mv = sketch_model(seq, w, k)
    n = len(mv)
    keep = [True] * n
    if n > mid_occ and q_occ_frac > 0 and mid_occ > 0:
        from collections import Counter
        cnt = Counter(int(x) for x in mv[:, 0])
        for i in range(n):
            c = cnt[int(mv[i, 0])]
            if c > mid_occ and np.float32(c) > np.float32(n) * np.float32(q_occ_frac):
                keep[i] = False
    surv = [i for i in range(n) if keep[i]]
    seeds = []  # (n_occ, q_pos, occ list, tandem)
    for t, i in enumerate(surv):
        mz = int(mv[i, 0]) >> 8
        occ = ix.get(mz)
        if len(occ) == 0:
            continue
        td = (t > 0 and int(mv[surv[t - 1], 0]) >> 8 == mz) or (t + 1 < len(surv) and int(mv[surv[t + 1], 0]) >> 8 == mz)
        seeds.append([len(occ), int(mv[i, 1]) & 0xffffffff, occ, td])
    m = len(seeds)
    flt = [0] * m
    for i in range(m):
        ni = seeds[i][0]
        if occ_dist > 0 and max_max_occ > mid_occ:
            if m >= 2 and ni > mid_occ:
                st, en, rank = i, i + 1, 0
                while st > 0 and seeds[st - 1][0] > mid_occ:
                    st -= 1
                    if seeds[st][0] <= ni:
                        rank += 1
                while en < m and seeds[en][0] > mid_occ:
                    if seeds[en][0] < ni:
                        rank += 1
                    en += 1
                ps = seeds[st - 1][1] >> 1 if st > 0 else 0
                pe = seeds[en][1] >> 1 if en < m else len(seq)
                mho = int((pe - ps) / occ_dist + .499)
                mho = min(mho, 128)
                flt[i] = 0 if (mho > 0 and rank < mho) else 1
                if ni > max_max_occ:
                    flt[i] = 1
        elif ni > mid_occ:
            flt[i] = 1
    rep = 0
    prev_en = 0
    for i in range(m):
        if flt[i]:
            en = (seeds[i][1] >> 1) + 1
            st = en - k
            rep += en - max(st, prev_en)
            prev_en = en
    a = []
    mp = []
    qlen = len(seq)
    for i in range(m):
        if flt[i]:
            continue
        nocc, qp, occ, td = seeds[i]
        fl = (1 << 42) if td else 0
        for r in occ:
            r = int(r)
            rpos = (r & 0xffffffff) >> 1
            if (r & 1) == (qp & 1):
                a.append(((r & 0xffffffff00000000) | rpos, k << 32 | qp >> 1 | fl))
            else:
                a.append((1 << 63 | (r & 0xffffffff00000000) | rpos, k << 32 | (qlen - ((qp >> 1) + 1 - k) - 1) | fl))
        mp.append(k << 32 | qp >> 1)
    perm = flag_sort_model([x for x, _ in a])
    a = np.array([a[i] for i in perm], dtype=np.uint64).reshape(-1, 2)
    return a, rep, np.array(mp, dtype=np.uint64)


def check_pipeline():
    import pyrefseed as rs
    rng = np.random.default_rng(11)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    # sort replay against the reference on keys with many ties
    bad = 0
    for n in (10, 64, 65, 200, 1000, 5000):
        for spread in (3, 50, 1 << 20, 1 << 40):
            x = rng.integers(0, spread, n).astype(np.uint64) * np.uint64(0x0101010101) + (rng.integers(0, 2, n).astype(np.uint64) << np.uint64(63))
            xy = np.stack([x, np.arange(n, dtype=np.uint64)], axis=1)
            ref = rs.radix_sort_128x(xy)
            perm = flag_sort_model([int(v) for v in x])
            if not np.array_equal(ref[:, 1], np.array(perm, dtype=np.uint64)):
                bad += 1
                print("sort MISMATCH n", n, "spread", spread)
    print("sort model mismatches:", bad)
    # a reference with repeat families and a tandem array; reads spanning them
    ref = bytearray(acgt[rng.integers(0, 4, 300000)].tobytes())
    unit = acgt[rng.integers(0, 4, 800)].tobytes()
    for p in rng.integers(0, 290000, 60):
        u = bytearray(unit)
        for q in rng.integers(0, 800, 20):
            u[q] = acgt[rng.integers(0, 4)]
        ref[p:p + 800] = u
    tu = acgt[rng.integers(0, 4, 53)].tobytes()
    ref[100000:100000 + 53 * 80] = tu * 80
    ref = bytes(ref)
    ix = rs.RefIndex([ref, ref[5000:60000][::-1]])
    bad = 0
    for mid_occ, dist, frac in ((10, 500, 0.01), (3, 500, 0.01), (3, 0, 0.01), (5, 100, 0.0), (2, 50, 0.002)):
        ix.field("mid_occ", mid_occ); ix.field("occ_dist", dist); ix.field("q_occ_frac", frac)
        for st, ln in ((1000, 3000), (99000, 8000), (100100, 2000), (50000, 20000), (0, 300), (200000, 12000)):
            read = bytearray(ref[st:st + ln])
            for q in rng.integers(0, ln, ln // 12):
                read[q] = acgt[rng.integers(0, 4)]
            if st % 2000 == 0:
                comp = bytes.maketrans(b"ACGT", b"TGCA")
                read = bytearray(bytes(read).translate(comp)[::-1])
            read = bytes(read)
            a0, rep0, mp0 = ix.seed(read)
            a1, rep1, mp1 = seed_model(ix, read, 10, 15, mid_occ, 4095, dist, np.float32(frac))
            ok = a0.shape == a1.shape and np.array_equal(a0, a1) and rep0 == rep1 and np.array_equal(mp0, mp1)
            if not ok:
                bad += 1
                same_set = a0.shape == a1.shape and np.array_equal(a0[np.lexsort((a0[:, 1], a0[:, 0]))], a1[np.lexsort((a1[:, 1], a1[:, 0]))])
                print("seed MISMATCH", mid_occ, dist, frac, st, ln, a0.shape, a1.shape, rep0, rep1, len(mp0), len(mp1), "same set" if same_set else "different set")
    print("seed model mismatches:", bad)


if __name__ == "__main__":
    check_pipeline()


# ---- second formulation of one American-flag pass (what k_seed_sort does since r6b) ------------------------------------------
# The walk only ever reads slots that still hold their original element (a bucket's cursor never passes a slot twice), so it can
# run on the ORIGINAL digits alone and emit, per moved element, its destination; the elements are then scattered in parallel.

def pass_dest_walk(dig):
    n = len(dig)
    cnt = [0] * 256
    for d in dig:
        cnt[d] += 1
    cur, en, run = [0] * 256, [0] * 256, 0
    for d in range(256):
        cur[d] = run
        run += cnt[d]
        en[d] = run
    dest = list(range(n))
    for kk in range(256):
        kb = cur[kk]
        while kb != en[kk]:
            d = dig[kb]
            if d == kk:
                kb += 1
                continue
            src = kb
            while True:
                pos = cur[d]
                cur[d] = pos + 1
                dest[src] = pos
                src = pos
                d = dig[pos]
                if d == kk:
                    break
            dest[src] = kb
            kb += 1
    return dest, en


def pass_dest_two(dig):
    """closed form for a pass with exactly two non-empty buckets"""
    n = len(dig)
    lo = min(dig)
    ca = sum(1 for d in dig if d == lo)
    dest = list(range(n))
    q = [x for x in range(ca, n) if dig[x] == lo]
    p = [x for x in range(0, ca) if dig[x] != lo]
    assert len(p) == len(q)
    t = len(p)
    for j in range(t):
        dest[p[j]] = (q[j - 1] + 1) if j else ca
        dest[q[j]] = p[j]
    qt = q[-1] if t else -1
    for x in range(ca, n):
        if dig[x] != lo and x < qt:
            dest[x] = x + 1
    return dest


def check_dest_forms():
    rng = np.random.default_rng(3)
    bad = 0
    for n in (2, 3, 10, 100, 1000):
        for nb in (2, 3, 7, 256):
            for _ in range(20):
                dig = [int(v) for v in rng.choice(rng.choice(256, nb, replace=False), n)]
                keys = [d << 56 for d in dig]
                # the reference order after ONE pass at shift 56 = flag pass on the digits; emulate with the word model
                W = list(range(n))
                cnt = [0] * 256
                for d in dig:
                    cnt[d] += 1
                cur, en, run = [0] * 256, [0] * 256, 0
                for d in range(256):
                    cur[d] = run; run += cnt[d]; en[d] = run
                D = list(dig)
                for kk in range(256):
                    kb = cur[kk]
                    while kb != en[kk]:
                        cw, cd = W[kb], D[kb]
                        if cd == kk:
                            kb += 1; continue
                        while True:
                            pos = cur[cd]; cur[cd] = pos + 1
                            ew, ed = W[pos], D[pos]
                            W[pos], D[pos] = cw, cd
                            cw, cd = ew, ed
                            if cd == kk:
                                break
                        W[kb], D[kb] = cw, cd
                        kb += 1
                dest, _ = pass_dest_walk(dig)
                W2 = [0] * n
                for s in range(n):
                    W2[dest[s]] = s
                if W2 != W:
                    bad += 1
                if len(set(dig)) == 2:
                    d2 = pass_dest_two(dig)
                    if d2 != dest:
                        bad += 1
                        print("two-bucket form differs", n, nb)
    print("dest-form mismatches:", bad)


if __name__ == "__main__":
    check_dest_forms()


# ---- third formulation: the walk with run jumps (tried in k_seed_sort, measured 2.5x SLOWER than the one-lane walk: kept as a model) ----
# An arrival into bucket d lands at cur[d] and pushes the run of d-residents behind it one slot up, until the next slot of d's
# region that holds a foreign element (which is evicted and carries on).  The run is a RANGE (dest[x] = x + 1 for all of it, found
# by scanning the digits 32 at a time on the device); the serial work is one step per foreign element only.  In the home phase of a
# bucket its in-place residents are passed over without moving.

def pass_dest_runs(dig):
    n = len(dig)
    cnt = [0] * 256
    for d in dig:
        cnt[d] += 1
    cur, en, run = [0] * 256, [0] * 256, 0
    for d in range(256):
        cur[d] = run
        run += cnt[d]
        en[d] = run
    dest = list(range(n))
    for kk in range(256):
        kb, ke = cur[kk], en[kk]
        while kb < ke:
            while kb < ke and dig[kb] == kk:      # home scan: residents in place
                kb += 1
            if kb >= ke:
                break
            frm, d = kb, dig[kb]
            while True:
                pos = cur[d]
                nm = pos
                while dig[nm] == d:               # the run of residents behind the landing slot moves up by one
                    dest[nm] = nm + 1
                    nm += 1
                    assert nm < en[d]
                dest[frm] = pos
                cur[d] = nm + 1
                frm, d = nm, dig[nm]
                if d == kk:
                    break
            dest[frm] = kb
            kb += 1
        cur[kk] = kb
    return dest


def check_runs_form():
    rng = np.random.default_rng(9)
    bad = 0
    for n in (2, 5, 40, 300, 2000):
        for nb in (2, 3, 6, 40, 256):
            for skew in (0, 1):
                for _ in range(10):
                    pool = rng.choice(256, nb, replace=False)
                    if skew:          # one dominant bucket: long runs of residents
                        p = np.full(nb, 0.1 / max(nb - 1, 1)); p[0] = 0.9 if nb > 1 else 1.0; p /= p.sum()
                        dig = [int(v) for v in rng.choice(pool, n, p=p)]
                    else:
                        dig = [int(v) for v in rng.choice(pool, n)]
                    a, _ = pass_dest_walk(dig)
                    b = pass_dest_runs(dig)
                    if a != b:
                        bad += 1
    print("run-jump form mismatches:", bad)


if __name__ == "__main__":
    check_runs_form()


# ---- fourth formulation: foreign elements only + one boundary per bucket (tried in k_seed_sort: bit-exact, slower; kept as a model) ---
# Arrivals into bucket l during the phases of earlier buckets consume l's foreign slots in order; when l's own phase starts its cursor
# stands right behind the last consumed one (B_l).  Every resident of l below B_l has been pushed up by exactly one slot, every
# resident at or above B_l stays: the shifts need no serial work at all.  The serial part visits foreign elements only (landing slot =
# the bucket's cursor, evicted element = the next foreign slot at or after it).

def pass_dest_foreign(dig):
    n = len(dig)
    cnt = [0] * 256
    for d in dig:
        cnt[d] += 1
    cur, st, en, run = [0] * 256, [0] * 256, [0] * 256, 0
    for d in range(256):
        cur[d] = st[d] = run
        run += cnt[d]
        en[d] = run
    def next_foreign(d, p):
        while p < en[d] and dig[p] == d:
            p += 1
        return p
    dest = list(range(n))
    bound = [0] * 256
    for kk in range(256):
        bound[kk] = cur[kk]                       # B_kk: where the home phase starts
        kb = next_foreign(kk, cur[kk])
        while kb < en[kk]:
            frm, d = kb, dig[kb]
            while True:
                pos = cur[d]
                nm = next_foreign(d, pos)
                assert nm < en[d]
                dest[frm] = pos
                cur[d] = nm + 1
                frm, d = nm, dig[nm]
                if d == kk:
                    break
            dest[frm] = kb
            kb = next_foreign(kk, kb + 1)
    for d in range(256):                          # the shifts, in parallel on the device
        for x in range(st[d], en[d]):
            if dig[x] == d and x < bound[d]:
                dest[x] = x + 1
    return dest


def check_foreign_form():
    rng = np.random.default_rng(19)
    bad = 0
    for n in (2, 5, 40, 300, 2000):
        for nb in (2, 3, 6, 40, 256):
            for skew in (0, 1):
                for _ in range(10):
                    pool = rng.choice(256, nb, replace=False)
                    if skew and nb > 1:
                        p = np.full(nb, 0.1 / (nb - 1)); p[0] = 0.9
                        dig = [int(v) for v in rng.choice(pool, n, p=p)]
                    else:
                        dig = [int(v) for v in rng.choice(pool, n)]
                    a, _ = pass_dest_walk(dig)
                    if pass_dest_foreign(dig) != a:
                        bad += 1
    print("foreign-only form mismatches:", bad)


if __name__ == "__main__":
    check_foreign_form()


# ---- homopolymer-compressed minimizers (MM_I_HPC, sketch.c:92-101) ---------------------------------------------------------------
# Every run of equal unambiguous bases is ONE element (code, position of its last base); every ambiguous base is an element of its own.
# mm_sketch's loop then does on the elements exactly what it does on positions without HPC, with the k-mer span = distance between the
# ends of the k-th previous element and this one (the sum of the last k run lengths, sketch.c:98-100), valid only below 256.

def hpc_elements(seq: bytes):
    code = [NT4.get(c, 4) for c in seq]
    n = len(code)
    el = []          # (code, end position)
    i = 0
    while i < n:
        c = code[i]
        if c < 4:
            j = i
            while j + 1 < n and code[j + 1] == c:
                j += 1
            el.append((c, j))
            i = j + 1
        else:
            el.append((4, i))
            i += 1
    return el


def sketch_model_hpc(seq: bytes, w: int, k: int, rid: int = 0):
    el = hpc_elements(seq)
    n = len(el)
    mask = (1 << (2 * k)) - 1
    ix = [NONE] * n
    iz = [0] * n
    l = [0] * n
    run = 0
    for e in range(n):
        run = run + 1 if el[e][0] < 4 else 0
        l[e] = run
        if run >= k:
            f = r = 0
            for t in range(k - 1, -1, -1):
                c = el[e - t][0]
                f = (f << 2 | c) & mask
                r = (r >> 2) | (3 ^ c) << (2 * (k - 1))
            span = el[e][1] - (el[e - k][1] if e - k >= 0 else -1)
            if f != r and span < 256:
                z = 0 if f < r else 1
                ix[e] = hash64(r if z else f, mask) << 8 | span
                iz[e] = z
    def X(p):
        return ix[p] if p >= 0 else NONE
    out = []
    T1 = w + k - 1
    for i in range(n):
        cur = ix[i]
        mprev_x, mprev_p = NONE, -1
        for d in range(w, 0, -1):
            x = X(i - d)
            if x <= mprev_x:
                mprev_x, mprev_p = x, i - d
        mode = 0
        mx, mp = NONE, -1
        if cur <= mprev_x:
            mode = 2
        elif mprev_p == i - w:
            mode = 3
            for d in range(w - 1, -1, -1):
                x = X(i - d)
                if x <= mx:
                    mx, mp = x, i - d
        def emit(p):
            out.append((ix[p], rid << 32 | el[p][1] << 1 | iz[p]))
        if l[i] == T1 and mprev_x != NONE:
            for d in range(w - 1, 0, -1):
                if X(i - d) == mprev_x and i - d != mprev_p:
                    emit(i - d)
        if mode == 2:
            if l[i] >= T1 + 1 and mprev_x != NONE:
                emit(mprev_p)
        elif mode == 3:
            if l[i] >= T1:
                emit(mprev_p)
            if l[i] >= T1 and mx != NONE:
                for d in range(w - 1, -1, -1):
                    if X(i - d) == mx and i - d != mp:
                        emit(i - d)
        if i == n - 1:
            fx, fp = (cur, i) if mode == 2 else (mx, mp) if mode == 3 else (mprev_x, mprev_p)
            if fx != NONE:
                emit(fp)
    return np.array(out, dtype=np.uint64).reshape(-1, 2)


def check_hpc():
    import pyrefseed as rs
    rng = np.random.default_rng(23)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    cases = [acgt[rng.integers(0, 4, n)].tobytes() for n in (1, 5, 19, 20, 40, 300, 3000)]
    # homopolymer-rich sequences
    def hp(n):
        out = bytearray()
        while len(out) < n:
            out += bytes([acgt[rng.integers(0, 4)]]) * int(rng.geometric(0.4))
        return bytes(out[:n])
    cases += [hp(n) for n in (50, 500, 4000)]
    s = bytearray(hp(3000))
    for p in (10, 11, 500, 1500, 1501, 2999):
        s[p] = ord("N")
    cases.append(bytes(s))
    cases.append(b"A" * 400 + hp(300) + b"C" * 300 + hp(500))      # runs longer than 255: span >= 256 invalidates k-mers
    cases.append(b"AC" * 300)
    bad = 0
    for w, k in ((10, 19), (5, 19), (10, 15), (19, 19)):
        for c in cases:
            a = rs.sketch(c, w, k, hpc=True)
            b = sketch_model_hpc(c, w, k)
            if a.shape != b.shape or not np.array_equal(a, b):
                bad += 1
                print("HPC MISMATCH w", w, "k", k, "len", len(c), a.shape, b.shape)
    print("hpc sketch model mismatches:", bad)


if __name__ == "__main__":
    check_hpc()
