"""Model check (CPU) of the batched lane-parallel chain walk of k_bt_walk_mid (bt_walk_body / flush_batch in
mm2-gb_b200/csrc/backtrack_kernels.cuh) against the reference's sequential mg_chain_backtrack (lchain.c:9-76, restated
literally in `sequential`).

The device collects the chain ends that are unclaimed when the scan meets them into batches of 32 (nothing is claimed in
between), walks every end of a batch independently against the claimed flags of that moment (<= 32 nodes), and commits the
results in visiting order: an end that has been claimed by an earlier end of the batch is dropped, an end whose evaluated
nodes were claimed in the meantime -- or whose path is longer than 32 nodes -- is walked again sequentially at its turn.
The chains (score, members) must be exactly those of the sequential algorithm."""
import random


def bk_end(max_drop, key, i0, f, p, t):
    """mg_chain_bk_end: the node the chain from i0 is cut at, and the last node the walk evaluated"""
    i, max_i, max_s = i0, i0, 0
    if i < 0 or t[i]:
        return i, 0
    while True:
        i = p[i]
        s = key if i < 0 else key - f[i]
        if s > max_s:
            max_s, max_i = s, i
        elif max_s - s > max_drop:
            break
        if i < 0 or t[i]:
            break
    return max_i, max_s


def take(end, key, f, p, t, min_cnt, min_sc, max_drop, chains):
    if t[end]:
        return
    cut, _ = bk_end(max_drop, key, end, f, p, t)
    mem, i = [], end
    while i != cut:
        mem.append(i)
        t[i] = 1
        i = p[i]
    sc = key if i < 0 else key - f[i]
    if sc >= min_sc and len(mem) > 0 and len(mem) >= min_cnt:
        chains.append((sc, tuple(mem)))


def sequential(z, f, p, min_cnt, min_sc, max_drop):
    t, chains = [0] * len(f), []
    for key, end in reversed(z):
        take(end, key, f, p, t, min_cnt, min_sc, max_drop, chains)
    return chains


def batched(z, f, p, min_cnt, min_sc, max_drop, width=32, maxlen=32):
    t, chains, pend = [0] * len(f), [], []

    def flush():
        spec = []
        for key, end in pend:                       # every "lane" on its own, against the flags of now
            path, cur, long_ = [], end, False
            for step in range(maxlen):
                path.append(cur)
                if step >= 1 and (cur < 0 or t[cur]):
                    break
                if step == maxlen - 1:
                    long_ = True
                    break
                cur = p[cur]
            cut, maxs, nev = 0, 0, 0
            if not long_:
                for j in range(1, len(path)):
                    s = key if path[j] < 0 else key - f[path[j]]
                    nev = j
                    if s > maxs:
                        maxs, cut = s, j
                    elif maxs - s > max_drop:
                        break
            spec.append((path, long_, cut, maxs, nev))
        for (key, end), (path, long_, cut, maxs, nev) in zip(pend, spec):   # commit in visiting order
            if t[end]:
                continue
            changed = any(path[j] >= 0 and t[path[j]] for j in range(0, min(nev, len(path) - 2) + 1)) if not long_ else True
            if long_ or changed:
                take(end, key, f, p, t, min_cnt, min_sc, max_drop, chains)
                continue
            if cut > 0:
                for j in range(cut):
                    t[path[j]] = 1
                if maxs >= min_sc and cut >= min_cnt:
                    chains.append((maxs, tuple(path[:cut])))
        pend.clear()

    k = len(z) - 1
    while k >= 0:                                   # the scan: 32 sorted ends at a time
        window = [z[j] for j in range(k, max(k - 32, -1), -1)]
        unc = [(key, end) for key, end in window if not t[end]]
        if len(pend) + len(unc) > width:
            flush()
            continue                                # claims changed: look at this window again
        pend.extend(unc)
        k -= 32
    if pend:
        flush()
    return chains


def forest(rng, n, kind):
    p = [i - rng.randint(1, 5) if rng.random() < 0.85 and i > 0 else -1 for i in range(n)]
    p = [q if q >= 0 else -1 for q in p]
    if kind == "chainlike":
        f = [15 * (i + 1) - rng.randint(0, 40) if rng.random() > 0.15 else rng.randint(15, 200) for i in range(n)]
    elif kind == "few":
        f = [rng.choice([40, 41, 55, 70, 300]) for _ in range(n)]
    else:
        f = [rng.randint(40, 4000) for _ in range(n)]
    return f, p


def test_batched_walk_equals_sequential_backtrack():
    rng = random.Random(7)
    for trial in range(400):
        n = rng.choice([1, 2, 5, 40, 200, 1000])
        f, p = forest(rng, n, rng.choice(["chainlike", "few", "wide"]))
        min_cnt, min_sc, max_drop = rng.choice([(3, 40, 500), (1, 1, 500), (2, 60, 20), (3, 40, 1 << 30)])
        z = sorted(((f[i], i) for i in range(n) if f[i] >= min_sc), key=lambda e: (e[0], rng.random()))
        a = sequential(z, f, p, min_cnt, min_sc, max_drop)
        b = batched(z, f, p, min_cnt, min_sc, max_drop)
        assert a == b, (trial, n)
