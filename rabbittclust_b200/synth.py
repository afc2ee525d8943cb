"""Seeded synthetic inputs of the shapes BASELINE.json names."""
import numpy as np


def family_sketches(n_families, per_family, s=1000, seed=7, keep_lo=0.5, keep_hi=1.0, ragged=None):
    """Config-3 model: family pool = s random u64; a member keeps a uniform keep_lo..keep_hi share
    of the pool plus private random u64 up to exactly s distinct values, sorted ascending.
    ragged=(lo, hi): member sizes log-uniform in [lo, hi] instead (config-4 style)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_families):
        pool_size = s if ragged is None else ragged[1]
        pool = np.unique(rng.integers(0, 2**64 - 4, pool_size, dtype=np.uint64))
        for _ in range(per_family):
            size = s if ragged is None else int(np.exp(rng.uniform(np.log(ragged[0]), np.log(ragged[1]))))
            keep = int(min(len(pool), size) * rng.uniform(keep_lo, keep_hi))
            part = rng.choice(pool, keep, replace=False)
            while True:
                priv = rng.integers(0, 2**64 - 4, size - keep + 8, dtype=np.uint64)
                sk = np.unique(np.concatenate([part, priv]))
                if len(sk) >= size:
                    break
            # drop private extras so that the sketch has exactly `size` values incl. all of `part`
            extra = np.setdiff1d(sk, part)[: size - keep]
            out.append(np.sort(np.concatenate([part, extra])).astype(np.uint64))
    return out


def family_sketches_fast(n_families, per_family, s=1000, seed=7):
    """Vectorised generator of the same model for large N (bench): returns (hashes, off).
    Pool values are even, private values odd (so they never collide); each member keeps a
    uniform 50-100 % random subset of its family pool; rows sorted ascending, distinct."""
    rng = np.random.default_rng(seed)
    n = n_families * per_family
    out = np.empty((n, s), np.uint64)
    for f in range(n_families):
        while True:
            pool = rng.integers(0, 2**63 - 2, s, dtype=np.uint64) * np.uint64(2)
            if len(np.unique(pool)) == s:
                break
        priv = rng.integers(0, 2**63 - 2, (per_family, s), dtype=np.uint64) * np.uint64(2) + np.uint64(1)
        keep = rng.integers(s // 2, s + 1, per_family)
        order = np.argsort(rng.random((per_family, s), dtype=np.float32), axis=1)
        mask = np.arange(s)[None, :] < keep[:, None]
        rows = np.where(mask, pool[order], priv)
        rows.sort(axis=1)
        out[f * per_family:(f + 1) * per_family] = rows
    # private values are random 63-bit numbers: a duplicate inside one row is a ~1e-13 event; fix if it happens
    dup = np.nonzero((out[:, 1:] == out[:, :-1]).any(axis=1))[0]
    for r in dup:
        while len(np.unique(out[r])) < s:
            out[r] = np.sort(np.where(np.r_[False, out[r][1:] == out[r][:-1]],
                                      rng.integers(0, 2**63 - 2, s, dtype=np.uint64) * np.uint64(2) + np.uint64(1),
                                      out[r]))
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(s)
    return np.ascontiguousarray(out.reshape(-1)), off


def random_genome(length, rng):
    return np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, length)].tobytes()


def mutate(genome, rate, rng):
    g = np.frombuffer(genome, np.uint8).copy()
    idx = np.nonzero(rng.random(len(g)) < rate)[0]
    g[idx] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, len(idx))]
    return g.tobytes()


def genome_families(n_families, per_family, length, seed=12345, rates=(0.001, 0.005, 0.01, 0.02, 0.05)):
    """Config-1 model: seed genome iid uniform ACGT; member m>0 = substitutions at rate rates[m%len]."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_families):
        base = random_genome(length, rng)
        out.append(base)
        for m in range(1, per_family):
            out.append(mutate(base, rates[m % len(rates)], rng))
    return out


def incremental_tie_set(seed=1, n_reps=2000, n_queries=400, size=99, query_size=None, threshold=0.08, kmer=18,
                        containment=False, hash_bits=31):
    """Sketch set that stresses the tie rule of the reference's incremental greedy clustering (the winner among equal
    distances is the first candidate in the iteration order of its phmap candidate table): n_reps mutually unrelated
    representatives of `size` hashes that all hold one common hash (so every query touches ALL of them and the table
    grows past its reserved capacity), followed by n_queries sketches that share EXACTLY the same number of hashes with
    2-4 randomly chosen representatives (equal sizes -> equal distances), sometimes with unequal shares, sometimes with
    none (they become representatives, and later queries tie between those too).
    -> (list of ascending np.uint64 sketches, n_reps)"""
    import math
    rng = np.random.default_rng(seed)
    qs = size if query_size is None else query_size
    x = math.exp(-threshold * kmer)
    jmin = x / (2.0 - x)
    common_hash = 7
    want = n_reps * (size - 1) + n_queries * qs
    pool = np.unique(rng.integers(100, 1 << hash_bits, size=want + want // 4 + 1000, dtype=np.int64))
    assert len(pool) >= want
    pool = rng.permutation(pool)[:want].astype(np.uint64)
    reps = [np.sort(np.concatenate([[common_hash], pool[i * (size - 1):(i + 1) * (size - 1)]]).astype(np.uint64)) for i in range(n_reps)]
    priv = pool[n_reps * (size - 1):]
    out = list(reps)
    sources = list(range(n_reps))                 # sketches a query may draw from (representatives; later also new ones)
    p = 0
    for qi in range(n_queries):
        m = int(rng.choice([0, 2, 2, 3, 4]))
        chosen = rng.choice(len(sources), size=m, replace=False) if m else []
        # smallest share that passes the filter for these sizes (Mash: c / (|q| + |r| - c) >= jmin; containment: c / min)
        if containment:
            cmin = int(math.ceil(jmin * min(qs, size))) + 1
        else:
            cmin = int(math.ceil(jmin * (qs + size) / (1.0 + jmin))) + 1
        cmax = (qs - 1) // max(m, 1)
        parts = [[common_hash]]
        if m and cmax >= cmin:
            c = int(rng.integers(cmin, cmax + 1))
            unequal = rng.random() < 0.25
            for j, si in enumerate(chosen):
                src = out[sources[int(si)]]
                src = src[src != common_hash]
                cj = c - (j if unequal else 0)
                parts.append(rng.choice(src, size=min(cj, len(src)), replace=False))
        have = sum(len(x) for x in parts)
        need = qs - have
        parts.append(priv[p:p + need])
        p += need
        q = np.unique(np.concatenate([np.asarray(x, np.uint64) for x in parts]))
        out.append(q)
        if m == 0:
            sources.append(len(out) - 1)
    return out, n_reps


def kssd_like_families(seed=3, n_fam=12, per_fam=10, pool=480, keep_lo=0.5, keep_hi=1.0, hash_bits=24):
    """KSSD-like sketch families in a small hash space (2^24, as k = 18 / drlevel 3 gives): members keep a random
    keep_lo..keep_hi share of a family pool of `pool` hashes plus up to 60 private ones -- many equal sizes and equal
    Jaccard values, i.e. ties for the greedy rules.  -> list of ascending np.uint64 sketches, family by family"""
    rng = np.random.default_rng(seed)
    sk = []
    for f in range(n_fam):
        p = rng.choice(1 << hash_bits, pool, replace=False)
        for m in range(per_fam):
            keep = rng.choice(p, int(pool * rng.uniform(keep_lo, keep_hi)), replace=False)
            priv = rng.integers(0, 1 << hash_bits, int(rng.integers(0, 60)))
            sk.append(np.unique(np.concatenate([keep, priv])).astype(np.uint64))
    return sk
