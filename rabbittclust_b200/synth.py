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
