"""GPU parity of the KSSD (--fast) sketch kernel (K2) against the reference's golden kssd.hash.sketch
and the CPU oracle (oracle/tclust_oracle.c:oc_kssd_sketch, pinned to the reference binary)."""
import os
import struct

import numpy as np
import pytest

from rabbittclust_b200 import fasta, sketch_io, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _records(name):
    return [r[2] for r in fasta.read_fasta(os.path.join(GOLD, "mst_genomes", name + ".fa.gz"))]


def test_kssd_matches_reference_golden(ctx, oracle):
    data = open(os.path.join(GOLD, "kssd", "kssd.hash.sketch"), "rb").read()
    _, half_k, half_subk, drlevel, ngen = struct.unpack_from("<iiiii", data, 0)
    pos, want = 20, []
    for _ in range(ngen):
        (m,) = struct.unpack_from("<Q", data, pos); pos += 8
        want.append(np.frombuffer(data, "<u4", m, pos).astype(np.uint64)); pos += 4 * m
    genomes = [_records("g%02d" % i) for i in range(ngen)]
    got = ctx.sketch_kssd(genomes, 2 * half_k, drlevel, oracle.kssd_shuffled_dim(half_subk))
    for i in range(ngen):
        assert np.array_equal(got[i], want[i]), i


@pytest.mark.parametrize("kmer,drlevel", [(18, 3), (17, 3), (20, 3), (24, 3), (26, 3), (18, 2), (22, 4), (16, 1)])
def test_kssd_vs_oracle_parameters(ctx, oracle, kmer, drlevel):
    # covers 32-bit and 64-bit (half_k - drlevel > 8) hashes, odd k (rounded up), several reduction levels
    rng = np.random.default_rng(kmer * 10 + drlevel)
    n = 400000 if drlevel >= 3 else 60000
    g = [[synth.random_genome(n, rng)],
         [synth.random_genome(n // 2, rng).lower(), synth.random_genome(n // 3, rng)],
         [synth.random_genome(5000, rng)] * 3]
    _, half_subk, _ = oracle.kssd_params(kmer, drlevel)
    got = ctx.sketch_kssd(g, kmer, drlevel, oracle.kssd_shuffled_dim(half_subk), max_hashes=8192)
    for recs, h in zip(g, got):
        want = oracle.kssd_sketch(recs, kmer, drlevel)
        assert np.array_equal(h, want), (len(h), len(want))
    assert len(got[0]) > 0


def test_kssd_non_acgt_resets_and_edge_cases(ctx, oracle):
    rng = np.random.default_rng(9)
    base = synth.random_genome(300000, rng)
    b = bytearray(base)
    for p in rng.integers(0, len(b), 400):
        b[p] = ord("N")
    for p in rng.integers(0, len(b), 50):
        b[p] = rng.integers(1, 256)
    g = [
        [bytes(b)],                                   # N and arbitrary bytes restart the window
        [b"ACGT"],                                    # shorter than the k-mer
        [b"", synth.random_genome(17, rng), synth.random_genome(18, rng)],
        [b"N" * 1000],
        [base[:100000], base[:100000]],               # duplicates across records collapse
        [synth.random_genome(2500000, rng)],          # several segments
    ]
    got = ctx.sketch_kssd(g, 18, 3, oracle.kssd_shuffled_dim(6), max_hashes=4096)
    for recs, h in zip(g, got):
        want = oracle.kssd_sketch(recs, 18, 3)
        assert np.array_equal(h, want), (len(h), len(want))
    assert len(got[1]) == 0 and len(got[3]) == 0


def test_kssd_overflow_is_an_error(ctx, oracle, tc):
    rng = np.random.default_rng(2)
    g = [[synth.random_genome(400000, rng)]]
    with pytest.raises(tc.TclustError):
        ctx.sketch_kssd(g, 18, 3, oracle.kssd_shuffled_dim(6), max_hashes=16)


def test_kssd_packed_transport_is_exact(tc, oracle, monkeypatch):
    """TCLUST_SKETCH_PACK=1 (2 bits per base over PCIe + exception list, rebuilt on the device): lower case, N runs (which
    reset the KSSD window), other bytes, several batches -- same sketches as the ASCII transport and as the oracle."""
    rng = np.random.default_rng(19)
    genomes = []
    for g in range(4):
        recs = []
        for r in range(1 + g % 2):
            b = bytearray(synth.random_genome(int(rng.integers(400000, 700000)), rng))
            if g % 2 == 0:
                b[2000:9000] = bytes(b[2000:9000]).lower()
            for p in rng.integers(0, len(b) - 200, 8):
                b[p:p + int(rng.integers(1, 120))] = b"N" * 1
            for p in rng.integers(0, len(b), 30):
                b[p] = b"RYKMnx-\x00\xff"[int(rng.integers(0, 9))]
            recs.append(bytes(b))
        genomes.append(recs)
    sd = oracle.kssd_shuffled_dim(6)
    monkeypatch.setenv("TCLUST_SKETCH_CHUNK_MB", "1")
    monkeypatch.setenv("TCLUST_SKETCH_PACK", "0")
    with tc.Context(0) as c:
        plain = c.sketch_kssd(genomes, 18, 3, sd)
    monkeypatch.setenv("TCLUST_SKETCH_PACK", "1")
    monkeypatch.setenv("TCLUST_PACK_THREADS", "5")
    with tc.Context(0) as c:
        packed = c.sketch_kssd(genomes, 18, 3, sd)
    for a, b in zip(plain, packed):
        assert np.array_equal(a, b)
    assert np.array_equal(packed[0], oracle.kssd_sketch(genomes[0], 18, 3))


def _fasta_records(path):
    recs, cur = [], None
    for line in open(path, "rb").read().split(b"\n"):
        if line.startswith(b">"):
            cur = []
            recs.append(cur)
        elif cur is not None:
            cur.append(line)
    return [b"".join(r) for r in recs]


@pytest.mark.parametrize("k", [13, 15, 16])
def test_kssd_per_record_mode_matches_reference_golden(ctx, oracle, k):
    """tc_sketch_kssd_seq against the reference's own clust-mst --fast -i run (tools/make_golden_kssd_seq.py, --drlevel 0:
    every window is in the sketch): N runs of many lengths, valid runs of k-2..k+2 bases, records shorter than k.  For
    the odd sizes the first window after every reset keeps one stale base and the shift uses the un-rounded k."""
    d = os.path.join(GOLD, "kssd_seq")
    gold = sketch_io.read_kssd_hash_sketch(os.path.join(d, "k%d" % k, "kssd.hash.sketch"))
    recs = _fasta_records(os.path.join(d, "multi.fa"))
    assert len(recs) == len(gold["sketches"])
    got = ctx.sketch_kssd([[r] for r in recs], k, 0, oracle.kssd_shuffled_dim(6), max_hashes=4096, per_record=True)
    for r, a, b in zip(recs, got, gold["sketches"]):
        assert np.array_equal(a, b), (len(r), len(a), len(b))


@pytest.mark.parametrize("k,drlevel", [(17, 3), (19, 3), (21, 2), (15, 0), (31, 3), (18, 3)])
def test_kssd_per_record_mode_vs_oracle(ctx, oracle, monkeypatch, k, drlevel):
    """larger records (several segments per record, several batches): gaps at segment borders, at the start and the end of
    records, long gaps (the stale base lies far back), lower case, other codes; odd and even k"""
    monkeypatch.setenv("TCLUST_SKETCH_CHUNK_MB", "1")
    rng = np.random.default_rng(100 + k)
    recs = []
    for r in range(5):
        b = bytearray(synth.random_genome(int(rng.integers(300000, 900000)), rng))
        for p in rng.integers(0, len(b) - 5000, 60):
            L = int(rng.choice([1, 2, k - 1, k, k + 1, 100, 3000]))
            b[p:p + L] = b"N" * L
        for p in rng.integers(0, len(b), 40):
            b[p] = b"RYKMnx-"[int(rng.integers(0, 7))]
        if r % 2:
            b[1000:50000] = bytes(b[1000:50000]).lower()
        if r == 1:
            b[:7] = b"NNNNNNN"
        if r == 2:
            b[-9:] = b"N" * 9
        # gaps exactly around multiples of the segment width the planner uses (>= 16384 positions, 2048-aligned)
        for q in (16384, 32768, 65536, 131072, 262144):
            if q + 40 < len(b):
                b[q - 3:q - 1] = b"NN"
                b[q + k - 2:q + k - 1] = b"N"
        recs.append(bytes(b))
    recs += [b"ACGT" * 3, synth.random_genome(k - 1, rng), synth.random_genome(k, rng), synth.random_genome(k + 1, rng), b"N" * 50]
    _, half_subk, _ = oracle.kssd_params(k, drlevel)
    cap = 1 << 20 if drlevel == 0 else 1 << 16
    got = ctx.sketch_kssd([[r] for r in recs], k, drlevel, oracle.kssd_shuffled_dim(half_subk), max_hashes=cap, per_record=True)
    for i, (r, a) in enumerate(zip(recs, got)):
        want = oracle.kssd_sketch_seq(r, k, drlevel)
        assert np.array_equal(a, want), (i, len(r), len(a), len(want))
    if k % 2 == 0:
        same = ctx.sketch_kssd([[r] for r in recs], k, drlevel, oracle.kssd_shuffled_dim(half_subk), max_hashes=cap)
        assert all(np.array_equal(a, b) for a, b in zip(got, same))
