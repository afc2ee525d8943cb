"""GPU parity of the KSSD (--fast) sketch kernel (K2) against the reference's golden kssd.hash.sketch
and the CPU oracle (oracle/tclust_oracle.c:oc_kssd_sketch, pinned to the reference binary)."""
import os
import struct

import numpy as np
import pytest

from rabbittclust_b200 import fasta, synth

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
