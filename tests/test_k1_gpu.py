"""GPU parity of MinHash sketch construction (K1) against the reference's golden sketches and the
CPU oracle."""
import glob
import os

import numpy as np
import pytest

from rabbittclust_b200 import fasta, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_genomes():
    out = {}
    for d in ("genomes", "ecoli"):
        for p in sorted(glob.glob(os.path.join(GOLD, d, "*.fa.gz"))):
            nm = os.path.basename(p).split(".")[0].replace("_150k", "")
            out[nm] = [r[2] for r in fasta.read_fasta(p)]
    return out


@pytest.mark.parametrize("k,s", [(21, 1000), (15, 200), (32, 64), (17, 1000)])
def test_sketch_matches_reference_golden(ctx, k, s):
    gold = np.load(os.path.join(GOLD, "sketches.npz"))
    gen = golden_genomes()
    names = sorted(gen)
    got = ctx.sketch_minhash([gen[n] for n in names], k=k, s=s)     # one batch, 8 genomes
    for nm, h in zip(names, got):
        assert np.array_equal(h, gold["%s_k%d_s%d" % (nm, k, s)]), (nm, k, s)


@pytest.mark.parametrize("k", [8, 11, 16, 19, 24, 25, 31])
def test_sketch_all_word_counts_vs_oracle(ctx, oracle, k):
    rng = np.random.default_rng(k)
    g = [[synth.random_genome(7000, rng)], [synth.random_genome(3000, rng).lower(), b"ACGTNNRYKMacgtn" * 30]]
    got = ctx.sketch_minhash(g, k=k, s=150)
    for recs, h in zip(g, got):
        assert np.array_equal(h, oracle.minhash_sketch(recs, k, 150))


def test_sketch_edge_cases(ctx, oracle):
    rng = np.random.default_rng(77)
    g = [
        [b"ACGT"],                                             # shorter than k: empty sketch
        [synth.random_genome(21, rng)],                        # exactly one window
        [b"A" * 5000],                                         # one distinct k-mer repeated
        [synth.random_genome(200, rng)] * 5,                   # identical records: duplicates across records
        [bytes(rng.integers(0, 256, 4000, dtype=np.uint8).tolist()).replace(b"\0", b"x")],   # arbitrary bytes
        [synth.random_genome(100000, rng)],                    # several segments, merged
        [synth.random_genome(40, rng), b"", synth.random_genome(25, rng)],
    ]
    got = ctx.sketch_minhash(g, k=21, s=1000)
    for recs, h in zip(g, got):
        want = oracle.minhash_sketch(recs, 21, 1000)
        assert np.array_equal(h, want), (len(h), len(want))
    assert len(got[0]) == 0 and len(got[1]) == 1 and len(got[2]) == 1


def test_sketch_variable_and_large_sizes(ctx, oracle):
    # containment mode: sketch size per genome = max(bytes/compress, 100); includes a size that needs
    # the global-memory candidate buffer
    rng = np.random.default_rng(5)
    lens = [12000, 150000, 600000, 60000]
    sizes = [100, 3000, 30000, 120]
    g = [[synth.random_genome(L, rng)] for L in lens]
    got = ctx.sketch_minhash(g, k=19, s=sizes)
    for recs, h, s in zip(g, got, sizes):
        assert np.array_equal(h, oracle.minhash_sketch(recs, 19, s))


def test_sketch_then_cluster_matches_reference_run(ctx, oracle):
    """config-1 shape in miniature: the reference's own clust-mst run on tests/golden/mst_genomes."""
    from rabbittclust_b200 import sketch_io
    d = os.path.join(GOLD, "mst_list")
    hs = sketch_io.read_hash_sketch(os.path.join(d, "hash.sketch"))
    _, info = sketch_io.read_info_sketch(os.path.join(d, "info.sketch"))
    genomes = [[r[2] for r in fasta.read_fasta(os.path.join(GOLD, "mst_genomes", rec["name"] + ".fa.gz"))]
               for rec in info]
    got = ctx.sketch_minhash(genomes, k=15, s=1000)
    for a, b in zip(got, hs["sketches"]):
        assert np.array_equal(a, b)
    ctx.upload(got)
    mst = ctx.mst(15, 0.05)
    ref = sketch_io.read_edge_mst(os.path.join(d, "edge.mst"))
    assert np.array_equal(np.sort(mst["dist"]), np.sort(ref["dist"]))
    _, labels, _ = oracle.clusters_from_mst(mst, len(info), 0.05)
    mine = sorted(sorted(info[i]["name"] for i in np.nonzero(labels == c)[0]) for c in range(labels.max() + 1))
    ref_c = sketch_io.parse_cluster_file(os.path.join(d, "clusters.txt"))
    assert mine == sorted(sorted(n for _, n in c) for c in ref_c)


def test_sketch_pipelined_batches(ctx, oracle, monkeypatch):
    """host-buffer entry point with 1 MB batches: 9 genomes of odd lengths -> several batches over the two staging
    slots (copy of batch c+1 overlaps the kernels of batch c); offsets deliberately not 16-byte aligned"""
    monkeypatch.setenv("TCLUST_SKETCH_CHUNK_MB", "1")
    rng = np.random.default_rng(123)
    lens = [400003, 700001, 5, 1300007, 250000, 999, 1048577, 31, 610001]
    g = [[synth.random_genome(L, rng)] if L > 1000 else [synth.random_genome(L, rng), synth.random_genome(77, rng)]
         for L in lens]
    got = ctx.sketch_minhash(g, k=21, s=500)
    for recs, h in zip(g, got):
        assert np.array_equal(h, oracle.minhash_sketch(recs, 21, 500))
    # the same input in one batch gives the same sketches
    monkeypatch.setenv("TCLUST_SKETCH_CHUNK_MB", "4096")
    again = ctx.sketch_minhash(g, k=21, s=500)
    assert all(np.array_equal(a, b) for a, b in zip(got, again))


def test_packed_transport_is_exact(tc, oracle, monkeypatch):
    """TCLUST_SKETCH_PACK=1: the host packs every batch to 2 bits per base + an exception list, the device rebuilds the
    bytes (pack2bit.cpp, k1_unpack_kernel).  Lower case, N runs, IUPAC codes, arbitrary bytes, several records per genome
    and several batches: the sketches are the ones of the ASCII transport and of the oracle."""
    rng = np.random.default_rng(9)
    genomes = []
    for g in range(5):
        recs = []
        for r in range(1 + g % 3):
            b = bytearray(synth.random_genome(int(rng.integers(200000, 500000)), rng))
            if g % 2:
                b[1000:5000] = bytes(b[1000:5000]).lower()
            for p in rng.integers(0, len(b) - 100, 6):
                b[p:p + int(rng.integers(1, 60))] = b"N" * 1
            for p in rng.integers(0, len(b), 40):
                b[p] = b"RYKMSWnxz*-\x00\xff"[int(rng.integers(0, 13))]
            recs.append(bytes(b))
        genomes.append(recs)
    monkeypatch.setenv("TCLUST_SKETCH_CHUNK_MB", "1")
    monkeypatch.setenv("TCLUST_SKETCH_PACK", "0")
    with tc.Context(0) as c:
        plain = c.sketch_minhash(genomes, k=21, s=1000)
    monkeypatch.setenv("TCLUST_SKETCH_PACK", "1")
    monkeypatch.setenv("TCLUST_PACK_THREADS", "3")
    with tc.Context(0) as c:
        l0 = c.launches
        packed = c.sketch_minhash(genomes, k=21, s=1000)
        assert c.launches > l0
    for a, b in zip(plain, packed):
        assert np.array_equal(a, b)
    for recs, h in list(zip(genomes, packed))[:2]:
        assert np.array_equal(h, oracle.minhash_sketch(recs, 21, 1000))
