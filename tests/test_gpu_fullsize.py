"""BASELINE.json-sized cases.  C1 (10 Mbp) and a C5-shaped batch are compared bit-exactly with the
C oracle given the GPU-sampled table; the human-chromosome-sized contig is checked through
size-independent properties (length bookkeeping, VCF replay on a window, idempotence)."""
import numpy as np
import pytest

from tests.helpers import engine_for, recs_to_muts
from tests.test_gpu_sample import args_ranges, check_invariants, random_contigs, sample

pytestmark = pytest.mark.gpu


def test_c1_full_size_bit_exact_against_c_oracle():
    from oracle import c_oracle
    L = 10_000_000
    contigs = random_contigs([L], seed=12345)
    eng, genome, goff, _ = engine_for(contigs)
    ranges = args_ranges([L], [0.01, 0.001, 0.001, 0, 0, 0], [1, 1, 1, 2, 1, 1, 1], [1, 10, 10, 3, 2, 2, 2])
    assert ranges[0]["k"] == 120000                      # 120 000 candidates
    recs, lit = sample(eng, ranges, [1] * 7, 0.5, seed=42)
    assert 119_000 < len(recs) < 120_000                 # reference: 119 473 written (0.44 % lost to greedy rejection)
    check_invariants(recs, [L], [1] * 7)
    fb, vb = eng.apply()
    want_fa, want_vcf = c_oracle.mutate_genome(contigs, recs_to_muts(recs, lit, goff, seed=42))
    assert eng.fasta() == want_fa
    assert eng.vcf() == want_vcf
    eng.close()


def test_many_small_contigs_bit_exact_against_c_oracle():
    """C5 shape (5 kbp contigs, all mutation types), 4 000 contigs: every piece of the file image holds several contigs."""
    from oracle import c_oracle
    lens = [5000] * 3990 + [1, 2, 3, 59, 60, 61, 4999, 5001, 16, 15]
    contigs = random_contigs(lens, seed=5)
    eng, genome, goff, _ = engine_for(contigs)
    ranges = [r for r in args_ranges(lens, [0.01, 0.001, 0.001, 0.0005, 0.0005, 0.0005], [1, 1, 1, 2, 1, 1, 1],
                                     [1, 10, 10, 50, 50, 50, 50]) if r["k"] > 0 and r["stop"] - r["k"] > 0]
    recs, lit = sample(eng, ranges, [1] * 7, 2.0 / 3.0, seed=7)
    check_invariants(recs, lens, [1] * 7)
    eng.apply()
    want_fa, want_vcf = c_oracle.mutate_genome(contigs, recs_to_muts(recs, lit, goff, seed=7))
    assert eng.fasta() == want_fa
    assert eng.vcf() == want_vcf
    eng.close()


def test_chromosome_sized_contig_properties():
    """A 250 Mbp contig (GRCh38 chr1 size): output length bookkeeping, window replay, idempotence."""
    from mutation_simulator_b200.engine import Engine
    L = 250_000_000
    eng = Engine(0)
    eng.synth_genome(3, [L], [60], [b"chr1 big"], [b"chr1"], n_fraction=0.03, telomere_n=10000)
    ranges = args_ranges([L], [0.01, 0.001, 0.001, 0.0005, 0.0005, 0.0005], [1, 1, 1, 2, 1, 1, 1], [1, 10, 10, 50, 50, 50, 50])
    eng.set_ranges(ranges, [1] * 7, 1, 2.0 / 3.0)
    eng.sample(9)
    fb, vb = eng.apply()
    recs = eng.records()
    check_invariants(recs, [L], [1] * 7)
    delta = int(recs["prod"].astype(np.int64).sum() - recs["cons"].astype(np.int64).sum())
    out_len = int(eng.contig_out_len()[0])
    assert out_len == L + delta
    hdr = len(b">chr1 big\n")
    assert fb == hdr + out_len + out_len // 60            # single contig: no trailing separator
    # out positions are consistent with a host-side exclusive scan of the deltas
    d = recs["prod"].astype(np.int64) - recs["cons"].astype(np.int64)
    out = recs["pos"].astype(np.int64) + np.concatenate(([0], np.cumsum(d)[:-1]))
    assert np.array_equal(out, recs["out"].astype(np.int64))
    # spot-check the image: every 61st byte is a line break, nothing else is
    img = eng.download(0)
    body = img[hdr:]
    nl = np.flatnonzero(body == 10)
    assert len(nl) == out_len // 60 and np.array_equal(nl, np.arange(60, len(body), 61)[:len(nl)])
    # SNP records show their ALT base at out (first 2 Mbp window)
    snp = recs[(recs["kind"] == 1) & (recs["out"] < 2_000_000)]
    o = snp["out"].astype(np.int64)
    assert np.array_equal(body[o + o // 60], snp["alt"])
    first = img[:1_000_000].copy()
    eng.apply()                                           # idempotent
    assert np.array_equal(eng.download(0)[:1_000_000], first)
    st = eng.stats()
    assert st["counts"][5] == st["counts"][6] > 0
    eng.close()
