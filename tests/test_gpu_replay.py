"""Gate A on the GPU: replaying the reference's own mutation tables through the
CUDA path (C ABI -> ms_load_records -> ms_apply) must reproduce the reference's
FASTA and VCF byte for byte; randomized tables are checked against the C oracle."""
import numpy as np
import pytest

from tests.helpers import MS_CASES, engine_for, load_case, load_muts, vcf_body

pytestmark = pytest.mark.gpu


def _tables(case):
    return [[dict(key=m.key, type=m.type, start=m.start, stop=m.stop, reverse=m.reverse, alt=m.alt, insert=m.insert)
             for m in muts] for muts in load_muts(case)]


@pytest.mark.parametrize("case", MS_CASES)
def test_replay_reference_tables_bit_exact(case):
    from mutation_simulator_b200 import records as R
    d, contigs = load_case(case)
    eng, genome, goff, lens = engine_for(contigs)
    recs, lit = R.build_records(genome, goff, lens, _tables(case))
    eng.load_records(recs, lit)
    fb, vb = eng.apply()
    want_fa = (d / "out.fa").read_bytes()
    want_vcf = vcf_body((d / "out.vcf").read_bytes())
    assert fb == len(want_fa) and vb == len(want_vcf)
    assert eng.fasta() == want_fa
    assert eng.vcf() == want_vcf
    # idempotence: applying again gives the same image
    eng.apply()
    assert eng.fasta() == want_fa
    eng.close()


@pytest.mark.parametrize("bpl,alphabet,n0", [(60, b"ACGT", 1_500_000), (7, b"ACGT", 200_000),
                                            (16, b"ACGTNRYKMSWBDHV", 300_000), (61, b"ACGTN", 400_000),
                                            (100000, b"ACGT", 300_000)])
def test_replay_randomized_against_c_oracle(bpl, alphabet, n0):
    from mutation_simulator_b200 import records as R
    from oracle import c_oracle, pyref
    from tests.test_emu import oracle_tables
    rng = np.random.default_rng(bpl)
    seqs = [bytes(rng.choice(np.frombuffer(alphabet, np.uint8), n)) for n in (n0, 33_333, 17, 1, 5_000)]
    contigs = [(b"c%d" % i, b"c%d some description" % i, s, bpl) for i, s in enumerate(seqs)]
    rates = [0.02, 0.004, 0.004, 0.003, 0.003, 0.003, 0.003]
    tables = oracle_tables(seqs, rates, [1, 1, 1, 2, 1, 1, 1], [1, 12, 40, 30, 25, 20, 20], [1] * 7, seed=bpl)
    muts = [[pyref.Mut(key=t["key"], type=t["type"], start=t["start"], stop=t["stop"], reverse=t["reverse"],
                       alt=t["alt"], insert=t["insert"]) for t in tb] for tb in tables]
    want_fa, want_vcf = c_oracle.mutate_genome(contigs, muts)
    eng, genome, goff, lens = engine_for(contigs)
    recs, lit = R.build_records(genome, goff, lens, tables)
    eng.load_records(recs, lit)
    eng.apply()
    assert eng.fasta() == want_fa
    assert eng.vcf() == want_vcf
    assert np.array_equal(eng.contig_out_len(), [len(pyref.walk(s, b"x", m)[0]) for s, m in zip(seqs, muts)])
    eng.close()


def test_no_mutations_is_a_rewrap_of_the_input():
    from oracle import pyref
    rng = np.random.default_rng(3)
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), n)) for n in (100_000, 60, 61, 0, 1234)]
    contigs = [(b"n%d" % i, b"n%d" % i, s, 60) for i, s in enumerate(seqs)]
    eng, *_ = engine_for(contigs)
    from mutation_simulator_b200.records import REC_DTYPE
    eng.load_records(np.zeros(0, dtype=REC_DTYPE))
    eng.apply()
    want, _ = pyref.mutate_genome(contigs, [[] for _ in contigs])
    assert eng.fasta() == want
    assert eng.vcf() == b""
    eng.close()


def test_overlapping_records_are_rejected():
    from mutation_simulator_b200 import records as R
    from mutation_simulator_b200._lib import MutSimError
    contigs = [(b"a", b"a", b"ACGT" * 100, 60)]
    eng, genome, goff, lens = engine_for(contigs)
    tables = [[dict(key=10, type="DE", start=10, stop=30, reverse=False, alt=None, insert=None),
               dict(key=20, type="SN", start=20, stop=20, reverse=False, alt=b"A", insert=None)]]
    recs, lit = R.build_records(genome, goff, lens, tables)
    eng.load_records(recs, lit)
    with pytest.raises(MutSimError):
        eng.apply()
    eng.close()
