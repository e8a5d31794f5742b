"""Tile-sharded apply (ms_apply_window): the chunking of contigs that are larger than one GPU's share.
Every part runs on a context that holds the same genome and the same record table; the parts' byte
ranges, put together, must be the unsharded output byte for byte — on any number of parts, with cuts
that fall inside contigs, inside lines and between the records of one contig."""
import numpy as np
import pytest

from mutation_simulator_b200.engine import BUF_FASTA, BUF_VCF
from tests.helpers import engine_for

pytestmark = pytest.mark.gpu

RATES = [0.01, 0.002, 0.002, 0.001, 0.001, 0.001, 0.001]


def _ranges(lens):
    cdf = np.cumsum(np.array(RATES) / sum(RATES))
    cdf = (cdf / cdf[-1]).tolist()
    return [dict(contig=i, start=0, stop=int(n) - 1, k=int(int(n) * sum(RATES)), limit=int(n), cdf=cdf,
                 minlen=[1, 1, 1, 2, 1, 1, 1], maxlen=[1, 10, 10, 40, 40, 40, 40]) for i, n in enumerate(lens) if n > 200]


@pytest.mark.parametrize("lens,bpl", [([700_000], 60), ([150_000, 90_001, 33, 260_000], 61), ([40_000] * 9, 1000)])
@pytest.mark.parametrize("n_parts", [2, 3, 8])
def test_window_parts_assemble_to_the_unsharded_output(lens, bpl, n_parts):
    rng = np.random.default_rng(len(lens) * 100 + bpl)
    contigs = [(b"c%d" % i, b"c%d window test" % i, bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), n)), bpl)
               for i, n in enumerate(lens)]
    eng, genome, goff, L = engine_for(contigs)
    eng.set_ranges(_ranges(lens), [1] * 7, 1, 0.5)
    eng.sample(99)
    fb, vb = eng.apply()
    want_fa, want_vcf = eng.fasta(), eng.vcf()
    got_fa, got_vcf = bytearray(fb), bytearray(vb)
    covered_f = covered_v = 0
    for part in range(n_parts):
        eng.sample(99)                                   # the same table on every "rank"
        w = eng.apply_window(part, n_parts)
        assert (w["fasta_bytes"], w["vcf_bytes"]) == (fb, vb)
        (f0, f1), (v0, v1) = w["fasta"], w["vcf"]
        assert f0 == covered_f and v0 == covered_v       # the windows tile the outputs in order
        got_fa[f0:f1] = eng.download(BUF_FASTA)[f0:f1].tobytes()
        got_vcf[v0:v1] = eng.download(BUF_VCF)[v0:v1].tobytes()
        covered_f, covered_v = f1, v1
    assert (covered_f, covered_v) == (fb, vb)
    assert bytes(got_fa) == want_fa
    assert bytes(got_vcf) == want_vcf
    eng.close()
