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


def test_c3_shaped_rmt_hot_cold_ranges_and_blocked_centromeres(tmp_path):
    """BASELINE config 3 shape, scaled: 3 contigs / 60 Mbp, ~1 200 alternating hot/cold ranges and a blocked (None)
    centromere per contig, through the real RMT parser and range planner.  Checks per-range rates, zero starts and
    zero extents in blocked ranges, and the length bookkeeping."""
    from mutation_simulator_b200 import plan
    from mutation_simulator_b200.engine import Engine
    from mutation_simulator_b200.rmt import SimulationSettings
    lens = [30_000_000, 20_000_000, 10_000_000]
    rng = np.random.default_rng(8)
    lines = ["titv=2.0", "std", "it None", "sn 0.001", ""]
    blocked, hot = [], []
    for ci, L in enumerate(lens):
        lines.append(f"chr {ci+1}")
        cen_lo, cen_hi = int(L * 0.4), int(L * 0.43)
        pos = 10_000
        k = 0
        while pos + 120_000 < L:
            n = int(rng.integers(10_000, 100_000))
            if pos < cen_hi and pos + n > cen_lo:          # the centromere: one None range
                lines.append(f"{cen_lo+1}-{cen_hi} None")
                blocked.append((ci, cen_lo, cen_hi - 1))
                pos = cen_hi + int(rng.integers(1, 5000))
                continue
            if k % 2 == 0:
                lines.append(f"{pos+1}-{pos+n} sn 0.05 in 0.005 inmin 1 inmax 10 de 0.005 demin 1 demax 60 tl 0.004 tlmin 5 tlmax 40")
                hot.append((ci, pos, pos + n - 1))
            else:
                lines.append(f"{pos+1}-{pos+n} sn 0.0001")
            pos += n + int(rng.integers(0, 3000))
            k += 1
    p = tmp_path / "c3.rmt"
    p.write_text("\n".join(lines) + "\n")

    class FakeFasta:                                   # the settings model only needs names and lengths
        def keys(self):
            return [f"c{i}" for i in range(len(lens))]

        def __getitem__(self, i):
            n = lens[i if isinstance(i, int) else int(i[1:])]
            return type("R", (), {"__len__": lambda s: n})()
    sim = SimulationSettings.from_rmt(p, FakeFasta(), True)
    arr, n_ranges = plan.build_ranges(sim, lens)
    assert n_ranges > 1000
    eng = Engine(0)
    names = [b"c0", b"c1", b"c2"]
    eng.synth_genome(11, lens, [60] * 3, names, names)
    eng.set_ranges_array(arr, n_ranges, plan.block_list(sim), min(sim.mut_block.values()), plan.p_transition(sim.titv))
    eng.sample(4)
    eng.apply()
    recs = eng.records()
    check_invariants(recs, lens, [1] * 7)
    pos = recs["pos"].astype(np.int64)
    ext = pos + np.maximum(recs["cons"].astype(np.int64), 1)
    for ci, lo, hi in blocked:
        m = recs["contig"] == ci
        assert not ((pos[m] >= lo) & (pos[m] <= hi)).any(), "start inside a blocked range"
        assert not ((pos[m] < lo) & (ext[m] > lo)).any(), "extent reaches into a blocked range"
    # hot ranges carry ~0.064 candidates per base, a few percent of them lost to first-come rejection
    got = exp = 0
    for ci, lo, hi in hot[::7]:
        m = (recs["contig"] == ci) & (pos >= lo) & (pos <= hi)
        got += int(m.sum()); exp += int((hi - lo + 1) * 0.064)
    assert 0.80 * exp < got <= exp, (got, exp)
    d = recs["prod"].astype(np.int64) - recs["cons"].astype(np.int64)
    for ci, L in enumerate(lens):
        assert int(eng.contig_out_len()[ci]) == L + int(d[recs["contig"] == ci].sum())
    eng.close()


def _compare_with_oracle(eng, names, lens, bpl, seed):
    """Every contig of the resident genome through the C oracle with the device's own records; FASTA image and VCF body
    must be byte-identical.  Bases come back from HBM one contig at a time."""
    import ctypes as C
    from tests.helpers import oracle_contig_from_recs
    recs, lit = eng.records(), eng.literals()
    image, vcf = eng.download(0), eng.download(1)
    goff = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    written = C.c_int64(0)
    fpos = vpos = 0
    lo = np.searchsorted(recs["contig"], np.arange(len(lens)), "left"); hi = np.searchsorted(recs["contig"], np.arange(len(lens)), "right")
    for ci, L in enumerate(lens):
        seq = eng.read_genome(int(goff[ci]), int(L)).tobytes()
        fa, lines = oracle_contig_from_recs(seq, names[ci], names[ci], bpl, recs[lo[ci]:hi[ci]], lit, int(goff[ci]), seed, ci, written)
        assert image[fpos:fpos + len(fa)].tobytes() == fa, f"FASTA of contig {ci} differs from the oracle"
        assert vcf[vpos:vpos + len(lines)].tobytes() == lines, f"VCF of contig {ci} differs from the oracle"
        fpos += len(fa); vpos += len(lines)
    assert fpos == len(image) and vpos == len(vcf)


def test_chromosome_sized_contig_bit_exact_against_c_oracle():
    """GRCh38 chr1 size (250 Mbp, N runs, all mutation types) sampled and applied on the GPU, then walked by the C
    oracle from the same records: both files byte for byte (full-size oracle compare, not only invariants)."""
    from mutation_simulator_b200.engine import Engine
    L = 250_000_000
    eng = Engine(0)
    eng.synth_genome(3, [L], [60], [b"chr1"], [b"chr1"], n_fraction=0.03, telomere_n=10000)
    ranges = args_ranges([L], [0.01, 0.001, 0.001, 0.0005, 0.0005, 0.0005], [1, 1, 1, 2, 1, 1, 1], [1, 10, 10, 50, 50, 50, 50])
    eng.set_ranges(ranges, [1] * 7, 1, 2.0 / 3.0)
    eng.sample(21)
    eng.apply()
    _compare_with_oracle(eng, [b"chr1"], [L], 60, 21)
    eng.close()


def test_c3_full_size_rmt_bit_exact_against_c_oracle():
    """BASELINE config 3 in full: the GRCh38-shaped 3.09 Gbp genome, 10 k hot / cold RMT ranges + blocked centromeres through
    from_rmt -> plan.build_ranges (bench.py's generator), sampled and applied on the GPU; every contig is then walked by
    the C oracle from the device's records and compared byte for byte; no record starts or reaches into a None range."""
    import bench
    from mutation_simulator_b200 import plan
    from mutation_simulator_b200.engine import Engine
    from mutation_simulator_b200.rmt import SimulationSettings
    wl = bench.WORKLOADS["c3"]
    lens, names = list(wl["lengths"]), [n.encode() for n in wl["names"]]
    text, n_explicit = bench.c3_rmt_text(lens, wl["n_rmt_ranges"], wl["n_fraction"])
    assert n_explicit > 9000
    import tempfile, os
    with tempfile.NamedTemporaryFile("w", suffix=".rmt", delete=False) as fh:
        fh.write(text)
    try:
        sim = SimulationSettings.from_rmt(__import__("pathlib").Path(fh.name), bench.LenFasta(wl["names"], lens), True)
    finally:
        os.unlink(fh.name)
    arr, n = plan.build_ranges(sim, lens)
    eng = Engine(0)
    eng.synth_genome(bench.GENOME_SEED, lens, [60] * len(lens), names, names, wl["n_fraction"], wl["telomere"])
    eng.set_ranges_array(arr, n, plan.block_list(sim), min(sim.mut_block.values()), plan.p_transition(sim.titv))
    eng.sample(5)
    eng.apply()
    recs = eng.records()
    pos = recs["pos"].astype(np.int64); ext = pos + np.maximum(recs["cons"].astype(np.int64), 1)
    for ci, L in enumerate(lens):                       # the centromere of every contig is a None range
        lo, hi = (L * 2) // 5, (L * 2) // 5 + int(wl["n_fraction"] * L)
        m = recs["contig"] == ci
        assert not ((pos[m] >= lo) & (pos[m] < hi)).any() and not ((pos[m] < lo) & (ext[m] > lo)).any()
    _compare_with_oracle(eng, names, lens, 60, 5)
    eng.close()
