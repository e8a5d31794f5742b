"""CPU emulation of the CUDA kernels' shared host/device logic (tests/emu/emu.cpp)
against known answers and the golden vectors.  This is how index math is checked
in the build container, which has no GPU; the product itself has no CPU path."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from mutation_simulator_b200 import records as R
from tests.helpers import MS_CASES, load_case, load_muts, vcf_body

HERE = Path(__file__).resolve().parent
SRC = HERE / "emu" / "emu.cpp"
LIB = HERE / "emu" / "_build" / "libemu.so"


@pytest.fixture(scope="session")
def emu():
    from tests.helpers import emu_lib
    return emu_lib()


def philox(emu, ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    emu.emu_philox(c, k, o)
    return list(o)


def test_philox_known_answers(emu):
    """Random123 kat_vectors for philox4x32-10."""
    assert philox(emu, [0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox(emu, [0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox(emu, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def prp(emu, seed, n, count, contig=0, idx=0):
    out = (C.c_uint32 * count)()
    emu.emu_prp(C.c_uint64(seed), C.c_uint32(contig), C.c_uint32(1), C.c_uint64(idx), C.c_uint32(n), C.c_uint32(count), out)
    return np.array(out)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 16, 17, 1000, 65537, 1_000_003])
def test_prp_is_a_permutation(emu, n):
    v = prp(emu, 123, n, n)
    assert np.array_equal(np.sort(v), np.arange(n))


def test_prp_small_domains_are_uniform(emu):
    """Every arrangement of the first two images must be about equally likely over keys."""
    from scipy import stats
    for n in (2, 3, 4, 5, 7):
        cnt = {}
        trials = 6000
        for s in range(trials):
            v = prp(emu, 1000 + s, n, min(n, 2))
            cnt[tuple(v)] = cnt.get(tuple(v), 0) + 1
        cells = n * (n - 1) if n > 1 else 1
        assert len(cnt) == cells
        chi2, p = stats.chisquare(list(cnt.values()))
        assert p > 1e-4, (n, cnt, p)


def test_prp_subset_is_uniform_over_positions(emu):
    from scipy import stats
    n, k = 100000, 2000
    hist = np.zeros(20)
    for s in range(20):
        v = prp(emu, s, n, k, contig=s)
        hist += np.histogram(v, bins=20, range=(0, n))[0]
    chi2, p = stats.chisquare(hist)
    assert p > 1e-3


def test_decimal_digit_count(emu):
    """POS / END / SVLEN widths of the VCF line size pass (vcf_writer.py:118-126 formats them with str())."""
    rng = np.random.default_rng(1)
    v = np.concatenate([[0, 1, 9, 2**32 - 1, 2**31], [10**e + d for e in range(1, 10) for d in (-1, 0, 1)],
                        [2**b + d for b in range(1, 32) for d in (-1, 0, 1)], rng.integers(0, 2**32, 20000)]).astype(np.uint32)
    out = np.zeros(len(v), np.uint32)
    emu.emu_ndigits(C.c_uint32(len(v)), v.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, [len(str(int(x))) for x in v])


def hypergeom_draws(emu, seed, N, K, n, count):
    out = (C.c_uint32 * count)()
    emu.emu_hypergeom(C.c_uint64(seed), C.c_uint32(N), C.c_uint32(K), C.c_uint32(n), C.c_uint32(count), out)
    return np.array(out, dtype=np.int64)


@pytest.mark.parametrize("N,K,n", [
    (20, 7, 5), (100, 50, 10), (1000, 3, 900), (56_000, 28_000, 768), (56_001, 27_999, 769), (2**31, 2**30, 400),
    (300_000, 100_000, 3_000),                                   # variance 444: exact walk, asymmetric split
    (2**31, 2**31 - 7, 2**30), (2**31, 2**30, 2**31 - 40),        # walked values beyond single precision
    (248_956_422, 124_478_211, 3_360_000), (1_000_000, 500_001, 40_000), (10_000_000, 3_333_333, 123_457),  # normal regime
])
def test_hypergeometric_split_sampler_matches_the_exact_law(emu, N, K, n):
    """The bucket counts of the in-order position sampler (util.py:104: random.sample(range(n), k)) rest on this
    sampler: chi-square against scipy's exact pmf in the inversion regime and in the rounded-normal regime."""
    from scipy import stats
    draws = 40_000
    x = hypergeom_draws(emu, 99, N, K, n, draws)
    lo, hi = max(0, n + K - N), min(n, K)
    assert x.min() >= lo and x.max() <= hi
    rv = stats.hypergeom(N, K, n)
    mean, sd = rv.mean(), rv.std()
    assert abs(x.mean() - mean) < 5 * sd / np.sqrt(draws) + 1e-9
    # cells: about 20 equal-probability groups of adjacent values
    support = np.arange(max(lo, int(mean - 8 * sd) - 1), min(hi, int(mean + 8 * sd) + 1) + 1)
    pmf = rv.pmf(support)
    edges, acc = [support[0]], 0.0
    for v, p in zip(support, pmf):
        acc += p
        if acc >= 0.05:
            edges.append(v + 1); acc = 0.0
    edges[-1] = support[-1] + 1 if edges[-1] <= support[-1] else edges[-1]
    if len(edges) < 3:
        return
    obs = np.histogram(x, bins=edges)[0].astype(float)
    exp = np.array([rv.cdf(edges[i + 1] - 1) - rv.cdf(edges[i] - 1) for i in range(len(edges) - 1)])
    exp[0] += rv.cdf(edges[0] - 1); exp[-1] += rv.sf(edges[-1] - 1)
    obs[0] += (x < edges[0]).sum(); obs[-1] += (x >= edges[-1]).sum()
    chi2, p = stats.chisquare(obs, exp * draws)
    assert p > 1e-4, (N, K, n, p, obs, exp * draws)


def test_split_tree_bucket_counts(emu):
    """Counts add up to k, respect every bucket's capacity, and each bucket's count has the hypergeometric marginal."""
    from scipy import stats
    nb, n, k = 13, 50_000, 5_000
    vlo = np.array([(i * n + nb - 1) // nb for i in range(nb)] + [n], dtype=np.uint32)
    trials = 3000
    cnts = np.zeros((trials, nb), dtype=np.int64)
    for s in range(trials):
        out = (C.c_uint32 * nb)()
        emu.emu_split_counts(C.c_uint64(s), C.c_uint32(3), C.c_uint32(17), C.c_uint32(nb), vlo.ctypes.data_as(C.c_void_p), C.c_uint32(k), out)
        cnts[s] = out
    assert (cnts.sum(axis=1) == k).all()
    width = np.diff(vlo.astype(np.int64))
    assert (cnts <= width).all()
    for b in (0, 5, 12):
        rv = stats.hypergeom(n, int(width[b]), k)
        z = (cnts[:, b].mean() - rv.mean()) / (rv.std() / np.sqrt(trials))
        assert abs(z) < 4.5, (b, z)
        assert 0.9 < cnts[:, b].std() / rv.std() < 1.1
    # neighbouring buckets are negatively correlated like a multivariate hypergeometric: cov = -k w_i w_j (n-k) / (n^2 (n-1))
    c01 = np.cov(cnts[:, 0], cnts[:, 1])[0, 1]
    exp = -k * width[0] * width[1] * (n - k) / (n * n * (n - 1))
    assert abs(c01 - exp) < 6 * rv.var() / np.sqrt(trials)
    # degenerate: k == n fills every bucket
    out = (C.c_uint32 * nb)()
    emu.emu_split_counts(C.c_uint64(1), C.c_uint32(0), C.c_uint32(0), C.c_uint32(nb), vlo.ctypes.data_as(C.c_void_p), C.c_uint32(n), out)
    assert list(out) == list(width)


def run_emu_apply(emu, contigs, tables, tile_bytes=4096, vcf_text=None):
    seqs = [c[2] for c in contigs]
    genome, goff = R.pack_genome(seqs)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    bpl = np.array([c[3] for c in contigs], dtype=np.int32)
    hdr = b"".join(c[1] for c in contigs)
    hoff = np.cumsum([0] + [len(c[1]) for c in contigs]).astype(np.int64)
    names = b"".join(c[0] for c in contigs)
    noff = np.cumsum([0] + [len(c[0]) for c in contigs]).astype(np.int64)
    if vcf_text is not None:
        recs, lit = R.records_from_vcf(vcf_text, [c[0] for c in contigs], goff, lens)
    else:
        recs, lit = R.build_records(genome, goff, lens, tables)
    cap_f = int(lens.sum() * 3 + recs["prod"].sum() * 2 + 4096 + len(hdr) * 2)
    cap_v = int(64 * len(recs) + 4 * (recs["prod"].sum() + recs["cons"].sum()) + len(names) * len(recs) + 4096)
    fa = np.zeros(cap_f, dtype=np.uint8)
    vcf = np.zeros(cap_v, dtype=np.uint8)
    fl, vl, nf, ns = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emu.emu_apply(P(genome), C.c_int32(len(contigs)), P(goff), P(lens), P(bpl), hdr, P(hoff), names, P(noff),
                       P(recs), C.c_int64(len(recs)), P(lit), P(fa), C.c_int64(cap_f), C.byref(fl),
                       P(vcf), C.c_int64(cap_v), C.byref(vl), C.c_int64(tile_bytes), C.byref(nf), C.byref(ns))
    assert rc == 0
    return fa[:fl.value].tobytes(), vcf[:vl.value].tobytes(), nf.value, ns.value


def run_emu_tiles(emu, contigs, tables, tile_bytes=16384):
    """FASTA image through the second-generation tile kernel's core (ms_tile_core.h) emulated on the CPU."""
    seqs = [c[2] for c in contigs]
    genome, goff = R.pack_genome(seqs)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    bpl = np.array([c[3] for c in contigs], dtype=np.int32)
    hdr = b"".join(c[1] for c in contigs)
    hoff = np.cumsum([0] + [len(c[1]) for c in contigs]).astype(np.int64)
    recs, lit = R.build_records(genome, goff, lens, tables)
    cap_f = int(lens.sum() * 3 + recs["prod"].sum() * 2 + 4096 + len(hdr) * 2)
    fa = np.zeros(cap_f, dtype=np.uint8)
    fl, nc, nd, nf = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emu.emu_apply_tiles(P(genome), C.c_int32(len(contigs)), P(goff), P(lens), P(bpl), hdr, P(hoff),
                             P(recs), C.c_int64(len(recs)), P(lit), P(fa), C.c_int64(cap_f), C.byref(fl), C.c_int64(tile_bytes),
                             C.byref(nc), C.byref(nd), C.byref(nf))
    assert rc == 0
    return fa[:fl.value].tobytes(), nc.value, nd.value, nf.value


@pytest.mark.parametrize("case", MS_CASES)
@pytest.mark.parametrize("tile", [1024, 16384])
def test_emulated_tile_kernel_matches_reference(emu, case, tile):
    d, contigs = load_case(case)
    fa, nc, nd, nf = run_emu_tiles(emu, contigs, tables_from_golden(case), tile)
    assert fa == (d / "out.fa").read_bytes()
    if case in ("args_all", "c1_small"):
        assert nc > 2 * nd and nf == 0     # copies far outnumber generated payload pieces; no tile needs the generic path


def tables_from_golden(case):
    out = []
    for muts in load_muts(case):
        out.append([dict(key=m.key, type=m.type, start=m.start, stop=m.stop, reverse=m.reverse,
                         alt=m.alt, insert=m.insert) for m in muts])
    return out


@pytest.mark.parametrize("case", MS_CASES)
@pytest.mark.parametrize("tile", [256, 4096])
def test_emulated_splice_and_vcf_match_reference(emu, case, tile):
    d, contigs = load_case(case)
    fa, vcf, nf, ns = run_emu_apply(emu, contigs, tables_from_golden(case), tile)
    assert fa == (d / "out.fa").read_bytes()
    assert vcf == vcf_body((d / "out.vcf").read_bytes())
    if case in ("args_all", "c1_small"):
        assert nf > 2 * ns  # most groups take the vector path


def assert_fasta_equal_up_to_silent_gap_snps(got: bytes, want: bytes, case: str):
    """The reference's VCF determines its FASTA except in one case: a SNP drawn on a gap character '-' substitutes
    conv('-') = 'N' -> 'N' (mutator.py:75, :449-455), which changes the FASTA ('-' becomes 'N') but is not written
    to the VCF because REF == ALT (vcf_writer.py:123).  A replay from the VCF file keeps the '-'.  Those bytes — and
    only those — may differ, and each must be a SNP of the reference's own table (muts.json)."""
    assert len(got) == len(want)
    if got == want:
        return
    a, b = np.frombuffer(got, np.uint8), np.frombuffer(want, np.uint8)
    diff = np.flatnonzero(a != b)
    assert set(a[diff]) == {ord("-")} and set(b[diff]) == {ord("N")}, case
    n_silent = sum(1 for muts in load_muts(case) for m in muts if m.type == "SN" and m.alt == b"N")
    assert 0 < len(diff) <= n_silent, (case, len(diff), n_silent)


@pytest.mark.parametrize("case", MS_CASES)
def test_replay_of_the_reference_vcf_file_reproduces_fasta_and_vcf(emu, case):
    """Gate A from the reference's FILES (north star, part one): in.fa + out.vcf -> out.fa, and the VCF re-emitted
    from the parsed records is the reference's VCF byte for byte (records.records_from_vcf; CPU emulation of the
    kernels' cores — tests/test_gpu_replay.py does the same through the C ABI on the GPU)."""
    d, contigs = load_case(case)
    text = (d / "out.vcf").read_bytes()
    fa, vcf, _, _ = run_emu_apply(emu, contigs, None, 1024, vcf_text=text)
    assert vcf == vcf_body(text)
    assert_fasta_equal_up_to_silent_gap_snps(fa, (d / "out.fa").read_bytes(), case)


def test_vcf_loader_rejects_malformed_records():
    names, goff, lens = [b"c"], np.array([0, 100]), [100]
    ok = "c\t5\t.\tA\tAGG\t.\t.\tSVTYPE=INS;END=5;SVLEN=2\tGT\t1\n"
    recs, lit = R.records_from_vcf(ok, names, goff, lens)
    assert (int(recs[0]["pos"]), int(recs[0]["prod"]), int(recs[0]["cons"])) == (5, 2, 0) and bytes(lit[:2]) == b"GG"
    first = "c\t1\t.\tA\tGGA\t.\t.\tSVTYPE=INS;END=1;SVLEN=2\tGT\t1\n"       # pos 0: insert + REF (mutator.py:351-354)
    recs, lit = R.records_from_vcf(first, names, goff, lens)
    assert int(recs[0]["pos"]) == 0 and bytes(lit[:2]) == b"GG"
    for bad in ("x\t5\t.\tA\tC\t.\t.\t.\tGT\t1\n",                                # unknown contig
                "c\t5\t.\tA\tTGG\t.\t.\tSVTYPE=INS;END=5;SVLEN=2\tGT\t1\n",      # ALT does not extend REF
                "c\t500\t.\tA\tC\t.\t.\t.\tGT\t1\n",                              # outside the contig
                "c\t5\t.\tAC\tG\t.\t.\tSVTYPE=DEL;END=6;SVLEN=1\tGT\t1\n",       # ALT is not REF[0]
                "c\t5\t.\tA\tC\t.\t.\tSVTYPE=CNV;END=6;SVLEN=1\tGT\t1\n"):
        with pytest.raises(ValueError):
            R.records_from_vcf(bad, names, goff, lens)


def test_bedpe_loader_inverts_the_writer():
    from mutation_simulator_b200.bedpe_writer import breakpoints_from_rows, rows
    from tests.helpers import GOLDEN
    import json
    for case in ("it_basic", "rmt_it"):
        d = GOLDEN / case
        bp = json.loads((d / "bp.json").read_text())
        from tests.helpers import read_fasta_simple
        src = d / ("out.fa" if (d / "out.fa").exists() else "in.fa")      # rmt_it: IT runs on the mutated genome
        contigs = read_fasta_simple(src)
        names, lens = [c[0] for c in contigs], [len(c[2]) for c in contigs]
        partners, bps = breakpoints_from_rows((d / "out.bedpe").read_bytes(), names, lens)
        assert partners == {int(k): v for k, v in bp["partners"].items() if k in bp["breakpoints"]}
        for k, v in bp["breakpoints"].items():
            assert list(bps[int(k)]["self"]) == v["self"] and list(bps[int(k)]["partner"]) == v["partner"], (case, k)
        # and forward again
        text = b"".join(rows(names[c].decode(), bps[c]["self"], lens[c], names[partners[c]].decode(), bps[c]["partner"],
                             lens[partners[c]]) for c in sorted(bps))
        assert sorted(text.splitlines()) == sorted((d / "out.bedpe").read_bytes().splitlines())


def oracle_tables(seqs, rates, minlen, maxlen, block, seed, titv=1.0):
    """Sample with the C oracle's sampler; return per-contig typed tables."""
    from oracle import c_oracle
    tables = []
    cdf = np.cumsum(np.array(rates) / sum(rates)).tolist()
    for ci, s in enumerate(seqs):
        L = len(s)
        k = int(L * sum(rates))
        rng = dict(start=0, stop=L - 1, k=k, cdf=cdf, minlen=minlen, maxlen=maxlen)
        m, pool = c_oracle.sample_contig(bytes(s), [rng], block, min(block), titv, seed + ci)
        t = []
        for r in m:
            d = dict(key=int(r["key"]), type=R.TYPE_NAME[int(r["type"])], start=int(r["start"]), stop=int(r["stop"]),
                     reverse=bool(r["reverse"]), alt=bytes([r["alt"]]), insert=None)
            if d["type"] == "IN":
                n = d["stop"] - d["start"] + 1
                d["insert"] = pool[int(r["lit_off"]):int(r["lit_off"]) + n]
            t.append(d)
        tables.append(t)
    return tables


@pytest.mark.parametrize("bpl,alphabet", [(60, b"ACGT"), (7, b"ACGT"), (16, b"ACGTNRYKMSWBDHV"), (61, b"ACGTN"), (1000, b"ACGT")])
def test_emulated_apply_matches_c_oracle_randomized(emu, bpl, alphabet):
    from oracle import c_oracle, pyref
    rng = np.random.default_rng(bpl)
    seqs = [bytes(rng.choice(np.frombuffer(alphabet, np.uint8), n)) for n in (120_000, 33_333, 17, 5_000)]
    contigs = [(b"c%d" % i, b"c%d some description" % i, s, bpl) for i, s in enumerate(seqs)]
    rates = [0.02, 0.004, 0.004, 0.003, 0.003, 0.003, 0.003]
    tables = oracle_tables(seqs, rates, [1, 1, 1, 2, 1, 1, 1], [1, 12, 40, 30, 25, 20, 20], [1] * 7, seed=bpl)
    muts = [[pyref.Mut(key=d["key"], type=d["type"], start=d["start"], stop=d["stop"], reverse=d["reverse"],
                       alt=d["alt"], insert=d["insert"]) for d in t] for t in tables]
    want_fa, want_vcf = c_oracle.mutate_genome(contigs, muts)
    fa, vcf, nf, ns = run_emu_apply(emu, contigs, tables, tile_bytes=1024)
    assert fa == want_fa
    assert vcf == want_vcf
    for tile in (1024, 16384):
        fa2, nc, nd, nfb = run_emu_tiles(emu, contigs, tables, tile)
        assert fa2 == want_fa, (bpl, tile)
        if tile == 1024:
            assert nfb == 0                # (at 16 KiB these very dense tables exceed the per-tile record pool)
