"""Gate B/C on the GPU: fresh sampling (ms_sample) — structural invariants,
statistical parity with the reference's own sampling runs (tests/golden/stats_*.json),
determinism, and GPU-apply of GPU-sampled records against the C oracle."""
import json

import numpy as np
import pytest
from scipy import stats

from tests.helpers import GOLDEN, engine_for, recs_to_muts

pytestmark = pytest.mark.gpu

TYPES = ["SN", "IN", "DE", "IV", "DU", "TL", "TLI"]


def args_ranges(lens, rates6, minlen, maxlen):
    """One range per contig (ARGS mode, rmt.py:376-392); rates6 in dict order SN,IN,DE,IV,DU,TL (tl not yet halved)."""
    rates = list(rates6[:5]) + [rates6[5] / 2, rates6[5] / 2]        # rmt.py:91-94
    total = sum(rates)
    cdf = (np.cumsum(np.array(rates) / total) / np.cumsum(np.array(rates) / total)[-1]).tolist()
    out = []
    for ci, L in enumerate(lens):
        k = int(((L - 1 - 0) + 1) * total)                            # mutator.py:225
        out.append(dict(contig=ci, start=0, stop=L - 1, k=k, limit=L, cdf=cdf, minlen=minlen, maxlen=maxlen))
    return out


def random_contigs(lens, seed=0, bpl=60, alphabet=b"ACGT"):
    rng = np.random.default_rng(seed)
    return [(b"chr%d" % (i + 1), b"chr%d" % (i + 1), bytes(rng.choice(np.frombuffer(alphabet, np.uint8), n)), bpl)
            for i, n in enumerate(lens)]


def sample(eng, ranges, block, p_ti, seed):
    eng.set_ranges(ranges, block, min(block), p_ti)
    eng.sample(seed)
    return eng.records(), eng.literals()


def check_invariants(recs, lens, block):
    """SURVEY.md §4 Gate C: no overlaps, TL == TLI per contig, spacing."""
    for ci, L in enumerate(lens):
        r = recs[recs["contig"] == ci]
        pos = r["pos"].astype(np.int64)
        assert (np.diff(pos) > 0).all()
        ext = pos + np.maximum(r["cons"].astype(np.int64), 1)
        assert (ext <= L).all()
        assert (pos[1:] >= ext[:-1]).all(), "overlapping extents"
        assert (np.diff(pos) >= min(block) + 1).all()
        assert (r["type"] == 5).sum() == (r["type"] == 6).sum()


def test_sampling_invariants_and_apply_matches_oracle():
    from oracle import c_oracle
    lens = [400_000, 150_000, 2_000, 37]
    contigs = random_contigs(lens, seed=1, alphabet=b"ACGTN")
    eng, genome, goff, _ = engine_for(contigs)
    block = [1, 1, 1, 1, 1, 1, 1]
    ranges = args_ranges(lens, [0.01, 0.002, 0.002, 0.001, 0.001, 0.002], [1, 1, 1, 2, 1, 1, 1], [1, 10, 10, 30, 30, 20, 20])
    recs, lit = sample(eng, ranges, block, 2.0 / 3.0, seed=5)
    st = eng.stats()
    assert st["n_candidates"] == sum(r["k"] for r in ranges)
    assert 0.9 * st["n_candidates"] < len(recs) <= st["n_candidates"]
    check_invariants(recs, lens, block)
    eng.apply()
    want_fa, want_vcf = c_oracle.mutate_genome(contigs, recs_to_muts(recs, lit, goff, seed=5))
    assert eng.fasta() == want_fa
    assert eng.vcf() == want_vcf
    # determinism: same seed -> identical records; different seed -> different
    recs2, _ = sample(eng, ranges, block, 2.0 / 3.0, seed=5)
    assert recs2.tobytes() == recs.tobytes()
    recs3, _ = sample(eng, ranges, block, 2.0 / 3.0, seed=6)
    assert recs3.tobytes() != recs.tobytes()
    eng.close()


def test_results_do_not_depend_on_contig_partition():
    """RNG is keyed by the global contig id: sampling a contig alone gives the same records."""
    lens = [120_000, 80_000, 50_000]
    contigs = random_contigs(lens, seed=2)
    ranges = args_ranges(lens, [0.01, 0.002, 0.002, 0.001, 0.001, 0.002], [1, 1, 1, 2, 1, 1, 1], [1, 10, 10, 30, 30, 20, 20])
    eng, *_ = engine_for(contigs)
    recs, _ = sample(eng, ranges, [1] * 7, 0.5, seed=9)
    eng.close()
    from mutation_simulator_b200.engine import Engine
    for ci in (1, 2):
        e2 = Engine(0)
        c = contigs[ci]
        e2.upload_genome(np.frombuffer(c[2], np.uint8), [lens[ci]], [c[3]], [c[1]], [c[0]], gid=[ci])
        r1 = dict(ranges[ci]); r1["contig"] = 0
        part, _ = sample(e2, [r1], [1] * 7, 0.5, seed=9)
        whole = recs[recs["contig"] == ci]
        for f in ("pos", "cons", "prod", "kind", "type", "alt"):
            assert np.array_equal(part[f], whole[f]), f
        e2.close()


def gate(pvalues: dict, alpha: float = 0.01):
    """The north star's tolerance — chi-square / KS at p > 0.01 — applied FAMILY-WISE: the m tests of one
    configuration pass together iff every p exceeds alpha / m (Bonferroni), so that a correct sampler fails a
    configuration with probability <= 1 % however many statistics are checked."""
    m = len(pvalues)
    bad = {k: v for k, v in pvalues.items() if not v > alpha / m}
    assert not bad, (f"family of {m} tests at family-wise alpha={alpha}: threshold {alpha / m:.2e}", bad)


@pytest.mark.parametrize("name,rates6,minlen,maxlen,block", [
    ("stats_c1", [0.01, 0.001, 0.001, 0, 0, 0], [1, 1, 1, 2, 1, 1, 1], [1, 10, 10, 3, 2, 2, 2], [1] * 7),
    ("stats_all", [0.01, 0.001, 0.001, 0.0005, 0.0005, 0.0005], [1, 1, 1, 2, 1, 1, 1], [1, 10, 10, 50, 50, 50, 50], [1] * 7),
    ("stats_dense", [0.05, 0.02, 0.02, 0.01, 0.01, 0.04], [1, 1, 1, 2, 1, 1, 1], [1, 20, 40, 40, 40, 30, 30], [2, 1, 10, 1, 1, 1, 1]),
])
def test_statistics_match_reference_runs(name, rates6, minlen, maxlen, block):
    """Per-type counts (chi-square), SV length histograms (chi-square), TL/TLI reversed fraction (Fisher) — against the
    reference's own runs — and positional uniformity of the candidates of EVERY contig (Kolmogorov-Smirnov against
    util.py:94-109's uniform k-subset), at family-wise p > 0.01 (see gate())."""
    g = json.loads((GOLDEN / f"{name}.json").read_text())
    lens = g["lengths"]
    contigs = random_contigs(lens, seed=7)
    eng, *_ = engine_for(contigs)
    ranges = args_ranges(lens, rates6, minlen, maxlen)
    for ci in range(len(lens)):
        assert ranges[ci]["k"] == g["runs"][0]["contigs"][ci]["candidates"]
    titv = 2.0 if name == "stats_all" else 1.0
    ref_counts = np.zeros(7)
    ref_lens = {t: {} for t in TYPES}
    ref_rev = ref_tli = 0
    for run in g["runs"]:
        for c in run["contigs"]:
            ref_counts += [c["counts"].get(t, 0) for t in TYPES]
            for t, h in c["lens"].items():
                for k, v in h.items():
                    ref_lens[t][int(k)] = ref_lens[t].get(int(k), 0) + v
            ref_rev += c["tli_reversed"]
            ref_tli += c["counts"].get("TLI", 0)
    my_counts = np.zeros(7)
    my_lens = {t: {} for t in TYPES}
    my_rev = my_tli = 0
    pv = {}
    goff = np.concatenate(([0], np.cumsum(lens)))
    for seed in range(len(g["runs"])):
        recs, _ = sample(eng, ranges, block, titv * (1 / (titv + 1)), seed=100 + seed)
        if seed == 0:      # start positions of the candidates (before rejection): uniform over each contig
            gpos, _, _, _ = eng.debug_candidates()
            for ci, L in enumerate(lens):
                pos = gpos[(gpos >= goff[ci]) & (gpos < goff[ci + 1])] - goff[ci]
                assert len(pos) == ranges[ci]["k"]
                pv[f"KS positions contig {ci}"] = stats.kstest((pos + 0.5) / L, "uniform").pvalue
        check_invariants(recs, lens, block)
        my_counts += np.bincount(recs["type"], minlength=8)[:7]
        for ti, t in enumerate(TYPES):
            if t == "SN":
                continue
            sel = recs[recs["type"] == ti]
            ln = np.where(np.isin(sel["type"], [2, 3, 5]), sel["cons"], sel["prod"])
            for k, v in zip(*np.unique(ln, return_counts=True)):
                my_lens[t][int(k)] = my_lens[t].get(int(k), 0) + int(v)
        tli = recs[recs["type"] == 6]
        my_tli += len(tli)
        my_rev += int((tli["kind"] == 5).sum())
    keep = (ref_counts + my_counts) > 0
    pv["type counts"] = stats.chi2_contingency(np.vstack([ref_counts[keep], my_counts[keep]]))[1]
    for t in TYPES[1:]:
        if t == "TLI" or not ref_lens[t]:
            continue
        keys = sorted(set(ref_lens[t]) | set(my_lens[t]))
        a = np.array([ref_lens[t].get(k, 0) for k in keys]); b = np.array([my_lens[t].get(k, 0) for k in keys])
        pv[f"lengths {t}"] = stats.chi2_contingency(np.vstack([a, b]) + 1e-9)[1]
    if ref_tli:
        pv["TLI reversed"] = stats.fisher_exact([[ref_rev, ref_tli - ref_rev], [my_rev, my_tli - my_rev]])[1]
    gate(pv)
    eng.close()


def test_titv_ratio_matches_reference():
    g = json.loads((GOLDEN / "stats_titv.json").read_text())
    L = 400_000
    contigs = random_contigs([L], seed=11)
    eng, genome, goff, _ = engine_for(contigs)
    ranges = args_ranges([L], [0.05, 0, 0, 0, 0, 0], [1, 1, 1, 2, 1, 1, 1], [1, 2, 2, 3, 2, 2, 2])
    trans = {ord("A"): ord("G"), ord("G"): ord("A"), ord("C"): ord("T"), ord("T"): ord("C")}
    pv = {}
    for titv_s, res in g["result"].items():
        titv = float(titv_s)
        recs, _ = sample(eng, ranges, [1] * 7, titv * (1 / (titv + 1)), seed=int(titv * 10) + 1)
        sn = recs[recs["type"] == 0]
        assert (sn["ref"] == genome[sn["pos"]]).all()
        ti = int(sum((sn["alt"][i] == trans[sn["ref"][i]]) for i in range(len(sn))))
        ref_ti = ref_n = 0
        for base in "ACGT":
            for c in res[base]:
                ref_ti += c.get(chr(trans[ord(base)]), 0)
                ref_n += sum(c.values())
        pv[f"titv {titv}"] = stats.fisher_exact([[ref_ti, ref_n - ref_ti], [ti, len(sn) - ti]])[1]
        # ... and against the nominal probability itself (mutator.py:436: p_ti = titv * (1 / (titv + 1)))
        pv[f"titv {titv} nominal"] = stats.binomtest(ti, len(sn), titv * (1 / (titv + 1))).pvalue
        assert (sn["alt"] != sn["ref"]).all()
    gate(pv)
    eng.close()


def test_random_insert_bases_are_uniform_in_the_gpu_output():
    """mutator.py:466-471: every inserted base is an independent uniform draw from [A, T, G, C].  The bases are read
    back from the FASTA image the GPU wrote (at the records' output positions), not from a host re-computation:
    composition (chi-square, 3 dof), independence of neighbouring bases (16 cells) and uniformity of the first base
    of an insert — family-wise p > 0.01."""
    L = 3_000_000
    contigs = random_contigs([L], seed=21, bpl=100_000_000)       # one line: output base index == file offset - header
    eng, *_ = engine_for(contigs)
    ranges = args_ranges([L], [0.0, 0.02, 0, 0, 0, 0], [1, 1, 1, 2, 1, 1, 1], [1, 12, 2, 3, 2, 2, 2])
    recs, _ = sample(eng, ranges, [1] * 7, 0.5, seed=4)
    eng.apply()
    body = np.frombuffer(eng.fasta(), np.uint8)[len(contigs[0][1]) + 2:]
    recs = eng.records()
    ins = recs[recs["type"] == 1]
    assert len(ins) > 40_000 and (ins["prod"] >= 1).all() and (ins["prod"] <= 12).all()
    first = body[ins["out"]]
    allb = np.concatenate([body[o:o + n] for o, n in zip(ins["out"].astype(np.int64), ins["prod"].astype(np.int64))])
    assert set(np.unique(allb)) == set(b"ACGT")
    code = np.zeros(256, np.int64); code[list(b"ACGT")] = range(4)
    pv = {"composition": stats.chisquare(np.bincount(code[allb], minlength=4)).pvalue,
          "first base": stats.chisquare(np.bincount(code[first], minlength=4)).pvalue}
    two = ins[ins["prod"] >= 2]
    pairs = code[body[two["out"]]] * 4 + code[body[two["out"].astype(np.int64) + 1]]
    pv["neighbouring bases"] = stats.chisquare(np.bincount(pairs, minlength=16)).pvalue
    # lengths are uniform on [min, max] (mutator.py:229-265)
    pv["insert lengths"] = stats.chisquare(np.bincount(ins["prod"], minlength=13)[1:]).pvalue
    gate(pv)
    eng.close()


def test_rmt_ranges_blocked_regions_and_rates():
    """Multiple ranges per contig incl. blocked gaps: zero starts in blocked regions, SVs never
    extend into them (SURVEY.md Q3 stance), per-range accepted counts near the reference's."""
    g = json.loads((GOLDEN / "stats_rmt.json").read_text())
    L = g["lengths"][0]
    contigs = random_contigs([L], seed=13)
    eng, *_ = engine_for(contigs)

    def rng(start, stop, rates, minlen, maxlen, limit):
        tot = sum(rates)
        cdf = np.cumsum(np.array(rates) / tot); cdf = (cdf / cdf[-1]).tolist()
        return dict(contig=0, start=start, stop=stop, k=int(((stop - start) + 1) * tot), limit=limit, cdf=cdf,
                    minlen=minlen, maxlen=maxlen)
    one = [1, 1, 1, 2, 1, 1, 1]
    ranges = [
        rng(100000, 299999, [0.05, 0.005, 0, 0, 0, 0, 0], one, [1, 10, 1, 2, 1, 1, 1], 400000),
        rng(300000, 399999, [0.0001, 0, 0, 0, 0, 0, 0], one, [1, 1, 1, 2, 1, 1, 1], 400000),
        rng(450000, 699999, [0, 0, 0.01, 0, 0, 0.005, 0.005], one, [1, 1, 100, 2, 1, 50, 50], L),
        rng(700000, L - 1, [0.001, 0, 0, 0, 0, 0, 0], one, [1, 1, 1, 2, 1, 1, 1], L),
    ]
    ranges[2]["minlen"] = [1, 1, 1, 2, 1, 5, 5]
    ref_acc = np.zeros(len(ranges))
    for run in g["runs"]:
        rr = [r for r in run["contigs"][0]["ranges"] if r["k"] > 0]
        assert [r["k"] for r in rr] == [r["k"] for r in ranges]
        ref_acc += [r["accepted"] for r in rr]
    mine = np.zeros(len(ranges))
    for seed in range(len(g["runs"])):
        eng.set_ranges(ranges, [1] * 7, 1, 0.5)
        eng.sample(seed)
        gpos, typ, ln, acc = eng.debug_candidates()
        assert len(gpos) == sum(r["k"] for r in ranges)
        a = gpos[acc == 1]
        for i, r in enumerate(ranges):
            mine[i] += ((a >= r["start"]) & (a <= r["stop"])).sum()
        recs = eng.records()
        pos = recs["pos"].astype(np.int64)
        ext = pos + np.maximum(recs["cons"].astype(np.int64), 1)
        for lo, hi in [(0, 100000), (400000, 450000)]:
            assert not ((pos >= lo) & (pos < hi)).any()
            assert not ((pos < lo) & (ext > lo)).any()
    assert np.allclose(mine / ref_acc, 1.0, atol=0.02), (mine, ref_acc)
    eng.close()


def test_sample_too_dense_raises_like_random_sample():
    from mutation_simulator_b200._lib import MS_ERR_SAMPLE, MutSimError
    contigs = random_contigs([1000], seed=1)
    eng, *_ = engine_for(contigs)
    r = dict(contig=0, start=0, stop=999, k=600, limit=1000, cdf=[1.0] * 7, minlen=[1] * 7, maxlen=[1] * 7)
    with pytest.raises(MutSimError) as ei:
        eng.set_ranges([r], [1] * 7, 1, 0.5)
    assert ei.value.code == MS_ERR_SAMPLE
    eng.close()


def test_it_breakpoints_match_sampler_contract():
    """it_mutator.py:108-111: n sorted breakpoints in [1, len-1] with gaps >= 2 on each member."""
    lens = [30_000, 26_000, 41_000, 19_000]
    contigs = random_contigs(lens, seed=4)
    eng, *_ = engine_for(contigs)
    a, b = eng.it_breakpoints(3, [0, 1], [2, 3], [120, 80])
    assert len(a) == 200 and len(b) == 200
    for arr, L in ((a[:120], lens[0]), (b[:120], lens[2]), (a[120:], lens[1]), (b[120:], lens[3])):
        arr = arr.astype(np.int64)
        assert arr[0] >= 1 and arr[-1] <= L - 1
        assert (np.diff(arr) >= 2).all()
    a2, b2 = eng.it_breakpoints(3, [0, 1], [2, 3], [120, 80])
    assert np.array_equal(a, a2) and np.array_equal(b, b2)
    eng.close()


def test_long_random_inserts_match_oracle():
    """Inserts longer than the 32 bases cached in the record (K_RAND beyond its cache -> Philox blocks 0, 1, ...)."""
    from oracle import c_oracle
    lens = [150_000, 40_000]
    contigs = random_contigs(lens, seed=21)
    eng, genome, goff, _ = engine_for(contigs)
    ranges = args_ranges(lens, [0.004, 0.004, 0.001, 0, 0, 0], [1, 20, 1, 2, 1, 1, 1], [1, 150, 10, 3, 2, 2, 2])
    recs, lit = sample(eng, ranges, [1] * 7, 0.5, seed=31)
    ins = recs[recs["type"] == 1]
    assert (ins["prod"] > 64).sum() > 50 and (ins["prod"] <= 32).sum() > 10
    eng.apply()
    want_fa, want_vcf = c_oracle.mutate_genome(contigs, recs_to_muts(recs, lit, goff, seed=31))
    assert eng.fasta() == want_fa
    assert eng.vcf() == want_vcf
    eng.close()


def test_long_blocking_spans_use_the_scan_path_and_agree_with_chain_semantics():
    """maxlen far above the spacing (span > 4096 -> prefix-max scan path) and huge blocks: acceptance must still be
    the reference's first-come rule (mutator.py:184-213), checked by replaying the rule on the host."""
    lens = [600_000]
    contigs = random_contigs(lens, seed=33)
    eng, *_ = engine_for(contigs)
    for maxlen, block in ((6000, [1, 1, 50, 1, 1, 1, 1]), (30, [3, 1, 200, 7, 1, 1, 1])):
        ranges = args_ranges(lens, [0.002, 0.0005, 0.001, 0.0005, 0.0005, 0.001], [1, 1, 1, 2, 1, 1, 1],
                             [1, 5, maxlen, maxlen, maxlen, maxlen, maxlen])
        eng.set_ranges(ranges, block, min(block), 0.5)
        eng.sample(3)
        gpos, typ, ln, acc = eng.debug_candidates()
        last_hi = -1
        want = np.zeros(len(gpos), np.uint8)
        for j in range(len(gpos)):
            t = int(typ[j])
            if t == 255 or gpos[j] < last_hi:
                continue
            want[j] = 1
            p = int(gpos[j])
            if t in (0, 1):
                last_hi = p + 1 + block[t]
            elif t == 6:
                last_hi = 1 + block[6]
            else:
                last_hi = min(p + int(ln[j]) + block[t], lens[0])
        assert np.array_equal(acc, want)
    eng.close()


@pytest.mark.parametrize("group_min", [0, 1, 120_000])
def test_streamed_run_is_byte_identical_to_upload_sample_apply_download(group_min):
    """ms_mutate_streamed = ms_genome_upload + ms_sample + ms_apply + ms_download with the copies overlapped: same
    bytes for every contig grouping, lower-case input, palindromic inversions and N SNPs (records the VCF omits)."""
    from mutation_simulator_b200.engine import BUF_FASTA, BUF_VCF, Engine
    lens = [300_000, 70_001, 50_000, 3, 90_017, 1, 15, 220_000]
    contigs = random_contigs(lens, seed=5, alphabet=b"ACGTNacgtnRY", bpl=70)
    # a stretch of AT repeats makes reverse-complement palindromes (REF == ALT inversions are not written)
    c0 = bytearray(contigs[0][2]); c0[1000:9000] = b"AT" * 4000
    contigs[0] = (contigs[0][0], contigs[0][1], bytes(c0), 70)
    ranges = args_ranges(lens, [0.02, 0.004, 0.004, 0.006, 0.002, 0.004], [1, 1, 1, 2, 1, 1, 1], [1, 40, 10, 4, 30, 20, 20])
    block = [1, 1, 1, 1, 1, 1, 1]
    eng, genome, goff, _ = engine_for(contigs)
    eng.set_ranges(ranges, block, 1, 2 / 3)
    eng.sample(77)
    fb, vb = eng.apply()
    want_fa, want_vcf = eng.download(BUF_FASTA).tobytes(), eng.download(BUF_VCF).tobytes()
    recs_want = eng.records(include_dead=True)
    eng.close()

    eng = Engine(0)
    eng.declare_genome(lens, [70] * len(lens), [c[1] for c in contigs], [c[0] for c in contigs])
    eng.set_ranges(ranges, block, 1, 2 / 3)
    raw = np.frombuffer(b"".join(c[2] for c in contigs), dtype=np.uint8).copy()
    fa = np.zeros(fb + 100, np.uint8)
    vcf = np.zeros(vb + 100, np.uint8)
    for _ in range(2):   # second run reuses every buffer and event
        got = eng.mutate_streamed(77, raw, fa, vcf, group_min)
        assert got == (fb, vb)
        assert fa[:fb].tobytes() == want_fa
        assert vcf[:vb].tobytes() == want_vcf
    assert (eng.records(include_dead=True) == recs_want).all()
    assert eng.download(BUF_FASTA).tobytes() == want_fa       # the device copies are complete as well
    with pytest.raises(Exception):
        eng.mutate_streamed(77, raw, fa[:1000], vcf, group_min)
    eng.close()


def test_streamed_run_with_nothing_to_mutate():
    """Rates too low for a single candidate (and no ranges at all): the streamed call still returns the wrapped,
    upper-cased genome and an empty VCF body."""
    from mutation_simulator_b200.engine import Engine
    lens = [1000, 37]
    contigs = random_contigs(lens, seed=2, alphabet=b"ACGTacgt", bpl=50)
    raw = np.frombuffer(b"".join(c[2] for c in contigs), dtype=np.uint8).copy()
    want = b"".join(b">" + c[1] + b"\n" + b"\n".join(c[2].upper()[o:o + 50] for o in range(0, len(c[2]), 50)) + b"\n" for c in contigs)
    want = want[:-1]    # no line break after a partial last line at the end of the file (fasta_writer.py:44-47)
    for ranges in ([], args_ranges(lens, [1e-5, 0, 0, 0, 0, 0], [1] * 7, [1] * 7)):
        ranges = [r for r in ranges if r["k"] > 0]
        eng = Engine(0)
        eng.declare_genome(lens, [50, 50], [c[1] for c in contigs], [c[0] for c in contigs])
        eng.set_ranges(ranges, [1] * 7, 1, 0.5)
        fa, vcf = np.zeros(4096, np.uint8), np.zeros(64, np.uint8)
        fb, vb = eng.mutate_streamed(5, raw, fa, vcf)
        assert vb == 0 and fa[:fb].tobytes() == want
        eng.close()


def test_position_sampler_is_a_uniform_subset_in_every_regime():
    """ms_sample_positions = util.sample_with_minimum_distance (util.py:94-109) for many ranges at once.  The values are
    generated in order (hypergeometric bucket counts, then a uniform subset inside each sort bucket): distinct, sorted,
    inside the range, spaced by min_dist; saturated ranges (k == population) return every position; over many small
    ranges every 2-subset of a 6-value span is equally likely; a single-bucket sparse range, a dense multi-bucket range
    and one whose buckets are sparse (bitonic path) are uniform by KS."""
    from mutation_simulator_b200.engine import Engine
    eng = Engine(0)
    # saturated: k == n for bitonic-sized, small-kernel-sized and bitmap-sized populations, and k = n - 1
    n = np.array([1, 2, 5, 63, 64, 100, 129, 500, 2000, 2000], dtype=np.uint32)
    k = n.copy(); k[-1] = 1999
    out = eng.sample_positions(11, np.arange(len(n)), np.zeros(len(n)), n, k, 0)     # values in [start, stop) = [0, n)
    o = 0
    for ni, ki in zip(n, k):
        v = out[o:o + ki]; o += int(ki)
        assert (np.diff(v.astype(np.int64)) > 0).all() and v.min() >= 0 and v.max() < ni
        if ki == ni:
            assert np.array_equal(v, np.arange(ni))
    # min distance: sample_with_minimum_distance(start, stop, k, d): gaps > d, all inside [start, stop)
    v = eng.sample_positions(5, [0], [100], [100 + 50_000], [5_000], 3).astype(np.int64)
    assert v.min() >= 100 and v.max() < 100 + 50_000 and (np.diff(v) > 3).all()
    # all 15 two-subsets of a 6-value span, over 6000 independent ranges (keyed by contig id)
    R = 6000
    v = eng.sample_positions(7, np.arange(R), np.zeros(R), np.full(R, 6), np.full(R, 2), 0).reshape(R, 2)
    cells = np.bincount(v[:, 0] * 6 + v[:, 1], minlength=36)
    cells = cells[cells > 0]
    assert len(cells) == 15
    assert stats.chisquare(cells).pvalue > 1e-3
    # KS against the uniform law: sparse single bucket, dense multi-bucket, multi-bucket with sparse buckets
    for gid, (pop, kk) in enumerate([(2_000_000_000, 300), (5_000_000, 400_000), (2_000_000_000, 100_000)]):
        v = eng.sample_positions(3, [gid], [0], [pop], [kk], 0).astype(np.float64)
        assert len(np.unique(v)) == kk
        assert stats.kstest(v / pop, "uniform").pvalue > 1e-3, (pop, kk)
        # bucket counts are hypergeometric, not equal: the spread of per-bucket counts matches sqrt(mean)
        if kk >= 100_000:
            nb = -(-kk // 384)
            cnt = np.histogram(v, bins=nb, range=(0, pop))[0]
            ratio = cnt.var() / (cnt.mean() * (1 - 1 / nb) * (pop - kk) / (pop - 1))
            assert 0.7 < ratio < 1.4, ratio
    eng.close()
