"""The drop-in class API and CLI on the GPU: Mutator / ITMutator / main()."""
import json
import shutil
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import pyref
from tests.helpers import GOLDEN, read_fasta_simple, vcf_body, vcf_head

pytestmark = pytest.mark.gpu


def run_main(argv):
    from mutation_simulator_b200.__main__ import main
    old = sys.argv
    sys.argv = ["mutation-simulator"] + [str(a) for a in argv]
    try:
        main()
    finally:
        sys.argv = old


def replay_plain(contigs, vcf_text):
    """Reference-independent check of a produced (FASTA, VCF) pair: REF->ALT replay."""
    recs = {}
    for line in vcf_body(vcf_text).splitlines():
        f = line.split(b"\t")
        recs.setdefault(f[0], []).append((int(f[1]), f[3], f[4]))
    fa = pyref.FastaOut()
    for name, long_name, seq, bpl in contigs:
        fa.set_bpl(bpl)
        fa.write_header(long_name)
        pos, out = 0, []
        for p, ref, alt in recs.get(name, []):
            assert p - 1 >= pos, "overlapping VCF records"
            out.append(seq[pos:p - 1])
            assert seq[p - 1:p - 1 + len(ref)] == ref
            out.append(alt)
            pos = p - 1 + len(ref)
        out.append(seq[pos:])
        fa.write_multi(b"".join(out))
    return fa.getvalue()


def test_cli_args_mode_is_self_consistent_and_reproducible(tmp_path):
    shutil.copy(GOLDEN / "args_all" / "in.fa", tmp_path / "in.fa")
    cmd = json.loads((GOLDEN / "args_all" / "cmd.json").read_text())["argv"]
    argv = [tmp_path / "in.fa", "-o", tmp_path / "in", "-q", "--seed", "11"] + cmd[4:]
    run_main(argv)
    fa1, vcf1 = (tmp_path / "in_ms.fa").read_bytes(), (tmp_path / "in_ms.vcf").read_bytes()
    # header identical to the reference's for the same command (minus the wall-clock line)
    head = b"".join(l for l in vcf_head(vcf1).splitlines(keepends=True) if not l.startswith(b"##filedate="))
    assert head == vcf_head((GOLDEN / "args_all" / "out.vcf").read_bytes())
    contigs = read_fasta_simple(tmp_path / "in.fa")
    assert replay_plain(contigs, vcf1) == fa1
    assert len(vcf_body(vcf1).splitlines()) > 100
    run_main(argv)
    assert (tmp_path / "in_ms.fa").read_bytes() == fa1 and (tmp_path / "in_ms.vcf").read_bytes() == vcf1
    argv[argv.index("11")] = "12"
    run_main(argv)
    assert (tmp_path / "in_ms.fa").read_bytes() != fa1


def test_cli_rmt_mode_respects_blocked_ranges(tmp_path):
    shutil.copy(GOLDEN / "rmt_ranges" / "in.fa", tmp_path / "in.fa")
    shutil.copy(GOLDEN / "rmt_ranges" / "in.rmt", tmp_path / "in.rmt")
    run_main([tmp_path / "in.fa", "-o", tmp_path / "in", "-q", "--seed", "5", "rmt", tmp_path / "in.rmt"])
    vcf = (tmp_path / "in_ms.vcf").read_bytes()
    assert b'species="test species"' in vcf and vcf_head(vcf).splitlines()[-1].endswith(b"\tsample")
    contigs = read_fasta_simple(tmp_path / "in.fa")
    assert replay_plain(contigs, vcf) == (tmp_path / "in_ms.fa").read_bytes()
    n = 0
    for line in vcf_body(vcf).splitlines():
        f = line.split(b"\t")
        if f[0] == b"seq1":
            p = int(f[1])
            end = p + len(f[3])
            for lo, hi in ((1, 500), (3001, 4000)):   # None ranges of chr 1 (1-based inclusive)
                assert not (p <= hi and end - 1 >= lo and len(f[3]) > 1 and p >= lo), line
                assert not (lo <= p <= hi and f[7] == b"."), line
            n += 1
    assert n > 50


def test_it_replay_of_reference_breakpoints_is_bit_exact():
    """Gate A for IT: the reference's breakpoints + partners -> the reference's *_it.fa and .bedpe."""
    from argparse import Namespace
    import tempfile
    from mutation_simulator_b200 import ITMutator, SimulationSettings, load_fasta
    d = GOLDEN / "it_basic"
    bp = json.loads((d / "bp.json").read_text())
    with tempfile.TemporaryDirectory() as t:
        t = Path(t)
        shutil.copy(d / "in.fa", t / "in.fa")
        fasta = load_fasta(t / "in.fa")
        sim = SimulationSettings.from_it(0.004, fasta, True)
        args = Namespace(outfastait=t / "o.fa", outbedpe=t / "o.bedpe", ignore_warnings=True, no_color=True, seed=1, device=0)
        it = ITMutator(args, fasta, sim)
        it._partners = {int(k): v for k, v in bp["partners"].items()}
        bps = {int(k): {"self": np.array(v["self"], np.uint32), "partner": np.array(v["partner"], np.uint32)}
               for k, v in bp["breakpoints"].items()}
        it._generate_all_breakpoints = lambda eng: bps
        it.mutate()
        it.close()
        assert (t / "o.fa").read_bytes() == (d / "out_it.fa").read_bytes()
        assert (t / "o.bedpe").read_bytes() == (d / "out.bedpe").read_bytes()


def test_cli_it_mode_and_rmt_with_it(tmp_path):
    shutil.copy(GOLDEN / "it_basic" / "in.fa", tmp_path / "in.fa")
    run_main([tmp_path / "in.fa", "-o", tmp_path / "in", "-q", "--seed", "3", "it", "0.004"])
    contigs = read_fasta_simple(tmp_path / "in.fa")
    out = read_fasta_simple(tmp_path / "in_ms_it.fa")
    bed = (tmp_path / "in_ms_it.bedpe").read_bytes().splitlines()
    assert [c[1] for c in out] == [c[1] for c in contigs]
    assert sum(len(c[2]) for c in out) == sum(len(c[2]) for c in contigs)   # swaps conserve bases
    assert len(bed) > 4
    # rebuild the output from the BEDPE rows with the oracle
    names = {c[0]: i for i, c in enumerate(contigs)}
    bps, partners = {}, {}
    for row in bed:
        a, s1, e1, b, s2, e2 = row.split(b"\t")
        i, j = names[a], names[b]
        partners[i] = j
        d = bps.setdefault(i, {"self": [], "partner": []})
        for x, y in ((int(s1), int(s2)), (int(e1), int(e2))):
            if x != len(contigs[i][2]):
                d["self"].append(x); d["partner"].append(y)
    want, _ = pyref.it_genome(contigs, bps, partners)
    assert want == (tmp_path / "in_ms_it.fa").read_bytes()
    # combined mode: Mutator first, IT on the mutated genome (__main__.py:88-102)
    shutil.copy(GOLDEN / "rmt_it" / "in.fa", tmp_path / "k.fa")
    shutil.copy(GOLDEN / "rmt_it" / "in.rmt", tmp_path / "k.rmt")
    run_main([tmp_path / "k.fa", "-o", tmp_path / "k", "-q", "--seed", "4", "rmt", tmp_path / "k.rmt"])
    ms = read_fasta_simple(tmp_path / "k_ms.fa")
    it = read_fasta_simple(tmp_path / "k_ms_it.fa")
    assert sum(len(c[2]) for c in it) == sum(len(c[2]) for c in ms)
    assert (tmp_path / "k_ms_it.bedpe").stat().st_size > 0


def test_chained_rmt_then_it_in_hbm_equals_reload_of_the_written_file(tmp_path, monkeypatch):
    """__main__.py:88-95 re-loads *_ms.fa before ITMutator; the chained path strips the FASTA image on the device
    instead.  Both must give the same IT FASTA and BEDPE, including pyfaidx's line width rule for a re-loaded
    contig shorter than one line."""
    rng = np.random.default_rng(11)
    seqs = [rng.choice(list(b"ACGT"), size=n).astype(np.uint8).tobytes() for n in (5000, 41, 7300, 3, 6100, 2500)]
    with open(tmp_path / "g.fa", "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">c%d some text\n" % i)
            for o in range(0, len(s), 50):
                f.write(s[o:o + 50] + b"\n")
    rmt = ["std", "it 0.002",
           "sn 0.01 in 0.004 inmin 1 inmax 9 de 0.004 demin 1 demax 30 du 0.002 dumin 2 dumax 12",
           "chr 2", "it 0.05", "1-41 de 0.1 demin 1 demax 3"]
    (tmp_path / "g.rmt").write_text("\n".join(rmt) + "\n")
    outs = {}
    for mode in ("chain", "reload"):
        if mode == "reload":
            monkeypatch.setenv("MS_NO_CHAIN", "1")
        run_main([tmp_path / "g.fa", "-o", tmp_path / mode, "-q", "-w", "--seed", "21", "rmt", tmp_path / "g.rmt"])
        outs[mode] = [(tmp_path / f"{mode}{suffix}").read_bytes() for suffix in ("_ms.fa", "_ms.vcf", "_ms_it.fa", "_ms_it.bedpe")]
    assert outs["chain"] == outs["reload"]
    assert len(outs["chain"][3].splitlines()) >= 2
    ms = read_fasta_simple(tmp_path / "chain_ms.fa")
    it = read_fasta_simple(tmp_path / "chain_ms_it.fa")
    assert sum(len(c[2]) for c in it) == sum(len(c[2]) for c in ms)
    assert len(ms[1][2]) < 41   # the short contig lost bases, so its re-loaded line width shrank


def _write_fasta(path, recs, width, tail=b"\n"):
    with open(path, "wb") as f:
        for i, (hdr, seq) in enumerate(recs):
            f.write(b">" + hdr + b"\n")
            for o in range(0, len(seq), width[i]):
                f.write(seq[o:o + width[i]] + b"\n")
        if tail != b"\n":
            f.seek(-1, 2); f.truncate(); f.write(tail)


def test_device_fasta_ingest_matches_the_host_loader(tmp_path):
    """ms_fasta_ingest_fd/_index/_commit (util.py:77-91 on the GPU): same index, names, deflines and upper-cased bases
    as the host parser; irregular files are handed back to it."""
    from mutation_simulator_b200 import load_fasta
    rng = np.random.default_rng(3)
    alpha = np.frombuffer(b"ACGTNacgtnRYKM", np.uint8)
    lens = [100_003, 60, 1, 0, 7_000, 61, 120]
    recs = [(b"c%d desc %d" % (i, i) if i % 2 else b"c%d" % i, bytes(rng.choice(alpha, n))) for i, n in enumerate(lens)]
    widths = [60, 60, 60, 60, 70, 61, 40]
    for name, tail in (("lf.fa", b"\n"), ("noeol.fa", b""), ("blank_end.fa", b"\n\n\n")):
        p = tmp_path / name
        _write_fasta(p, recs, widths, tail)
        host = load_fasta(p)
        dev = load_fasta(p, device=0)
        assert dev.engine is not None, name
        assert dev.names == host.names and dev.long_names == host.long_names
        assert (dev.lengths == host.lengths).all() and (dev.bpl == host.bpl).all()
        for nm in host.names:
            a, b = dev.faidx.index[nm], host.faidx.index[nm]
            assert (a.rlen, a.offset, a.lenc, a.lenb) == (b.rlen, b.offset, b.lenc, b.lenb), (name, nm)
        want = np.concatenate([host.gather([i]) for i in range(len(lens))])
        want = np.where((want >= 97) & (want <= 122), want - 32, want).astype(np.uint8)
        assert (dev.engine.download_genome() == want).all()
        assert str(dev[0][5:15]) == str(host[0][5:15]) and len(dev[4]) == 7_000
        dev.close(); host.close()
    # irregular layouts: the device path declines, the host parser decides (and raises what pyfaidx would)
    ragged = tmp_path / "ragged.fa"
    ragged.write_bytes(b">a\nACGTACGT\nACG\nACGTACGT\n")
    from mutation_simulator_b200.fasta import FastaIndexingError
    with pytest.raises(FastaIndexingError):
        load_fasta(ragged, device=0)
    crlf = tmp_path / "crlf.fa"
    crlf.write_bytes(b">a x\r\nACGT\r\nAC\r\n>b\r\nGG\r\n")
    f = load_fasta(crlf, device=0)
    assert f.engine is None and f.names == ["a", "b"] and list(f.lengths) == [6, 2]
    dup = tmp_path / "dup.fa"
    dup.write_bytes(b">a\nACGT\n>a\nAC\n")
    from mutation_simulator_b200 import FastaDuplicateHeaderError
    with pytest.raises(FastaDuplicateHeaderError):
        load_fasta(dup, device=0)


# ---- Gate A from the reference's FILES (north star, part one) ------------------------------------------------------
@pytest.mark.parametrize("case", ["args_all", "iupac", "c1_small", "rmt_ranges", "rmt_it", "edges", "tiny"])
def test_replay_of_the_reference_vcf_file_through_the_class_api(case, tmp_path):
    """in.fa + the reference's out.vcf -> Mutator.load_vcf().mutate() must write the reference's out.fa, and the VCF
    re-emitted from the parsed records must be the reference's VCF body byte for byte."""
    from mutation_simulator_b200 import Mutator, SimulationSettings, load_fasta
    from tests.test_emu import assert_fasta_equal_up_to_silent_gap_snps
    from mutation_simulator_b200 import get_args
    d = GOLDEN / case
    shutil.copy(d / "in.fa", tmp_path / "in.fa")
    if (d / "in.rmt").exists():
        shutil.copy(d / "in.rmt", tmp_path / "in.rmt")
    argv = json.loads((d / "cmd.json").read_text())["argv"]
    if argv == ["crafted"]:      # hand-crafted tables pushed through the reference's walk: no command line
        argv = ["in.fa", "-o", "in", "-q", "args", "-sn", "0.01"]
    argv = [str(tmp_path / a) if a in ("in.fa", "in.rmt", "in") else a for a in argv]
    args = get_args(argv)
    args.outfasta, args.outvcf, args.device = tmp_path / "o.fa", tmp_path / "o.vcf", 0
    fasta = load_fasta(args.infile, device=0)
    sim = SimulationSettings.from_rmt(args.rmtfile, fasta, True) if args.mode == "rmt" else SimulationSettings.from_args(args, fasta, True)
    m = Mutator(args, fasta, sim)
    m.load_vcf((d / "out.vcf").read_bytes())
    m.mutate()
    m.close()
    fasta.close()
    assert vcf_body((tmp_path / "o.vcf").read_bytes()) == vcf_body((d / "out.vcf").read_bytes())
    assert_fasta_equal_up_to_silent_gap_snps((tmp_path / "o.fa").read_bytes(), (d / "out.fa").read_bytes(), case)


@pytest.mark.parametrize("case", ["it_basic", "rmt_it"])
def test_replay_of_the_reference_bedpe_file_through_the_class_api(case, tmp_path):
    """(input FASTA, the reference's out.bedpe) -> ITMutator.load_bedpe().mutate() must write the reference's
    out_it.fa and the same BEDPE rows.  rmt_it: the IT step runs on the mutated genome (__main__.py:88-95)."""
    from argparse import Namespace
    from mutation_simulator_b200 import ITMutator, SimulationSettings, load_fasta
    d = GOLDEN / case
    shutil.copy(d / ("out.fa" if case == "rmt_it" else "in.fa"), tmp_path / "in.fa")
    fasta = load_fasta(tmp_path / "in.fa", device=0)
    sim = SimulationSettings.from_it(0.001, fasta, True)
    args = Namespace(outfastait=tmp_path / "o.fa", outbedpe=tmp_path / "o.bedpe", ignore_warnings=True, no_color=True, seed=1, device=0)
    it = ITMutator(args, fasta, sim)
    it.load_bedpe((d / "out.bedpe").read_bytes())
    it.mutate()
    it.close()
    fasta.close()
    assert (tmp_path / "o.fa").read_bytes() == (d / "out_it.fa").read_bytes()
    assert (tmp_path / "o.bedpe").read_bytes() == (d / "out.bedpe").read_bytes()
