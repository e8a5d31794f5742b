"""Host-side pieces of the drop-in API that need no GPU: VCF header, BEDPE rows,
range planning, output naming, C-ABI symbol table."""
import ctypes as C
import json
import re
import sys
from datetime import datetime
from pathlib import Path

import numpy as np
import pytest

from mutation_simulator_b200 import _lib, bedpe_writer, plan, vcf_writer
from mutation_simulator_b200.argument_parser import get_args
from mutation_simulator_b200.fasta import Fasta, FastaIndexingError, FastaNotFoundError
from mutation_simulator_b200.rmt import SimulationSettings
from tests.helpers import GOLDEN, vcf_head

REPO = Path(__file__).resolve().parent.parent


def test_cabi_library_exports_every_declared_symbol():
    hdr = (REPO / "include" / "mutsim_b200.h").read_text()
    declared = set(re.findall(r"\b(ms_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ms_abi_version() == 1
    assert lib.ms_stage_name(8) == b"splice_emit_fasta"


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mutation_simulator_b200.engine import Engine
    with pytest.raises(_lib.MutSimError):
        Engine(0)


def test_vcf_header_matches_reference():
    f = Fasta(GOLDEN / "args_all" / "in.fa", build_index=False)
    txt = vcf_writer.header_text("in.fa", [(f[n].name, len(f[n])) for n in f.keys()], "asmX", "Some species", "smp1",
                                 now=datetime(2026, 1, 5))
    assert "##filedate=202615\n" in txt           # un-padded YYYYMD (vcf_writer.py:85)
    txt = "".join(l for l in txt.splitlines(keepends=True) if not l.startswith("##filedate="))
    assert txt.encode() == vcf_head((GOLDEN / "args_all" / "out.vcf").read_bytes())


@pytest.mark.parametrize("case", ["it_basic", "rmt_it"])
def test_bedpe_rows_match_reference(case):
    d = GOLDEN / case
    f = Fasta(d / ("out.fa" if (d / "out.fa").exists() else "in.fa"), build_index=False)
    bp = json.loads((d / "bp.json").read_text())
    out = b""
    for i in range(len(f.names)):
        if str(i) in bp["breakpoints"]:
            p = bp["partners"][str(i)]
            b = bp["breakpoints"][str(i)]
            out += bedpe_writer.rows(f[i].name, b["self"], len(f[i]), f[p].name, b["partner"], len(f[p]))
    assert out == (d / "out.bedpe").read_bytes()


def _args(argv):
    old = sys.argv
    sys.argv = ["mutation-simulator"] + argv
    try:
        return get_args()
    finally:
        sys.argv = old


def test_output_names_follow_reference():
    a = _args(["dir/genome.fa", "it", "0.1"])
    assert (str(a.outfasta), str(a.outvcf), str(a.outfastait), str(a.outbedpe)) == \
        ("genome_ms.fa", "genome_ms.vcf", "genome_ms_it.fa", "genome_ms_it.bedpe")
    a = _args(["dir/genome.fasta", "-o", "out/base", "-q", "it", "0.1"])
    assert (str(a.outfasta), str(a.outvcf), str(a.outfastait), str(a.outbedpe)) == \
        ("out/base_ms.fasta", "out/base_ms.vcf", "out/base_ms_it.fasta", "out/base_ms_it.bedpe")
    assert a.ignore_warnings and a.no_progress and a.seed is None


def test_plan_candidate_counts_match_reference_runs():
    g = json.loads((GOLDEN / "stats_all.json").read_text())
    lengths = g["lengths"]

    class F:  # minimal fasta stand-in for from_args
        def keys(self):
            return [f"c{i}" for i in range(len(lengths))]

        def __getitem__(self, i):
            return "x" * 0 if False else type("R", (), {"__len__": lambda s: lengths[i if isinstance(i, int) else int(i[1:])]})()
    a = _args(["g.fa", "args"] + g["args"])
    sim = SimulationSettings.from_args(a, F(), True)
    arr, n = plan.build_ranges(sim, lengths)
    assert n == len(lengths)
    for i in range(n):
        assert arr[i].k == g["runs"][0]["contigs"][i]["candidates"]
        assert arr[i].start == 0 and arr[i].stop == lengths[i] - 1 and arr[i].limit == lengths[i]
        assert abs(arr[i].cdf[6] - 1.0) < 1e-12
    assert plan.block_list(sim) == [1] * 7
    assert plan.p_transition(2.0) == 2.0 * (1 / 3.0)


def test_plan_rmt_limits_and_overlap(tmp_path):
    f = Fasta(GOLDEN / "rmt" / "g.fa", build_index=False)
    sim = SimulationSettings.from_rmt(GOLDEN / "rmt" / "ok_ranges.rmt".replace("ok_ranges", "ok_unsorted_ranges"), f, True)
    arr, n = plan.build_ranges(sim, f.lengths)
    rows = [(arr[i].contig, arr[i].start, arr[i].stop, arr[i].limit) for i in range(n)]
    # chr 2: 1-100 sn 0.1 | 101-200 std | 201-300 None: SVs of the first two ranges may not reach base 200
    assert (1, 0, 99, 200) in rows and (1, 100, 199, 200) in rows
    p = tmp_path / "o.rmt"
    p.write_text("std\nit None\nsn 0.01\nchr 1\n10-200 sn 0.1\n150-300 sn 0.1\n")
    sim = SimulationSettings.from_rmt(p, f, True)
    with pytest.raises(plan.RangeOverlapError):
        plan.build_ranges(sim, f.lengths)


def test_fasta_loader_errors(tmp_path):
    with pytest.raises(FastaNotFoundError):
        Fasta(tmp_path / "nope.fa")
    p = tmp_path / "ragged.fa"
    p.write_text(">a\nACGT\nAC\nACGT\n")
    with pytest.raises(FastaIndexingError):
        Fasta(p)
    p = tmp_path / "dup.fa"
    p.write_text(">a x\nACGT\n>a y\nAC\n")
    from mutation_simulator_b200.util import FastaDuplicateHeaderError, load_fasta
    with pytest.raises(FastaDuplicateHeaderError):
        load_fasta(p)
    p = tmp_path / "ok.fa"
    p.write_text(">a desc\nacgtn\nAC\n>b\nGG\n")
    f = load_fasta(p)
    assert f.keys() == ["a", "b"] and str(f["a"]) == "ACGTNAC" and f[1].long_name == "b"
    assert f.faidx.index["a"].lenc == 5 and (tmp_path / "ok.fa.fai").exists()
    assert f["a"][1:3] == "CG" and f["a"][0:0] == "ACGTNAC" and list(f["a"]) == ["ACGTN", "AC"]


def test_cli_without_a_gpu_fails_loudly_with_the_reference_error_convention(tmp_path):
    """No CPU fallback: on a box without a CUDA device the CLI prints `ERROR: ...` and exits 1 (util.py:21-31)."""
    import subprocess
    import sys
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    (tmp_path / "a.fa").write_text(">a\nACGTACGTAC\n")
    r = subprocess.run([sys.executable, "-m", "mutation_simulator_b200", str(tmp_path / "a.fa"), "-o", str(tmp_path / "o"),
                        "args", "-sn", "0.1"], capture_output=True, text=True, cwd=str(Path(__file__).resolve().parent.parent))
    assert r.returncode == 1
    assert r.stderr.startswith("ERROR: ") and "no CPU fallback" in r.stderr
    assert not (tmp_path / "o_ms.fa").exists() or (tmp_path / "o_ms.fa").stat().st_size == 0


def test_plan_column_wise_path_equals_the_general_loop(monkeypatch):
    """Many single-range contigs (C5) take a vectorised path through plan.build_ranges; it must build the same table."""
    from mutation_simulator_b200.fasta import FastaRecord

    class LengthsOnly:
        def __init__(self, lens):
            self.names = [f"c{i}" for i in range(len(lens))]
            self.lengths = np.array(lens, np.int64)
            self._r = [FastaRecord(n, n, None, 60, length=int(l)) for n, l in zip(self.names, lens)]
        def __getitem__(self, k): return self._r[k] if isinstance(k, (int, np.integer)) else self._r[self.names.index(k)]
        def keys(self): return self.names
        def __len__(self): return len(self.names)

    rng = np.random.default_rng(1)
    lens = rng.integers(30, 9000, size=3000).tolist()
    fa = LengthsOnly(lens)
    a = get_args(["x.fa", "args", "-sn", "0.013", "-in", "0.001", "-inmax", "10", "-de", "0.0007", "-demax", "9", "-tl", "0.0005", "-tlmax", "50"])
    sim = SimulationSettings.from_args(a, fa, True)
    ids = [i for i in range(len(lens)) if i % 3]          # a rank's share
    fast, nf = plan.build_ranges(sim, fa.lengths, ids)
    monkeypatch.setattr(plan, "_build_single_range", lambda *a, **k: None)
    slow, ns = plan.build_ranges(sim, fa.lengths, ids)
    assert nf == ns > 1000
    assert bytes(memoryview(fast))[:nf * C.sizeof(_lib.MsRange)] == bytes(memoryview(slow))[:ns * C.sizeof(_lib.MsRange)]


def test_it_pairing_reproduces_the_reference_pair_count():
    """it_mutator.py:59-70 iterates a list it is shrinking: ceil(n/3) pairs, not floor(n/2).
    tests/golden/it_pairs.json = the unmodified reference's counts (tests/golden/make_it_pairs.py)."""
    import random
    from mutation_simulator_b200.it_mutator import assign_partners
    want = json.loads((GOLDEN / "it_pairs.json").read_text())
    for n in range(2, 41):
        for seed in range(25):
            p = assign_partners(list(range(100, 100 + n)), random.Random(seed))
            assert all(p[p[a]] == a and p[a] != a for a in p)
            assert set(p) <= set(range(100, 100 + n))
            assert len(p) // 2 == want[str(n)]["pairs"], (n, seed)
    # every contig is equally likely to end up paired (the list is shuffled first)
    hits = np.zeros(24)
    for seed in range(3000):
        for a in assign_partners(list(range(24)), random.Random(seed)):
            hits[a] += 1
    from scipy import stats
    assert stats.chisquare(hits).pvalue > 0.01


def test_overlapping_ranges_fail_before_the_output_files_are_touched(tmp_path):
    """the overlap check runs in Mutator.__init__ before the writers open (and truncate) the outputs."""
    from argparse import Namespace
    from mutation_simulator_b200 import Mutator
    from mutation_simulator_b200.plan import RangeOverlapError
    fa = tmp_path / "g.fa"
    fa.write_text(">c1\n" + "ACGT" * 500 + "\n")
    rmt = tmp_path / "g.rmt"
    rmt.write_text("std\nit None\nsn 0.01\n\nchr 1\n1-1200 sn 0.05\n400-1500 sn 0.05\n")
    fasta = Fasta(str(fa))
    sim = SimulationSettings.from_rmt(rmt, fasta, True)
    out_fa, out_vcf = tmp_path / "o.fa", tmp_path / "o.vcf"
    out_fa.write_text("precious")
    args = Namespace(outfasta=out_fa, outvcf=out_vcf, infile=fa, ignore_warnings=True, no_color=True, no_progress=True, seed=1, device=0)
    with pytest.raises(RangeOverlapError):
        Mutator(args, fasta, sim)
    assert out_fa.read_text() == "precious" and not out_vcf.exists()


def test_fai_is_rebuilt_when_older_than_the_fasta(tmp_path):
    import os
    fa = tmp_path / "g.fa"
    fa.write_text(">c1\nACGTACGT\nACG\n")
    Fasta(str(fa))
    fai = tmp_path / "g.fa.fai"
    assert fai.read_text().split("\t")[1] == "11"
    fa.write_text(">c1\nACGTACGT\nACGTA\n")
    os.utime(fai, (1, 1))                       # the index predates the file
    Fasta(str(fa))
    assert fai.read_text().split("\t")[1] == "13"
