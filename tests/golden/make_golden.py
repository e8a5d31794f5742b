#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (/root/reference, Mutation-Simulator 3.0.2) in this container.

The reference cannot travel to the GPU box, so its outputs are committed as
small fixtures.  ``pyfaidx`` is replaced by the in-memory stand-in under
oracle/pyfaidx_standin (test fixture).  The reference has no seed flag; we seed
``random`` and ``numpy.random`` before every run, which makes its output
byte-reproducible.

Run:  python tests/golden/make_golden.py        (only where /root/reference exists)

What is captured per case (directory tests/golden/<case>/):
  in.fa        input FASTA
  out.fa       reference output FASTA            (Mutator, mutator.py:105-142)
  out.vcf      reference output VCF, ##filedate line removed (wall clock, vcf_writer.py:83-85)
  muts.json    the mutation table the reference's walk actually applied: the
               ``muts`` dict handed to Mutator.__mutate_sequence (mutator.py:318)
               plus the SNP ALT bases / insert strings its RNG produced
  out_it.fa / out.bedpe / bp.json   for IT cases (it_mutator.py)
  cmd.json     the command line and seed
"""
from __future__ import annotations

import io
import json
import os
import random
import shutil
import sys
from argparse import Namespace
from contextlib import redirect_stderr, redirect_stdout
from pathlib import Path

import numpy

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REF))
sys.path.insert(0, str(REPO / "oracle" / "pyfaidx_standin"))

import mutation_simulator  # noqa: E402  (the reference)
from mutation_simulator import __main__ as ref_main  # noqa: E402
from mutation_simulator.it_mutator import ITMutator  # noqa: E402
from mutation_simulator.mut_types import MutType  # noqa: E402
from mutation_simulator.mutator import Mutation, Mutator  # noqa: E402
from mutation_simulator.rmt import SimulationSettings  # noqa: E402
from mutation_simulator.util import load_fasta  # noqa: E402

assert str(Path(mutation_simulator.__file__).resolve()).startswith(str(REF))

SPANNING = {"DE", "TL", "IV", "DU"}


# ----------------------------------------------------------------------------
# capture hooks (wrap, never alter, the reference's behaviour)
# ----------------------------------------------------------------------------
class Capture:
    def __init__(self):
        self.contigs = []      # list of dict(name, muts=[...])
        self.snp_alts = []
        self.inserts = []
        self.breakpoints = None
        self.partners = None

    def install(self):
        cap = self
        self._orig_walk = Mutator._Mutator__mutate_sequence
        self._orig_snp = Mutator.__dict__["_Mutator__get_snp"]
        self._orig_ins = Mutator.__dict__["_Mutator__get_insert"]
        self._orig_bp = ITMutator._ITMutator__generate_all_breakpoints

        def walk(self_, sequence, muts, titv):
            cap.snp_alts, cap.inserts = [], []
            snapshot = {k: (v.type.name, v.start, v.stop, bool(v.trans_reverse),
                            v.trans_insert_pos) for k, v in muts.items()}
            cap._orig_walk(self_, sequence, muts, titv)
            cap.contigs.append(cap._finish(sequence.name, len(sequence), snapshot))

        orig_snp_f = self._orig_snp.__func__

        def snp(cls, base, ti_tv):
            alt = orig_snp_f(cls, base, ti_tv)
            cap.snp_alts.append(alt)
            return alt

        orig_ins_f = self._orig_ins.__func__

        def ins(leng):
            s = orig_ins_f(leng)
            cap.inserts.append(s)
            return s

        def bps(self_):
            bp = cap._orig_bp(self_)
            cap.breakpoints = {str(k): v for k, v in bp.items()}
            cap.partners = {str(k): v for k, v in self_._ITMutator__partners.items()}
            return bp

        Mutator._Mutator__mutate_sequence = walk
        Mutator._Mutator__get_snp = classmethod(snp)
        Mutator._Mutator__get_insert = staticmethod(ins)
        ITMutator._ITMutator__generate_all_breakpoints = bps

    def uninstall(self):
        Mutator._Mutator__mutate_sequence = self._orig_walk
        Mutator._Mutator__get_snp = self._orig_snp
        Mutator._Mutator__get_insert = self._orig_ins
        ITMutator._ITMutator__generate_all_breakpoints = self._orig_bp

    def _finish(self, name, length, snapshot):
        """Replay the walk's visiting order (mutator.py:332-425) to find which
        dict entries were actually visited, and attach RNG outputs."""
        out = []
        pos_limit = -1  # last position consumed by a spanning mutation
        si = ii = 0
        for key in sorted(snapshot):
            typ, start, stop, rev, tip = snapshot[key]
            if key <= pos_limit or key >= length:
                continue  # swallowed by an earlier DE/TL/IV/DU span
            rec = {"key": key, "type": typ, "start": start, "stop": stop,
                   "reverse": rev, "insert_pos": tip}
            if typ == "SN":
                rec["alt"] = self.snp_alts[si]
                si += 1
            elif typ == "IN":
                rec["insert"] = self.inserts[ii]
                ii += 1
            if typ in SPANNING:
                pos_limit = stop
            out.append(rec)
        assert si == len(self.snp_alts) and ii == len(self.inserts), (
            name, si, len(self.snp_alts), ii, len(self.inserts))
        return {"name": name, "length": length, "muts": out}


def strip_filedate(text: str) -> str:
    return "".join(l for l in text.splitlines(keepends=True)
                   if not l.startswith("##filedate="))


def run_cli(argv, seed):
    """python -m mutation_simulator <argv> with seeded global RNGs."""
    random.seed(seed)
    numpy.random.seed(seed)
    old = sys.argv
    sys.argv = ["mutation-simulator"] + [str(a) for a in argv]
    err = io.StringIO()
    try:
        with redirect_stdout(io.StringIO()), redirect_stderr(err):
            ref_main.main()
    finally:
        sys.argv = old
    return err.getvalue()


# ----------------------------------------------------------------------------
# input generators
# ----------------------------------------------------------------------------
def write_fasta(path, contigs):
    """contigs: list of (header, sequence, line_width)"""
    with open(path, "w") as fh:
        for hdr, seq, w in contigs:
            fh.write(f">{hdr}\n")
            for i in range(0, len(seq), w):
                fh.write(seq[i:i + w] + "\n")


def rand_seq(rng, n, alphabet="ACGT", p=None):
    idx = rng.choice(len(alphabet), size=n, p=p)
    return "".join(alphabet[i] for i in idx)


def case_dir(name):
    d = HERE / name
    if d.exists():
        shutil.rmtree(d)
    d.mkdir(parents=True)
    return d


def finish_case(d, cap, argv, seed, stderr_text, it=False, ms=True):
    stem = "in_ms"
    if ms:
        (d / "out.fa").write_bytes((d / f"{stem}.fa").read_bytes())
        (d / "out.vcf").write_text(strip_filedate((d / f"{stem}.vcf").read_text()))
        (d / f"{stem}.fa").unlink()
        (d / f"{stem}.vcf").unlink()
        (d / "muts.json").write_text(json.dumps(cap.contigs, indent=0, separators=(",", ":")))
    if it:
        (d / "out_it.fa").write_bytes((d / f"{stem}_it.fa").read_bytes())
        (d / "out.bedpe").write_bytes((d / f"{stem}_it.bedpe").read_bytes())
        (d / f"{stem}_it.fa").unlink()
        (d / f"{stem}_it.bedpe").unlink()
        (d / "bp.json").write_text(json.dumps(
            {"breakpoints": cap.breakpoints, "partners": cap.partners}))
    (d / "cmd.json").write_text(json.dumps(
        {"argv": [str(a) for a in argv[2:]], "seed": seed, "stderr": stderr_text}))


def run_case(name, contigs, mode_argv, seed, it=False, ms=True, rmt_text=None):
    d = case_dir(name)
    write_fasta(d / "in.fa", contigs)
    if rmt_text is not None:
        (d / "in.rmt").write_text(rmt_text)
        mode_argv = ["rmt", d / "in.rmt"]
    argv = [d / "in.fa", "-o", d / "in", "-q"] + mode_argv
    cap = Capture()
    cap.install()
    try:
        err = run_cli(argv, seed)
    finally:
        cap.uninstall()
    # keep argv relative
    rel = ["in.fa", "-o", "in", "-q"] + [("in.rmt" if isinstance(a, Path) else a) for a in mode_argv]
    finish_case(d, cap, ["x", "y"] + rel, seed, err, it=it, ms=ms)
    return d


# ----------------------------------------------------------------------------
# hand-crafted mutation tables pushed through the reference's own walk
# ----------------------------------------------------------------------------
def run_crafted(name, contigs, crafted, seed, titv=1.0):
    """crafted: dict contig_index -> list of Mutation kwargs (key = dict key)."""
    d = case_dir(name)
    write_fasta(d / "in.fa", contigs)
    fasta = load_fasta(d / "in.fa")
    args = Namespace(outfasta=d / "in_ms.fa", outvcf=d / "in_ms.vcf",
                     infile=Path("in.fa"), no_progress=True, ignore_warnings=True,
                     no_color=True, snp=0.01, insert=0.0, deletion=0.0, inversion=0.0,
                     duplication=0.0, translocation=0.0,
                     insertminlength=1, insertmaxlength=2, deletionminlength=1,
                     deletionmaxlength=2, inversionminlength=2, inversionmaxlength=3,
                     duplicationminlength=1, duplicationmaxlength=2,
                     translocationminlength=1, translocationmaxlength=2,
                     snpblock=1, insertblock=1, deletionblock=1, inversionblock=1,
                     duplicationblock=1, translocationblock=1,
                     transitionstransversions=titv, species="Unknown",
                     assembly="Unknown", sample="Unknown")
    sim = SimulationSettings.from_args(args, fasta, True)
    per_chrom = {}
    for ci, lst in crafted.items():
        m = {}
        for kw in lst:
            key = kw.pop("key")
            m[key] = Mutation(MutType[kw.pop("type")], **kw)
        per_chrom[ci] = m

    orig_get = Mutator._Mutator__get_mutations
    state = {"i": -1}

    def fake_get(self_, rng, chrom_len):
        state["i"] += 1
        return per_chrom.get(state["i"], {}), [], []

    cap = Capture()
    cap.install()
    Mutator._Mutator__get_mutations = fake_get
    random.seed(seed)
    numpy.random.seed(seed)
    try:
        with redirect_stdout(io.StringIO()), redirect_stderr(io.StringIO()):
            mut = Mutator(args, fasta, sim)
            mut.mutate()
            mut.close()
    finally:
        Mutator._Mutator__get_mutations = orig_get
        cap.uninstall()
    finish_case(d, cap, ["x", "y", "crafted"], seed, "")
    return d


# ----------------------------------------------------------------------------
# statistics of fresh sampling (Gate B).  Runs only the sampling half of the
# reference (mutator.py:144-316), which is ~2 % of its runtime.
# ----------------------------------------------------------------------------
def sampling_stats(name, lengths, args_list, seeds, rmt_text=None):
    tmp = HERE / "_tmp_stats"
    if tmp.exists():
        shutil.rmtree(tmp)
    tmp.mkdir()
    rng = numpy.random.default_rng(99)
    write_fasta(tmp / "g.fa", [(f"chr{i+1}", rand_seq(rng, n), 60) for i, n in enumerate(lengths)])
    fasta = load_fasta(tmp / "g.fa")
    if rmt_text is not None:
        (tmp / "g.rmt").write_text(rmt_text)
        sim = SimulationSettings.from_rmt(tmp / "g.rmt", fasta, True)
    else:
        old = sys.argv
        sys.argv = ["x", str(tmp / "g.fa"), "-q", "args"] + [str(a) for a in args_list]
        from mutation_simulator.argument_parser import get_args
        a = get_args()
        sys.argv = old
        sim = SimulationSettings.from_args(a, fasta, True)
    args = Namespace(outfasta=tmp / "o.fa", outvcf=tmp / "o.vcf", infile=Path("g.fa"),
                     no_progress=True, ignore_warnings=True, no_color=True)
    runs = []
    for seed in seeds:
        random.seed(seed)
        numpy.random.seed(seed)
        m = Mutator(args, fasta, sim)
        run = {"seed": seed, "contigs": []}
        for chrom in sim.chromosomes:
            muts, tls, tlis = {}, [], []
            n_cand = 0
            per_range = []
            for rd in chrom.range_definitions:
                if rd.mutation_settings.has_mutations:
                    k = int(((rd.stop - rd.start) + 1) * sum(rd.mutation_settings.mut_rates.values()))
                    n_cand += k
                    rm, rt, rti = m._Mutator__get_mutations(rd, len(fasta[chrom.number]))
                    per_range.append({"start": rd.start, "stop": rd.stop, "k": k, "accepted": len(rm)})
                    muts.update(rm)
                    tls.extend(rt)
                    tlis.extend(rti)
                else:
                    per_range.append({"start": rd.start, "stop": rd.stop, "k": 0, "accepted": 0})
            n_tl_before, n_tli_before = len(tls), len(tlis)
            if tls:
                muts = m._Mutator__link_tls(muts, tls, tlis)
            counts = {}
            lens = {}
            rev = 0
            positions = sorted(muts)
            for p in positions:
                mu = muts[p]
                t = mu.type.name
                counts[t] = counts.get(t, 0) + 1
                if t != "SN":
                    ln = mu.stop - mu.start + 1
                    lens.setdefault(t, {})
                    lens[t][ln] = lens[t].get(ln, 0) + 1
                if t == "TLI" and mu.trans_reverse:
                    rev += 1
            gaps = numpy.diff(numpy.array(positions)) if len(positions) > 1 else numpy.array([1 << 30])
            L = len(fasta[chrom.number])
            dec = numpy.histogram(numpy.array(positions), bins=10, range=(0, L))[0].tolist()
            run["contigs"].append({
                "length": L, "candidates": n_cand, "counts": counts,
                "lens": {t: {str(k): v for k, v in sorted(d_.items())} for t, d_ in lens.items()},
                "tl_before_link": n_tl_before, "tli_before_link": n_tli_before,
                "tli_reversed": rev, "min_gap": int(gaps.min()), "pos_deciles": dec,
                "ranges": per_range if len(per_range) <= 64 else None,
            })
        m.close()
        runs.append(run)
    shutil.rmtree(tmp)
    (HERE / f"{name}.json").write_text(json.dumps(
        {"lengths": lengths, "args": [str(a) for a in args_list], "rmt": rmt_text, "runs": runs}))


def titv_stats(name, titvs, n, seeds):
    """ti/tv behaviour of Mutator.__get_snp (mutator.py:429-463) on each base."""
    out = {}
    f = Mutator.__dict__["_Mutator__get_snp"].__func__
    for titv in titvs:
        res = {}
        for seed in seeds:
            random.seed(seed)
            for base in "ACGTN":
                c = {}
                for _ in range(n):
                    a = f(Mutator, base, titv)
                    c[a] = c.get(a, 0) + 1
                res.setdefault(base, []).append(c)
        out[str(titv)] = res
    (HERE / f"{name}.json").write_text(json.dumps({"n": n, "result": out}))


# ----------------------------------------------------------------------------
# settings model goldens: RMT texts -> normalised dump or (error class, message)
# ----------------------------------------------------------------------------
def dump_sim(sim):
    chroms = []
    for c in sim.chromosomes:
        rds = []
        for rd in c.range_definitions:
            ms = rd.mutation_settings
            rds.append({
                "start": rd.start, "stop": rd.stop,
                "rates": None if ms.mut_rates is None else [[t.name, r] for t, r in ms.mut_rates.items()],
                "chances": None if ms.mut_chances is None else [[t.name, r] for t, r in ms.mut_chances.items()],
                "min": None if not ms.mut_lengs else [[t.name, v] for t, v in ms.mut_lengs["min"].items()],
                "max": None if not ms.mut_lengs else [[t.name, v] for t, v in ms.mut_lengs["max"].items()],
                "has_mutations": ms.has_mutations,
            })
        chroms.append({"number": c.number, "it_rate": c.it_rate, "ranges": rds})
    return {
        "chromosomes": chroms,
        "mut_block": [[t.name, v] for t, v in sim.mut_block.items()],
        "fasta": sim.fasta, "md5": sim.md5, "titv": sim.titv,
        "species_name": sim.species_name, "assembly_name": sim.assembly_name,
        "sample_name": sim.sample_name,
        "has_mutations": sim.has_mutations, "has_it": sim.has_it,
    }


RMT_CASES = {
    "ok_basic": "std\nit None\nsn 0.01\n",
    "ok_meta": ("fasta=Genome.FA\nmd5=ABCDEF\nspecies_name=Homo Sapiens\nassembly_name=GRCh38\n"
                "sample_name=S1\ntitv=2.5\nsn_block=3\nin_block=0\ntl_block=7\n"
                "std\nit 0.001\nsn 0.01 in 0.002 inmin 1 inmax 5\n"),
    "ok_ranges": ("std\nit None\nsn 0.001\n\nchr 1 #first\n11-50 None\n101-200 sn 0.05 in 0.01 inmin 2 inmax 4\n"
                  "301-END de 0.02 demin 1 demax 3 # tail\nchr 3\nit 0.01\n1-100 None\n"),
    "ok_all_types": ("std\nit 0.0\nsn 0.01 in 0.001 inmin 1 inmax 3 de 0.001 demin 1 demax 3 "
                     "iv 0.001 ivmin 2 ivmax 4 du 0.001 dumin 1 dumax 3 tl 0.002 tlmin 1 tlmax 3\n"),
    "ok_unsorted_ranges": "std\nit None\nsn 0.01\nchr 2\n201-300 None\n1-100 sn 0.1\n",
    "ok_std_none": "std\nit None\nNone\nchr 1\n1-100 sn 0.1\n",
    "ok_unknown_kw": "std\nit None\nsn 0.01 foo 3 bar 4\n",
    "ok_it_only": "std\nit 0.01\nNone\n",
    "ok_chr_it_none": "std\nit 0.01\nsn 0.01\nchr 1\nit None\nchr 2\nit 0.2\n",
    "ok_overlap_small": "std\nit None\nsn 0.01\nchr 1\n10-50 None\n45-90 None\n",
    "ok_end_equals_len": "std\nit None\nsn 0.01\nchr 1\n1-401 None\n",
    "ok_block_negative": "sn_block=-4\nde_block=2\nstd\nit None\nsn 0.01\n",
    "err_no_std": "chr 1\n1-100 None\n",
    "err_std_three": "std\nit None\nsn 0.01\nsn 0.02\n",
    "err_rate_high": "std\nit None\nsn 0.6\n",
    "err_rate_high_single": "std\nit None\nsn 1.5\n",
    "err_rate_zero": "std\nit None\nsn 0\n",
    "err_rate_negative": "std\nit None\nsn -0.1 in 0.3 inmin 1 inmax 2\n",
    "err_missing_len": "std\nit None\nin 0.01\n",
    "err_missing_max": "std\nit None\nin 0.01 inmin 1\n",
    "err_min_gt_max": "std\nit None\nde 0.01 demin 5 demax 2\n",
    "err_min_low": "std\nit None\ndu 0.01 dumin 0 dumax 2\n",
    "err_iv_min_low": "std\nit None\niv 0.01 ivmin 1 ivmax 2\n",
    "err_it_high": "std\nit 0.7\nsn 0.01\n",
    "err_it_negative": "std\nit -0.1\nsn 0.01\n",
    "err_it_malformed": "std\nit\nsn 0.01\n",
    "err_it_text": "std\nit abc\nsn 0.01\n",
    "err_titv": "titv=-1\nstd\nit None\nsn 0.01\n",
    "err_titv_text": "titv=abc\nstd\nit None\nsn 0.01\n",
    "err_block_text": "sn_block=x\nstd\nit None\nsn 0.01\n",
    "err_odd_tokens": "std\nit None\nsn 0.01 in\n",
    "err_bad_float": "std\nit None\nsn abc\n",
    "err_bad_int": "std\nit None\nin 0.01 inmin x inmax 2\n",
    "err_chr_missing": "std\nit None\nsn 0.01\nchr 9\n1-10 None\n",
    "err_chr_index": "std\nit None\nsn 0.01\nchr x\n1-10 None\n",
    "err_range_oob": "std\nit None\nsn 0.01\nchr 1\n1-100000 None\n",
    "err_range_zero": "std\nit None\nsn 0.01\nchr 1\n0-10 None\n",
    "err_range_malformed": "std\nit None\nsn 0.01\nchr 1\n1-2-3 None\n",
    "err_range_text": "std\nit None\nsn 0.01\nchr 1\na-b None\n",
    "err_two_spaces": "std\nit None\nsn 0.01\nchr 1\n1-10  None\n",
    "err_it_not_enough": "std\nit None\nsn 0.01\nchr 1\nit 0.1\n",
    "err_it_all_zero": "std\nit 0\nsn 0.01\nchr 1\nit 0.0\nchr 2\nit 0.1\n",
    "err_meta_no_eq": "species_name\nstd\nit None\nsn 0.01\n",
}


def settings_goldens():
    d = case_dir("rmt")
    rng = numpy.random.default_rng(5)
    contigs = [("chrA desc", rand_seq(rng, 400), 60), ("chrB", rand_seq(rng, 300), 60),
               ("chrC", rand_seq(rng, 2), 60), ("chrD", rand_seq(rng, 250), 50)]
    write_fasta(d / "g.fa", contigs)
    fasta = load_fasta(d / "g.fa")
    res = {}
    for name, text in RMT_CASES.items():
        p = d / f"{name}.rmt"
        p.write_text(text)
        err = io.StringIO()
        try:
            with redirect_stderr(err):
                sim = SimulationSettings.from_rmt(p, fasta, False)
            res[name] = {"ok": dump_sim(sim), "stderr": err.getvalue()}
        except Exception as e:  # noqa: BLE001
            res[name] = {"error": type(e).__name__,
                         "message": str(e).replace(str(p), "<PATH>")}
    # from_args / from_it
    for name, argv in {
        "args_default_sn": ["-sn", "0.01"],
        "args_all": ["-sn", "0.01", "-in", "0.001", "-inmax", "10", "-de", "0.001", "-demax", "10",
                     "-du", "0.0005", "-dumax", "50", "-iv", "0.0005", "-ivmax", "50", "-tl", "0.0005",
                     "-tlmax", "50", "-titv", "2.0", "-snb", "3", "-deb", "0", "-a", "asm", "-s", "sp", "-n", "smp"],
        "args_none": [],
        "args_too_high": ["-sn", "0.3", "-in", "0.3"],
        "args_min_gt_max": ["-de", "0.1", "-demin", "5", "-demax", "2"],
        "args_titv_neg": ["-sn", "0.1", "-titv", "-2"],
    }.items():
        old = sys.argv
        sys.argv = ["x", str(d / "g.fa"), "args"] + argv
        from mutation_simulator.argument_parser import get_args
        a = get_args()
        sys.argv = old
        err = io.StringIO()
        try:
            with redirect_stderr(err):
                sim = SimulationSettings.from_args(a, fasta, False)
            res[name] = {"argv": argv, "ok": dump_sim(sim), "stderr": err.getvalue()}
        except Exception as e:  # noqa: BLE001
            res[name] = {"argv": argv, "error": type(e).__name__, "message": str(e)}
    for name, rate in {"it_ok": 0.01, "it_high": 0.6, "it_zero": 0.0, "it_neg": -0.5}.items():
        try:
            sim = SimulationSettings.from_it(rate, fasta, False)
            res[name] = {"rate": rate, "ok": dump_sim(sim)}
        except Exception as e:  # noqa: BLE001
            res[name] = {"rate": rate, "error": type(e).__name__, "message": str(e)}
    (d / "expected.json").write_text(json.dumps(res, indent=1))


# ----------------------------------------------------------------------------
def main():
    rng = numpy.random.default_rng(2026)

    # 1. ARGS mode, all seven mutation types, three contigs with different line widths
    contigs = [("chr1 first contig", rand_seq(rng, 6000), 60),
               ("chr2", rand_seq(rng, 4500), 70),
               ("chr3 third", rand_seq(rng, 3000), 50)]
    run_case("args_all", contigs,
             ["args", "-sn", "0.01", "-titv", "2.0", "-in", "0.003", "-inmax", "8", "-de", "0.003",
              "-demax", "8", "-du", "0.002", "-dumax", "12", "-iv", "0.002", "-ivmax", "12",
              "-tl", "0.004", "-tlmax", "10", "-a", "asmX", "-s", "Some species", "-n", "smp1"], seed=7)

    # 2. IUPAC codes, soft-masked (lower-case) input, N runs, '-' gaps
    iu = "ACGTNRYKMSWBDHVacgtnrykm-"
    p = numpy.array([18, 18, 18, 18, 6, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 3, 3, 3, 3, 1, 0.5, 0.5, 0.5, 0.5, 0.5])
    p = p / p.sum()
    c1 = rand_seq(rng, 5000, iu, p)
    c1 = c1[:1000] + "N" * 300 + c1[1300:]
    contigs = [("ctgI iupac", c1, 80), ("ctgII", rand_seq(rng, 2500, iu, p), 61)]
    run_case("iupac", contigs,
             ["args", "-sn", "0.03", "-in", "0.004", "-inmax", "6", "-de", "0.004", "-demax", "9",
              "-du", "0.004", "-dumax", "9", "-iv", "0.004", "-ivmax", "9", "-tl", "0.006",
              "-tlmax", "9", "-tlmin", "1"], seed=11)

    # 3. BASELINE config 1 shape, scaled to 60 kbp (full config 1 is a stats/bench case)
    run_case("c1_small", [("chr1", rand_seq(rng, 60000), 60)],
             ["args", "-sn", "0.01", "-in", "0.001", "-inmin", "1", "-inmax", "10",
              "-de", "0.001", "-demin", "1", "-demax", "10"], seed=42)

    # 4. RMT mode: hot/cold ranges, blocked (None) ranges, per-type blocks, titv
    rmt = ("titv=3.0\nspecies_name=Test Species\nassembly_name=T1\nsample_name=SAMPLE\n"
           "sn_block=2\nde_block=5\niv_block=3\n"
           "std\nit None\nsn 0.002\n\n"
           "chr 1\n1-500 None\n501-2500 sn 0.05 in 0.01 inmin 1 inmax 5 de 0.01 demin 2 demax 20\n"
           "3001-4000 None\n4001-6000 iv 0.01 ivmin 2 ivmax 30 du 0.01 dumin 1 dumax 15 tl 0.02 tlmin 3 tlmax 25\n"
           "chr 2\n1-1000 sn 0.0001\n2001-END tl 0.03 tlmin 1 tlmax 4 sn 0.01\n")
    contigs = [("seq1", rand_seq(rng, 8000), 60), ("seq2 second", rand_seq(rng, 5000), 60),
               ("seq3", rand_seq(rng, 1200), 40)]
    run_case("rmt_ranges", contigs, None, seed=3, rmt_text=rmt)

    # 5. IT mode, 4 contigs -> 2 pairs
    contigs = [(f"c{i+1} it", rand_seq(rng, n), w) for i, (n, w) in
               enumerate([(3000, 60), (2600, 50), (4100, 70), (1900, 60)])]
    run_case("it_basic", contigs, ["it", "0.004"], seed=5, it=True, ms=False)

    # 6. RMT with mutations AND it (Mutator -> reload -> ITMutator, __main__.py:88-102)
    rmt = ("std\nit 0.003\nsn 0.01 in 0.002 inmin 1 inmax 4 de 0.002 demin 1 demax 4\n"
           "chr 2\n1-400 None\n")
    contigs = [(f"k{i+1}", rand_seq(rng, n), 60) for i, n in enumerate([2400, 3000, 1800, 2100])]
    run_case("rmt_it", contigs, None, seed=9, it=True, ms=True, rmt_text=rmt)

    # 7. hand-crafted edge cases through the reference's own walk
    s1 = rand_seq(rng, 200)           # exact multiple of bpl=50
    s2 = rand_seq(rng, 137)
    s3 = "ACGT" * 10 + "NNNNNNNNNN" + rand_seq(rng, 100) + "GAATTC" + rand_seq(rng, 44)  # palindrome at 150
    s4 = rand_seq(rng, 90)
    s5 = "RYKMSWBDHVN-acgt" * 8
    s6 = rand_seq(rng, 120)
    contigs = [("e1 ins@0", s1, 50), ("e2 del@0", s2, 60), ("e3", s3, 40), ("e4 untouched", s4, 30),
               ("e5 iupac", s5, 16), ("e6 tli@0", s6, 60)]
    crafted = {
        0: [dict(key=0, type="IN", start=0, stop=3),               # INS at pos 0 (alt anchoring)
            dict(key=10, type="SN", start=10, stop=10),
            dict(key=49, type="IN", start=49, stop=49),            # insert right before a line end
            dict(key=120, type="DU", start=120, stop=131),
            dict(key=190, type="DE", start=190, stop=199)],        # deletion running to the contig end
        1: [dict(key=0, type="DE", start=0, stop=4),               # DEL at pos 0
            dict(key=20, type="IV", start=20, stop=33),
            dict(key=60, type="TL", start=60, stop=66),
            dict(key=100, type="TLI", start=60, stop=66, trans_reverse=True, trans_insert_pos=100),
            dict(key=130, type="DU", start=130, stop=136)],        # dup of the contig tail
        2: [dict(key=42, type="SN", start=42, stop=42),            # SNP on N -> no VCF line
            dict(key=3, type="SN", start=3, stop=3),
            dict(key=150, type="IV", start=150, stop=155),         # palindromic inversion -> REF==ALT
            dict(key=170, type="DE", start=170, stop=170)],
        4: [dict(key=1, type="SN", start=1, stop=1),               # SNP on IUPAC (Y)
            dict(key=8, type="SN", start=8, stop=8),               # SNP on H
            dict(key=17, type="IN", start=17, stop=19),            # anchor base is IUPAC (R)
            dict(key=33, type="DE", start=33, stop=40),
            dict(key=50, type="IV", start=50, stop=60),
            dict(key=70, type="DU", start=70, stop=80),            # DU keeps raw IUPAC
            dict(key=90, type="TL", start=90, stop=99),
            dict(key=110, type="TLI", start=90, stop=99, trans_reverse=False, trans_insert_pos=110),
            dict(key=122, type="SN", start=122, stop=122)],
        5: [dict(key=0, type="TLI", start=50, stop=58, trans_reverse=False, trans_insert_pos=0),
            dict(key=50, type="TL", start=50, stop=58),
            dict(key=100, type="TL", start=100, stop=100),
            dict(key=80, type="TLI", start=100, stop=100, trans_reverse=False, trans_insert_pos=80)],
    }
    run_crafted("edges", contigs, crafted, seed=13, titv=1.0)

    # 8. single-line contigs, bpl larger than contig, 1-base contig
    contigs = [("t1", rand_seq(rng, 35), 60), ("t2", "A", 60), ("t3", rand_seq(rng, 61), 61),
               ("t4", rand_seq(rng, 300), 10)]
    crafted = {0: [dict(key=5, type="IN", start=5, stop=40)],       # pushes a 1-line contig over bpl
               2: [dict(key=60, type="SN", start=60, stop=60)],
               3: [dict(key=0, type="DU", start=0, stop=24),
                   dict(key=100, type="DE", start=100, stop=180),
                   dict(key=250, type="IV", start=250, stop=298)]}
    run_crafted("tiny", contigs, crafted, seed=17)

    # ------------------------------------------------------------------
    settings_goldens()

    # statistics (Gate B)
    sampling_stats("stats_c1", [2_000_000],
                   ["-sn", "0.01", "-in", "0.001", "-inmin", "1", "-inmax", "10",
                    "-de", "0.001", "-demin", "1", "-demax", "10"], seeds=[1, 2, 3, 4])
    sampling_stats("stats_all", [1_000_000, 500_000],
                   ["-sn", "0.01", "-titv", "2.0", "-in", "0.001", "-inmax", "10", "-de", "0.001", "-demax", "10",
                    "-du", "0.0005", "-dumax", "50", "-iv", "0.0005", "-ivmax", "50", "-tl", "0.0005", "-tlmax", "50"],
                   seeds=[1, 2, 3, 4])
    sampling_stats("stats_dense", [200_000],
                   ["-sn", "0.05", "-in", "0.02", "-inmax", "20", "-de", "0.02", "-demax", "40",
                    "-du", "0.01", "-dumax", "40", "-iv", "0.01", "-ivmax", "40", "-tl", "0.04", "-tlmax", "30",
                    "-deb", "10", "-snb", "2"], seeds=[1, 2, 3, 4])
    rmt = ("std\nit None\nsn 0.001\nchr 1\n1-100000 None\n100001-300000 sn 0.05 in 0.005 inmin 1 inmax 10\n"
           "300001-400000 sn 0.0001\n400001-450000 None\n450001-700000 de 0.01 demin 1 demax 100 tl 0.01 tlmin 5 tlmax 50\n")
    sampling_stats("stats_rmt", [1_000_000], [], seeds=[1, 2, 3, 4], rmt_text=rmt)
    titv_stats("stats_titv", [0.0, 0.5, 1.0, 2.0, 10.0], 20000, seeds=[1, 2])
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
