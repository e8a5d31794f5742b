"""CPU restatement (pure Python) of the reference's mutation-injection path.

TEST INFRASTRUCTURE — NOT A PRODUCT PATH.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this module.  The product package
(mutation_simulator_b200) never imports anything under oracle/ and fails loudly
when its CUDA library is missing.

Parity status: PINNED against golden vectors produced by running the unmodified
reference in the build container (tests/golden/make_golden.py; checked by
tests/test_oracle_golden.py).  The reference ships no tests of its own
(SURVEY.md §4), so those reference-run outputs are the pin.

Every function cites the reference file:line it restates
(paths relative to /root/reference/mutation_simulator/).
"""
from __future__ import annotations

import random as _random
from dataclasses import dataclass
from typing import Iterable, Optional

# mutator.py:75-77
NON_AMBIGUOUS = bytes.maketrans(b"KSYMWRBDHV-", b"GCCAAACAAAN")
COMPLEMENT = bytes.maketrans(b"ACGTUMRWSYKVHDB", b"TGCAAKYWSRMBDHV")
TRANSITIONS = bytes.maketrans(b"AGTC", b"GACT")
# mutator.py:449-455 (letters outside the table raise KeyError in the reference;
# the build maps them to themselves)
TRANSVERSIONS = {ord("A"): b"TC", ord("G"): b"CT", ord("T"): b"GA", ord("C"): b"AG", ord("N"): b"NN"}

SN, IN, DE, DU, IV, TL, TLI = "SN", "IN", "DE", "DU", "IV", "TL", "TLI"


@dataclass
class Mut:
    """mutator.py:26-47 — one entry of the walk's ``muts`` dict.

    key     dict key = position at which the walk applies it
    start/stop  0-based inclusive; for TLI these are the linked TL's extent
    alt     SN only: the substituted base the RNG chose (mutator.py:338)
    insert  IN only: the random insert the RNG chose (mutator.py:344)
    """
    key: int
    type: str
    start: int
    stop: int
    reverse: bool = False
    alt: Optional[bytes] = None
    insert: Optional[bytes] = None


def conv(b: bytes) -> bytes:
    """mutator.py:474-479 __convert_ambiguous"""
    return b.translate(NON_AMBIGUOUS)


def revcomp(b: bytes) -> bytes:
    """mutator.py:383,406 — ``[::-1].translate(complement)``"""
    return b[::-1].translate(COMPLEMENT)


def snp_alt(ref: int, titv: float, rng: _random.Random) -> int:
    """mutator.py:429-463 __get_snp/__get_ti_Base/__get_tv_Base"""
    p_ti = titv * (1 / (titv + 1))
    p = rng.uniform(0, 1)
    if p <= p_ti:
        return bytes([ref]).translate(TRANSITIONS)[0]
    tv = TRANSVERSIONS.get(ref)
    if tv is None:
        return ref
    return tv[rng.randint(0, 1)]


def vcf_line(name: bytes, start: int, ref: bytes, alt: bytes, svtype: str, end: int, length: int) -> bytes:
    """vcf_writer.py:44-52,118-126 — empty when REF == ALT"""
    if ref == alt:
        return b""
    info = b"." if svtype == "sn" else b"SVTYPE=%s;END=%d;SVLEN=%d" % (svtype.encode(), end, length)
    return b"%s\t%d\t.\t%s\t%s\t.\t.\t%s\tGT\t1\n" % (name, start, ref, alt, info)


def walk(seq: bytes, name: bytes, muts: Iterable[Mut]):
    """mutator.py:318-426 __mutate_sequence, restated over slices.

    ``seq`` is the upper-cased contig (util.py:87 sequence_always_upper).
    Returns (mutated bases without line breaks, VCF body bytes).
    Mutations whose key falls inside an earlier DE/TL/IV/DU span are skipped,
    exactly as the reference's ``pos`` jump does (mutator.py:376,386,398).
    """
    out = []
    vcf = []
    pos = 0
    L = len(seq)
    for m in sorted(muts, key=lambda x: x.key):
        p = m.key
        if p < pos or p >= L:
            continue
        out.append(seq[pos:p])
        if m.type == SN:  # :334-341
            ref = conv(seq[p:p + 1])
            alt = m.alt
            out.append(alt)
            vcf.append(vcf_line(name, p + 1, ref, alt, "sn", 0, 0))
            pos = p + 1
        elif m.type == IN:  # :343-358
            ins = m.insert
            if p > 0:
                ref = conv(seq[p - 1:p])
                vcf.append(vcf_line(name, p, ref, ref + ins, "INS", p, len(ins)))
            else:
                ref = conv(seq[0:1])
                vcf.append(vcf_line(name, 1, ref, ins + ref, "INS", 1, len(ins)))
            out.append(ins + seq[p:p + 1])
            pos = p + 1
        elif m.type in (DE, TL):  # :360-377
            sv = "DEL" if m.type == DE else "DEL:ME"
            end = m.stop + 1
            if p > 0:
                ref = conv(seq[p - 1:end])
                vcf.append(vcf_line(name, p, ref, ref[0:1], sv, end, m.stop - p + 1))
            else:
                end += 1
                ref = conv(seq[0:end])
                vcf.append(vcf_line(name, 1, ref, ref[-1:], sv, end, m.stop - p + 1))
            pos = m.stop + 1
        elif m.type == IV:  # :379-387
            ref = conv(seq[p:m.stop + 1])
            alt = revcomp(ref)
            out.append(alt)
            vcf.append(vcf_line(name, p + 1, ref, alt, "INV", m.stop + 1, 0))
            pos = m.stop + 1
        elif m.type == DU:  # :389-399
            dupe = seq[p:m.stop + 1]
            out.append(dupe * 2)
            vcf.append(vcf_line(name, p + 1, dupe, dupe * 2, "DUP", p + len(dupe), len(dupe)))
            pos = m.stop + 1
        elif m.type == TLI:  # :401-421
            ins = conv(seq[m.start:m.stop + 1])
            if m.reverse:
                ins = revcomp(ins)
            if p > 0:
                ref = conv(seq[p - 1:p])
                vcf.append(vcf_line(name, p, ref, ref + ins, "INS:ME", p, len(ins)))
            else:
                ref = conv(seq[p:p + 1])
                vcf.append(vcf_line(name, 1, ref, ins + ref, "INS:ME", 1, len(ins)))
            out.append(ins + seq[p:p + 1])
            pos = p + 1
        else:
            raise ValueError(m.type)
    out.append(seq[pos:])
    return b"".join(out), b"".join(vcf)


class FastaOut:
    """fasta_writer.py:13-65 — line wrapping state machine, restated on bytes."""

    def __init__(self):
        self.parts = []
        self.written = 0
        self.bpl = 60

    def set_bpl(self, bpl: int):  # :34-38
        self.bpl = bpl

    def write_header(self, header: bytes):  # :40-47
        if self.written != 0:
            self.parts.append(b"\n")
        self.parts.append(b">" + header + b"\n")
        self.written = 0

    def write_multi(self, bases: bytes):  # :49-65, many bases at once
        i = 0
        n = len(bases)
        while i < n:
            room = self.bpl - self.written
            chunk = bases[i:i + room]
            self.parts.append(chunk)
            self.written += len(chunk)
            i += len(chunk)
            if self.written == self.bpl:
                self.parts.append(b"\n")
                self.written = 0

    def getvalue(self) -> bytes:
        return b"".join(self.parts)


def vcf_header(reference: str, contigs, assembly: str, species: str, sample: str, filedate: Optional[str]) -> bytes:
    """vcf_writer.py:74-116.  ``contigs`` = [(name, length)].  filedate None omits the line
    (the goldens have it stripped because it is wall clock)."""
    h = ["##fileformat=VCFv4.3\n"]
    if filedate is not None:
        h.append(f"##filedate={filedate}\n")
    h.append("##source=Mutation-Simulator\n")
    h.append(f"##reference={reference}\n")
    for n, ln in contigs:
        h.append(f"##contig=<ID={n},length={ln},assembly={assembly},species=\"{species}\">\n")
    h.append("##INFO=<ID=SVTYPE,Number=1,Type=String,Description=\"Type of structural variant\">\n")
    h.append("##INFO=<ID=END,Number=1,Type=Integer,Description=\"End position of the variant described in this record\">\n")
    h.append("##INFO=<ID=SVLEN,Number=.,Type=Integer,Description=\"Difference in length between REF and ALT alleles\">\n")
    h.append("##ALT=<ID=INS,Description=\"Insert\">\n")
    h.append("##ALT=<ID=DEL,Description=\"Deletion\">\n")
    h.append("##ALT=<ID=DUP,Description=\"Duplication\">\n")
    h.append("##ALT=<ID=INV,Description=\"Inversion\">\n")
    h.append("##ALT=<ID=DEL:ME,Description=\"Deletion of mobile element\">\n")
    h.append("##ALT=<ID=INS:ME,Description=\"Insertion of mobile element\">\n")
    h.append("##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n")
    h.append(f"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t{sample}\n")
    return "".join(h).encode()


def mutate_genome(contigs, muts_per_contig):
    """mutator.py:105-142 Mutator.mutate, given the per-contig mutation tables.

    contigs: [(name, long_name, seq_upper, bpl)], muts_per_contig: [list[Mut]].
    Returns (fasta bytes, vcf body bytes)."""
    fa = FastaOut()
    vcf = []
    for (name, long_name, seq, bpl), muts in zip(contigs, muts_per_contig):
        fa.set_bpl(bpl)
        fa.write_header(long_name)
        body, lines = walk(seq, name, muts)
        fa.write_multi(body)
        vcf.append(lines)
    return fa.getvalue(), b"".join(vcf)


# ----------------------------------------------------------------------------
# interchromosomal translocations
# ----------------------------------------------------------------------------
def it_genome(contigs, breakpoints, partners):
    """it_mutator.py:120-142,185-213 — alternate own/partner intervals.

    contigs: [(name, long_name, seq, bpl)]; breakpoints: {idx: {"self": [...], "partner": [...]}};
    partners: {idx: idx}.  Contigs without breakpoints follow SURVEY.md Q1's stance
    (one header, normally wrapped body) — the reference's duplicated header /
    unwrapped body for that case is treated as a bug, not reproduced.
    Returns (fasta bytes, bedpe bytes)."""
    fa = FastaOut()
    bed = []
    for i, (name, long_name, seq, bpl) in enumerate(contigs):
        fa.set_bpl(bpl)
        fa.write_header(long_name)
        if i in breakpoints:
            pi = partners[i]
            pname, _, pseq, _ = contigs[pi]
            a = [0] + list(breakpoints[i]["self"]) + [len(seq)]
            b = [0] + list(breakpoints[i]["partner"]) + [len(pseq)]
            for j in range(len(a) - 1):
                if j % 2:
                    fa.write_multi(pseq[b[j]:b[j + 1]])
                else:
                    fa.write_multi(seq[a[j]:a[j + 1]])
            bed.append(bedpe_rows(name, breakpoints[i]["self"], len(seq), pname,
                                  breakpoints[i]["partner"], len(pseq)))
        else:
            fa.write_multi(seq)
    return fa.getvalue(), b"".join(bed)


def bedpe_rows(chrom: bytes, bp, length, partner: bytes, bpp, plength) -> bytes:
    """bedpe_writer.py:36-55"""
    rows = []
    n = len(bp)
    for i in range(0, n, 2):
        if i != n - 1:
            rows.append(b"%s\t%d\t%d\t%s\t%d\t%d\n" % (chrom, bp[i], bp[i + 1], partner, bpp[i], bpp[i + 1]))
        elif n % 2:
            rows.append(b"%s\t%d\t%d\t%s\t%d\t%d\n" % (chrom, bp[i], length, partner, bpp[i], plength))
    return b"".join(rows)


# ----------------------------------------------------------------------------
# sampling half (used for statistics only; streams need not match)
# ----------------------------------------------------------------------------
def sample_with_minimum_distance(rng: _random.Random, start: int, stop: int, k: int, d: int):
    """util.py:94-109"""
    s = sorted(rng.sample(range(start, stop - (k - 1) * d), k))
    return [v + d * i for i, v in enumerate(s)]


def candidate_count(start: int, stop: int, rate_sum: float) -> int:
    """mutator.py:225"""
    return int(((stop - start) + 1) * rate_sum)


def stop_position(typ, start, minlen, maxlen, chrom_len, rng):
    """mutator.py:229-265.  Returns stop or None (IV that does not fit)."""
    if typ == SN:
        return start
    if typ == IV:
        if start + maxlen[IV] >= chrom_len - 1:
            return None
        return rng.randint(start + minlen[IV] - 1, start + maxlen[IV] - 1)
    if typ == IN:
        return rng.randint(start + minlen[IN] - 1, start + maxlen[IN] - 1)
    if typ in (DU, DE, TL):
        s = rng.randint(start + minlen[typ] - 1, start + maxlen[typ] - 1)
        return min(s, chrom_len - 1)
    return 0  # TLI: untouched (mutator.py:238-265 has no branch)


def greedy_accept(cands, block, minlen, maxlen, chrom_len, rng):
    """mutator.py:184-213.  cands: [(pos, type)] ascending.  Returns accepted
    [(pos, type, stop)], tls, tlis."""
    acc, tls, tlis = [], [], []
    last_hi = -1  # exclusive upper end of last_mut_range; its low end is <= every later pos
    last_lo = 0
    for pos, typ in cands:
        if last_lo <= pos < last_hi:
            continue
        stop = stop_position(typ, pos, minlen, maxlen, chrom_len, rng)
        if stop is None:
            continue
        acc.append((pos, typ, stop))
        last_lo = pos
        last_hi = (pos if typ in (SN, IN) else stop) + 1 + block[typ]
        if typ == TL:
            tls.append(pos)
        if typ == TLI:
            tlis.append(pos)
    return acc, tls, tlis
