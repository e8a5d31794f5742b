"""Simulation settings and the RMT parser.

Host-side restatement of the reference's rmt.py (the object model consumed by the
engines and the Random-Mutation-Table file format).  It defines the work list of
the GPU path — per range the exact candidate count, the type probabilities, the
length bounds and the block values — so it is kept behaviour-identical: same
attribute names, same validation order, same exception classes and messages,
same quirks (values are lower-cased, unknown keywords are ignored, the bounds
check allows stop == length, overlapping ranges yield negative fillers).
Differentially tested against outputs of the reference (tests/golden/rmt).
"""
from __future__ import annotations

from typing import Optional

from .defaults import Defaults
from .mut_types import MutType
from .util import print_warning

_TYPES_WITH_RATE = ("sn", "in", "de", "iv", "du", "tl")
_MAX_KEYS = {f"{t}max": t for t in _TYPES_WITH_RATE if t != "sn"}
_MIN_KEYS = {f"{t}min": t for t in _TYPES_WITH_RATE if t != "sn"}
_META_TEXT = ("fasta", "md5", "species_name", "assembly_name", "sample_name")
_META_NUMBER = ("titv",)
_META_BLOCK = {f"{t}_block": t for t in _TYPES_WITH_RATE}


class RatesTooHighError(Exception):
    pass


class RatesTooLowError(Exception):
    pass


class ItRateTooHighError(Exception):
    pass


class ItRateTooLowError(Exception):
    pass


class ITNotEnoughAvailChromsError(Exception):
    pass


class TitvTooLowError(Exception):
    pass


class RMTParseError(Exception):
    pass


class MissingLengthError(Exception):
    pass


class MinimumLengthTooLowError(Exception):
    pass


class MinimumLengthHigherThanMaximumError(Exception):
    pass


class ChromNotExistError(Exception):
    pass


class RangeDefinitionOutOfBoundsError(Exception):
    pass


class MutationSettings:
    """Rates, lengths and derived type probabilities of one range (rmt.py:79-163)."""

    def __init__(self, mut_rates: Optional[dict], mut_lengs: Optional[dict]):
        self.mut_rates = mut_rates
        self.mut_lengs = mut_lengs
        self._check()
        if self.mut_rates and MutType.TL in self.mut_rates:
            half = self.mut_rates[MutType.TL] / 2          # a translocation = one cut + one paste
            self.mut_rates[MutType.TL] = half
            self.mut_rates[MutType.TLI] = half
        if self.mut_rates:
            total = sum(self.mut_rates.values())
            self.mut_chances = {t: r / total for t, r in self.mut_rates.items()}
        else:
            self.mut_chances = None

    def _check(self):
        rates = self.mut_rates
        if not rates:
            return
        if any(r < 0 for r in rates.values()) or sum(rates.values()) <= 0:
            names = ", ".join(t.name for t, r in rates.items() if r <= 0)
            raise RatesTooLowError(f"Mutation rate/s too low for: {names}. If this was intentional use the None keyword instead")
        if sum(rates.values()) > 0.5:
            names = ", ".join(t.name for t, r in rates.items() if r > 1)
            raise RatesTooHighError(f"Mutation rate/s too high for: {names}" if names else "Sum of mutation rates too high")
        for t in rates:
            if t is MutType.SN:
                continue
            lengs = self.mut_lengs
            if not lengs or t not in lengs["min"] or t not in lengs["max"]:
                raise MissingLengthError(f"Missing length keyword for: {t.name}")
            lo, hi = lengs["min"][t], lengs["max"][t]
            if lo > hi:
                raise MinimumLengthHigherThanMaximumError(f"Minimum length is greater and maximum length for: {t.name}")
            if lo < (2 if t is MutType.IV else 1):
                raise MinimumLengthTooLowError(f"Minimum length too low for: {t.name}")

    def __repr__(self):
        return f"Rates: {self.mut_rates}, Chances: {self.mut_chances}, Lengs: {self.mut_lengs}"

    @property
    def has_mutations(self) -> bool:
        return bool(self.mut_rates) and any(self.mut_rates.values())


class RangeDefinition:
    """0-based inclusive range with its settings (rmt.py:166-189)."""

    def __init__(self, start: int, stop, mutation_settings: MutationSettings):
        self.start, self.stop, self.mutation_settings = start, stop, mutation_settings

    def __repr__(self):
        return f"{self.start}-{self.stop} {self.mutation_settings}"

    def eval_end(self, chrom_length: int):
        if self.stop == "end":
            self.stop = chrom_length - 1


class ChromosomeSettings:
    """it rate and ranges of one contig (rmt.py:192-258)."""

    def __init__(self, number: int, it_rate: Optional[float], range_definitions: list):
        self.number = number
        self.it_rate = it_rate
        if it_rate and it_rate > 0.5:
            raise ItRateTooHighError("Interchromosomal translocation rate too high")
        if it_rate and it_rate < 0:
            raise ItRateTooLowError("Interchromosomal translocation rate too low")
        self.range_definitions = range_definitions

    def __repr__(self):
        return f"{self.number}\nit={self.it_rate}\n{self.range_definitions}\n"

    def fill_missing_ranges(self, chrom_length: int, std: MutationSettings):
        """Gap filling with the standard settings; mirrors the reference's order of
        insertion (head, tail, then gaps) so that overlapping input yields the same
        (negative-length) fillers."""
        rds = self.range_definitions
        if not rds:
            rds.append(RangeDefinition(0, chrom_length - 1, std))
            return
        if rds[0].start != 0:
            rds.insert(0, RangeDefinition(0, rds[0].start - 1, std))
        if rds[-1].stop != chrom_length - 1:
            rds.append(RangeDefinition(rds[-1].stop + 1, chrom_length - 1, std))
        i = 0
        while i < len(rds) - 1:
            if rds[i].stop + 1 != rds[i + 1].start:
                rds.insert(i + 1, RangeDefinition(rds[i].stop + 1, rds[i + 1].start - 1, std))
            i += 1


class SimulationSettings:
    """Everything the engines need (rmt.py:261-557)."""

    def __init__(self, std: MutationSettings, std_it: Optional[float], chromosomes: list, mut_block: Optional[dict],
                 fasta=None, md5=None, titv: float = Defaults.TITV, species_name: str = Defaults.SPECIES_NAME,
                 assembly_name: str = Defaults.ASSEMBLY_NAME, sample_name: str = Defaults.SAMPLE_NAME,
                 ignore_warnings: bool = Defaults.IGNORE_WARNINGS, no_color: bool = Defaults.NO_COLOR):
        self._std = std
        self._std_it = std_it
        self.chromosomes = chromosomes
        self.mut_block = self._normalise_blocks(mut_block, ignore_warnings, no_color)
        self.fasta = fasta
        self.md5 = md5
        self.titv = titv
        if self.titv < 0:
            raise TitvTooLowError("Titv value is below 0")
        self.species_name = species_name
        self.assembly_name = assembly_name
        self.sample_name = sample_name

    def __repr__(self):
        return (f"[META]\nfasta={self.fasta}\nmd5={self.md5}\ntitv={self.titv}\nspecies_name={self.species_name}\n"
                f"assembly_name={self.assembly_name}\nsample_name={self.sample_name}\nmut_block={self.mut_block}\n\n"
                f"[STD]\nit={self._std_it}\n{self._std}\n\n[RD]\n{self.chromosomes}")

    @staticmethod
    def _normalise_blocks(blocks, ignore_warnings, no_color):
        if not blocks:
            return Defaults.MUT_BLOCK
        for t in MutType:
            if t is MutType.TLI:
                continue
            if t not in blocks:
                blocks[t] = 1
            elif blocks[t] < 1:
                blocks[t] = 1
                if not ignore_warnings:
                    print_warning(f"'{t.name}' block was set to 1", no_color)
        blocks[MutType.TLI] = blocks[MutType.TL]
        return blocks

    # -- completion against the FASTA -------------------------------------
    def _validate_it(self, fasta):
        avail = [c.number for c in self.chromosomes if c.it_rate is not None and len(fasta[c.number]) > 2]
        if len(avail) < 2:
            raise ITNotEnoughAvailChromsError("Not enought available chromosomes for interchromosomal translocations")
        if sum(self.chromosomes[n].it_rate for n in avail) == 0:
            raise ItRateTooLowError("Interchromosomal translocation rates are too low")

    def _check_chroms_exist(self, fasta):
        n = len(fasta.keys())
        for c in self.chromosomes:
            if not (-n <= c.number < n):
                raise ChromNotExistError(f"Chromosome {c.number+1} does not exist in the fasta file")

    def _eval_chrom_ends(self, fasta):
        for c in self.chromosomes:
            if c.range_definitions:
                c.range_definitions[-1].eval_end(len(fasta[c.number]))

    def _fill_missing_chroms(self, fasta):
        have = {c.number for c in self.chromosomes}
        for idx in range(len(fasta.keys())):
            if idx not in have:
                self.chromosomes.append(ChromosomeSettings(idx, self._std_it, [RangeDefinition(0, len(fasta[idx]) - 1, self._std)]))

    def _sort(self):
        self.chromosomes = sorted(self.chromosomes, key=lambda c: c.number)
        for c in self.chromosomes:
            c.range_definitions = sorted(c.range_definitions, key=lambda rd: rd.start)

    def _check_bounds(self, fasta):
        names = fasta.keys()
        for c in self.chromosomes:
            if not c.range_definitions:
                continue
            if c.range_definitions[0].start < 0:
                raise RangeDefinitionOutOfBoundsError(f"A range definition of chromosome {c.number+1} is starting at 0")
            if c.range_definitions[-1].stop > len(fasta[names[c.number]]):
                raise RangeDefinitionOutOfBoundsError(f"A range definition of chromosome {c.number+1} is longer than the chromosome")

    def _fill_missing_ranges(self, fasta):
        for c in self.chromosomes:
            c.fill_missing_ranges(len(fasta[c.number]), self._std)

    # -- constructors -------------------------------------------------------
    @classmethod
    def from_args(cls, args, fasta, ignore_warnings: bool) -> "SimulationSettings":
        T = MutType
        rates = {T.SN: args.snp, T.IN: args.insert, T.DE: args.deletion, T.IV: args.inversion, T.DU: args.duplication,
                 T.TL: args.translocation}
        lengs = {"min": {T.IN: args.insertminlength, T.DE: args.deletionminlength, T.IV: args.inversionminlength,
                         T.DU: args.duplicationminlength, T.TL: args.translocationminlength},
                 "max": {T.IN: args.insertmaxlength, T.DE: args.deletionmaxlength, T.IV: args.inversionmaxlength,
                         T.DU: args.duplicationmaxlength, T.TL: args.translocationmaxlength}}
        blocks = {T.SN: args.snpblock, T.IN: args.insertblock, T.DE: args.deletionblock, T.IV: args.inversionblock,
                  T.DU: args.duplicationblock, T.TL: args.translocationblock}
        sim = cls(MutationSettings(rates, lengs), None, [], blocks, titv=args.transitionstransversions,
                  species_name=args.species, assembly_name=args.assembly, sample_name=args.sample,
                  ignore_warnings=ignore_warnings)
        sim._fill_missing_chroms(fasta)
        sim._sort()
        return sim

    @classmethod
    def from_it(cls, it_rate: float, fasta, ignore_warnings: bool) -> "SimulationSettings":
        sim = cls(MutationSettings(None, None), it_rate, [], None, ignore_warnings=ignore_warnings)
        sim._fill_missing_chroms(fasta)
        sim._sort()
        sim._validate_it(fasta)
        return sim

    @classmethod
    def from_rmt(cls, path, fasta, ignore_warnings: bool) -> "SimulationSettings":
        sections = _split_sections(_read_lines(path))
        if len(sections["std"]) != 2:
            raise RMTParseError(f"Standard section not defined or malformed. Occurred while reading {path}")
        try:
            meta, blocks = _parse_meta(sections["meta"])
            std_it = _parse_it(sections["std"][0])
            std = _parse_settings(sections["std"][1])
            chroms = _parse_ranges(sections["rd"], std_it)
            sim = cls(std, std_it, chroms, blocks, ignore_warnings=ignore_warnings, **meta)
            sim._check_chroms_exist(fasta)
            sim._eval_chrom_ends(fasta)
            sim._fill_missing_chroms(fasta)
            sim._sort()
            sim._check_bounds(fasta)
            sim._fill_missing_ranges(fasta)
            if sim.has_it:
                sim._validate_it(fasta)
        except Exception as e:  # noqa: BLE001 - the reference re-raises every type with the file appended
            raise type(e)(f"{e}. Occurred while reading {path}")
        return sim

    @property
    def has_mutations(self) -> bool:
        return any(rd.mutation_settings.has_mutations for c in self.chromosomes for rd in c.range_definitions)

    @property
    def has_it(self) -> bool:
        return any(c.it_rate for c in self.chromosomes)


# ---- RMT text -------------------------------------------------------------
def _read_lines(path) -> list:
    out = []
    with open(path, "r") as fh:
        for line in fh.readlines():
            if line.startswith("#"):
                continue
            line = line.split("#")[0].strip()
            if line:
                out.append(line)
    return out


def _split_sections(lines) -> dict:
    sections = {"meta": [], "std": [], "rd": []}
    where = "meta"
    for line in lines:
        line = line.lower()
        if line == "std":
            where = "std"
            continue
        if line.startswith("chr"):
            where = "rd"
        sections[where].append(line)
    return sections


def _parse_meta(lines):
    meta, blocks = {}, {}
    for line in lines:
        key, val = [tok.strip() for tok in line.split("=")]  # a line without exactly one '=' is a ValueError, as in the reference
        if key in _META_TEXT:
            meta[key] = val
        elif key in _META_NUMBER:
            try:
                meta[key] = float(val)
            except ValueError:
                raise RMTParseError(f"{key.capitalize()} value of '{val}' is not representable as a float")
        elif key in _META_BLOCK:
            try:
                blocks[MutType[_META_BLOCK[key].upper()]] = int(val)
            except ValueError:
                raise RMTParseError(f"Mut block value of {key} is not representable as an integer")
    return meta, blocks


def _tokens(line: str) -> list:
    return [t.strip() for t in line.split(" ") if t.strip()]


def _parse_it(line: str) -> Optional[float]:
    tok = _tokens(line)
    if tok[0] != "it" or len(tok) != 2:
        raise RMTParseError("Malformed interchromosomal translocation rate setting")
    if tok[1] == "none":
        return None
    try:
        return float(tok[1])
    except ValueError:
        raise RMTParseError("Malformed interchromosomal translocation rate setting")


def _parse_settings(line: str) -> MutationSettings:
    if line == "none":
        return MutationSettings(None, None)
    tok = _tokens(line)
    if len(tok) % 2:
        raise RMTParseError("Malformed mutation settings")
    rates, lengs = {}, {"min": {}, "max": {}}
    for i, word in enumerate(tok[:-1]):  # every token is tried as a keyword, exactly like the reference
        try:
            if word in _TYPES_WITH_RATE:
                rates[MutType[word.upper()]] = float(tok[i + 1])
            elif word in _MAX_KEYS:
                lengs["max"][MutType[_MAX_KEYS[word].upper()]] = int(tok[i + 1])
            elif word in _MIN_KEYS:
                lengs["min"][MutType[_MIN_KEYS[word].upper()]] = int(tok[i + 1])
        except ValueError:
            raise RMTParseError("Malformed mutation settings")
    return MutationSettings(rates, lengs)


def _parse_range(text: str):
    parts = [t.strip() for t in text.split("-") if t.strip()]
    if len(parts) != 2 or text.count("-") != 1:
        raise RMTParseError("Malformed range in range definitions")
    try:
        start = int(parts[0]) - 1
        stop = parts[1] if parts[1] == "end" else int(parts[1]) - 1
    except ValueError:
        raise RMTParseError("Malformed range in range definitions")
    return start, stop


def _parse_ranges(lines, std_it) -> list:
    table = {}
    cur = None
    for row in lines:
        if row.startswith("chr"):
            try:
                cur = int(row.split(" ")[-1]) - 1
            except ValueError:
                raise RMTParseError(f"Chromosome index {row} is invalid")
            table[cur] = {"rds": []}
        elif row.startswith("it"):
            table[cur]["it"] = _parse_it(row)
        else:
            rng = row.partition(" ")[0]
            start, stop = _parse_range(rng)
            table[cur]["rds"].append(RangeDefinition(start, stop, _parse_settings(row.removeprefix(rng + " "))))
    return [ChromosomeSettings(idx, e.get("it", std_it) if "it" in e else std_it, e["rds"]) for idx, e in table.items()]
