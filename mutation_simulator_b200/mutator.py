"""``Mutator`` — the reference's class API (mutator.py:73-142) on the B200 path.

``Mutator(args, fasta, sim).mutate()`` samples positions, types and lengths for every
range of every contig, resolves overlaps, links translocations, splices the genome and
formats FASTA + VCF — all on the GPU through libmutsim_b200 — then writes the two files.
Nothing here touches a base on the host."""
from __future__ import annotations

import secrets
import sys

import numpy as np

from .engine import BUF_FASTA, BUF_VCF, Engine
from .fasta_writer import FastaWriter
from .plan import block_list, build_ranges, p_transition
from .util import format_warning
from .vcf_writer import VcfWriter


def run_seed(args) -> int:
    """--seed if given; otherwise a fresh 64-bit seed (the reference's RNGs are unseeded too)."""
    s = getattr(args, "seed", None)
    return secrets.randbits(63) if s is None else int(s)


class Mutator:
    def __init__(self, args, fasta, sim):
        self._args, self._fasta, self._sim = args, fasta, sim
        self._fasta_writer = FastaWriter(args.outfasta)
        self._vcf_writer = VcfWriter(args.outvcf)
        self._vcf_writer.write_header(args.infile.name, fasta, sim.assembly_name, sim.species_name, sim.sample_name)
        self._engine = None
        self.stats = None

    def close(self):
        self._fasta_writer.close()
        self._vcf_writer.close()
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def mutate(self):
        """Creates random mutations and writes them to a Fasta and VCF file."""
        args, fasta, sim = self._args, self._fasta, self._sim
        eng = self._engine = Engine(getattr(args, "device", 0))
        fasta.upload(eng)
        ranges, n = build_ranges(sim, fasta.lengths)
        blocks = block_list(sim)
        eng.set_ranges_array(ranges, n, blocks, min(sim.mut_block.values()), p_transition(sim.titv))
        eng.sample(run_seed(args))
        eng.apply()
        recs_contig = eng.records()["contig"]
        if not args.ignore_warnings:
            per = np.bincount(recs_contig, minlength=len(fasta.names))
            for i in np.flatnonzero(per == 0):
                print(format_warning(f"No mutations could be generated on sequence {int(i)+1} (mutation rates too low)",
                                     args.no_color), file=sys.stderr)
        self._fasta_writer.write_image(eng.download(BUF_FASTA))
        self._vcf_writer.write_body(eng.download(BUF_VCF))
        self.stats = eng.stats()
