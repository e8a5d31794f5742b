"""``Mutator`` — the reference's class API (mutator.py:73-142) on the B200 path.

``Mutator(args, fasta, sim).mutate()`` samples positions, types and lengths for every
range of every contig, resolves overlaps, links translocations, splices the genome and
formats FASTA + VCF — all on the GPU through libmutsim_b200 — then writes the two files.
Nothing here touches a base on the host."""
from __future__ import annotations

import secrets
import sys

import numpy as np

from . import distributed as D
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
        self._rank, self._world = D.init()
        self._engine = None
        self.stats = None
        if self._world > 1:      # one process per GPU: files are assembled by write_partitioned()
            self._fasta_writer = self._vcf_writer = None
            return
        self._fasta_writer = FastaWriter(args.outfasta)
        self._vcf_writer = VcfWriter(args.outvcf)
        self._vcf_writer.write_header(args.infile.name, fasta, sim.assembly_name, sim.species_name, sim.sample_name)

    def close(self):
        if self._fasta_writer is not None:
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
        world = self._world
        n_contigs = len(fasta.names)
        my_ids = D.lpt_partition(fasta.lengths, world)[self._rank] if world > 1 else list(range(n_contigs))
        seed = D.broadcast_object(run_seed(args))
        eng = self._engine = Engine(D.local_device(getattr(args, "device", 0)) if world > 1 else getattr(args, "device", 0))
        fasta.upload(eng, my_ids if world > 1 else None)
        ranges, n = build_ranges(sim, fasta.lengths, my_ids)
        eng.set_ranges_array(ranges, n, block_list(sim), min(sim.mut_block.values()), p_transition(sim.titv))
        eng.sample(seed)
        eng.apply()
        if not args.ignore_warnings:
            per = eng.contig_records()
            for i in np.flatnonzero(per == 0):
                print(format_warning(f"No mutations could be generated on sequence {my_ids[int(i)]+1} (mutation rates too low)",
                                     args.no_color), file=sys.stderr)
        if world == 1:
            self._fasta_writer.write_image(eng.download(BUF_FASTA))
            self._vcf_writer.write_body(eng.download(BUF_VCF))
        else:
            from .vcf_writer import header_text
            chunks, vcf_off = D.fasta_chunks(eng, my_ids, n_contigs)
            D.write_partitioned(args.outfasta, my_ids, chunks, n_contigs)
            head = header_text(args.infile.name, [(fasta[k].name, len(fasta[k])) for k in fasta.keys()], sim.assembly_name,
                               sim.species_name, sim.sample_name).encode("latin-1")
            D.write_partitioned(args.outvcf, my_ids, D.vcf_chunks(eng, my_ids, vcf_off), n_contigs, prefix=head)
        self.stats = eng.stats()
