"""``python -m mutation_simulator_b200`` / ``mutation-simulator``: the reference's
entry point (__main__.py:34-107) — same modes, messages and exit codes."""
from __future__ import annotations

import os
from timeit import default_timer as timer

from . import (BedpeWriterError, ChromNotExistError, FastaDuplicateHeaderError, FastaIndexingError, FastaNotFoundError,
               FastaWriterError, ITMutator, ITNotEnoughAvailChromsError, ItRateTooHighError, ItRateTooLowError,
               MinimumLengthHigherThanMaximumError, MinimumLengthTooLowError, MissingLengthError, Mutator, MutSimError,
               RangeDefinitionOutOfBoundsError, RangeOverlapError, RatesTooHighError, RatesTooLowError, RMTParseError,
               SimulationSettings, TitvTooLowError, VcfWriterError, exit_with_error, get_args, get_md5, load_fasta,
               print_success, print_warning)

_SETUP_ERRORS = (FileNotFoundError, ITNotEnoughAvailChromsError, RatesTooHighError, RatesTooLowError, FastaIndexingError,
                 FastaNotFoundError, ItRateTooHighError, ItRateTooLowError, RMTParseError, MissingLengthError,
                 MinimumLengthTooLowError, TitvTooLowError, ChromNotExistError, RangeDefinitionOutOfBoundsError,
                 FastaDuplicateHeaderError, MinimumLengthHigherThanMaximumError)


def initialize():
    args = get_args()
    try:
        from . import distributed
        # the file is parsed on the GPU and stays there for the engines; with several GPUs every
        # rank ingests it and later keeps its share of the contigs
        multi = distributed.rank_world()[1] > 1
        fasta = load_fasta(args.infile, device=distributed.local_device(getattr(args, "device", 0)) if multi
                           else getattr(args, "device", 0))
        if args.mode == "args":
            sim = SimulationSettings.from_args(args, fasta, args.ignore_warnings)
        elif args.mode == "it":
            sim = SimulationSettings.from_it(args.interchromosomalrate, fasta, args.ignore_warnings)
        else:
            sim = SimulationSettings.from_rmt(args.rmtfile, fasta, args.ignore_warnings)
    except _SETUP_ERRORS + (MutSimError,) as e:   # MutSimError: no usable CUDA device / library (there is no CPU fallback)
        exit_with_error(e, args.no_color)
    if not args.ignore_warnings:
        warn_user(args, sim)
    return args, fasta, sim


def warn_user(args, sim):
    if sim.fasta and args.infile.name != sim.fasta:
        print_warning("Fasta filename does not match RMT", args.no_color)
    if sim.md5 and get_md5(args.infile) != sim.md5:
        print_warning("Fasta md5 hash does not match RMT", args.no_color)


def main():
    start = timer()
    args, fasta, sim = initialize()
    from . import distributed
    resident = None
    if sim.has_mutations:
        try:
            mutator = Mutator(args, fasta, sim)
            mutator.mutate()
            # the mutated genome stays in HBM for the IT step (MS_NO_CHAIN=1 forces the reference's reload of *_ms.fa);
            # with several GPUs every rank keeps the contigs it mutated and the IT step reuses the partition
            if sim.has_it and not os.environ.get("MS_NO_CHAIN"):
                world = distributed.rank_world()[1]
                ok = mutator._engine is not None and mutator.shard == "contigs" and len(mutator.my_ids) > 0 \
                    and int(mutator._engine.contig_out_len().min()) > 0
                if world > 1:
                    ok = all(distributed.all_gather_object(bool(ok)))
                if ok:
                    from .fasta import ResidentFasta
                    resident = ResidentFasta(mutator.detach_engine(), fasta, mutator.my_ids if world > 1 else None, mutator.parts)
                    fasta.detach_engine()
            mutator.close()
            fasta.close()
        except (FastaWriterError, VcfWriterError, MutSimError, RangeOverlapError) as e:
            exit_with_error(e, args.no_color)
    if sim.has_it:
        if resident is not None:
            fasta = resident
        elif sim.has_mutations:   # IT runs on the mutated genome (__main__.py:88-95)
            distributed.barrier()   # all ranks have written their slices of *_ms.fa
            try:
                fasta = load_fasta(args.outfasta)
            except (FastaDuplicateHeaderError, FastaIndexingError, FastaNotFoundError) as e:
                exit_with_error(e, args.no_color)
        try:
            it_mutator = ITMutator(args, fasta, sim)
            it_mutator.mutate()
            it_mutator.close()
            fasta.close()
        except (FastaWriterError, BedpeWriterError, MutSimError) as e:
            exit_with_error(e, args.no_color)
    runtime = round(timer() - start, 4)
    if distributed.rank_world()[0] != 0:
        return
    if not args.quiet:
        print_success(f"Mutation-Simulator finished in: {runtime}s", args.no_color)


if __name__ == "__main__":
    main()
