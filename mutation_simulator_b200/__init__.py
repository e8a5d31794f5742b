"""mutation_simulator_b200 — B200-native drop-in for Mutation-Simulator's mutation-injection
path.  Re-exports the names of the reference's package (__init__.py:1-18); the engines run
on the GPU through libmutsim_b200.so and raise if it (or a CUDA device) is missing."""
from ._lib import MutSimError
from ._version import __version__
from .argument_parser import get_args
from .bedpe_writer import BedpeWriterError
from .colors import Colors
from .fasta import Fasta, FastaIndexingError, FastaNotFoundError
from .fasta_writer import FastaWriterError
from .it_mutator import ITMutator
from .mut_types import MutType
from .mutator import Mutator
from .plan import RangeOverlapError
from .rmt import (ChromNotExistError, ITNotEnoughAvailChromsError, ItRateTooHighError, ItRateTooLowError,
                  MinimumLengthHigherThanMaximumError, MinimumLengthTooLowError, MissingLengthError,
                  RangeDefinitionOutOfBoundsError, RatesTooHighError, RatesTooLowError, RMTParseError, SimulationSettings,
                  TitvTooLowError)
from .util import (FastaDuplicateHeaderError, exit_with_error, format_warning, get_md5, load_fasta, print_success,
                   print_warning)
from .vcf_writer import VcfWriterError
