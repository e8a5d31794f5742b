"""Import alias: put ``<repo>/dropin`` (and ``<repo>``) on PYTHONPATH and existing
``import mutation_simulator`` code runs on the B200 path unchanged."""
from mutation_simulator_b200 import *  # noqa: F401,F403
from mutation_simulator_b200 import (Mutator, ITMutator, SimulationSettings, MutType, get_args, load_fasta,  # noqa: F401
                                     __version__)
