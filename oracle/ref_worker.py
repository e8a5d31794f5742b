#!/usr/bin/env python3
"""Runs the UNMODIFIED reference CLI (oracle/_ref, see make_ref.py) in this process, with `random` and `numpy.random`
seeded and `pyfaidx` served by the in-memory stand-in (oracle/pyfaidx_standin).  TEST / BENCH INFRASTRUCTURE ONLY:
bench.py times this as the CPU baseline (`kind: "reference"`); nothing in the product imports it.

    python oracle/ref_worker.py SEED <mutation-simulator argv...>
"""
import random
import runpy
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent


def main():
    seed = int(sys.argv[1])
    sys.path[:0] = [str(HERE / "_ref"), str(HERE / "pyfaidx_standin")]
    import numpy
    random.seed(seed)
    numpy.random.seed(seed)
    sys.argv = ["mutation-simulator"] + sys.argv[2:]
    runpy.run_module("mutation_simulator", run_name="__main__")


if __name__ == "__main__":
    main()
