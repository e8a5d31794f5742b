"""Where the file-to-file second goes: cProfile over load_fasta + Mutator.mutate() on the C2 genome in /dev/shm.
usage (GPU box): python profiles/f2f_profile.py > gpurun_out/f2f_profile.txt"""
import cProfile
import pstats
import shutil
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import bench  # noqa: E402
from mutation_simulator_b200 import Mutator, SimulationSettings, get_args, load_fasta  # noqa: E402
from mutation_simulator_b200.engine import BUF_FASTA, Engine  # noqa: E402
from mutation_simulator_b200.records import REC_DTYPE  # noqa: E402

wl = bench.WORKLOADS["c2"]
lengths = list(wl["lengths"])
names = wl["names"] or [f"ctg{i}" for i in range(len(lengths))]
d = Path("/dev/shm/ms_f2f_profile")
d.mkdir(parents=True, exist_ok=True)
try:
    g = Engine(0)
    g.synth_genome(4242, lengths, [60] * len(lengths), [n.encode() for n in names], [n.encode() for n in names],
                   wl["n_fraction"], wl["telomere"])
    g.load_records(np.zeros(0, dtype=REC_DTYPE))
    g.apply()
    with open(d / "genome.fa", "wb") as fh:
        fh.write(memoryview(g.download(BUF_FASTA)))
        fh.write(b"\n")
    g.close()
    argv = [str(d / "genome.fa"), "-o", str(d / "out"), "-q", "--seed", "7"] + bench.reference_cli_args(wl)

    def once():
        a = get_args(argv)
        a.device = 0
        t0 = time.perf_counter()
        fasta = load_fasta(a.infile, device=0)
        t1 = time.perf_counter()
        sim = SimulationSettings.from_args(a, fasta, True)
        m = Mutator(a, fasta, sim)
        t2 = time.perf_counter()
        m.mutate()
        t3 = time.perf_counter()
        m.close()
        fasta.close()
        t4 = time.perf_counter()
        print(f"load_fasta {t1 - t0:.3f}  settings+Mutator() {t2 - t1:.3f}  mutate() {t3 - t2:.3f}  close {t4 - t3:.3f}  total {t4 - t0:.3f}")

    once()                 # first run: .fai is built
    once()
    pr = cProfile.Profile()
    pr.enable()
    once()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
finally:
    shutil.rmtree(d, ignore_errors=True)
