"""world_size-2 test (gloo, CPU) of the multi-GPU host logic: contig partitioning and the
max/sum reductions bench.py uses for device-timed numbers."""
import json
import subprocess
import sys
from pathlib import Path

import bench

HERE = Path(__file__).resolve().parent


def test_lpt_partition_properties():
    for n in (1, 2, 4, 8):
        parts = bench.lpt_partition(bench.GRCH38, n)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(24))
        loads = [sum(bench.GRCH38[i] for i in p) for p in parts]
        assert max(loads) <= 1.06 * sum(bench.GRCH38) / n     # near-perfect balance on the GRCh38 shape
    parts = bench.lpt_partition([5000] * 1000, 8)
    assert sorted(len(p) for p in parts) == [125] * 8


def test_two_rank_gloo_reduction(tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(HERE / "dist_worker.py"), str(tmp_path)]
    subprocess.run(cmd, check=True, timeout=300, capture_output=True)
    r0 = json.loads((tmp_path / "rank0.json").read_text())
    r1 = json.loads((tmp_path / "rank1.json").read_text())
    assert sorted(r0["mine"] + r1["mine"]) == list(range(24)) and not set(r0["mine"]) & set(r1["mine"])
    total = float(sum(bench.GRCH38))
    assert r0["sum"] == r1["sum"] == [total, 24.0]
    l0 = sum(bench.GRCH38[i] for i in r0["mine"]); l1 = sum(bench.GRCH38[i] for i in r1["mine"])
    assert r0["max"] == r1["max"] == [float(max(l0, l1)), 1.0]


def test_shard_mode_switches_to_tiles_when_contigs_cannot_fill_the_ranks(monkeypatch):
    from mutation_simulator_b200.distributed import shard_mode
    monkeypatch.delenv("MS_SHARD", raising=False)
    assert shard_mode(bench.GRCH38, 1) == "contigs"
    assert shard_mode(bench.GRCH38, 8) == "contigs"
    assert shard_mode([1_000_000_000], 8) == "tiles"               # README.md:441: one contig, eight GPUs
    assert shard_mode([5_000_000] * 3, 8) == "tiles"               # fewer contigs than ranks
    assert shard_mode([900_000_000] + [1_000_000] * 23, 8) == "tiles"
    assert shard_mode([5000] * 200_000, 8) == "contigs"
    monkeypatch.setenv("MS_SHARD", "tiles")
    assert shard_mode(bench.GRCH38, 8) == "tiles"
