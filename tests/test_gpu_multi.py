"""N-GPU output == 1-GPU output, byte for byte (needs >= 2 GPUs; run with `gpurun --gpus 2`).
ARGS mode shards contigs with no collective; IT mode reads partner contigs from the owning GPU."""
import os
import shutil
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


def _ngpu():
    import torch
    return torch.cuda.device_count()


def cli(argv, nproc, port, **extra_env):
    env = dict(os.environ, PYTHONPATH=str(REPO), **extra_env)
    if nproc == 1:
        cmd = [sys.executable, "-m", "mutation_simulator_b200"] + argv
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
               f"--master-port={port}", "-m", "mutation_simulator_b200"] + argv
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def make_genome(path, lens, seed, bpl=60):
    rng = np.random.default_rng(seed)
    with open(path, "w") as fh:
        for i, n in enumerate(lens):
            s = rng.choice(np.frombuffer(b"ACGT", np.uint8), n).tobytes().decode()
            fh.write(f">ctg{i+1} len={n}\n")
            for j in range(0, n, bpl):
                fh.write(s[j:j + bpl] + "\n")


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_partitioned_runs_equal_single_gpu(tmp_path):
    n = min(_ngpu(), 4)
    lens = [90_000, 70_011, 65_000, 40_000, 33_333, 20_000, 7, 1_000]
    make_genome(tmp_path / "g.fa", lens, 1)
    common = ["-q", "--seed", "77", "args", "-sn", "0.01", "-in", "0.002", "-inmax", "9", "-de", "0.002", "-demax", "9", "-du", "0.001",
              "-dumax", "30", "-iv", "0.001", "-ivmax", "30", "-tl", "0.002", "-tlmax", "20"]
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    cli([str(tmp_path / "g.fa"), "-o", str(tmp_path / "a" / "g")] + common, 1, 0)
    cli([str(tmp_path / "g.fa"), "-o", str(tmp_path / "b" / "g")] + common, n, 29541)
    for f in ("g_ms.fa", "g_ms.vcf"):
        a, b = (tmp_path / "a" / f).read_bytes(), (tmp_path / "b" / f).read_bytes()
        if f.endswith(".vcf"):
            a = b"".join(l for l in a.splitlines(keepends=True) if not l.startswith(b"##filedate"))
            b = b"".join(l for l in b.splitlines(keepends=True) if not l.startswith(b"##filedate"))
        assert a == b, f
    # IT: several seeds so that at least one pairing straddles the GPUs; the partner's intervals reach the splice
    # kernel in place over NVLink (direct, the default), through a copy-engine pull, or over NCCL — same files
    for seed, route in ((1, "direct"), (2, "direct"), (3, "direct"), (2, "pull"), (3, "nccl")):
        for d in ("a", "b"):
            for f in (tmp_path / d).glob("*_it*"):
                f.unlink()
        it = ["-q", "--seed", str(seed), "it", "0.001"]
        cli([str(tmp_path / "g.fa"), "-o", str(tmp_path / "a" / "g")] + it, 1, 0)
        cli([str(tmp_path / "g.fa"), "-o", str(tmp_path / "b" / "g")] + it, n, 29542 + seed + (10 if route != "direct" else 0),
            MS_IT_EXCHANGE=route)
        for f in ("g_ms_it.fa", "g_ms_it.bedpe"):
            assert (tmp_path / "a" / f).read_bytes() == (tmp_path / "b" / f).read_bytes(), (seed, route, f)
        assert (tmp_path / "a" / "g_ms_it.bedpe").stat().st_size > 0


def _same(a: Path, b: Path, f: str):
    x, y = (a / f).read_bytes(), (b / f).read_bytes()
    if f.endswith(".vcf"):
        x = b"".join(l for l in x.splitlines(keepends=True) if not l.startswith(b"##filedate"))
        y = b"".join(l for l in y.splitlines(keepends=True) if not l.startswith(b"##filedate"))
    assert x == y, f


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_one_contig_is_cut_into_tiles_over_the_gpus(tmp_path):
    """A genome with fewer contigs than GPUs (README.md:441: one contig) is sharded at tile boundaries of the
    output (ms_apply_window); the assembled files equal the 1-GPU run.  Also: the same sharding forced on a
    many-contig genome, and IT mode with too few contigs for a contig partition."""
    n = min(_ngpu(), 8)
    common = ["-q", "--seed", "5", "args", "-sn", "0.01", "-in", "0.002", "-inmax", "9", "-de", "0.002", "-demax", "9", "-du", "0.001",
              "-dumax", "30", "-iv", "0.001", "-ivmax", "30", "-tl", "0.002", "-tlmax", "20"]
    for name, lens, env in (("one", [1_500_000], {}), ("many", [90_000, 70_011, 65_000, 40_000, 33_333, 20_000, 7, 1_000], {"MS_SHARD": "tiles"})):
        make_genome(tmp_path / f"{name}.fa", lens, 3, bpl=70)
        (tmp_path / f"{name}_a").mkdir(); (tmp_path / f"{name}_b").mkdir()
        cli([str(tmp_path / f"{name}.fa"), "-o", str(tmp_path / f"{name}_a" / "g")] + common, 1, 0)
        cli([str(tmp_path / f"{name}.fa"), "-o", str(tmp_path / f"{name}_b" / "g")] + common, n, 29561, **env)
        for f in ("g_ms.fa", "g_ms.vcf"):
            _same(tmp_path / f"{name}_a", tmp_path / f"{name}_b", f)
    make_genome(tmp_path / "three.fa", [300_000, 200_000, 100_000], 4)
    (tmp_path / "it_a").mkdir(); (tmp_path / "it_b").mkdir()
    it = ["-q", "--seed", "11", "it", "0.0005"]
    cli([str(tmp_path / "three.fa"), "-o", str(tmp_path / "it_a" / "g")] + it, 1, 0)
    cli([str(tmp_path / "three.fa"), "-o", str(tmp_path / "it_b" / "g")] + it, n, 29562, MS_SHARD="tiles")
    for f in ("g_ms_it.fa", "g_ms_it.bedpe"):
        _same(tmp_path / "it_a", tmp_path / "it_b", f)
    assert (tmp_path / "it_a" / "g_ms_it.bedpe").stat().st_size > 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_rmt_with_it_chained_in_hbm_on_several_gpus(tmp_path):
    """RMT mode with mutations AND interchromosomal translocations (__main__.py:88-102): every rank keeps the contigs
    it mutated in HBM and the IT step runs on them with the same partition; all four output files equal the 1-GPU
    run, and the run that reloads *_ms.fa instead (MS_NO_CHAIN=1)."""
    n = min(_ngpu(), 4)
    lens = [90_000, 70_011, 65_000, 40_000, 33_333, 20_000, 3_000, 1_000]
    make_genome(tmp_path / "g.fa", lens, 9)
    (tmp_path / "g.rmt").write_text("titv=2.0\nstd\nit 0.0008\nsn 0.01 in 0.002 inmin 1 inmax 9 de 0.002 demin 1 demax 9 tl 0.002 tlmin 2 tlmax 20\n\n"
                                    "chr 1\n1000-20000 None\n30001-60000 sn 0.05\nchr 3\n1-2000 None\n")
    runs = {"a": (1, {}), "b": (n, {}), "c": (n, {"MS_NO_CHAIN": "1"})}
    for d, (k, env) in runs.items():
        (tmp_path / d).mkdir()
        cli([str(tmp_path / "g.fa"), "-o", str(tmp_path / d / "g"), "-q", "--seed", "31", "rmt", str(tmp_path / "g.rmt")], k,
            29571 + ord(d), **env)
    for d in ("b", "c"):
        for f in ("g_ms.fa", "g_ms.vcf", "g_ms_it.fa", "g_ms_it.bedpe"):
            _same(tmp_path / "a", tmp_path / d, f)
    assert (tmp_path / "a" / "g_ms_it.bedpe").stat().st_size > 0
