"""Worker for tests/test_dist_cpu.py: the N>1 host logic of bench.py on the gloo backend."""
import json
import os
import sys
from pathlib import Path

import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def main():
    out = Path(sys.argv[1])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lengths = bench.GRCH38
    mine = bench.lpt_partition(lengths, world)[rank]
    load = float(sum(lengths[i] for i in mine))
    mx = bench.reduce_over_ranks([load, float(rank)], "max", world)
    sm = bench.reduce_over_ranks([load, float(len(mine))], "sum", world)
    dist.barrier()
    (out / f"rank{rank}.json").write_text(json.dumps({"mine": mine, "max": mx, "sum": sm}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
