#!/usr/bin/env python3
"""Pin the pair count of the reference's ITMutator.__assign_parters (it_mutator.py:59-70).

The reference iterates ``__avail_chroms`` while removing from the same list through the
alias ``remain_chr``; the loop therefore ends early and forms ceil(n/3) pairs, not
floor(n/2).  This script runs the UNMODIFIED reference method for n = 2..40 eligible
contigs and 200 seeds each and stores the (seed-independent) pair count and the number
of contigs left unpaired in tests/golden/it_pairs.json.

Run:  python tests/golden/make_it_pairs.py        (only where /root/reference exists)
"""
from __future__ import annotations

import json
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(REPO / "oracle" / "pyfaidx_standin"))

from mutation_simulator.it_mutator import ITMutator  # noqa: E402  (the reference)

class _Closed:
    def close(self):
        pass


out = {}
for n in range(2, 41):
    counts = set()
    for seed in range(200):
        random.seed(seed)
        it = ITMutator.__new__(ITMutator)
        it._ITMutator__fasta_writer = it._ITMutator__bedpe_writer = _Closed()   # __del__ closes them
        it._ITMutator__avail_chroms = list(range(n))
        it._ITMutator__assign_parters()
        partners = it._ITMutator__partners
        assert all(partners[partners[a]] == a and partners[a] != a for a in partners)
        counts.add(len(partners) // 2)
    assert len(counts) == 1, (n, counts)
    out[str(n)] = {"pairs": counts.pop()}
(HERE / "it_pairs.json").write_text(json.dumps(out, indent=0) + "\n")
print(out)
