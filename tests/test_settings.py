"""Differential test of the host-side settings model / RMT parser against outputs of
the reference's rmt.py (tests/golden/rmt/expected.json, made by make_golden.py)."""
import io
import json
import sys
from contextlib import redirect_stderr

import pytest

from mutation_simulator_b200 import rmt as R
from mutation_simulator_b200.fasta import Fasta
from tests.helpers import GOLDEN

EXPECTED = json.loads((GOLDEN / "rmt" / "expected.json").read_text())


@pytest.fixture(scope="module")
def fasta():
    return Fasta(GOLDEN / "rmt" / "g.fa", build_index=False)


def dump(sim):
    chroms = []
    for c in sim.chromosomes:
        rds = []
        for rd in c.range_definitions:
            ms = rd.mutation_settings
            rds.append({"start": rd.start, "stop": rd.stop,
                        "rates": None if ms.mut_rates is None else [[t.name, r] for t, r in ms.mut_rates.items()],
                        "chances": None if ms.mut_chances is None else [[t.name, r] for t, r in ms.mut_chances.items()],
                        "min": None if not ms.mut_lengs else [[t.name, v] for t, v in ms.mut_lengs["min"].items()],
                        "max": None if not ms.mut_lengs else [[t.name, v] for t, v in ms.mut_lengs["max"].items()],
                        "has_mutations": ms.has_mutations})
        chroms.append({"number": c.number, "it_rate": c.it_rate, "ranges": rds})
    return {"chromosomes": chroms, "mut_block": [[t.name, v] for t, v in sim.mut_block.items()], "fasta": sim.fasta,
            "md5": sim.md5, "titv": sim.titv, "species_name": sim.species_name, "assembly_name": sim.assembly_name,
            "sample_name": sim.sample_name, "has_mutations": sim.has_mutations, "has_it": sim.has_it}


def check(name, build, path_token=None):
    exp = EXPECTED[name]
    err = io.StringIO()
    try:
        with redirect_stderr(err):
            sim = build()
    except Exception as e:  # noqa: BLE001
        assert "error" in exp, f"{name}: unexpected {type(e).__name__}: {e}"
        assert type(e).__name__ == exp["error"]
        msg = str(e).replace(path_token, "<PATH>") if path_token else str(e)
        assert msg == exp["message"]
        return
    assert "ok" in exp, f"{name}: expected {exp.get('error')}"
    assert dump(sim) == exp["ok"]
    # The reference printed these while the goldens were generated (its print_warning binds sys.stderr at
    # import time, so make_golden's redirect could not capture them into expected.json; they were seen on
    # the console: rmt.py:336-341).
    assert err.getvalue() == WARNINGS.get(name, "")


WARNINGS = {"ok_meta": "WARNING: 'IN' block was set to 1\n", "ok_block_negative": "WARNING: 'SN' block was set to 1\n",
            "args_all": "WARNING: 'DE' block was set to 1\n"}
RMT_NAMES = sorted(p.stem for p in (GOLDEN / "rmt").glob("*.rmt"))


@pytest.mark.parametrize("name", RMT_NAMES)
def test_from_rmt_matches_reference(name, fasta):
    p = GOLDEN / "rmt" / f"{name}.rmt"
    check(name, lambda: R.SimulationSettings.from_rmt(p, fasta, False), str(p))


@pytest.mark.parametrize("name", [k for k in EXPECTED if k.startswith("args_")])
def test_from_args_matches_reference(name, fasta):
    from mutation_simulator_b200.argument_parser import get_args
    old = sys.argv
    sys.argv = ["x", str(GOLDEN / "rmt" / "g.fa"), "args"] + EXPECTED[name]["argv"]
    try:
        a = get_args()
    finally:
        sys.argv = old
    check(name, lambda: R.SimulationSettings.from_args(a, fasta, False))


@pytest.mark.parametrize("name", [k for k in EXPECTED if k.startswith("it_")])
def test_from_it_matches_reference(name, fasta):
    check(name, lambda: R.SimulationSettings.from_it(EXPECTED[name]["rate"], fasta, False))
