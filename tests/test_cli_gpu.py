"""The batched command line driver writes, byte for byte, what the reference's `pyjess` CLI
writes (goldens: `python -m pyjess ...` of the reference run from the repository root on the
list files in tests/golden/, see the commands below)."""
import io
import os

import pytest

from _util import GOLD
from pyjess_b200.cli import main

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LISTS = ["-T", "tests/golden/cli_templates.txt", "-Q", "tests/golden/cli_queries.txt"]

CASES = {
    # python -m pyjess -T ... -Q ... -R 2 -D 3 -M 3
    "cli_default.txt": ["-R", "2", "-D", "3", "-M", "3"],
    # ... -R 2 -D 4 -M 4 -b -n --no-reorder -c 5 -q --ignore-res-chain -e
    "cli_flags.txt": ["-R", "2", "-D", "4", "-M", "4", "-b", "-n", "--no-reorder", "-c", "5", "-q",
                      "--ignore-res-chain", "-e"],
    # ... -R 1.5 -D 4 -M 4 -c 3
    "cli_maxcand.txt": ["-R", "1.5", "-D", "4", "-M", "4", "-c", "3"],
}


@pytest.mark.parametrize("golden", sorted(CASES))
@pytest.mark.parametrize("batch", [256, 1])
def test_cli_output_is_byte_identical(engine, monkeypatch, golden, batch):
    monkeypatch.chdir(ROOT)
    out, err = io.StringIO(), io.StringIO()
    rc = main(LISTS + CASES[golden] + ["--batch", str(batch)], stdout=out, stderr=err)
    assert rc == 0, err.getvalue()
    with open(os.path.join(GOLD, golden)) as f:
        want = f.read()
    assert out.getvalue() == want


def test_cli_reports_errors_like_the_reference(engine, monkeypatch, tmp_path):
    monkeypatch.chdir(ROOT)
    missing = tmp_path / "q.txt"
    missing.write_text("does/not/exist.pdb\n")
    out, err = io.StringIO(), io.StringIO()
    rc = main(["-T", "tests/golden/cli_templates.txt", "-Q", str(missing), "-R", "2", "-D", "3", "-M", "3"],
              stdout=out, stderr=err)
    assert rc != 0 and err.getvalue().startswith("Error: ")


@pytest.mark.parametrize("golden", sorted(CASES))
@pytest.mark.parametrize("mode", ["all", "best"])
def test_cli_best_sites_matches_filter_output_script(engine, monkeypatch, golden, mode):
    """`--best-sites`: the text vendor/jess/filter_output.py writes when fed the reference
    command line's output (goldens: tests/golden/make_golden_sites.py)."""
    monkeypatch.chdir(ROOT)
    out, err = io.StringIO(), io.StringIO()
    rc = main(LISTS + CASES[golden] + ["--best-sites", mode], stdout=out, stderr=err)
    assert rc == 0, err.getvalue()
    with open(os.path.join(GOLD, "sites_" + golden.replace(".txt", "") + "_" + mode + ".txt")) as f:
        want = f.read()
    assert out.getvalue() == want
