#!/usr/bin/env python3
"""Run the REFERENCE's own unit-test suite (src/pyjess/tests of the PyJess checkout) UNCHANGED
against this repository's drop-in classes.

The tests import `from .._jess import Atom, Molecule, Template, Jess ...`; a three-line shim
package named `pyjess` re-exports the classes of `pyjess_b200` under those names.  Nothing of
the reference is stored in the repository: `--stage` copies the reference's test directory from
the read-only checkout into the untracked, git-ignored `.refshim_tmp/` (so that it travels to a
GPU box with `gpurun`), `--run` executes it there, `--clean` removes it.

    python tools/reference_suite.py --stage                  # in the build container
    gpurun -- 'python tools/reference_suite.py --run'        # on a B200 (query tests need the GPU)
    python tools/reference_suite.py --clean

Round 1 result on a B200: 82 tests, OK (11 skipped: gemmi / biopython / biotite are not
installed -- the same skips the reference itself reports in this image).

`--cif-shim` additionally stages an empty stand-in module named `gemmi`: the reference gates its
mmCIF tests on `import gemmi` (test_molecule.py:17-20,254-397), while this package reads mmCIF
without it (`pyjess_b200/_cif.py`), so with the stand-in those tests run UNCHANGED against the
gemmi-free loader; the one test that needs the real library (`test_from_gemmi`) skips itself
through the stand-in's `read_pdb_string`.  `test_molecule` needs no GPU:

    python tools/reference_suite.py --stage --cif-shim --run test_molecule

With the stand-in on a B200: 82 tests, OK, 3 skipped (biopython, biotite and `test_from_gemmi`).
"""
import argparse
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, ".refshim_tmp")
REF = os.environ.get("JESS_REFERENCE", "/root/reference")
EXPORTS = ("from pyjess_b200 import Atom, Mat4, Molecule, Template, TemplateAtom\n"
           "from pyjess_b200._jess import Jess, Query, Hit\n")
MODULES = ["test_atom", "test_molecule", "test_template", "test_template_atom", "test_jess", "test_hit"]


GEMMI_STANDIN = '''"""Stand-in for `gemmi` (not installed): lets the reference's gemmi-gated mmCIF tests run against
pyjess_b200's own reader.  Anything that needs the real library skips the test."""
import unittest


def read_pdb_string(*args, **kwargs):
    raise unittest.SkipTest("the real gemmi library is not installed")
'''


def stage(cif_shim=False):
    clean()
    os.makedirs(os.path.join(SHIM, "pyjess"))
    for name in ("__init__.py", "_jess.py"):
        with open(os.path.join(SHIM, "pyjess", name), "w") as f:
            f.write(EXPORTS + ('__version__ = "0.10.1"\n' if name == "__init__.py" else ""))
    if cif_shim:
        with open(os.path.join(SHIM, "gemmi.py"), "w") as f:
            f.write(GEMMI_STANDIN)
    shutil.copytree(os.path.join(REF, "src", "pyjess", "tests"), os.path.join(SHIM, "pyjess", "tests"), symlinks=False)


def run(modules):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIM, ROOT, os.environ.get("PYTHONPATH", "")]))
    return subprocess.call([sys.executable, "-m", "unittest"] + [f"pyjess.tests.{m}" for m in modules], cwd=SHIM, env=env)


def clean():
    shutil.rmtree(SHIM, ignore_errors=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--clean", action="store_true")
    ap.add_argument("--cif-shim", action="store_true", help="stage a stand-in `gemmi` module (see the module docstring)")
    ap.add_argument("modules", nargs="*", default=MODULES)
    args = ap.parse_args()
    rc = 0
    if args.stage:
        stage(args.cif_shim)
    if args.run:
        rc = run(args.modules)
    if args.clean:
        clean()
    sys.exit(rc)
